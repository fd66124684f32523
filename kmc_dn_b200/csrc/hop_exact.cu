// hop_exact.cu -- REPLAY hop loops: op-for-op device restatements of the reference's two CPU
// loops, for deterministic replay under an injected random stream (north-star check 1).
// Compiled with -fmad=false; every arithmetic step that the reference rounds is an explicit
// round-to-nearest intrinsic, so nothing is fused or re-associated.
//
//   MODE_GO_SIMULATE    goSimulation/simulation.go:194-325  (fp32, incremental makeJump energies)
//   MODE_GO_RECORDPLUS  goSimulation/simulation.go:327-432  (fp32, energies from scratch per hop)
//   MODE_PY             kmc_dopant_networks.py:33-135       (fp64 numba loop)
//
// One warp per trajectory.  Rates are evaluated lane-parallel into a per-member scratch list in
// global memory; the cumulative list is then built by ONE lane in the reference's sequential
// order (a float sum is order-dependent; this is the price of bit-exactness) and the event is
// found with a lane-parallel first-index search, which returns the same index as
// getRandomEvent's lower-bound walk (simulation.go:163-188) and as the linear scan at
// kmc_dopant_networks.py:107-110 because the list is non-decreasing.
// These kernels are validation paths: correctness over speed.
#include "kmc_internal.cuh"

namespace kmcb200 {

#define FULL 0xffffffffu

__device__ __forceinline__ bool transition_possible(int i, int j, int N, const uint8_t *occ) {
    // simulation.go:40-55 == kmc_dopant_networks.py:137-163
    if (i == j) return false;
    if (i >= N && j >= N) return false;
    if (i >= N) return !occ[j];
    if (j >= N) return occ[i];
    return occ[i] && !occ[j];
}

// first index k in [0,len) with pred(k) true, lane-parallel; -1 if none
template <typename Pred>
__device__ __forceinline__ int warp_first(int len, int lane, Pred pred) {
    for (int base = 0; base < len; base += 32) {
        const int k = base + lane;
        const bool hit = (k < len) && pred(k);
        const uint32_t bal = __ballot_sync(FULL, hit);
        if (bal) return base + __ffs(bal) - 1;
    }
    return -1;
}

// ------------------------------------------------------------------ Go semantics (fp32)
template <bool RECORDPLUS>
__global__ void __launch_bounds__(32) kmc_exact_go_kernel(const LayoutDev L, const EnsembleDev E) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = L.N, P = L.P, S = L.S, Lp = L.L;
    float *se = reinterpret_cast<float *>(smem_raw);       // [S] site energies
    float *E32 = se + S;                                   // [N] narrowed E_constant
    uint8_t *occ = reinterpret_cast<uint8_t *>(E32 + N);   // [N]
    const int lane = threadIdx.x;
    const int64_t m = blockIdx.x;
    const float nu = L.nu32, I_0 = L.I032, R = L.R32;
    const float kT = (float)E.kT[m];
    const float IR = __fmul_rn(I_0, R);
    const float *d = L.d32, *tc = L.tc32;
    float *list = reinterpret_cast<float *>(E.scratch + m * (int64_t)S * S);  // rates, then probList
    const int64_t total_hops = E.prehops + E.hops;

    for (int i = lane; i < N; i += 32) {
        occ[i] = E.occupation0 ? (E.occupation0[m * N + i] != 0) : 0;
        double e0;
        if (E.E_constant) e0 = E.E_constant[m * N + i];
        else {
            e0 = E.basis[(int64_t)P * N + i];
            for (int p = 0; p < P; ++p) e0 += E.electrode_v[m * P + p] * E.basis[(int64_t)p * N + i];
        }
        E32[i] = (float)e0;  // simulationWrapper.go:50-56
    }
    for (int i = N + lane; i < S; i += 32) se[i] = (float)E.electrode_v[m * P + (i - N)];
    __syncwarp();

    auto energies_from_scratch = [&]() {  // simulation.go:226-234 / :378-386
        for (int i = lane; i < N; i += 32) {
            float acc = 0.0f;
            for (int j = 0; j < N; ++j)
                if (j != i && !occ[j]) acc = __fadd_rn(acc, __fdiv_rn(1.0f, d[i * S + j]));
            se[i] = __fsub_rn(E32[i], __fmul_rn(IR, acc));
        }
        __syncwarp();
    };
    if (!RECORDPLUS) energies_from_scratch();

    double time = 0.0;
    double eo_mine = 0.0;  // lane p < P owns electrode p (electrode_occupation is float64 in Go)
    double occtime[8];     // lane owns acceptors lane, lane+32, ... (N <= 256)
#pragma unroll
    for (int q = 0; q < 8; ++q) occtime[q] = 0.0;

    for (int64_t h = 0; h < total_hops; ++h) {
        if (h == E.prehops) {
            time = 0.0;
            eo_mine = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) occtime[q] = 0.0;
        }
        if (RECORDPLUS) energies_from_scratch();
        // simulation.go:58-80
        for (int k = lane; k < Lp; k += 32) {
            const int from = L.pairs[k].x, to = L.pairs[k].y;
            float rate = 0.0f;
            if (transition_possible(from, to, N, occ)) {
                float dE;
                if (from < N && to < N)
                    dE = __fsub_rn(__fsub_rn(se[to], se[from]), __fdiv_rn(IR, d[from * S + to]));
                else
                    dE = __fsub_rn(se[to], se[from]);
                if (dE > 0.0f) rate = __fmul_rn(nu, (float)exp((double)__fdiv_rn(-dE, kT)));
                else rate = nu;
                rate = __fmul_rn(rate, tc[from * S + to]);
            }
            list[k] = rate;
        }
        __syncwarp();
        // simulation.go:270-276: sequential float32 running sum
        if (lane == 0) {
            float acc = 0.0f;
            for (int k = 0; k < Lp; ++k) {
                acc = (k == 0) ? list[0] : __fadd_rn(acc, list[k]);
                list[k] = acc;
            }
        }
        __syncwarp();
        const float total = list[Lp - 1];
        const double ek = E.stream_e[m * total_hops + h];
        const float uk = E.stream_u[m * total_hops + h];
        const double time_step = __ddiv_rn(ek, (double)total);  // :297
        time = __dadd_rn(time, time_step);
        const float eventRand = __fmul_rn(uk, total);           // :164
        int event = warp_first(Lp, lane, [&](int k) { return list[k] >= eventRand; });
        if (event < 0) event = Lp - 1;  // unreachable for finite lists; keeps indices in range
        const int from = L.pairs[event].x, to = L.pairs[event].y;

        if (h >= E.prehops) {
            if (E.trace && lane == 0) {
                int32_t *tp = E.trace + (m * E.hops + (h - E.prehops)) * 2;
                tp[0] = from;
                tp[1] = to;
            }
            if (!RECORDPLUS) {  // simulation.go:309-317 (simulateRecordPlus never records, :164-165)
                if (E.traffic && lane == 0) {
                    double *tr = E.traffic + m * (int64_t)S * S;
                    tr[from * S + to] += 1.0;
                    tr[to * S + from] -= 1.0;
                }
                for (int i = lane, q = 0; i < N; i += 32, ++q)
                    if (occ[i]) occtime[q] = __dadd_rn(occtime[q], time_step);
            }
        }
        __syncwarp();
        // hop: simulation.go:107-130 (makeJump) / :420-429
        if (from < N) {
            if (!RECORDPLUS)
                for (int j = lane; j < N; j += 32)
                    if (j != from) se[j] = __fsub_rn(se[j], __fmul_rn(IR, __fdiv_rn(1.0f, d[j * S + from])));
        } else if (lane == from - N) eo_mine -= 1.0;
        __syncwarp();
        if (lane == 0 && from < N) occ[from] = 0;
        if (to < N) {
            if (!RECORDPLUS)
                for (int j = lane; j < N; j += 32)
                    if (j != to) se[j] = __fadd_rn(se[j], __fmul_rn(IR, __fdiv_rn(1.0f, d[j * S + to])));
        } else if (lane == to - N) eo_mine += 1.0;
        __syncwarp();
        if (lane == 0 && to < N) occ[to] = 1;
        __syncwarp();
    }

    if (lane == 0) E.time[m] = time;
    if (lane < P) E.electrode_occ[m * P + lane] = (int64_t)eo_mine;
    for (int i = lane, q = 0; i < N; i += 32, ++q) {
        if (E.occupation_out) E.occupation_out[m * N + i] = occ[i];
        if (E.avg_occupation) E.avg_occupation[m * N + i] = occtime[q];
    }
    if (E.site_energies_out)
        for (int i = lane; i < S; i += 32) E.site_energies_out[m * S + i] = (double)se[i];
}

// ------------------------------------------------------------------ numba semantics (fp64)
__global__ void __launch_bounds__(32) kmc_exact_py_kernel(const LayoutDev L, const EnsembleDev E) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = L.N, P = L.P, S = L.S;
    double *se = reinterpret_cast<double *>(smem_raw);    // [S]
    double *Ec = se + S;                                  // [N]
    uint8_t *occ = reinterpret_cast<uint8_t *>(Ec + N);   // [N]
    const int lane = threadIdx.x;
    const int64_t m = blockIdx.x;
    const double nu = L.nu64, I_0 = L.I064, R = L.R64;
    const double kT = E.kT[m];
    const double *d = L.d64, *tc = L.tc64;
    double *list = E.scratch + m * (int64_t)S * S;
    const int64_t total_hops = E.prehops + E.hops;
    const int SS = S * S;

    for (int i = lane; i < N; i += 32) {
        occ[i] = E.occupation0 ? (E.occupation0[m * N + i] != 0) : 0;
        double e0;
        if (E.E_constant) e0 = E.E_constant[m * N + i];
        else {
            e0 = E.basis[(int64_t)P * N + i];
            for (int p = 0; p < P; ++p) e0 += E.electrode_v[m * P + p] * E.basis[(int64_t)p * N + i];
        }
        Ec[i] = e0;
    }
    for (int i = N + lane; i < S; i += 32) se[i] = E.electrode_v[m * P + (i - N)];
    __syncwarp();

    const double mIR = __dmul_rn(-I_0, R);  // (-I_0*R) of kmc_dopant_networks.py:66
    const double IR = __dmul_rn(I_0, R);    // I_0*R of :76
    double time = 0.0;
    long long eo_mine = 0;
    double occtime[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) occtime[q] = 0.0;

    for (int64_t h = 0; h < total_hops; ++h) {
        if (h == E.prehops) {
            time = 0.0;
            eo_mine = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) occtime[q] = 0.0;
        }
        // :58-66
        for (int i = lane; i < N; i += 32) {
            double acc = 0.0;
            for (int j = 0; j < N; ++j)
                if (j != i && !occ[j]) acc = __dadd_rn(acc, __ddiv_rn(1.0, d[i * S + j]));
            se[i] = __dadd_rn(Ec[i], __dmul_rn(mIR, acc));
        }
        __syncwarp();
        // :68-87
        for (int k = lane; k < SS; k += 32) {
            const int i = k / S, j = k - i * S;
            double t = 0.0;
            if (transition_possible(i, j, N, occ)) {
                double dE;
                if (i < N && j < N) dE = __dsub_rn(__dsub_rn(se[j], se[i]), __ddiv_rn(IR, d[k]));
                else dE = __dsub_rn(se[j], se[i]);
                if (dE > 0.0) t = __dmul_rn(nu, exp(__ddiv_rn(-dE, kT)));
                else t = nu;
            }
            list[k] = __dmul_rn(tc[k], t);
        }
        __syncwarp();
        // :91-97 sequential cumulative sum
        if (lane == 0) {
            double acc = list[0];
            for (int k = 1; k < SS; ++k) {
                acc = __dadd_rn(list[k], acc);
                list[k] = acc;
            }
        }
        __syncwarp();
        const double total = list[SS - 1];
        const double u1 = E.stream_u64[m * 2 * total_hops + 2 * h];
        const double u2 = E.stream_u64[m * 2 * total_hops + 2 * h + 1];
        const double hop_time = __dmul_rn(-log(__dsub_rn(1.0, u1)), __ddiv_rn(1.0, total));  // :100
        int event = warp_first(SS, lane, [&](int k) { return __ddiv_rn(list[k], total) >= u2; });  // :103-110
        if (event < 0) event = 0;
        const int from = event / S, to = event - from * S;  // :113
        __syncwarp();
        // :115-123
        if (from < N) { if (lane == 0) occ[from] = 0; }
        else if (lane == from - N) eo_mine -= 1;
        __syncwarp();
        if (to < N) { if (lane == 0) occ[to] = 1; }
        else if (lane == to - N) eo_mine += 1;
        __syncwarp();
        if (h >= E.prehops) {
            if (E.trace && lane == 0) {
                int32_t *tp = E.trace + (m * E.hops + (h - E.prehops)) * 2;
                tp[0] = from;
                tp[1] = to;
            }
            // :126-130 (post-hop occupation, plain count)
            if (E.traffic && lane == 0) E.traffic[m * (int64_t)SS + from * S + to] += 1.0;
            if (E.avg_occupation)
                for (int i = lane, q = 0; i < N; i += 32, ++q)
                    if (occ[i]) occtime[q] = __dadd_rn(occtime[q], hop_time);
        }
        time = __dadd_rn(time, hop_time);  // :133
    }

    if (lane == 0) E.time[m] = time;
    if (lane < P) E.electrode_occ[m * P + lane] = eo_mine;
    for (int i = lane, q = 0; i < N; i += 32, ++q) {
        if (E.occupation_out) E.occupation_out[m * N + i] = occ[i];
        if (E.avg_occupation) E.avg_occupation[m * N + i] = occtime[q];
    }
    if (E.site_energies_out)
        for (int i = lane; i < S; i += 32) E.site_energies_out[m * S + i] = se[i];
}

cudaError_t launch_exact(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    if (E.B <= 0) return cudaSuccess;
    if (L.N > 256 || L.P > 32) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)E.B;
    if (E.mode == 3) {
        const size_t smem = sizeof(double) * (L.S + L.N) + L.N + 16;
        kmc_exact_py_kernel<<<grid, 32, smem, st>>>(L, E);
    } else {
        const size_t smem = sizeof(float) * (L.S + L.N) + L.N + 16;
        if (E.mode == 2) kmc_exact_go_kernel<true><<<grid, 32, smem, st>>>(L, E);
        else kmc_exact_go_kernel<false><<<grid, 32, smem, st>>>(L, E);
    }
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace kmcb200
