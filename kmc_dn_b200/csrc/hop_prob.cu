// hop_prob.cu -- mean-field "probability" solver (KMCB200_MODE_PROB): the deterministic pre-screen the
// reference's searches run before KMC (dn_search.py:49-52 strategy 0 -> wrapperSimulateProbability,
// goSimulation/simulationWrapper.go:218-233 -> probSimulate, probabilitySimulation.go:53-157).
//
// fp64 throughout, no random numbers.  One warp per ensemble member; lane i owns rows i, i+32, ...:
//   site energies   probabilitySimulation.go:84-93   (same j-order per row -> bit-identical sums)
//   rates           calcProbTransitions :8-38, probTransitionPossible :40-50
//   step limiter    :99-116 (max_change, time_step = 0.98*max_change/tot_rates)
//   update/clamp    :123-155
// The reference accumulates tot_rates / difference / occupation pair by pair in (i,j) order; here each lane sums
// its row and its column and the warp reduces, so results agree to fp64 rounding (tested to 1e-9), not bit for bit.
// Compiled with -fmad=false.
#include "kmc_internal.cuh"

namespace kmcb200 {

#define FULL 0xffffffffu

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, d));
    return v;
}

__global__ void __launch_bounds__(32) kmc_prob_kernel(const LayoutDev L, const EnsembleDev E) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = L.N, P = L.P, S = L.S;
    double *occ = reinterpret_cast<double *>(smem_raw);  // [N]
    double *se = occ + N;                                // [S]
    double *Ec = se + S;                                 // [N]
    const int lane = threadIdx.x;
    const int64_t m = blockIdx.x;
    const double nu = L.nu64, kT = E.kT[m], IR = __dmul_rn(L.I064, L.R64);
    const double *d = L.d64, *tc = L.tc64;
    double *T = E.scratch ? E.scratch + m * (int64_t)S * S : nullptr;  // pair rates of this step (record only)

    for (int i = lane; i < N; i += 32) {
        occ[i] = 0.5;  // simulationWrapper.go:226-228
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

    auto rate = [&](int i, int j) -> double {  // calcProbTransitions :14-29
        double base;
        if (i >= N && j >= N) base = 0.0;
        else if (i >= N) base = 1.0 - occ[j];
        else if (j >= N) base = occ[i];
        else base = __dmul_rn(1.0 - occ[j], occ[i]);
        double dE = __dsub_rn(se[j], se[i]);
        if (i < N && j < N) dE = __dsub_rn(dE, __ddiv_rn(IR, d[i * S + j]));
        const double t = (dE > 0.0) ? __dmul_rn(__dmul_rn(base, nu), exp(__ddiv_rn(-dE, kT))) : __dmul_rn(base, nu);
        return __dmul_rn(t, tc[i * S + j]);
    };

    double time = 0.0;
    double eo[8], avg[8];  // lane owns electrodes / acceptors lane, lane+32, ... (S <= 256+32)
#pragma unroll
    for (int q = 0; q < 8; ++q) eo[q] = avg[q] = 0.0;

    for (int64_t h = 0; h < E.hops; ++h) {
        for (int i = lane; i < N; i += 32) {  // :84-93
            double acc = 0.0;
            for (int j = 0; j < N; ++j)
                if (j != i) acc = __dadd_rn(acc, __ddiv_rn(1.0 - occ[j], d[i * S + j]));
            se[i] = __dsub_rn(Ec[i], __dmul_rn(IR, acc));
        }
        __syncwarp();
        double out[9], in[9];  // rows owned by this lane: i = lane + 32 q
        double tot = 0.0;
        int nq = 0;
        for (int i = lane; i < S; i += 32, ++nq) {
            double so = 0.0, si = 0.0;
            for (int j = 0; j < S; ++j) {
                const double tij = rate(i, j), tji = rate(j, i);
                so = __dadd_rn(so, tij);
                si = __dadd_rn(si, tji);
                if (T) T[i * S + j] = tij;
            }
            out[nq] = so; in[nq] = si;
            tot = __dadd_rn(tot, so);
        }
        tot = warp_sum(tot);
        double mc = 1.0;  // :99-114
        nq = 0;
        for (int i = lane; i < S; i += 32, ++nq) {
            if (i >= N) continue;
            const double diff = __ddiv_rn(__dsub_rn(in[nq], out[nq]), tot);
            const double nv = __dadd_rn(occ[i], diff);
            if (nv < 0.0) mc = fmin(mc, __ddiv_rn(occ[i], -diff));
            if (nv > 1.0) mc = fmin(mc, __ddiv_rn(1.0 - occ[i], diff));
        }
        mc = warp_min(mc);
        const double time_step = __ddiv_rn(__dmul_rn(0.98, mc), tot);
        time = __dadd_rn(time, time_step);
        const double f = __ddiv_rn(mc, tot);
        __syncwarp();
        nq = 0;
        for (int i = lane, qa = 0, qe = 0; i < S; i += 32, ++nq) {
            const double delta = __dmul_rn(__dsub_rn(in[nq], out[nq]), f);
            if (i < N) {
                if (E.avg_occupation) avg[qa] = __dadd_rn(avg[qa], __dmul_rn(occ[i], time_step));
                double v = __dadd_rn(occ[i], delta);
                occ[i] = fmin(fmax(v, 0.0), 1.0);  // :146-155
                ++qa;
            } else {
                eo[qe] = __dadd_rn(eo[qe], delta);
                ++qe;
            }
        }
        if (T && E.traffic) {  // :140-143 -> traffic[i][j] += rate_ij - rate_ji (electrode-electrode pairs skipped)
            double *tr = E.traffic + m * (int64_t)S * S;
            __syncwarp();
            for (int k = lane; k < S * S; k += 32) {
                const int i = k / S, j = k - i * S;
                if (i >= N && j >= N) continue;
                tr[k] = __dadd_rn(tr[k], __dmul_rn(__dsub_rn(T[k], T[j * S + i]), f));
            }
        }
        __syncwarp();
    }

    if (lane == 0) E.time[m] = time;
    for (int i = lane, qa = 0, qe = 0; i < S; i += 32) {
        if (i < N) {
            if (E.prob_occupation) E.prob_occupation[m * N + i] = occ[i];
            if (E.avg_occupation) E.avg_occupation[m * N + i] = avg[qa];
            ++qa;
        } else {
            if (E.prob_electrode_occ) E.prob_electrode_occ[m * P + (i - N)] = eo[qe];
            ++qe;
        }
        if (E.site_energies_out) E.site_energies_out[m * S + i] = se[i];
    }
}

cudaError_t launch_prob(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    if (E.B <= 0) return cudaSuccess;
    if (L.S > 288) return cudaErrorInvalidValue;
    const size_t smem = sizeof(double) * (size_t)(2 * L.N + L.S);
    kmc_prob_kernel<<<(unsigned)E.B, 32, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace kmcb200
