// hop_reforder.cu -- REFERENCE-ORDER variant of the warp-per-trajectory hop loop (KMCB200_MODE_FAST_REFORDER).
// First-generation production kernel, kept because it preserves the reference's row-major event order and
// therefore follows the Go loop hop for hop under an injected stream (tests/test_gpu_parity.py).  The
// production kernel (hop_fast.cu) uses the same rate arithmetic with a cheaper lane-major event order.
//
// Reference semantics being accelerated (MUTUEL/kmc_dn, paths relative to the reference tree):
//   site energies      goSimulation/simulation.go:226-234  (E_const - I0*R*sum_{j empty} 1/d_ij)
//   incremental update goSimulation/simulation.go:107-130  (makeJump)
//   allowed pairs      goSimulation/simulation.go:40-55
//   Miller-Abrahams    goSimulation/simulation.go:58-80
//   cumulative list    goSimulation/simulation.go:267-276  (row-major (from,to) order -- kept here)
//   dwell time / pick  goSimulation/simulation.go:297-299, 163-188
//   tallies            goSimulation/simulation.go:306-319
//
// B200 design:
//   * rows of the rate matrix (the "from" sites) live on lanes: lane l owns rows l, l+32, ...
//     Only ALLOWED targets are visited: the empty acceptors (bit-loop over the warp-uniform
//     occupation mask) and the electrodes.  Disallowed pairs are never evaluated.
//   * layout table in shared memory, shared by all warps of the CTA, indexed [target][source]
//     as float2 {nu*tc, I0*R/d}; pitch 32*SLOTS+1 float2 => conflict-free for both the
//     row-parallel sweep and the column-parallel second-level pick.
//   * site energies: fp64 master copy per row, updated incrementally by +-(double)kd32 -- sums
//     of fp32 values in fp64 are exact here, so the incremental energy equals the from-scratch
//     energy (no drift, unlike simulation.go:113,124); rounded once per hop to fp32 for the rates.
//   * rates fp32 with MUFU.EX2; row sums fp32; prefix over rows, event pick and elapsed time fp64.
//   * two-level pick: warp-shuffle inclusive scan over row sums -> row, then the row's targets
//     are re-evaluated lane-parallel and scanned -> column.  Row-major order == reference order.
//   * Philox4x32-10, counter = (64-hop block, lane | member), key = seed: one call per lane
//     yields the two 32-bit variates of 64 hops for the whole warp.
#include "kmc_device.cuh"
#include "kmc_internal.cuh"

namespace kmcb200 {

template <int SLOTS, bool RECORD>
__global__ void __launch_bounds__(256) kmc_reforder_kernel(const LayoutDev L, const EnsembleDev E) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tbl = reinterpret_cast<float2 *>(smem_raw);
    const int N = L.N, P = L.P, S = L.S, pitch2 = L.pitch2;
    float *eps_base = reinterpret_cast<float *>(tbl + S * pitch2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    for (int i0 = 0; i0 < S * pitch2; i0 += blockDim.x)  // (thread-independent trip counts: see the staging loop of hop_lanes.cu)
        if (i0 + tid < S * pitch2) tbl[i0 + tid] = L.tbl[i0 + tid];
    __syncthreads();

    const int64_t m = (int64_t)blockIdx.x * nwarps + warp;
    if (m >= E.B) return;
    float *eps = eps_base + warp * (32 * SLOTS);

    // ---- static masks per row slot
    uint32_t accm[SLOTS], elm[SLOTS], occ[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int lo = 32 * k;
        accm[k] = (N >= lo + 32) ? ~0u : (N > lo ? ((1u << (N - lo)) - 1u) : 0u);
        const uint32_t sm = (S >= lo + 32) ? ~0u : (S > lo ? ((1u << (S - lo)) - 1u) : 0u);
        elm[k] = sm & ~accm[k];
    }

    // ---- initial state: occupation, E_constant (optionally by superposition), energies
    double eps64[SLOTS];
    float eps32[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int i = lane + 32 * k;
        bool o = false;
        double e0 = 0.0;
        if (i < N) {
            if (E.occupation0) o = E.occupation0[m * N + i] != 0;
            if (E.E_constant) e0 = E.E_constant[m * N + i];
            else {
                e0 = E.basis[(int64_t)P * N + i];
                for (int p = 0; p < P; ++p) e0 += E.electrode_v[m * P + p] * E.basis[(int64_t)p * N + i];
            }
            e0 = (double)(float)e0;  // simulationWrapper.go:50-56 narrows E_constant to float32
        } else if (i < S) {
            e0 = (double)(float)E.electrode_v[m * P + (i - N)];
        }
        occ[k] = __ballot_sync(FULL, o);
        eps64[k] = e0;
    }
#pragma unroll
    for (int kw = 0; kw < SLOTS; ++kw) {
        uint32_t mm = ~occ[kw] & accm[kw];
        while (mm) {
            const int j = kw * 32 + __ffs(mm) - 1;
            mm &= mm - 1;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) eps64[k] -= (double)tbl[j * pitch2 + lane + 32 * k].y;
        }
    }

    const float negbeta = (float)E.kT[m];  // kT (the Boltzmann factor takes kT itself: kmc_device.cuh boltz)
    const uint64_t gm = E.member_index0 + (uint64_t)m;
    const uint2 key = make_uint2((uint32_t)E.seed, (uint32_t)(E.seed >> 32));
    const bool inject = E.stream_e != nullptr;
    const int64_t total_hops = E.prehops + E.hops;

    uint4 rnd = make_uint4(0, 0, 0, 0);
    double t_acc = 0.0;
    int eoc = 0;
    double occtime[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) occtime[k] = 0.0;
    bool dead = false;

    for (int64_t h = 0; h < total_hops; ++h) {
        if (h == E.prehops) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
            t_acc = 0.0;
            eoc = 0;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) occtime[k] = 0.0;
        }
        // ---- publish fp32 energies to the warp's mirror
        __syncwarp();
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            eps32[k] = (float)eps64[k];
            eps[lane + 32 * k] = eps32[k];
        }
        __syncwarp();

        bool act[SLOTS];
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) act[k] = ((occ[k] | elm[k]) >> lane) & 1u;

        // ---- sweep: row sums over allowed targets (empty acceptors, then electrodes)
        float rs[SLOTS];
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) rs[k] = 0.0f;
#pragma unroll
        for (int kw = 0; kw < SLOTS; ++kw) {
            uint32_t mm = ~occ[kw] & accm[kw];
            while (mm) {
                const int j = kw * 32 + __ffs(mm) - 1;
                mm &= mm - 1;
                const float ej = eps[j];
                const float2 *row = tbl + j * pitch2 + lane;
#pragma unroll
                for (int k = 0; k < SLOTS; ++k) {
                    const float r = ma_rate(row[32 * k], ej, eps32[k], negbeta);
                    if (act[k]) rs[k] += r;
                }
            }
        }
        for (int e = 0; e < P; ++e) {
            const int j = N + e;
            const float ej = eps[j];
            const float2 *row = tbl + j * pitch2 + lane;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                if (32 * k < N) {  // slots holding only electrode rows have no electrode targets
                    const float r = ma_rate(row[32 * k], ej, eps32[k], negbeta);
                    if (act[k]) rs[k] += r;
                }
            }
        }

        // ---- first level: fp64 prefix over rows in row-major order
        double pre[SLOTS];
        double base = 0.0;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            pre[k] = warp_incl_scan((double)rs[k], lane) + base;
            base = __shfl_sync(FULL, pre[k], 31);
        }
        const double total = base;
        if (!(total > 0.0)) {  // no transition possible (simulation.go:297 would divide by zero)
            dead = true;
            break;
        }

        // ---- random variates
        double r_pick;
        float dt;
        if (!inject) {
            if ((h & 63) == 0) {
                const uint64_t blk = (uint64_t)(h >> 6) * 32u + (uint64_t)lane;
                rnd = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)gm, (uint32_t)(gm >> 32)), key);
            }
            const int q = (int)(h & 63);
            const uint32_t a = (q & 1) ? rnd.z : rnd.x;
            const uint32_t b = (q & 1) ? rnd.w : rnd.y;
            const uint32_t x1 = __shfl_sync(FULL, a, q >> 1);
            const uint32_t x2 = __shfl_sync(FULL, b, q >> 1);
            const float u1 = fmaf((float)x1, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (x+0.5)/2^32
            dt = (-0.6931471805599453f * lg2_approx(u1)) * rcp_approx((float)total);
            const double ts = total * 2.3283064365386963e-10;
            r_pick = fma((double)x2, ts, 0.5 * ts);
            t_acc += (double)dt;
        } else {
            const double ek = E.stream_e[m * total_hops + h];
            const float uk = E.stream_u[m * total_hops + h];
            const double dt64 = ek / total;  // simulation.go:297
            dt = (float)dt64;
            r_pick = (double)uk * total;     // simulation.go:164
            t_acc += dt64;
        }

        // ---- pick the row
        int wslot = -1, wlane = 0;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const uint32_t bal = __ballot_sync(FULL, pre[k] >= r_pick);
            if (wslot < 0 && bal) {
                wslot = k;
                wlane = __ffs(bal) - 1;
            }
        }
        if (wslot < 0) {  // r_pick rounded above total: take the last row with a positive sum
#pragma unroll
            for (int k = SLOTS - 1; k >= 0; --k) {
                const uint32_t bal = __ballot_sync(FULL, rs[k] > 0.0f);
                if (wslot < 0 && bal) {
                    wslot = k;
                    wlane = 31 - __clz(bal);
                }
            }
        }
        const int from = wslot * 32 + wlane;
        double excl = 0.0;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k)
            if (k == wslot) excl = pre[k] - (double)rs[k];
        const double rres = r_pick - __shfl_sync(FULL, excl, wlane);

        // ---- second level: re-evaluate row `from` lane-parallel over its targets
        const float e_from = eps[from];
        int to = -1;
        double accum = 0.0;
        uint32_t nzA[SLOTS], nzE = 0;
#pragma unroll
        for (int kw = 0; kw < SLOTS; ++kw) {
            nzA[kw] = 0;
            if (accm[kw]) {
                const int j = lane + 32 * kw;
                float rr = 0.0f;
                if (((~occ[kw] & accm[kw]) >> lane) & 1u) rr = ma_rate(tbl[j * pitch2 + from], eps32[kw], e_from, negbeta);
                nzA[kw] = __ballot_sync(FULL, rr > 0.0f);
                if (to < 0 && nzA[kw]) {
                    const double s = warp_incl_scan((double)rr, lane) + accum;
                    const uint32_t bal = __ballot_sync(FULL, s >= rres) & nzA[kw];
                    if (bal) to = kw * 32 + __ffs(bal) - 1;
                    accum = __shfl_sync(FULL, s, 31);
                }
            }
        }
        if (from < N) {
            float rr = 0.0f;
            if (lane < P) rr = ma_rate(tbl[(N + lane) * pitch2 + from], eps[N + lane], e_from, negbeta);
            nzE = __ballot_sync(FULL, rr > 0.0f);
            if (to < 0 && nzE) {
                const double s = warp_incl_scan((double)rr, lane) + accum;
                const uint32_t bal = __ballot_sync(FULL, s >= rres) & nzE;
                if (bal) to = N + __ffs(bal) - 1;
            }
        }
        if (to < 0) {  // residual rounded past the row's end: last target with a positive rate
            if (nzE) to = N + 31 - __clz(nzE);
            else {
#pragma unroll
                for (int kw = SLOTS - 1; kw >= 0; --kw)
                    if (to < 0 && nzA[kw]) to = kw * 32 + 31 - __clz(nzA[kw]);
            }
        }
        if (to < 0) {
            dead = true;
            break;
        }

        // ---- tallies (simulation.go:309-317: pre-hop occupation, antisymmetric traffic)
        if (RECORD && h >= E.prehops) {
            const double dtd = inject ? (E.stream_e[m * total_hops + h] / total) : (double)dt;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k)
                if ((occ[k] >> lane) & 1u) occtime[k] += dtd;
            if (lane == 0) {
                if (E.traffic) {
                    double *tr = E.traffic + m * (int64_t)S * S;
                    tr[from * S + to] += 1.0;
                    tr[to * S + from] -= 1.0;
                }
                if (E.trace) {
                    int32_t *tp = E.trace + (m * E.hops + (h - E.prehops)) * 2;
                    tp[0] = from;
                    tp[1] = to;
                }
            }
        }

        // ---- apply the hop (simulation.go:107-130)
        if (from < N) {
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                if (k == (from >> 5)) occ[k] &= ~(1u << (from & 31));
                eps64[k] -= (double)tbl[from * pitch2 + lane + 32 * k].y;
            }
        } else if (lane == from - N) eoc -= 1;
        if (to < N) {
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                if (k == (to >> 5)) occ[k] |= (1u << (to & 31));
                eps64[k] += (double)tbl[to * pitch2 + lane + 32 * k].y;
            }
        } else if (lane == to - N) eoc += 1;
    }

    // ---- results
    if (dead) t_acc = __longlong_as_double(0x7ff0000000000000LL);  // +inf, as time_step = e/0 would give
    if (lane == 0) E.time[m] = t_acc;
    if (lane < P) E.electrode_occ[m * P + lane] = (int64_t)eoc;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int i = lane + 32 * k;
        if (i < N) {
            if (E.occupation_out) E.occupation_out[m * N + i] = (occ[k] >> lane) & 1u;
            if (RECORD && E.avg_occupation) E.avg_occupation[m * N + i] = occtime[k];
        }
        if (i < S && E.site_energies_out) E.site_energies_out[m * S + i] = eps64[k];
    }
}

template <int SLOTS>
static cudaError_t launch_reforder_t(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    const bool record = E.avg_occupation || E.traffic || E.trace;
    // warps per CTA: large enough to amortise the table copy, small enough to balance small ensembles
    int warps = 8;
    while (warps > 1 && (E.B + warps - 1) / warps < 2 * 148) warps >>= 1;
    const int threads = warps * 32;
    const size_t smem = (size_t)L.S * L.pitch2 * sizeof(float2) + (size_t)warps * 32 * SLOTS * sizeof(float);
    const unsigned grid = (unsigned)((E.B + warps - 1) / warps);
    auto kern = record ? kmc_reforder_kernel<SLOTS, true> : kmc_reforder_kernel<SLOTS, false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, threads, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_reforder(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    if (E.B <= 0) return cudaSuccess;
    switch (L.slots) {
        case 1: return launch_reforder_t<1>(L, E, st, launches);
        case 2: return launch_reforder_t<2>(L, E, st, launches);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace kmcb200
