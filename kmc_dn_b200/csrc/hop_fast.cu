// hop_fast.cu -- production KMC hop loop for sm_100a (KMCB200_MODE_FAST).  One warp = one trajectory.
//
// Reference semantics being accelerated (MUTUEL/kmc_dn, paths relative to the reference tree):
//   site energies      goSimulation/simulation.go:226-234  (E_const - I0*R*sum_{j empty} 1/d_ij)
//   incremental update goSimulation/simulation.go:107-130  (makeJump)
//   allowed pairs      goSimulation/simulation.go:40-55
//   Miller-Abrahams    goSimulation/simulation.go:58-80
//   cumulative list    goSimulation/simulation.go:267-276
//   dwell time / pick  goSimulation/simulation.go:297-299, 163-188
//   tallies            goSimulation/simulation.go:306-319
//
// B200 design.  The v1 kernel (hop_reforder.cu) was issue-bound at 589
// warp-instructions per hop (profiles/ncu_fast_r01_v1_details.txt); this one does the same physics in
// roughly a third of that:
//   * lanes own ACCEPTORS only (lane l <-> acceptor l, l+32, ...).  Every allowed pair is evaluated exactly
//     once per hop and nothing else is:
//       - acceptor->acceptor: warp-uniform bit-loop over the EMPTY sites j; lane i (occupied) evaluates i->j;
//       - acceptor<->electrode: loop over electrodes e; lane i evaluates i->e if it is occupied and e->i if it
//         is empty (exactly one of the two directions is allowed for every (i,e)), so all lanes work.
//     The event order of the cumulative list is therefore lane-major instead of the reference's row-major
//     (electrode->acceptor events sit in the acceptor's lane).  Any fixed order samples the same Markov
//     chain; the row-major variant is kept in hop_reforder.cu for lock-step replay against the oracle.
//   * a rate is tc * exp2(min(0, ((e_to - e_from) - kd) * (-log2e/kT))), in the reference's operation order
//     (simulation.go:66-77).  Rows that may not act as a source carry e_from = -1e30 instead of a predicate.
//   * table in shared memory, [target][source] float2: {nu*tc(i->j), I0*R/d_ij} for acceptor targets,
//     {nu*tc(i->e), nu*tc(e->i)} for electrode targets; pitch 32*AS+1 float2 (conflict-free row- and
//     column-wise).  fp32 narrowing as the cgo wrappers do (simulationWrapper.go:37-56).
//   * energies: fp64 master per acceptor, updated by +-(double)kd32.  Sums of fp32 values in fp64 are exact
//     here, so the incremental energy equals the from-scratch energy bit for bit (no drift, unlike
//     simulation.go:113,124) and the rate list is a pure function of the occupation.
//   * rates and within-lane sums fp32; prefix over lanes, event threshold and elapsed time fp64.
//   * two-level pick: fp64 warp-shuffle scan over lane sums -> lane; the lane's targets are re-evaluated
//     lane-parallel and scanned (fp32) -> target.
//   * Philox4x32-10, counter = (64-hop block | lane, member), key = seed.  Once per 64 hops every lane turns
//     one Philox call into two (unit-exponential, 32-bit uniform) pairs and parks them in shared memory, so
//     the per-hop cost of the generator and of the logarithm is one broadcast LDS.64.
#include "kmc_device.cuh"
#include "kmc_internal.cuh"

namespace kmcb200 {

#define BIGS 1.0e30f

__device__ __forceinline__ float warp_incl_scan_f(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// first lane whose inclusive prefix reaches thr among lanes with a positive rate; if rounding put thr past
// the end, the last positive lane; -1 if the group is empty.
__device__ __forceinline__ int pick_in_group(float rr, float thr, int lane) {
    const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
    if (!nz) return -1;
    const float s = warp_incl_scan_f(rr, lane);
    const uint32_t bal = __ballot_sync(FULL, s >= thr) & nz;
    return bal ? (__ffs(bal) - 1) : (31 - __clz(nz));
}

template <int AS, int PT, bool DBG, bool GT>
__global__ void __launch_bounds__(256) kmc_fast_kernel(const LayoutDev L, const EnsembleDev E) {
    constexpr int PITCH = 32 * AS + 1;
    constexpr int MIRW = 32 * AS + 32;  // per-warp mirror: acceptor energies [0,32*AS), electrode energies after
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // GT: the pair table stays in global memory (L1/L2-resident; N > 64: 2*S^2 floats exceed shared memory)
    const float2 *tbl = GT ? L.tblf : reinterpret_cast<const float2 *>(smem_raw);
    const int N = L.N, S = L.S;
    const int P = PT > 0 ? PT : L.P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float *mir_base = reinterpret_cast<float *>(smem_raw + (GT ? 0 : (((size_t)S * PITCH * sizeof(float2) + 15) & ~size_t(15))));
    uint2 *rng_base = reinterpret_cast<uint2 *>(mir_base + nwarps * MIRW);

    if (!GT) {
        float2 *stage = reinterpret_cast<float2 *>(smem_raw);
        for (int i0 = 0; i0 < S * PITCH; i0 += blockDim.x)  // (thread-independent trip counts: see the staging loop of hop_lanes.cu)
            if (i0 + tid < S * PITCH) stage[i0 + tid] = L.tblf[i0 + tid];
    }
    __syncthreads();

    const int64_t m = (int64_t)blockIdx.x * nwarps + warp;
    if (m >= E.B) return;
    float *mir = mir_base + warp * MIRW;
    uint2 *rngbuf = rng_base + warp * 64;

    // ---- static masks
    uint32_t accm[AS], occ[AS];
#pragma unroll
    for (int k = 0; k < AS; ++k) {
        const int lo = 32 * k;
        accm[k] = (N >= lo + 32) ? ~0u : (N > lo ? ((1u << (N - lo)) - 1u) : 0u);
    }

    // ---- member parameters
    const float kT = (float)E.kT[m];
    const float nb = kT;  // (the Boltzmann factor takes kT itself: kmc_device.cuh boltz)
    float se_reg[PT > 0 ? PT : 1];
    if (lane < P) mir[32 * AS + lane] = (float)E.electrode_v[m * P + lane];
    __syncwarp();
    if (PT > 0) {
#pragma unroll
        for (int e = 0; e < (PT > 0 ? PT : 1); ++e) se_reg[e] = mir[32 * AS + e];
    }

    // ---- initial state: occupation, E_constant (optionally by superposition), energies
    double eps64[AS];
#pragma unroll
    for (int k = 0; k < AS; ++k) {
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
        }
        occ[k] = __ballot_sync(FULL, o);
        eps64[k] = e0;
    }
#pragma unroll
    for (int kw = 0; kw < AS; ++kw) {
        uint32_t mm = ~occ[kw] & accm[kw];
        while (mm) {
            const int j = kw * 32 + __ffs(mm) - 1;
            mm &= mm - 1;
#pragma unroll
            for (int k = 0; k < AS; ++k) eps64[k] -= (double)tbl[j * PITCH + lane + 32 * k].y;
        }
    }

    const uint64_t gm = E.member_index0 + (uint64_t)m;
    const uint2 key = make_uint2((uint32_t)E.seed, (uint32_t)(E.seed >> 32));
    const bool inject = DBG && E.stream_e != nullptr;
    const int64_t total_hops = E.prehops + E.hops;

    double t_acc = 0.0;
    float t_part = 0.0f;
    int eoc = 0;
    double occtime[AS];
#pragma unroll
    for (int k = 0; k < AS; ++k) occtime[k] = 0.0;
    bool dead = false;

    for (int64_t h = 0; h < total_hops; ++h) {
        if (h == E.prehops) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
            t_acc = 0.0;
            t_part = 0.0f;
            eoc = 0;
#pragma unroll
            for (int k = 0; k < AS; ++k) occtime[k] = 0.0;
        }
        if (!inject && (h & 63) == 0) {
            // 64 hops' worth of variates: lane l serves hops 2l and 2l+1 of this block
            t_acc += (double)t_part;
            t_part = 0.0f;
            const uint64_t blk = (uint64_t)(h >> 6) * 32u + (uint64_t)lane;
            const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)gm, (uint32_t)(gm >> 32)), key);
            // unit exponential from (x+0.5)/2^32
            const float e0 = -0.6931471805599453f * lg2_approx(fmaf((float)r.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
            const float e1 = -0.6931471805599453f * lg2_approx(fmaf((float)r.z, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
            __syncwarp();
            reinterpret_cast<uint4 *>(rngbuf)[lane] = make_uint4(__float_as_uint(e0), r.y, __float_as_uint(e1), r.w);
        }

        // ---- publish scaled energies; per-lane source terms
        float s_true[AS], src[AS], nbs[AS];
        const float *erow[AS];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < AS; ++k) {
            const bool o = (occ[k] >> lane) & 1u;
            s_true[k] = (float)eps64[k];
            mir[lane + 32 * k] = s_true[k];
            src[k] = o ? s_true[k] : -BIGS;         // only occupied acceptors emit to acceptors
            nbs[k] = o ? 1.0f : -1.0f;              // occupied: i->e, dE = V_e - e_i ; empty: e->i, dE = e_i - V_e
            erow[k] = reinterpret_cast<const float *>(tbl + N * PITCH + lane + 32 * k) + (o ? 0 : 1);
        }
        __syncwarp();

        // ---- sweep: every allowed pair exactly once
        float rsA[AS], rsE[AS];
#pragma unroll
        for (int k = 0; k < AS; ++k) rsA[k] = rsE[k] = 0.0f;
#pragma unroll
        for (int kw = 0; kw < AS; ++kw) {
            uint32_t mm = ~occ[kw] & accm[kw];
            while (mm) {
                const int j = kw * 32 + __ffs(mm) - 1;
                mm &= mm - 1;
                const float sj = mir[j];
                const float2 *row = tbl + j * PITCH + lane;
#pragma unroll
                for (int k = 0; k < AS; ++k) {
                    const float2 v = row[32 * k];
                    rsA[k] += ma_rate(v, sj, src[k], nb);
                }
            }
        }
        if (PT > 0) {
#pragma unroll
            for (int e = 0; e < (PT > 0 ? PT : 1); ++e) {
#pragma unroll
                for (int k = 0; k < AS; ++k) {
                    const float t = (se_reg[e] - s_true[k]) * nbs[k];
                    rsE[k] = fmaf(erow[k][e * 2 * PITCH], boltz(t, nb), rsE[k]);
                }
            }
        } else {
            for (int e = 0; e < P; ++e) {
                const float se = mir[32 * AS + e];
#pragma unroll
                for (int k = 0; k < AS; ++k) {
                    const float t = (se - s_true[k]) * nbs[k];
                    rsE[k] = fmaf(erow[k][e * 2 * PITCH], boltz(t, nb), rsE[k]);
                }
            }
        }

        // ---- first level: fp64 prefix over lanes (lane-major event order)
        float rs[AS];
        double pre[AS];
        double base = 0.0;
#pragma unroll
        for (int k = 0; k < AS; ++k) {
            rs[k] = rsA[k] + rsE[k];
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
        double dtd = 0.0;
        if (!inject) {
            const uint2 rv = rngbuf[h & 63];
            const float dt = __uint_as_float(rv.x) * rcp_approx((float)total);
            t_part += dt;
            if (DBG) dtd = (double)dt;
            const double ts = total * 2.3283064365386963e-10;
            r_pick = fma((double)rv.y, ts, 0.5 * ts);
        } else {
            dtd = E.stream_e[m * total_hops + h] / total;          // simulation.go:297
            r_pick = (double)E.stream_u[m * total_hops + h] * total;  // simulation.go:164
            t_acc += dtd;
        }

        // ---- pick the lane
        int wslot = -1, wlane = 0;
#pragma unroll
        for (int k = 0; k < AS; ++k) {
            const uint32_t bal = __ballot_sync(FULL, pre[k] >= r_pick);
            if (wslot < 0 && bal) {
                wslot = k;
                wlane = __ffs(bal) - 1;
            }
        }
        if (wslot < 0) {  // r_pick rounded above total: last lane with a positive sum
#pragma unroll
            for (int k = AS - 1; k >= 0; --k) {
                const uint32_t bal = __ballot_sync(FULL, rs[k] > 0.0f);
                if (wslot < 0 && bal) {
                    wslot = k;
                    wlane = 31 - __clz(bal);
                }
            }
        }
        const int istar = wslot * 32 + wlane;
        float rf_mine = 0.0f, rsA_mine = 0.0f;
        bool rowocc = false;
#pragma unroll
        for (int k = 0; k < AS; ++k)
            if (k == wslot) {
                rf_mine = (float)(r_pick - (pre[k] - (double)rs[k]));
                rsA_mine = rsA[k];
                rowocc = (occ[k] >> wlane) & 1u;
            }
        const float rf = __shfl_sync(FULL, rf_mine, wlane);
        const float s_star = mir[istar];

        // ---- second level: re-evaluate the winning lane's targets lane-parallel
        int from, to;
        const float *ecol = reinterpret_cast<const float *>(tbl + (N + lane) * PITCH + istar);  // electrode `lane` vs istar
        if (rowocc) {
            from = istar;
            to = -1;
            const float rsA_star = __shfl_sync(FULL, rsA_mine, wlane);
            const bool tryA = rf < rsA_star;
            if (tryA) {  // acceptor targets: istar -> empty j
                float thr = rf;
                int lastpos = -1;
#pragma unroll
                for (int kw = 0; kw < AS; ++kw) {
                    if (to < 0 && (~occ[kw] & accm[kw])) {
                        float rr = 0.0f;
                        if (((~occ[kw] & accm[kw]) >> lane) & 1u) {
                            const float2 v = tbl[(lane + 32 * kw) * PITCH + istar];
                            rr = ma_rate(v, s_true[kw], s_star, nb);
                        }
                        const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
                        if (nz) {
                            lastpos = kw * 32 + 31 - __clz(nz);
                            const float s = warp_incl_scan_f(rr, lane);
                            const uint32_t bal = __ballot_sync(FULL, s >= thr) & nz;
                            if (bal) to = kw * 32 + __ffs(bal) - 1;
                            thr -= __shfl_sync(FULL, s, 31);
                        }
                    }
                }
                if (to < 0) to = lastpos;  // residual rounded past the group's end: its last positive target
            }
            if (to < 0) {  // electrode targets: istar -> e   (or acceptor group empty / exhausted)
                float rr = 0.0f;
                if (lane < P) rr = ecol[0] * boltz(mir[32 * AS + lane] - s_star, nb);
                const int e = pick_in_group(rr, tryA ? BIGS : rf - rsA_star, lane);
                if (e >= 0) to = N + e;
            }
            if (to < 0 && !tryA) {  // electrode group empty although rf >= rsA (rounding): last acceptor target
#pragma unroll
                for (int kw = AS - 1; kw >= 0; --kw) {
                    if (to < 0 && (~occ[kw] & accm[kw])) {
                        float rr = 0.0f;
                        if (((~occ[kw] & accm[kw]) >> lane) & 1u) {
                            const float2 v = tbl[(lane + 32 * kw) * PITCH + istar];
                            rr = ma_rate(v, s_true[kw], s_star, nb);
                        }
                        const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
                        if (nz) to = kw * 32 + 31 - __clz(nz);
                    }
                }
            }
            if (to < 0) {
                dead = true;
                break;
            }
        } else {  // empty acceptor: events e -> istar
            to = istar;
            float rr = 0.0f;
            if (lane < P) rr = ecol[1] * boltz(s_star - mir[32 * AS + lane], nb);
            const int e = pick_in_group(rr, rf, lane);
            if (e < 0) {
                dead = true;
                break;
            }
            from = N + e;
        }

        // ---- tallies (simulation.go:309-317: pre-hop occupation, antisymmetric traffic)
        if (DBG && h >= E.prehops) {
#pragma unroll
            for (int k = 0; k < AS; ++k)
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
            for (int k = 0; k < AS; ++k) {
                if (k == (from >> 5)) occ[k] &= ~(1u << (from & 31));
                eps64[k] -= (double)tbl[from * PITCH + lane + 32 * k].y;
            }
        } else if (lane == from - N) eoc -= 1;
        if (to < N) {
#pragma unroll
            for (int k = 0; k < AS; ++k) {
                if (k == (to >> 5)) occ[k] |= (1u << (to & 31));
                eps64[k] += (double)tbl[to * PITCH + lane + 32 * k].y;
            }
        } else if (lane == to - N) eoc += 1;
    }

    // ---- results
    t_acc += (double)t_part;
    if (dead) t_acc = __longlong_as_double(0x7ff0000000000000LL);  // +inf, as time_step = e/0 would give
    if (lane == 0) E.time[m] = t_acc;
    if (lane < P) E.electrode_occ[m * P + lane] = (int64_t)eoc;
#pragma unroll
    for (int k = 0; k < AS; ++k) {
        const int i = lane + 32 * k;
        if (i < N) {
            if (E.occupation_out) E.occupation_out[m * N + i] = (occ[k] >> lane) & 1u;
            if (DBG && E.avg_occupation) E.avg_occupation[m * N + i] = occtime[k];
            if (E.site_energies_out) E.site_energies_out[m * S + i] = eps64[k];
        }
    }
    if (E.site_energies_out && lane < P) E.site_energies_out[m * S + N + lane] = (double)(float)E.electrode_v[m * P + lane];
}

template <int AS, int PT, bool GT>
static cudaError_t launch_fast_t(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    const bool dbg = E.avg_occupation || E.traffic || E.trace || E.stream_e;
    // warps per CTA: large enough to amortise the table copy, small enough to balance small ensembles
    int warps = AS >= 4 ? 4 : 8;
    while (warps > 1 && (E.B + warps - 1) / warps < 2 * 148) warps >>= 1;
    const int threads = warps * 32;
    const size_t smem = (GT ? 0 : (((size_t)L.S * (32 * AS + 1) * sizeof(float2) + 15) & ~size_t(15))) +
                        (size_t)warps * (32 * AS + 32) * sizeof(float) + (size_t)warps * 64 * sizeof(uint2);
    const unsigned grid = (unsigned)((E.B + warps - 1) / warps);
    auto kern = dbg ? kmc_fast_kernel<AS, PT, true, GT> : kmc_fast_kernel<AS, PT, false, GT>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, threads, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// N <= 256 acceptors.  (N <= 32 normally runs the memoised kernel of hop_memo.cu; this one is the general path.)
cudaError_t launch_fast(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    if (E.B <= 0) return cudaSuccess;
    const int as = (L.N + 31) / 32;
    if (L.pitchf != 32 * (as == 3 ? 4 : (as > 4 ? 8 : as)) + 1) return cudaErrorInvalidValue;
    if (as <= 1) {
        if (L.P == 8) return launch_fast_t<1, 8, false>(L, E, st, launches);
        if (L.P == 2) return launch_fast_t<1, 2, false>(L, E, st, launches);
        return launch_fast_t<1, 0, false>(L, E, st, launches);
    }
    if (as == 2) {
        if (L.P == 8) return launch_fast_t<2, 8, false>(L, E, st, launches);
        return launch_fast_t<2, 0, false>(L, E, st, launches);
    }
    if (as <= 4) return launch_fast_t<4, 0, true>(L, E, st, launches);
    if (as <= 8) return launch_fast_t<8, 0, true>(L, E, st, launches);
    return cudaErrorInvalidValue;
}

}  // namespace kmcb200
