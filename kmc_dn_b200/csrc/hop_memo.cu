// hop_memo.cu -- production KMC hop loop with STATE MEMOISATION (KMCB200_MODE_FAST, N <= 32 acceptors).
//
// Same physics, arithmetic and event order as hop_fast.cu (read its header first).  What is added is the
// GPU-native counterpart of the reference's state cache (goSimulation/simulation.go:222-223, 251-296,
// 351-412): a trajectory revisits a few dozen occupation states over and over (C3: 16 states cover 80-99 %
// of 1e5 hops), and in this kernel -- exactly as in simulateRecordPlus, which recomputes the energies from
// scratch on a miss (simulation.go:378-386) -- the cumulative rate structure is a PURE function of the
// occupation bit-mask, because the fp64 incremental energies are exact.  So every warp keeps a small
// direct-mapped cache in shared memory:
//
//     key   = 32-bit occupation mask              (slot = multiplicative hash, 2^LOGK slots)
//     value = the fp64 inclusive prefix over the 32 lane sums (256 B: one double per lane, lane-private column)
//
// Hit:  the whole sweep (every allowed pair) and the fp64 scan are skipped; the hop costs one lookup, the
//       first-level ballot, the second-level re-evaluation of ONE lane's targets, and the state update.
// Miss: sweep + scan as in hop_fast.cu, then the prefix is parked in the slot.
// Memoising a pure function cannot change a result: with the cache disabled (LOGK = -1 instantiation,
// KMCB200_FLAG_NO_MEMO) the kernel produces bit-identical trajectories (tests/test_gpu_parity.py).
//
// The reference's cache stores the full per-pair list per state (len(transitions) floats, up to 150e6 floats
// per trajectory).  Here the second level is recomputed instead of stored: 256 B per state keeps 16 states
// per warp on chip for 32 resident warps per SM.
#include "kmc_device.cuh"
#include "kmc_internal.cuh"

namespace kmcb200 {

#define BIGS 1.0e30f

// Kogge-Stone steps with the add predicated on the shuffle's in-range flag (SHFL + @p ADD, no select).
__device__ __forceinline__ float scan_step_f(float v, int d) {
    asm volatile(
        "{ .reg .pred p; .reg .f32 t;\n"
        "  shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n"
        "  @p add.f32 %0, %0, t; }"
        : "+f"(v)
        : "r"(d));
    return v;
}
__device__ __forceinline__ double scan_step_d(double v, int d) {
    asm volatile(
        "{ .reg .pred p; .reg .b32 lo, hi, tlo, thi; .reg .f64 t;\n"
        "  mov.b64 {lo, hi}, %0;\n"
        "  shfl.sync.up.b32 tlo|p, lo, %1, 0, 0xffffffff;\n"
        "  shfl.sync.up.b32 thi|p, hi, %1, 0, 0xffffffff;\n"
        "  mov.b64 t, {tlo, thi};\n"
        "  @p add.f64 %0, %0, t; }"
        : "+d"(v)
        : "r"(d));
    return v;
}
template <int STEPS>
__device__ __forceinline__ float scan_f(float v) {
#pragma unroll
    for (int s = 0; s < STEPS; ++s) v = scan_step_f(v, 1 << s);
    return v;
}
__device__ __forceinline__ double scan_d(double v) {
#pragma unroll
    for (int s = 0; s < 5; ++s) v = scan_step_d(v, 1 << s);
    return v;
}

// first lane whose inclusive prefix reaches thr among lanes with a positive rate; if rounding put thr past the
// end, the last positive lane; -1 if the group is empty.  STEPS = log2(lanes that can be positive).
template <int STEPS>
__device__ __forceinline__ int pick_group(float rr, float thr) {
    const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
    if (!nz) return -1;
    const float s = scan_f<STEPS>(rr);
    const uint32_t bal = __ballot_sync(FULL, s >= thr) & nz;
    return bal ? (__ffs(bal) - 1) : (31 - __clz(nz));
}

template <int PT, int LOGK, bool DBG>
__global__ void __launch_bounds__(256) kmc_memo_kernel(const LayoutDev L, const EnsembleDev E) {
    constexpr int PITCH = 33;
    constexpr int K = LOGK >= 0 ? (1 << LOGK) : 0;
    constexpr int ESTEPS = (PT > 0 && PT <= 2) ? 1 : (PT > 0 && PT <= 4) ? 2 : (PT > 0 && PT <= 8) ? 3 : 5;
    // per-warp shared memory (bytes): mirror 64 f32 | rng 64 x uint2 | keys K u32 (>=16 B) | cache K x 32 f64
    constexpr int WARP_BYTES = 256 + 512 + (K > 4 ? K * 4 : 16) + K * 256;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tbl = reinterpret_cast<float2 *>(smem_raw);
    const int N = L.N, S = L.S;
    const int P = PT > 0 ? PT : L.P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    for (int idx = tid; idx < S * PITCH; idx += blockDim.x) tbl[idx] = L.tblf[idx];
    __syncthreads();

    const int64_t m = (int64_t)blockIdx.x * nwarps + warp;
    if (m >= E.B) return;
    unsigned char *wbase = smem_raw + (((size_t)S * PITCH * sizeof(float2) + 15) & ~size_t(15)) + (size_t)warp * WARP_BYTES;
    float *mir = reinterpret_cast<float *>(wbase);               // [0,32) acceptors, [32,64) electrodes
    uint2 *rngbuf = reinterpret_cast<uint2 *>(wbase + 256);
    uint32_t *keys = reinterpret_cast<uint32_t *>(wbase + 768);
    double *cache = reinterpret_cast<double *>(wbase + 768 + (K > 4 ? K * 4 : 16));

    const uint32_t accm = (N >= 32) ? ~0u : ((1u << N) - 1u);

    // ---- member parameters
    const float kT = (float)E.kT[m];
    const float nb = -1.4426950408889634f / kT;  // energies are carried as s = eps*nb
    const float pb = -nb;
    if (lane < P) mir[32 + lane] = (float)E.electrode_v[m * P + lane] * nb;
    __syncwarp();
    float se_reg[PT > 0 ? PT : 1];
    if (PT > 0) {
#pragma unroll
        for (int e = 0; e < (PT > 0 ? PT : 1); ++e) se_reg[e] = mir[32 + e];
    }
    const float se_mine = (lane < P) ? mir[32 + lane] : 0.0f;  // electrode `lane` (second level)

    // ---- initial state
    bool o0 = false;
    double eps64 = 0.0;
    if (lane < N) {
        if (E.occupation0) o0 = E.occupation0[m * N + lane] != 0;
        if (E.E_constant) eps64 = E.E_constant[m * N + lane];
        else {
            eps64 = E.basis[(int64_t)P * N + lane];
            for (int p = 0; p < P; ++p) eps64 += E.electrode_v[m * P + p] * E.basis[(int64_t)p * N + lane];
        }
        eps64 = (double)(float)eps64;  // simulationWrapper.go:50-56 narrows E_constant to float32
    }
    uint32_t occ = __ballot_sync(FULL, o0);
    {
        uint32_t mm = ~occ & accm;
        while (mm) {
            const int j = __ffs(mm) - 1;
            mm &= mm - 1;
            eps64 -= (double)tbl[j * PITCH + lane].y;
        }
    }

    const uint64_t gm = E.member_index0 + (uint64_t)m;
    const uint2 key = make_uint2((uint32_t)E.seed, (uint32_t)(E.seed >> 32));
    const bool inject = DBG && E.stream_e != nullptr;
    const int64_t total_hops = E.prehops + E.hops;
    // byte offsets of this lane's table entries
    const float2 *col_acc = tbl + lane * PITCH;        // + istar : pair (source istar -> target acceptor `lane`)
    const float2 *col_el = tbl + (N + lane) * PITCH;   // + istar : pair (acceptor istar <-> electrode `lane`)
    const float2 *row_me = tbl + lane;                 // + j*PITCH : pair (source `lane` -> target j)

    double t_acc = 0.0;
    float t_part = 0.0f;
    int eoc = 0;
    double occtime = 0.0;
    uint32_t valid = 0;
    bool dead = false;
    long long n_miss = 0;

    for (int64_t h0 = 0; h0 < total_hops && !dead; h0 += 64) {
        const int nq = (total_hops - h0 < 64) ? (int)(total_hops - h0) : 64;
        const int qreset = (E.prehops >= h0 && E.prehops < h0 + 64) ? (int)(E.prehops - h0) : -1;
        if (!inject) {
            // 64 hops' worth of variates: lane l serves hops 2l and 2l+1 of this block
            t_acc += (double)t_part;
            t_part = 0.0f;
            const uint64_t blk = (uint64_t)(h0 >> 6) * 32u + (uint64_t)lane;
            const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)gm, (uint32_t)(gm >> 32)), key);
            const float e0 = -0.6931471805599453f * lg2_approx(fmaf((float)r.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
            const float e1 = -0.6931471805599453f * lg2_approx(fmaf((float)r.z, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
            __syncwarp();
            reinterpret_cast<uint4 *>(rngbuf)[lane] = make_uint4(__float_as_uint(e0), r.y, __float_as_uint(e1), r.w);
        }
        for (int q = 0; q < nq; ++q) {
            if (q == qreset) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
                t_acc = 0.0;
                t_part = 0.0f;
                eoc = 0;
                occtime = 0.0;
            }
            // ---- publish scaled energies
            const float s_me = (float)eps64 * nb;
            __syncwarp();
            mir[lane] = s_me;
            __syncwarp();

            // ---- cumulative structure of this state: cached or computed
            double pre;
            const uint32_t slot = LOGK > 0 ? ((occ * 0x9E3779B1u) >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u;
            const bool hit = (K > 0) && ((valid >> slot) & 1u) && (keys[slot] == occ);
            if (hit) {
                pre = cache[slot * 32 + lane];
            } else {
                if (DBG) ++n_miss;
                const bool o = (occ >> lane) & 1u;
                const float src = o ? s_me : BIGS;       // only occupied acceptors emit to acceptors
                const float esig = o ? 1.0f : -1.0f;     // occupied: i->e, t = s_e - s_i ; empty: e->i, t = s_i - s_e
                const float ea = o ? -s_me : s_me;
                const float *erow = reinterpret_cast<const float *>(tbl + N * PITCH + lane) + (o ? 0 : 1);
                float rs = 0.0f;
                uint32_t mm = ~occ & accm;
                while (mm) {
                    const int j = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const float sj = mir[j];
                    const float2 v = row_me[j * PITCH];
                    const float t = fmaf(v.y, pb, sj - src);
                    rs = fmaf(v.x, ex2_approx(fminf(t, 0.0f)), rs);
                }
                if (PT > 0) {
#pragma unroll
                    for (int e = 0; e < (PT > 0 ? PT : 1); ++e) {
                        const float t = fmaf(esig, se_reg[e], ea);
                        rs = fmaf(erow[e * 2 * PITCH], ex2_approx(fminf(t, 0.0f)), rs);
                    }
                } else {
                    for (int e = 0; e < P; ++e) {
                        const float t = fmaf(esig, mir[32 + e], ea);
                        rs = fmaf(erow[e * 2 * PITCH], ex2_approx(fminf(t, 0.0f)), rs);
                    }
                }
                pre = scan_d((double)rs);
                if (K > 0) {
                    cache[slot * 32 + lane] = pre;
                    if (lane == 0) keys[slot] = occ;
                    valid |= 1u << slot;
                }
            }
            const double total = __shfl_sync(FULL, pre, 31);
            if (!(total > 0.0)) {  // no transition possible (simulation.go:297 would divide by zero)
                dead = true;
                break;
            }

            // ---- random variates
            double r_pick;
            double dtd = 0.0;
            if (!inject) {
                const uint2 rv = rngbuf[q];
                const float dt = __uint_as_float(rv.x) * rcp_approx((float)total);
                t_part += dt;
                if (DBG) dtd = (double)dt;
                const double ts = total * 2.3283064365386963e-10;
                r_pick = fma((double)rv.y, ts, 0.5 * ts);
            } else {
                dtd = E.stream_e[m * total_hops + h0 + q] / total;          // simulation.go:297
                r_pick = (double)E.stream_u[m * total_hops + h0 + q] * total;  // simulation.go:164
                t_acc += dtd;
            }

            // ---- first level: the lane
            uint32_t bal = __ballot_sync(FULL, pre >= r_pick);
            if (!bal) bal = __ballot_sync(FULL, pre >= total);  // threshold rounded past the end: last positive lane
            const int istar = __ffs(bal) - 1;
            const double pprev = __shfl_sync(FULL, pre, istar > 0 ? istar - 1 : 0);
            const float rf = (float)(r_pick - (istar > 0 ? pprev : 0.0));
            const bool rowocc = (occ >> istar) & 1u;
            const float s_star = mir[istar];

            // ---- second level: re-evaluate the winning lane's targets lane-parallel
            int from, to;
            if (rowocc) {
                from = istar;
                to = -1;
                int lastA = -1;
                float sA = 0.0f;
                const uint32_t emp = ~occ & accm;
                if (emp) {  // acceptor targets: istar -> empty `lane`
                    float rr = 0.0f;
                    if ((emp >> lane) & 1u) {
                        const float2 v = col_acc[istar];
                        rr = v.x * ex2_approx(fminf(fmaf(v.y, pb, s_me - s_star), 0.0f));
                    }
                    const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
                    if (nz) {
                        const float s = scan_f<5>(rr);
                        const uint32_t b2 = __ballot_sync(FULL, s >= rf) & nz;
                        if (b2) to = __ffs(b2) - 1;
                        else {
                            lastA = 31 - __clz(nz);
                            sA = __shfl_sync(FULL, s, 31);
                        }
                    }
                }
                if (to < 0) {  // electrode targets: istar -> electrode `lane`
                    float rr = 0.0f;
                    if (lane < P) rr = col_el[istar].x * ex2_approx(fminf(se_mine - s_star, 0.0f));
                    const int e = pick_group<ESTEPS>(rr, rf - sA);
                    to = (e >= 0) ? N + e : lastA;  // electrode group empty (rounding): last acceptor target
                }
                if (to < 0) {
                    dead = true;
                    break;
                }
            } else {  // empty acceptor: events electrode `lane` -> istar
                to = istar;
                float rr = 0.0f;
                if (lane < P) rr = col_el[istar].y * ex2_approx(fminf(s_star - se_mine, 0.0f));
                const int e = pick_group<ESTEPS>(rr, rf);
                if (e < 0) {
                    dead = true;
                    break;
                }
                from = N + e;
            }

            // ---- tallies (simulation.go:309-317: pre-hop occupation, antisymmetric traffic)
            if (DBG && h0 + q >= E.prehops) {
                if ((occ >> lane) & 1u) occtime += dtd;
                if (lane == 0) {
                    if (E.traffic) {
                        double *tr = E.traffic + m * (int64_t)S * S;
                        tr[from * S + to] += 1.0;
                        tr[to * S + from] -= 1.0;
                    }
                    if (E.trace) {
                        int32_t *tp = E.trace + (m * E.hops + (h0 + q - E.prehops)) * 2;
                        tp[0] = from;
                        tp[1] = to;
                    }
                }
            }

            // ---- apply the hop (simulation.go:107-130)
            if (from < N) {
                occ &= ~(1u << from);
                eps64 -= (double)row_me[from * PITCH].y;
            } else if (lane == from - N) eoc -= 1;
            if (to < N) {
                occ |= (1u << to);
                eps64 += (double)row_me[to * PITCH].y;
            } else if (lane == to - N) eoc += 1;
        }
    }

    // ---- results
    t_acc += (double)t_part;
    if (dead) t_acc = __longlong_as_double(0x7ff0000000000000LL);  // +inf, as time_step = e/0 would give
    if (lane == 0) E.time[m] = t_acc;
    if (lane < P) E.electrode_occ[m * P + lane] = (int64_t)eoc;
    if (lane < N) {
        if (E.occupation_out) E.occupation_out[m * N + lane] = (occ >> lane) & 1u;
        if (DBG && E.avg_occupation) E.avg_occupation[m * N + lane] = occtime;
        if (E.site_energies_out) E.site_energies_out[m * S + lane] = eps64;
    }
    if (E.site_energies_out && lane < P) E.site_energies_out[m * S + N + lane] = (double)(float)E.electrode_v[m * P + lane];
    if (DBG && E.misses && lane == 0) E.misses[m] = n_miss;
}

template <int PT, int LOGK>
static cudaError_t launch_memo_t(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    constexpr int K = LOGK >= 0 ? (1 << LOGK) : 0;
    constexpr int WARP_BYTES = 256 + 512 + (K > 4 ? K * 4 : 16) + K * 256;
    const bool dbg = E.avg_occupation || E.traffic || E.trace || E.stream_e || E.misses;
    int warps = 8;
    while (warps > 1 && (E.B + warps - 1) / warps < 2 * 148) warps >>= 1;
    const size_t smem = (((size_t)L.S * 33 * sizeof(float2) + 15) & ~size_t(15)) + (size_t)warps * WARP_BYTES;
    const unsigned grid = (unsigned)((E.B + warps - 1) / warps);
    auto kern = dbg ? kmc_memo_kernel<PT, LOGK, true> : kmc_memo_kernel<PT, LOGK, false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, warps * 32, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

template <int PT>
static cudaError_t launch_memo_p(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches) {
    switch (logk) {
        case -1: return launch_memo_t<PT, -1>(L, E, st, launches);
        case 3: return launch_memo_t<PT, 3>(L, E, st, launches);
        case 5: return launch_memo_t<PT, 5>(L, E, st, launches);
        default: return launch_memo_t<PT, 4>(L, E, st, launches);
    }
}

// logk: log2(cache slots per warp); -1 disables the memoisation (same code path, every hop a miss)
cudaError_t launch_memo(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches) {
    if (E.B <= 0) return cudaSuccess;
    if (L.N > 32 || L.P > 32) return cudaErrorInvalidValue;
    if (L.P == 8) return launch_memo_p<8>(L, E, logk, st, launches);
    if (L.P == 2) return launch_memo_p<2>(L, E, logk, st, launches);
    return launch_memo_p<0>(L, E, logk, st, launches);
}

}  // namespace kmcb200
