// hop_wide.cu -- production KMC hop loop with STATE MEMOISATION for 32 <= N <= 256 acceptors (KMCB200_MODE_FAST).
//
// Reference semantics being accelerated (MUTUEL/kmc_dn, paths relative to the reference tree):
//   site energies      goSimulation/simulation.go:226-234  (E_const - I0*R*sum_{j empty} 1/d_ij)
//   incremental update goSimulation/simulation.go:107-130  (makeJump)
//   allowed pairs      goSimulation/simulation.go:40-55
//   Miller-Abrahams    goSimulation/simulation.go:58-80
//   cumulative list    goSimulation/simulation.go:267-276
//   dwell time / pick  goSimulation/simulation.go:297-299, 163-188
//   tallies            goSimulation/simulation.go:306-319
//   state cache        goSimulation/simulation.go:222-223, 251-296, 351-412
//
// One warp = one trajectory; lane l owns the AS acceptors l, l+32, ... (hop_fast.cu's layout, sweep and exact
// two-level pick -- read its header first).  What this kernel adds is hop_memo.cu's memoisation, generalised to
// multi-word occupation masks.  On the examples/scaling.py layout (N = 256) 95-97 % of the hops land in a state the
// trajectory has seen before, and the largest event of every lane carries > 99 % of the total rate, so per state
// the cache keeps
//
//     pre[32]   NORMALISED fp64 exclusive prefix over the lanes' TOP events (+inf for a lane without one)
//     aux[32]   the lane's top event: partner (9 bits: acceptor j | 256+e hole into electrode e | 320+e hole out of
//               electrode e), which of the lane's acceptors it starts from (3 bits), and the change it makes to the
//               state's Zobrist hash (upper 19 bits)
//     tail      1/total (fp32), normalised mass of the top events (fp64), total (fp64), the AS key words and a tag
//               (launch id, member index + 1): entries of earlier launches / members can never hit, nothing is reset
//
// in a direct-mapped first level in shared memory and a second level in global memory (L2).  Unlike the reference's
// getKey (simulation.go:29-38), which shifts all but the last 64 acceptors out of its uint64 key and therefore
// CONFUSES states for N > 64, the key here is the full mask; the slot index comes from a Zobrist hash that is
// maintained incrementally (one XOR per hop).
//
// Hit:  one compare of the lane's key word, the ballot prefix < uniform, a shuffle of the winner's aux word and a
//       branch-free update of the lane's mask word, the hash and the electrode tallies.  No energies, no rates.
// Miss: the fp64 energies are brought to the current mask from the mask of the last sweep (exact: fp64 sums of fp32
//       terms, so incremental == from scratch and the cached structure is a PURE function of the mask), then
//       hop_fast's sweep with per-acceptor (top, rest) tracking, prefix, normalisation, install in both levels.
// Rest: a uniform above the top events' mass (C5: ~0.5 % of the hops) takes the exact two-level pick over all
//       events except the lanes' top ones.
// With the cache disabled (LOGK = -1, KMCB200_FLAG_NO_MEMO) the same code runs every hop as a miss and produces
// bit-identical trajectories (tests/test_gpu_parity.py).
#include "kmc_device.cuh"
#include "kmc_internal.cuh"

namespace kmcb200 {

#define BIGW 1.0e30f
#define WENT 448u  // bytes per cache entry: 32 x f64 pre | 32 x u32 aux | tail 64 B
#define WT_RTOT 384u   // f32 1/total
#define WT_MTOP 392u   // f64 normalised mass of the top events
#define WT_TOTAL 400u  // f64 total rate
#define WT_KEY 408u    // u32 x 8 key words, then the tag {member + 1, launch id}
#define WT_GEN 440u
#define WT_LAUNCH 444u
#ifndef RING_D
#define RING_D 4       // rows in flight per warp in the cp.async ring of the GT sweep
#endif
#ifndef SPC
#define SPC 32
#endif
#ifndef SP_LOGK
#define SP_LOGK 4
#endif
#ifndef SP_MIN_CTAS
#define SP_MIN_CTAS 3
#endif
// SPC: pairs per lane parked between two evaluation rounds of the sparse sweep
#ifndef WIDE_MIN_CTAS
#define WIDE_MIN_CTAS 3  // 159 registers instead of 211: three CTAs per SM (C5: +20 %, profiles/r02/exp7_wide_occupancy.sh)
#endif

namespace {

__device__ __forceinline__ uint32_t wl_u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float wl_f(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double wl_d(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 wl_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void ws_u(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void ws_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void ws_d(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void ws_u4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ int w_bfind(uint32_t v) {
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t w_bit(uint32_t n) {  // 1 << (n & 31)
    uint32_t r;
    asm("bmsk.wrap.b32 %0, %1, 1;" : "=r"(r) : "r"(n));
    return r;
}

// Zobrist value of acceptor site i (murmur finaliser of i+1)
__device__ __forceinline__ uint32_t zob(uint32_t i) {
    uint32_t z = (i + 1u) * 0x9E3779B1u;
    z ^= z >> 15;
    z *= 0x85EBCA6Bu;
    z ^= z >> 13;
    return z;
}

__device__ __forceinline__ float w_scan_f(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// first lane whose inclusive prefix reaches thr among lanes with a positive rate; if rounding put thr past the end,
// the last positive lane; -1 if the group is empty.
__device__ __forceinline__ int w_pick_group(float rr, float thr, int lane) {
    const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
    if (!nz) return -1;
    const float s = w_scan_f(rr, lane);
    const uint32_t bal = __ballot_sync(FULL, s >= thr) & nz;
    return bal ? (__ffs(bal) - 1) : (31 - __clz(nz));
}

template <int AS>
struct SweepOut {
    float top[AS], rest[AS];  // per acceptor of this lane: its largest rate, the sum of its other rates
    int ptn[AS];              // partner site of the largest rate (acceptor j or N+e)
    float s_true[AS];         // fp32 energies of this lane's acceptors (as published to the mirror)
};

// Brings the fp64 energies from the mask of the last sweep (occ_sw) to the mask occ, publishes them, and evaluates
// every allowed pair of the state exactly once (hop_fast.cu's sweep).
//
// SP (sparse pair table: a layout built with a prune threshold, simulation.go:200-215 -- a pruned pair is an EXACT zero there
// and here): the sweep visits only the pairs whose table entry is not zero.  near[j][lane] holds, as bits k, which of the
// lane's acceptors lane + 32k have a non-zero constant towards target j (either direction for an electrode row).  For every
// target row the lane parks its (row, k) pairs in a list; when a list is nearly full, and at the end, all lanes evaluate their
// lists side by side (table entries straight from L1/L2, two loads ahead of the arithmetic) into per-lane accumulators in
// shared memory.  Per (lane, k) the pairs arrive in the dense sweep's order and a zero pair changes nothing there
// (rest += min(0, top); 0 > top is false), so the result is BIT-IDENTICAL to the dense sweep (tests/test_gpu_parity.py);
// a miss costs O(non-zero pairs of the state) -- C5 at prune_threshold 1e-7: ~330 pairs instead of 8 400.
template <int AS, bool GT, bool SP>
__device__ __forceinline__ void wide_sweep(const float2 *tbl, int PITCH, int N, int P, int lane, float nb, const uint32_t (&accm)[AS],
                                           const uint32_t (&occ)[AS], uint32_t (&occ_sw)[AS], double (&eps64)[AS],
                                           uint32_t a_mir, uint32_t a_ring, uint32_t a_near, SweepOut<AS> &o) {
    // ---- energies: flip the sites that differ (simulation.go:107-130, applied exactly)
#pragma unroll
    for (int kw = 0; kw < AS; ++kw) {
        uint32_t diff = occ[kw] ^ occ_sw[kw];
        while (diff) {
            const int b = __ffs(diff) - 1;
            diff &= diff - 1;
            const int j = kw * 32 + b;
            const bool now_occ = (occ[kw] >> b) & 1u;  // an EMPTY site j contributes -kd_ij to every other site
#pragma unroll
            for (int k = 0; k < AS; ++k) {
                const double kd = (double)tbl[j * PITCH + lane + 32 * k].y;
                eps64[k] += now_occ ? kd : -kd;
            }
        }
        occ_sw[kw] = occ[kw];
    }
    float src[AS], nbs[AS];
    const float *erow[AS];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < AS; ++k) {
        const bool oc = (occ[k] >> lane) & 1u;
        o.s_true[k] = (float)eps64[k];
        ws_f(a_mir + (lane + 32 * k) * 4, o.s_true[k]);
        src[k] = oc ? o.s_true[k] : -BIGW;       // only occupied acceptors emit to acceptors
        nbs[k] = oc ? 1.0f : -1.0f;              // occupied: i->e, dE = V_e - e_i ; empty: e->i, dE = e_i - V_e
        erow[k] = reinterpret_cast<const float *>(tbl + N * PITCH + lane + 32 * k) + (oc ? 0 : 1);
        o.top[k] = 0.0f; o.rest[k] = 0.0f; o.ptn[k] = 0;
    }
    __syncwarp();
    if (SP) {
        const uint32_t a_top = a_ring, a_rest = a_ring + AS * 128, a_ptn = a_ring + 2 * AS * 128;
        const uint32_t a_pl = a_ring + 3 * AS * 128;  // parked pairs [slot][lane], u16 = row << 3 | k
        const uint32_t a_list = a_pl + SPC * 64;      // u16 row indices: empties in ascending order, then N + e
        uint32_t occm8 = 0, valid8 = 0;
#pragma unroll
        for (int k = 0; k < AS; ++k) {
            occm8 |= ((occ[k] >> lane) & 1u) << k;
            valid8 |= ((accm[k] >> lane) & 1u) << k;
            ws_f(a_top + (k * 32 + lane) * 4, 0.0f);
            ws_f(a_rest + (k * 32 + lane) * 4, 0.0f);
            ws_u(a_ptn + (k * 32 + lane) * 4, 0u);
        }
        int n_emp = 0;
#pragma unroll
        for (int kw = 0; kw < AS; ++kw) {
            const uint32_t w = ~occ[kw] & accm[kw];
            if ((w >> lane) & 1u) {
                const int pos = n_emp + __popc(w & ((1u << lane) - 1u));
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_list + 2u * (uint32_t)pos), "h"((unsigned short)(kw * 32 + lane)));
            }
            n_emp += __popc(w);
        }
        if (lane < P) asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_list + 2u * (uint32_t)(n_emp + lane)), "h"((unsigned short)(N + lane)));
        __syncwarp();
        const int T = n_emp + P;
        auto row_of = [&](int t) {
            unsigned short jr;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(jr) : "r"(a_list + 2u * (uint32_t)t));
            return (int)jr;
        };
        int n = 0;
        auto fetch = [&](int e) {  // table entry of this lane's parked pair e
            float2 v = make_float2(0.0f, 0.0f);
            if (e < n) {
                unsigned short pr;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(pr) : "r"(a_pl + (uint32_t)(e * 32 + lane) * 2u));
                v = __ldg(tbl + row_of((int)pr >> 3) * PITCH + lane + 32 * ((int)pr & 7));
            }
            return v;
        };
        for (int t = 0;; ++t) {
            const bool done = t >= T;
            if (!done) {
                const int j = row_of(t);
                uint32_t m8;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(m8) : "r"(a_near + (uint32_t)j * 32u + (uint32_t)lane));
                m8 &= (t < n_emp) ? occm8 : valid8;
                while (m8) {
                    const int k = __ffs(m8) - 1;
                    m8 &= m8 - 1;
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_pl + (uint32_t)(n * 32 + lane) * 2u), "h"((unsigned short)((t << 3) | k)));
                    ++n;
                }
            }
            if (done || __any_sync(FULL, n > SPC - AS)) {
                const int nmax = __reduce_max_sync(FULL, n);
                float2 va = fetch(0), vb = fetch(1);
                for (int e = 0; e < nmax; ++e) {
                    const float2 v = va;
                    va = vb;
                    vb = fetch(e + 2);
                    if (e < n) {
                        unsigned short pr;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(pr) : "r"(a_pl + (uint32_t)(e * 32 + lane) * 2u));
                        const int k = (int)pr & 7, j = row_of((int)pr >> 3);
                        const float si = wl_f(a_mir + (lane + 32 * k) * 4);
                        float x;
                        if (j < N) x = ma_rate(v, wl_f(a_mir + j * 4), si, nb);
                        else {
                            const float se = wl_f(a_mir + (32 * AS + (j - N)) * 4);
                            const bool oc = (occm8 >> k) & 1u;  // occupied: i -> e, dE = V_e - e_i; empty: e -> i, dE = e_i - V_e
                            x = (oc ? v.x : v.y) * boltz(oc ? se - si : si - se, nb);
                        }
                        const uint32_t a = (uint32_t)(k * 32 + lane) * 4u;
                        const float top = wl_f(a_top + a);
                        ws_f(a_rest + a, wl_f(a_rest + a) + fminf(x, top));
                        if (x > top) {
                            ws_u(a_ptn + a, (uint32_t)j);
                            ws_f(a_top + a, x);
                        }
                    }
                }
                n = 0;
            }
            if (done) break;
        }
#pragma unroll
        for (int k = 0; k < AS; ++k) {
            o.top[k] = wl_f(a_top + (k * 32 + lane) * 4);
            o.rest[k] = wl_f(a_rest + (k * 32 + lane) * 4);
            o.ptn[k] = (int)wl_u(a_ptn + (k * 32 + lane) * 4);
        }
    } else if (GT) {
        // The pair table lives in global memory (L2): the rows of the state's targets -- its empty acceptors, then the
        // electrodes -- are streamed through a per-warp ring in shared memory with cp.async, RING_D rows ahead of the
        // arithmetic.  Every lane copies exactly the elements it consumes itself (columns lane + 32k), so
        // cp.async.wait_group is all the synchronisation there is.
        const uint32_t a_list = a_ring + RING_D * AS * 256;  // u16 row indices: empties in ascending order, then N + e
        int n_emp = 0;
#pragma unroll
        for (int kw = 0; kw < AS; ++kw) {
            const uint32_t w = ~occ[kw] & accm[kw];
            if ((w >> lane) & 1u) {
                const int pos = n_emp + __popc(w & ((1u << lane) - 1u));
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_list + 2u * (uint32_t)pos), "h"((unsigned short)(kw * 32 + lane)));
            }
            n_emp += __popc(w);
        }
        if (lane < P) asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_list + 2u * (uint32_t)(n_emp + lane)), "h"((unsigned short)(N + lane)));
        __syncwarp();
        const int T = n_emp + P;
        auto issue = [&](int t) {
            if (t < T) {
                unsigned short jr;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(jr) : "r"(a_list + 2u * (uint32_t)t));
                const float2 *row = tbl + (int)jr * PITCH + lane;
                const uint32_t dst = a_ring + (uint32_t)(t % RING_D) * (AS * 256) + lane * 8u;
#pragma unroll
                for (int k = 0; k < AS; ++k)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + k * 256u), "l"(row + 32 * k));
            }
            asm volatile("cp.async.commit_group;");
        };
#pragma unroll
        for (int t = 0; t < RING_D; ++t) issue(t);
        for (int t = 0; t < T; ++t) {
            asm volatile("cp.async.wait_group %0;" ::"n"(RING_D - 1));
            unsigned short jr;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(jr) : "r"(a_list + 2u * (uint32_t)t));
            const int j = (int)jr;
            const uint32_t srcb = a_ring + (uint32_t)(t % RING_D) * (AS * 256) + lane * 8u;
            const float sj = wl_f(a_mir + (j < N ? j : 32 * AS + (j - N)) * 4);
            float2 v[AS];
#pragma unroll
            for (int k = 0; k < AS; ++k)
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[k].x), "=f"(v[k].y) : "r"(srcb + k * 256u));
            if (t < n_emp) {  // acceptor target j: lane's occupied acceptors -> j
#pragma unroll
                for (int k = 0; k < AS; ++k) {
                    const float x = ma_rate(v[k], sj, src[k], nb);
                    o.rest[k] += fminf(x, o.top[k]);
                    if (x > o.top[k]) o.ptn[k] = j;
                    o.top[k] = fmaxf(x, o.top[k]);
                }
            } else {  // electrode j - N: acceptor -> electrode if occupied, electrode -> acceptor if empty
#pragma unroll
                for (int k = 0; k < AS; ++k) {
                    const float tc = (nbs[k] > 0.0f) ? v[k].x : v[k].y;
                    const float x = tc * boltz((sj - o.s_true[k]) * nbs[k], nb);
                    o.rest[k] += fminf(x, o.top[k]);
                    if (x > o.top[k]) o.ptn[k] = j;
                    o.top[k] = fmaxf(x, o.top[k]);
                }
            }
            issue(t + RING_D);
        }
        asm volatile("cp.async.wait_group 0;");
    } else {
#pragma unroll
    for (int kw = 0; kw < AS; ++kw) {
        uint32_t mm = ~occ[kw] & accm[kw];
        while (mm) {
            const int j = kw * 32 + __ffs(mm) - 1;
            mm &= mm - 1;
            const float sj = wl_f(a_mir + j * 4);
            const float2 *row = tbl + j * PITCH + lane;
#pragma unroll
            for (int k = 0; k < AS; ++k) {
                const float x = ma_rate(row[32 * k], sj, src[k], nb);
                o.rest[k] += fminf(x, o.top[k]);
                if (x > o.top[k]) o.ptn[k] = j;
                o.top[k] = fmaxf(x, o.top[k]);
            }
        }
    }
    for (int e = 0; e < P; ++e) {
        const float se = wl_f(a_mir + (32 * AS + e) * 4);
#pragma unroll
        for (int k = 0; k < AS; ++k) {
            const float x = erow[k][e * 2 * PITCH] * boltz((se - o.s_true[k]) * nbs[k], nb);
            o.rest[k] += fminf(x, o.top[k]);
            if (x > o.top[k]) o.ptn[k] = N + e;
            o.top[k] = fmaxf(x, o.top[k]);
        }
    }
    }
}

}  // namespace

template <int AS, int LOGK, bool GT, bool SP = false>
struct WideGeom {
    static constexpr int K = LOGK >= 0 ? (1 << LOGK) : 0;
    static constexpr int MIRW = 32 * AS + 32;  // per-warp mirror: acceptor energies [0,32*AS), electrode energies after
    // row ring + row-index list; sparse sweep: accumulators (top, rest, partner) + parked pairs + row-index list
    static constexpr int RINGB = SP ? 3 * AS * 128 + SPC * 64 + 2 * (32 * AS + 32) : (GT ? RING_D * AS * 256 + 2 * (32 * AS + 32) : 0);
    static constexpr int WARP_BYTES = MIRW * 4 + 1024 + (K > 0 ? K : 1) * (int)WENT + RINGB;  // mirror | variates | entries | ring
};

template <int AS, int LOGK, bool DBG, bool GT, bool SP>
__global__ void __launch_bounds__(128, SP ? SP_MIN_CTAS : WIDE_MIN_CTAS) kmc_wide_kernel(const LayoutDev L, const EnsembleDev E) {
    using G = WideGeom<AS, LOGK, GT, SP>;
    constexpr int K = G::K;
    constexpr int PITCH = 32 * AS + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // GT: the pair table stays in global memory (L1/L2-resident; N > 64: 2*S^2 floats exceed shared memory)
    const float2 *tbl = GT ? L.tblf : reinterpret_cast<const float2 *>(smem_raw);
    const int N = L.N, S = L.S, P = L.P;
    const int tid = threadIdx.x, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int warp = __shfl_sync(FULL, tid >> 5, 0);
    // (sparse sweep: the CTA's copy of the near masks takes the place of the table)
    const uint32_t tbl_bytes = SP ? (uint32_t)(((size_t)S * 32 + 15) & ~size_t(15))
                                  : (GT ? 0u : (uint32_t)((((size_t)S * PITCH * sizeof(float2)) + 15) & ~size_t(15)));
    if (SP) {
        uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(L.near);
        for (int i0 = 0; i0 < S * 8; i0 += blockDim.x)
            if (i0 + tid < S * 8) stage[i0 + tid] = src[i0 + tid];
    }
    if (!GT) {
        float2 *stage = reinterpret_cast<float2 *>(smem_raw);
        for (int i0 = 0; i0 < S * PITCH; i0 += blockDim.x)  // (thread-independent trip counts: see the staging loop of hop_lanes.cu)
            if (i0 + tid < S * PITCH) stage[i0 + tid] = L.tblf[i0 + tid];
    }
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t wb = sb + tbl_bytes + (uint32_t)warp * G::WARP_BYTES;
    const uint32_t a_mir = wb, a_rng = wb + G::MIRW * 4, a_cache = a_rng + 1024;
    const uint32_t a_ring = a_cache + (K > 0 ? K : 1) * WENT;
    // tags of the first level: 0 = never written (member tags start at 1)
    for (int s = lane; s < (K > 0 ? K : 1); s += 32) ws_u(a_cache + s * WENT + WT_GEN, 0u);
    __syncthreads();

    uint32_t accm[AS];
#pragma unroll
    for (int k = 0; k < AS; ++k) {
        const int lo = 32 * k;
        accm[k] = (N >= lo + 32) ? ~0u : (N > lo ? ((1u << (N - lo)) - 1u) : 0u);
    }
    // which word of the key this lane compares: lanes 0..29 the mask word (lane mod AS), lanes 30 / 31 the tag
    // (launch id / member + 1: entries of earlier launches and members can never hit, so nothing is ever reset)
    const int wl = (lane >= 30) ? 99 : (lane & (AS - 1));
    const uint32_t a_keyoff = (lane == 31) ? WT_GEN : (lane == 30 ? WT_LAUNCH : (WT_KEY + 4u * (uint32_t)(lane & (AS - 1))));
    const double INF = __longlong_as_double(0x7ff0000000000000LL);

    const int GLOG = (K > 0) ? E.gtab_log : 0;
    const int64_t wslot = (int64_t)blockIdx.x * nwarps + warp;
    unsigned char *gtab = (GLOG > 0) ? E.gtab + ((size_t)wslot << GLOG) * WENT : nullptr;

    for (;;) {
        unsigned long long mq = 0;
        if (lane == 0) mq = atomicAdd(E.queue, 1ULL);
        const int64_t m = (int64_t)__shfl_sync(FULL, mq, 0);
        if (m >= E.B) break;
        const uint32_t gen = (uint32_t)m + 1u;

        // ---- member parameters
        const float nb = (float)E.kT[m];  // kT, narrowed as the cgo wrappers do (simulationWrapper.go:92)
        const float ve_mine = (lane < P) ? (float)E.electrode_v[m * P + lane] : 0.0f;
        __syncwarp();
        ws_f(a_mir + (32 * AS + lane) * 4, ve_mine);

        // ---- initial state: occupation, E_constant (optionally by superposition), energies of the all-occupied mask
        uint32_t occ0[AS], occ_sw[AS];
        double eps64[AS];
        uint32_t H = 0;
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
            occ0[k] = __ballot_sync(FULL, o);
            if (o) H ^= zob((uint32_t)i);
            eps64[k] = e0;
            occ_sw[k] = accm[k];  // energies so far: no empty site; the first sweep subtracts the empty ones
        }
        H = __reduce_xor_sync(FULL, H);
        uint32_t occw = (lane == 30) ? E.launch_id : gen;  // lanes 30 / 31 carry the tag in place of a mask word
#pragma unroll
        for (int k = 0; k < AS; ++k)
            if (wl == k) occw = occ0[k];

        const uint64_t gm = E.member_index0 + (uint64_t)m;
        const uint2 key = make_uint2((uint32_t)E.seed, (uint32_t)(E.seed >> 32));
        const bool inject = DBG && E.stream_e != nullptr;
        const int64_t total_hops = E.prehops + E.hops, prehops = E.prehops;

        double t_acc = 0.0;
        float t_part = 0.0f;
        int eoc = 0;
        double occtime[AS];
#pragma unroll
        for (int k = 0; k < AS; ++k) occtime[k] = 0.0;
        bool dead = false;
        long long n_miss = 0;

        // loop-carried cache line of the CURRENT state
        uint32_t keyv = ~occw;  // first hop: miss
        double pre = 0.0, mtopn = 0.0;
        uint32_t aux = 0;
        float rtot = 0.0f;
        __syncwarp();

        int64_t h = 0;
        while (h < total_hops && !dead) {
            const int q0 = (int)(h & 63);
            int64_t hend = h - q0 + 64;
            if (hend > total_hops) hend = total_hops;
            if (h < prehops && hend > prehops) hend = prehops;
            const int q1 = q0 + (int)(hend - h);
            if (!inject) {
                t_acc += (double)t_part;
                t_part = 0.0f;
                const uint64_t blk = (uint64_t)(h >> 6) * 32u + (uint64_t)lane;
                const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)gm, (uint32_t)(gm >> 32)), key);
                const float e0 = -0.6931471805599453f * lg2_approx(fmaf((float)r.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
                const float e1 = -0.6931471805599453f * lg2_approx(fmaf((float)r.z, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
                const double u0 = ((double)r.y + 0.5) * 2.3283064365386963e-10;
                const double u1 = ((double)r.w + 0.5) * 2.3283064365386963e-10;
                __syncwarp();
                ws_u4(a_rng + lane * 32, make_uint4(__float_as_uint(e0), 0u, (uint32_t)__double2loint(u0), (uint32_t)__double2hiint(u0)));
                ws_u4(a_rng + lane * 32 + 16, make_uint4(__float_as_uint(e1), 0u, (uint32_t)__double2loint(u1), (uint32_t)__double2hiint(u1)));
                __syncwarp();
            }
            for (int q = q0; q < q1; ++q) {
                // ---- event structure of this state: cached, or computed and parked
                bool hit = false;
                if (K > 0) hit = __all_sync(FULL, keyv == occw);
                if (!hit) {
                    uint32_t occ[AS];
#pragma unroll
                    for (int k = 0; k < AS; ++k) occ[k] = __reduce_or_sync(FULL, lane == k ? occw : 0u);
                    const uint32_t Hu = __reduce_or_sync(FULL, H) & 0xffff0000u;
                    const uint32_t a_ent = a_cache + (LOGK > 0 ? (Hu >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u) * WENT;
                    double total = 0.0;
                    bool hit2 = false;
                    unsigned char *gent = nullptr;
                    if (GLOG > 0) {  // second level (global memory, L2): all loads in flight at once
                        gent = gtab + (size_t)((Hu >> 16) & ((1u << GLOG) - 1u)) * WENT;
                        const double g_pre = __ldcg(reinterpret_cast<const double *>(gent + lane * 8));
                        const uint32_t g_aux = __ldcg(reinterpret_cast<const uint32_t *>(gent + 256 + lane * 4));
                        const uint4 g0 = __ldcg(reinterpret_cast<const uint4 *>(gent + WT_RTOT));      // rtot | pad | mtopn
                        const double g_tot = __ldcg(reinterpret_cast<const double *>(gent + WT_TOTAL));
                        const uint32_t g_key = __ldcg(reinterpret_cast<const uint32_t *>(gent + a_keyoff));
                        hit2 = __all_sync(FULL, g_key == occw);
                        if (hit2) {
                            pre = g_pre;
                            aux = g_aux;
                            rtot = __uint_as_float(g0.x);
                            mtopn = __hiloint2double((int)g0.w, (int)g0.z);
                            total = g_tot;
                        }
                    }
                    if (!hit2) {
                        if (DBG) ++n_miss;
                        SweepOut<AS> sw;
                        wide_sweep<AS, GT, SP>(tbl, PITCH, N, P, lane, nb, accm, occ, occ_sw, eps64, a_mir, a_ring, sb, sw);
                        // the lane's top event and everything else
                        int ktop = 0;
                        float top = sw.top[0];
#pragma unroll
                        for (int k = 1; k < AS; ++k)
                            if (sw.top[k] > top) { top = sw.top[k]; ktop = k; }
                        double rsum = 0.0;
#pragma unroll
                        for (int k = 0; k < AS; ++k) rsum += (double)(sw.rest[k] + (k != ktop ? sw.top[k] : 0.0f));
                        const double incl = warp_incl_scan((double)top, lane);
                        const double mtop = __shfl_sync(FULL, incl, 31);
                        double ex = __shfl_up_sync(FULL, incl, 1);
                        if (lane == 0) ex = 0.0;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) rsum += __shfl_xor_sync(FULL, rsum, d);
                        total = mtop + rsum;
                        if (__all_sync(FULL, !(total > 0.0))) {  // no transition possible (simulation.go:297 would divide by zero)
                            dead = true;
                            break;
                        }
                        const double inv = 1.0 / total;
                        rtot = (float)inv;
                        mtopn = mtop * inv;
                        pre = (top > 0.0f) ? ex * inv : INF;
                        int ptop = sw.ptn[0];
#pragma unroll
                        for (int k = 1; k < AS; ++k)
                            if (k == ktop) ptop = sw.ptn[k];
                        const uint32_t site = (uint32_t)(lane + 32 * ktop);
                        bool osite = false;
#pragma unroll
                        for (int k = 0; k < AS; ++k)
                            if (k == ktop) osite = (occ[k] >> lane) & 1u;
                        const uint32_t code = (ptop < N) ? (uint32_t)ptop : (uint32_t)(ptop - N) + (osite ? 256u : 320u);
                        const uint32_t hd = zob(site) ^ ((ptop < N) ? zob((uint32_t)ptop) : 0u);
                        aux = (hd & 0xffffe000u) | ((uint32_t)ktop << 9) | code;
                        if (GLOG > 0) {
                            __stcg(reinterpret_cast<double *>(gent + lane * 8), pre);
                            __stcg(reinterpret_cast<uint32_t *>(gent + 256 + lane * 4), aux);
                            __stcg(reinterpret_cast<uint32_t *>(gent + a_keyoff), occw);
                            if (lane == 0) {
                                __stcg(reinterpret_cast<uint4 *>(gent + WT_RTOT),
                                       make_uint4(__float_as_uint(rtot), 0u, (uint32_t)__double2loint(mtopn), (uint32_t)__double2hiint(mtopn)));
                                __stcg(reinterpret_cast<double *>(gent + WT_TOTAL), total);
                            }
                        }
                    }
                    // install in the first level (without memoisation: a scratch entry that never hits)
                    __syncwarp();  // (the other lanes' prefetch reads of this slot are ordered before lane 0's writes)
                    ws_d(a_ent + lane * 8, pre);
                    ws_u(a_ent + 256 + lane * 4, aux);
                    ws_u(a_ent + a_keyoff, K > 0 ? occw : ~occw);
                    if (lane == 0) {
                        ws_u4(a_ent + WT_RTOT, make_uint4(__float_as_uint(rtot), 0u, (uint32_t)__double2loint(mtopn), (uint32_t)__double2hiint(mtopn)));
                        ws_d(a_ent + WT_TOTAL, total);
                    }
                    __syncwarp();
                }

                // ---- random variates: unit exponential for the dwell time (simulation.go:297), uniform for the pick (:164)
                double u;
                double dtd = 0.0;
                if (!inject) {
                    const uint4 rv = wl_u4(a_rng + q * 16);
                    const float dt = __uint_as_float(rv.x) * rtot;
                    t_part += dt;
                    if (DBG) dtd = (double)dt;
                    u = __hiloint2double((int)rv.w, (int)rv.z);
                } else {
                    const int64_t hh = h + (q - q0);
                    const uint32_t a_ent = a_cache + (LOGK > 0 ? ((H & 0xffff0000u) >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u) * WENT;
                    dtd = E.stream_e[m * total_hops + hh] / wl_d(a_ent + WT_TOTAL);
                    u = (double)E.stream_u[m * total_hops + hh];
                    t_acc += dtd;
                }

                if (DBG && h >= prehops) {  // occupied time of the PRE-hop state (simulation.go:314-316)
#pragma unroll
                    for (int k = 0; k < AS; ++k) {
                        const uint32_t ok = __reduce_or_sync(FULL, lane == k ? occw : 0u);
                        if ((ok >> lane) & 1u) occtime[k] += dtd;
                    }
                }

                int from = 0, to = 0;
                const uint32_t bal = __ballot_sync(FULL, pre < u);
                if (__builtin_expect(__all_sync(FULL, u < mtopn) && bal != 0u, 1)) {
                    // ---- the common case: one of the cached top events.  lane = the highest one whose interval starts below
                    //      the uniform.  Branch-free update (simulation.go:107-130) of this lane's mask word, the hash and
                    //      the electrode tallies.
                    const int istar = w_bfind(bal);
                    const uint32_t a = (uint32_t)__shfl_sync(FULL, (int)aux, istar);
                    const uint32_t code = a & 511u;
                    const uint32_t ks = (a >> 9) & 7u;
                    H ^= a;
                    uint32_t flip = (ks == (uint32_t)wl) ? w_bit((uint32_t)istar) : 0u;
                    if ((code >> 5) == (uint32_t)wl) flip |= w_bit(code);
                    occw ^= flip;
                    asm("{ .reg .pred p, q; .reg .u32 t;\n"
                        "  sub.u32 t, %1, %2;\n"
                        "  setp.eq.u32 p, t, 256;\n"
                        "  setp.eq.u32 q, t, 320;\n"
                        "  @p add.s32 %0, %0, 1;\n"
                        "  @q add.s32 %0, %0, -1; }"
                        : "+r"(eoc)
                        : "r"(code), "r"(lane));
                    if (DBG) {
                        const int site = istar + 32 * (int)ks;
                        if (code < 256u) { from = site; to = (int)code; }
                        else if (code < 320u) { from = site; to = N + (int)code - 256; }
                        else { from = N + (int)code - 320; to = site; }
                    }
                } else {
                    // ---- the rest of the list: exact two-level pick over all events EXCEPT the lanes' top ones
                    //      (rare enough that the sweep is simply repeated, even when this very hop already missed)
                    uint32_t occ[AS];
#pragma unroll
                    for (int k = 0; k < AS; ++k) occ[k] = __reduce_or_sync(FULL, lane == k ? occw : 0u);
                    const uint32_t Hu = __reduce_or_sync(FULL, H) & 0xffff0000u;
                    const uint32_t a_ent = a_cache + (LOGK > 0 ? (Hu >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u) * WENT;
                    if (DBG && hit) ++n_miss;
                    SweepOut<AS> sw;
                    wide_sweep<AS, GT, SP>(tbl, PITCH, N, P, lane, nb, accm, occ, occ_sw, eps64, a_mir, a_ring, sb, sw);
                    const double total = wl_d(a_ent + WT_TOTAL);
                    int ktop = 0;
                    {
                        float top = sw.top[0];
#pragma unroll
                        for (int k = 1; k < AS; ++k)
                            if (sw.top[k] > top) { top = sw.top[k]; ktop = k; }
                    }
                    // first level over (slot k, lane): fp64 prefix of the rows' masses without the lanes' top events
                    float rs[AS];
                    double pr[AS];
                    double base = 0.0;
#pragma unroll
                    for (int k = 0; k < AS; ++k) {
                        rs[k] = sw.rest[k] + (k != ktop ? sw.top[k] : 0.0f);
                        pr[k] = warp_incl_scan((double)rs[k], lane) + base;
                        base = __shfl_sync(FULL, pr[k], 31);
                    }
                    int wslotk = -1, wlane = 0;
                    const bool above = __all_sync(FULL, !(u < mtopn));
                    if (above && __all_sync(FULL, base > 0.0)) {
                        double rres = (u - mtopn) * total;
                        if (!(rres < base)) rres = base;
#pragma unroll
                        for (int k = 0; k < AS; ++k) {
                            const uint32_t b2 = __ballot_sync(FULL, pr[k] >= rres && rs[k] > 0.0f);
                            if (wslotk < 0 && b2) {
                                wslotk = k;
                                wlane = __ffs(b2) - 1;
                            }
                        }
                        if (wslotk < 0) {  // rounding: last row with a positive mass
#pragma unroll
                            for (int k = AS - 1; k >= 0; --k) {
                                const uint32_t b2 = __ballot_sync(FULL, rs[k] > 0.0f);
                                if (wslotk < 0 && b2) {
                                    wslotk = k;
                                    wlane = 31 - __clz(b2);
                                }
                            }
                        }
                    }
                    int skip = -1;
                    float rf = BIGW;
                    if (wslotk >= 0) {
                        float rf_mine = 0.0f;
                        int skip_mine = -1;
#pragma unroll
                        for (int k = 0; k < AS; ++k)
                            if (k == wslotk) {
                                rf_mine = (float)((u - mtopn) * total - (pr[k] - (double)rs[k]));
                                if (k == ktop) skip_mine = sw.ptn[k];
                            }
                        rf = __shfl_sync(FULL, rf_mine, wlane);
                        skip = __shfl_sync(FULL, skip_mine, wlane);
                    } else {
                        // no mass outside the top events (rounding), or a uniform of exactly 0 (injected stream): take the
                        // last (first) top event instead
                        const uint32_t posu = __ballot_sync(FULL, pre < INF);
                        if (!posu) {
                            dead = true;
                            break;
                        }
                        wlane = above ? 31 - __clz(posu) : __ffs(posu) - 1;
                        wslotk = __shfl_sync(FULL, ktop, wlane);
                        int ptop = sw.ptn[0];
#pragma unroll
                        for (int k = 1; k < AS; ++k)
                            if (k == ktop) ptop = sw.ptn[k];
                        skip = -2 - __shfl_sync(FULL, ptop, wlane);  // "take exactly this partner"
                    }
                    const int istar = wslotk * 32 + wlane;
                    bool rowocc = false;
#pragma unroll
                    for (int k = 0; k < AS; ++k)
                        if (k == wslotk) rowocc = (occ[k] >> wlane) & 1u;
                    const float s_star = wl_f(a_mir + istar * 4);
                    const float *ecol = reinterpret_cast<const float *>(tbl + (N + lane) * PITCH + istar);  // electrode `lane` vs istar
                    if (skip <= -2) {  // forced top event
                        const int p = -2 - skip;
                        if (rowocc) { from = istar; to = p; }
                        else { from = p; to = istar; }
                    } else if (rowocc) {
                        from = istar;
                        to = -1;
                        float thr = rf;
                        int lastpos = -1;
#pragma unroll
                        for (int kw = 0; kw < AS; ++kw) {
                            if (to < 0 && (~occ[kw] & accm[kw])) {
                                float rr = 0.0f;
                                if ((((~occ[kw] & accm[kw]) >> lane) & 1u) && (lane + 32 * kw) != skip) {
                                    const float2 v = tbl[(lane + 32 * kw) * PITCH + istar];
                                    rr = ma_rate(v, sw.s_true[kw], s_star, nb);
                                }
                                const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
                                if (nz) {
                                    lastpos = kw * 32 + 31 - __clz(nz);
                                    const float s = w_scan_f(rr, lane);
                                    const uint32_t b2 = __ballot_sync(FULL, s >= thr) & nz;
                                    if (b2) to = kw * 32 + __ffs(b2) - 1;
                                    thr -= __shfl_sync(FULL, s, 31);
                                }
                            }
                        }
                        if (to < 0) {  // electrode targets: istar -> e
                            float rr = 0.0f;
                            if (lane < P && (N + lane) != skip) rr = ecol[0] * boltz(ve_mine - s_star, nb);
                            const int e = w_pick_group(rr, thr, lane);
                            to = (e >= 0) ? N + e : lastpos;
                        }
                        if (to < 0 && skip >= 0) to = skip;  // rounding fallback: the row's top event
                    } else {  // empty acceptor: events e -> istar
                        to = istar;
                        float rr = 0.0f;
                        if (lane < P && (N + lane) != skip) rr = ecol[1] * boltz(s_star - ve_mine, nb);
                        const int e = w_pick_group(rr, rf, lane);
                        from = (e >= 0) ? N + e : (skip >= 0 ? skip : -1);
                    }
                    if (__any_sync(FULL, to < 0 || from < 0)) {
                        dead = true;
                        break;
                    }
                    // apply: mask word of this lane, hash, electrode tallies
                    uint32_t flip = 0u, hd = 0u;
                    if (from < N) {
                        if ((from >> 5) == wl) flip |= 1u << (from & 31);
                        hd ^= zob((uint32_t)from);
                    } else eoc -= (int)(lane == from - N);
                    if (to < N) {
                        if ((to >> 5) == wl) flip |= 1u << (to & 31);
                        hd ^= zob((uint32_t)to);
                    } else eoc += (int)(lane == to - N);
                    occw ^= flip;
                    H ^= hd;
                }

                // ---- tallies (simulation.go:309-317: antisymmetric traffic)
                if (DBG && h >= prehops && lane == 0) {
                    const int64_t hh = h + (q - q0);
                    if (E.traffic) {
                        double *tr = E.traffic + m * (int64_t)S * S;
                        tr[from * S + to] += 1.0;
                        tr[to * S + from] -= 1.0;
                    }
                    if (E.trace) {
                        int32_t *tp = E.trace + (m * E.hops + (hh - prehops)) * 2;
                        tp[0] = from;
                        tp[1] = to;
                    }
                }

                // ---- prefetch the next state's cache line
                if (K > 0) {
                    const uint32_t a_ent = a_cache + (LOGK > 0 ? ((H & 0xffff0000u) >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u) * WENT;
                    pre = wl_d(a_ent + lane * 8);
                    aux = wl_u(a_ent + 256 + lane * 4);
                    const uint4 t0 = wl_u4(a_ent + WT_RTOT);  // rtot | pad | mtopn
                    rtot = __uint_as_float(t0.x);
                    mtopn = __hiloint2double((int)t0.w, (int)t0.z);
                    keyv = wl_u(a_ent + a_keyoff);
                }
            }
            h = hend;
            if (h == prehops && prehops > 0 && !dead) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
                t_acc = 0.0;
                t_part = 0.0f;
                eoc = 0;
#pragma unroll
                for (int k = 0; k < AS; ++k) occtime[k] = 0.0;
            }
        }

        // ---- results
        t_acc += (double)t_part;
        if (dead) t_acc = INF;  // +inf, as time_step = e/0 would give
        if (lane == 0) E.time[m] = t_acc;
        if (lane < P) E.electrode_occ[m * P + lane] = (int64_t)eoc;
        {
            uint32_t occ[AS];
#pragma unroll
            for (int k = 0; k < AS; ++k) occ[k] = __reduce_or_sync(FULL, lane == k ? occw : 0u);
            if (E.site_energies_out) {  // energies of the final mask
                SweepOut<AS> sw;
                wide_sweep<AS, GT, SP>(tbl, PITCH, N, P, lane, nb, accm, occ, occ_sw, eps64, a_mir, a_ring, sb, sw);
            }
#pragma unroll
            for (int k = 0; k < AS; ++k) {
                const int i = lane + 32 * k;
                if (i < N) {
                    if (E.occupation_out) E.occupation_out[m * N + i] = (occ[k] >> lane) & 1u;
                    if (DBG && E.avg_occupation) E.avg_occupation[m * N + i] = occtime[k];
                    if (E.site_energies_out) E.site_energies_out[m * S + i] = eps64[k];
                }
            }
        }
        if (E.site_energies_out && lane < P) E.site_energies_out[m * S + N + lane] = (double)ve_mine;
        if (DBG && E.misses && lane == 0) E.misses[m] = n_miss;
        __syncwarp();
    }  // members of this warp slot
}

template <int AS, int LOGK, bool GT, bool SP>
static cudaError_t launch_wide_t(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches, MemoPlan *plan_only) {
    using G = WideGeom<AS, LOGK, GT, SP>;
    const bool dbg = E.avg_occupation || E.traffic || E.trace || E.stream_e || E.misses;
    int warps = 4;
    while (warps > 1 && (E.B + warps - 1) / warps < 2 * 148) warps >>= 1;
    const size_t smem = (SP ? (((size_t)L.S * 32 + 15) & ~size_t(15)) : (GT ? 0 : ((((size_t)L.S * (32 * AS + 1) * sizeof(float2)) + 15) & ~size_t(15)))) +
                        (size_t)warps * G::WARP_BYTES;
    auto kern = dbg ? kmc_wide_kernel<AS, LOGK, true, GT, SP> : kmc_wide_kernel<AS, LOGK, false, GT, SP>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (E.B + warps - 1) / warps;
    const unsigned grid = (unsigned)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
    if (plan_only) {
        plan_only->warp_slots = (int64_t)grid * warps;
        return cudaSuccess;
    }
    kern<<<grid, warps * 32, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

template <int AS, bool GT, bool SP = false>
static cudaError_t launch_wide_k(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches, MemoPlan *plan) {
    if (logk < 0) return launch_wide_t<AS, -1, GT, SP>(L, E, st, launches, plan);
    return launch_wide_t<AS, SP ? SP_LOGK : 4, GT, SP>(L, E, st, launches, plan);
}

// 32 <= N <= 256 acceptors (N <= 31 runs hop_memo.cu).  logk < 0 disables the memoisation (every hop a miss).
// plan != nullptr: only report the launch geometry; the caller sizes the second-level table
// E.gtab = warp_slots * 2^E.gtab_log * 448 bytes from it.
cudaError_t launch_wide(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches, MemoPlan *plan) {
    if (E.B <= 0) {
        if (plan) plan->warp_slots = 0;
        return cudaSuccess;
    }
    const int as = (L.N + 31) / 32;
    if (L.P > 32 || as > 8) return cudaErrorInvalidValue;
    if (L.pitchf != 32 * (as == 3 ? 4 : (as > 4 ? 8 : as)) + 1) return cudaErrorInvalidValue;
    if (as <= 1) return launch_wide_k<1, false>(L, E, logk, st, launches, plan);
    if (as == 2) return launch_wide_k<2, false>(L, E, logk, st, launches, plan);
    // a pair table that is mostly zeros (a prune threshold was given): the sweep walks the non-zero pairs only
    bool sparse = L.sparse != 0;
    if (const char *ev = getenv("KMCB200_WIDE_SPARSE")) sparse = atoi(ev) != 0 && L.near != nullptr;
    if (as <= 4) return sparse ? launch_wide_k<4, true, true>(L, E, logk, st, launches, plan) : launch_wide_k<4, true>(L, E, logk, st, launches, plan);
    return sparse ? launch_wide_k<8, true, true>(L, E, logk, st, launches, plan) : launch_wide_k<8, true>(L, E, logk, st, launches, plan);
}

}  // namespace kmcb200
