// hop_lanes.cu -- production KMC hop loop, ONE THREAD PER TRAJECTORY on the common path (KMCB200_MODE_FAST, N <= 31
// acceptors).
//
// Reference semantics being accelerated (MUTUEL/kmc_dn, paths relative to the reference tree):
//   site energies      goSimulation/simulation.go:226-234, 378-386
//   allowed pairs      goSimulation/simulation.go:40-55
//   Miller-Abrahams    goSimulation/simulation.go:58-80
//   cumulative list    goSimulation/simulation.go:267-276
//   dwell time / pick  goSimulation/simulation.go:297-299, 163-188
//   hop + tallies      goSimulation/simulation.go:107-130, 306-319
//   state cache        goSimulation/simulation.go:222-223, 251-296, 351-412
//
// A warp carries 32 trajectories in lock-step, one per thread.  A hop whose state is in the table is plain per-thread
// code: Philox, one hash, ONE 32-byte read (the state's key, -ln2/total and its 6 most likely events), a select tree,
// the mask update.  Only what is NOT in the table is warp-cooperative: the warp stops, evaluates the missing state of
// one of its trajectories (lane i = acceptor i; the sweep of memo_common.cuh) and goes on.
//
// Round-2 redesign (generation 9).  The round-1 kernel was bound by the L1TEX data pipe (two scattered
// 256-bit loads per thread and hop = 64 wavefronts per warp-hop, 74 % of the pipe's cycles) and by evaluations caused
// by CONFLICT misses of its direct-mapped table (a run of 16 seeds visits 60-600 distinct states in 1.6e6 hops, yet
// 0.26 % of the hops were evaluated).  Now:
//
// Entry (64 B = two 32-byte sectors), a pure function of (layout, E_constant, electrode energies, kT, occupation mask):
//     sector A:  u32 key (the mask) | f32 -ln2/total | 6 x u32 event words
//     sector B:  8 x u32 event words
//   event word k = F_k << 12 | code_k: the K = 14 LARGEST events of the state in decreasing order (candidates: the 2 or
//   3 largest events of every acceptor), F_k = q_0 + .. + q_k the cumulative probability in units of 2^-20 with
//   q_k = floor(rate_k / total * 2^20) -- rounded DOWN; code = acceptor | event << 5 (event: partner acceptor j |
//   32 + e hole into electrode e | 64 + e hole out of electrode e).  With X = x | 0xfff (x = 32 uniform bits),
//   X > word_k  <=>  (x >> 12) >= F_k: six compares and five selects resolve a hop.
//   What the floor leaves over (the events' residuals) and every other event form the TAIL, [F_13 * 2^12, 2^32): a hop
//   lands there with probability ~3e-5 (C3) and is resolved by evaluating the state again and an exact two-level pick
//   over the tail list (tail_pick below).  Every event keeps exactly its probability rate / total: front part + residual.
//
// Table.  Trajectories with identical parameters (the seeds of one voltage vector / temperature) SHARE one table: the
// warp detects the runs of consecutive identical members among its 32; every run gets an equal power-of-two share of
// the warp slot's sets.  A set is one 128-byte line = 2 ways, LRU: a hit in way 1 swaps the ways, an insertion pushes
// way 0 to way 1.  (Measured on the CPU model of this table, C3: 2-way LRU with 2048 entries evaluates 0.05 % of the
// hops where the direct-mapped table with 8192 evaluated 0.3 %.)  The keys are cleared when a warp takes a block of
// members: no tags, no launch ids, nothing to zero at allocation.
//
// Lock-step.  All 32 trajectories of a warp execute hop h together.  Hot path (inline): way-0 hit in sector A.
// Everything else goes through lanes_cold(): read phase (per thread: sector B / way 1), then -- behind a warp barrier,
// so that no read overlaps a write -- way swaps, tail picks and evaluations, one at a time, warp-cooperatively.
// With the table disabled (lanes_flags & 1) every hop is evaluated -- same arithmetic, bit-identical results (tested).
//
// Scheduling.  Work items are (block of 32 members, range of hops).  The first blocks run all their hops in one item;
// the LAST blocks of the queue are cut into slices of hops (state checkpointed in the output arrays), handed out
// slice-major with a per-block progress flag, so that the launch does not end on a few warps finishing whole blocks
// (round 1: SMs active 88 % of the launch).
//
// Halves.  An ensemble whose blocks would leave more than half of the device's warp slots empty gives every block two warps;
// a block splits where a run of identical members starts at member 16 (see the kernel).  Geometry: CTAs of 14 warps, two per
// SM (72 registers), which keeps shared memory at 117 KB per SM and L1 at 124 KB.
//
// The same file holds kmc_solo_kernel, the latency kernel for a FEW trajectories (one warp walks a trajectory's visited states
// as a graph in shared memory, a second warp does the bookkeeping): same evaluation, same entries, bit-identical results.
//
// RNG: Philox4x32-10 (key = seed, counter = (64-hop block * 32 + pair, global member index), two hops per call), so
// streams do not depend on batching, on the number of GPUs or on the kernel's geometry.  The round keys come
// precomputed from the host (EnsembleDev::rk, constant bank operands).
#include <algorithm>
#include <cstdlib>
#include "memo_common.cuh"

namespace kmcb200 {

#define LSETB 128u  // bytes per set (2 ways x 64 B)
// Launch geometry: 28 warps per SM at 72 registers (measured 1.60e11 hops/s on C3 against 1.26e11 with 32 warps / 64 registers --
// spills in the hop loop -- and 1.57e11 with 24 / 80), in CTAs of 14 warps: the layout tables (10 KB) are per CTA, and two CTAs of
// 14 warps need 117 KB of shared memory where seven of 4 need 168 -- the carve-out drops from 196 to 132 KB and L1 grows from 56
// to 124 KB.  C3 1 048 576 x 1e5: +1 %; 65 536 x 1e6: +8 %; 32 768 x 1e5: +6 %; C4: +6 %; one CTA of 28 warps: +2 % / +10 % / -20 % /
// -29 % (profiles/r02/exp8_lanes_l1.sh).
#ifndef LANES_WARPS
#define LANES_WARPS 14    // warps per CTA
#endif
#ifndef LANES_MIN_CTAS
#define LANES_MIN_CTAS 2  // resident CTAs per SM the register budget is set for
#endif
#ifndef LANES_RMAX
#define LANES_RMAX 8      // runs of a warp whose parameters live in shared memory (later runs: from global memory)
#endif
#define LANES_K 14        // events per entry
#define LN2F 0.6931471805599453

template <int PT>
struct LanesGeom {
    static constexpr int PV = PT > 0 ? PT : 32;             // electrode slots per trajectory
    static constexpr int MIRB = 256;                        // mirror: acceptor energies (128 B) | electrode energies (128 B)
    static constexpr int RUNB = 128 + 4 * PV + 16;          // per run: E_constant row (f32 x 32) | electrode energies | kT
    static constexpr int TALB = PV * 32 * 4;                // electrode tallies [electrode][trajectory]
    static constexpr int STASHW = 6;                        // words per trajectory parked in shared memory around an evaluation
    static constexpr int WARP_BYTES = MIRB + LANES_RMAX * RUNB + TALB + STASHW * 128;
};

struct Sector {
    uint32_t w[8];
};
__device__ __forceinline__ Sector ldg_sector(const unsigned char *p) {
    Sector s;
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]), "=r"(s.w[4]), "=r"(s.w[5]), "=r"(s.w[6]), "=r"(s.w[7])
                 : "l"(p)
                 : "memory");
    return s;
}
__device__ __forceinline__ uint32_t ldg_u32(const unsigned char *p) {
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg_u32(unsigned char *p, uint32_t v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// 1 << (n & 31) (BMSK: the code's low 5 bits are the acceptor, no masking needed)
__device__ __forceinline__ uint32_t bit_wrap(uint32_t n) {
    uint32_t r;
    asm("bmsk.wrap.b32 %0, %1, 1;" : "=r"(r) : "r"(n));
    return r;
}

// Philox4x32-10 with the round keys precomputed on the host (rk[2r], rk[2r+1] = key + r * Weyl constants)
__device__ __forceinline__ uint4 philox_rk(uint4 c, const uint32_t *__restrict__ rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ rk[2 * r], lo1, hi0 ^ c.w ^ rk[2 * r + 1], lo0);
    }
    return c;
}

// event k of a sector's words w[0..n) for X = x | 0xfff (the caller has checked X <= w[n-1], i.e. the event is here)
__device__ __forceinline__ uint32_t select6(uint32_t X, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f) {
    const uint32_t s01 = X > a ? b : a, s23 = X > c ? d : c, s45 = X > e ? f : e;
    return X > d ? s45 : (X > b ? s23 : s01);
}
__device__ __forceinline__ uint32_t select4(uint32_t X, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t s01 = X > a ? b : a, s23 = X > c ? d : c;
    return X > b ? s23 : s01;
}
__device__ __forceinline__ uint32_t select8(uint32_t X, const Sector &s) {
    const uint32_t s01 = X > s.w[0] ? s.w[1] : s.w[0], s23 = X > s.w[2] ? s.w[3] : s.w[2];
    const uint32_t s45 = X > s.w[4] ? s.w[5] : s.w[4], s67 = X > s.w[6] ? s.w[7] : s.w[6];
    const uint32_t lo = X > s.w[1] ? s23 : s01, hi = X > s.w[5] ? s67 : s45;
    return X > s.w[3] ? hi : lo;
}

// per-warp, per-item constants of the cold path (lives in local memory; the hot loop keeps its own copies in registers)
struct LaneCtx {
    uint32_t a_row_me, a_col_me, a_elF, a_elR, a_elF_e, a_elR_e, a_mir, a_run;
    uint32_t accm, lead_mask;
    int lane, N, P, hshift;
    unsigned char *wtab;  // this warp slot's sets
    int64_t base;         // first member of the block
    bool use_table;
};

// parameters of run r (leader = member base + ld): E_constant of acceptor `lane` (narrowed to float32,
// simulationWrapper.go:50-56), energy of electrode `lane`, kT (narrowed)
template <int PT>
__device__ __forceinline__ void run_params(const EnsembleDev &E, const LaneCtx &c, int r, int ld, double &E64, float &ve_mine, float &kTt) {
    using G = LanesGeom<PT>;
    const int lane = c.lane, N = c.N, P = c.P;
    if (r < LANES_RMAX) {
        const uint32_t a = c.a_run + (uint32_t)r * G::RUNB;
        E64 = (double)lds_f(a + lane * 4);
        ve_mine = (lane < P) ? lds_f(a + 128 + lane * 4) : 0.0f;
        kTt = lds_f(a + 128 + 4 * G::PV);
    } else {
        const int64_t mt = c.base + ld;
        double e64 = 0.0;
        if (lane < N) {
            if (E.E_constant) e64 = E.E_constant[mt * N + lane];
            else {
                e64 = E.basis[(int64_t)P * N + lane];
                for (int p = 0; p < P; ++p) e64 += E.electrode_v[mt * P + p] * E.basis[(int64_t)p * N + lane];
            }
        }
        E64 = (double)(float)e64;
        ve_mine = (lane < P) ? (float)E.electrode_v[mt * P + lane] : 0.0f;
        kTt = (float)E.kT[mt];
    }
}

// What an evaluation leaves in the registers of the warp: lane k < K holds event k (word = F_k << 12 | code), every lane
// (= acceptor) its energy, its total rate and its candidates.
template <int NR>
struct Eval {
    double total;      // total rate of the state (warp-uniform)
    float rtp;         // -ln2 / total
    uint32_t word;     // lane k < LANES_K: F_k << 12 | code_k, else 0xffffffff
    uint32_t F13;      // F of the last event (warp-uniform): the tail starts at F13 << 12
    uint32_t win_q;    // lane k < LANES_K: owner acceptor | q_k << 5
    float e_me, R;     // this acceptor's fp32 site energy, the sum of all its rates
};

// Evaluate state occu of a run: sweep, total, the K largest events, their quantised lengths.
template <int PT, int NR>
__device__ __forceinline__ bool evaluate_state(const LaneCtx &c, uint32_t occu, double E64, float ve_mine, float kTt, Eval<NR> &ev) {
    const int lane = c.lane, N = c.N, P = c.P;
    __syncwarp();
    if (lane < P) sts_f(c.a_mir + 128 + lane * 4, ve_mine);
    float rest;
    float tk[NR];
    int pk[NR];
    sweep_state<PT, NR>(occu, c.accm, E64, lane, N, P, kTt, c.a_row_me, c.a_mir, c.a_elF, c.a_elR, ev.e_me, tk, pk, rest);
    float R = rest;
#pragma unroll
    for (int r = NR - 1; r >= 0; --r) R += tk[r];
    if (lane >= N) R = 0.0f;
    ev.R = R;
    double total = (double)R;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(FULL, total, d);
    ev.total = total;
    if (!(total > 0.0)) return false;  // no transition possible (simulation.go:297 would divide by zero)
    ev.rtp = (float)(-LN2F / total);
    // candidates of this acceptor: key = rate bits (7 low mantissa bits dropped) | rank << 5 | lane, event codes packed
    const bool occ_a = (occu >> lane) & 1u;
    uint32_t keys[NR], evts = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const uint32_t evt = ((uint32_t)pk[r] < (uint32_t)N) ? (uint32_t)pk[r] : ((uint32_t)pk[r] - (uint32_t)N + (occ_a ? 32u : 64u));
        evts |= evt << (7 * r);
        keys[r] = (lane < N && tk[r] > 0.0f) ? ((__float_as_uint(tk[r]) & ~127u) | ((uint32_t)r << 5) | (uint32_t)lane) : 0u;
    }
    // the K largest candidates, one REDUX.MAX per event; lane k keeps event k
    uint32_t head = keys[0], mywk = 0, myev = 0;
    int ptr = 0;
#pragma unroll
    for (int k = 0; k < LANES_K; ++k) {
        const uint32_t wk = __reduce_max_sync(FULL, head);
        const int win = (int)(wk & 31u);
        const uint32_t e = __shfl_sync(FULL, evts, win);
        if (lane == k) { mywk = wk; myev = e; }
        if (lane == win) {
            ++ptr;
            head = 0u;
#pragma unroll
            for (int r = 1; r < NR; ++r)
                if (ptr == r) head = keys[r];
        }
    }
    // quantised length of event k: q = floor(rate / total * 2^20 * (1 - 2^-24)) from the truncated rate: never more than
    // the event's true share, and the sum stays below 2^20
    const double inv20 = 1048575.9375 / total;
    uint32_t q = 0, code = 0;
    if (lane < LANES_K && mywk) {
        const int rank = (int)((mywk >> 5) & 3u);
        code = (mywk & 31u) | (((myev >> (7 * rank)) & 127u) << 5);
        q = (uint32_t)__double2uint_rd((double)__uint_as_float(mywk & ~127u) * inv20);
    }
    uint32_t F = q;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, F, d);
        if (lane >= d) F += t;
    }
    ev.F13 = __shfl_sync(FULL, F, LANES_K - 1);
    ev.word = (lane < LANES_K) ? ((F << 12) | code) : 0xffffffffu;
    ev.win_q = (code & 31u) | (q << 5);
    return true;
}

// The tail: exact two-level pick over every event of the state with the front parts q_k * total / 2^20 of the K front
// events taken off (x >= F13 << 12).  A pure function of the state and x; warp-cooperative (lane = acceptor / target /
// electrode); returns the event code (acceptor | event << 5), warp-uniform.
template <int NR>
__device__ __noinline__ uint32_t tail_pick(const LaneCtx &c, const Eval<NR> &ev, uint32_t occu, float kTt, float ve_mine, uint32_t x) {
    const int lane = c.lane, N = c.N, P = c.P;
    const double unit = ev.total * 9.5367431640625e-07;  // total / 2^20
    const double u = ((double)(x - (ev.F13 << 12)) + 0.5) * 2.3283064365386963e-10 * ev.total;
    // first level: acceptors, each with the front parts of its own events removed
    double mine = (double)ev.R;
    const uint32_t evcode = ev.word & 4095u;
    for (int k = 0; k < LANES_K; ++k) {
        const uint32_t wq = __shfl_sync(FULL, ev.win_q, k);
        if ((int)(wq & 31u) == lane) mine -= (double)(wq >> 5) * unit;
    }
    if (mine < 0.0) mine = 0.0;
    const double incl = scan_d(mine);
    double ex = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) ex = 0.0;
    const uint32_t pos = __ballot_sync(FULL, mine > 0.0);
    // an event that is certainly allowed, for the cases rounding leaves without a pick: the largest event of the state
    const uint32_t fallback = __shfl_sync(FULL, evcode, 0);
    if (!pos) return fallback;
    uint32_t b2 = __ballot_sync(FULL, ex < u) & pos;
    if (!b2) b2 = pos & (0u - pos);
    const int istar = 31 - __clz(b2);
    const float rf = __shfl_sync(FULL, (float)(u - ex), istar);
    const bool rowocc = (occu >> istar) & 1u;
    const float e_star = __shfl_sync(FULL, ev.e_me, istar);
    // front parts of istar's events, by target: acceptor target j / electrode e <-> lane
    float offA = 0.0f, offE = 0.0f;
    for (int k = 0; k < LANES_K; ++k) {
        const uint32_t wq = __shfl_sync(FULL, ev.win_q, k);
        const uint32_t cd = __shfl_sync(FULL, evcode, k);
        if ((int)(wq & 31u) == istar) {
            const uint32_t evt = cd >> 5;
            const float part = (float)((double)(wq >> 5) * unit);
            if (evt < 32u) {
                if ((int)evt == lane) offA = part;
            } else if ((int)(evt & 31u) == lane) offE = part;
        }
    }
    int from, to;
    if (rowocc) {
        from = istar;
        to = -1;
        int lastA = -1;
        float sA = 0.0f;
        const uint32_t emp = ~occu & c.accm;
        if (emp) {  // acceptor targets: istar -> empty `lane`
            float rr = 0.0f;
            if ((emp >> lane) & 1u) {
                const float2 v = lds_f2(c.a_col_me + istar * 8);
                rr = fmaxf(ma(v.x, v.y, ev.e_me, e_star, kTt) - offA, 0.0f);
            }
            const uint32_t nz = __ballot_sync(FULL, rr > 0.0f);
            if (nz) {
                const float sc = scan_f<5>(rr);
                const uint32_t b3 = __ballot_sync(FULL, sc >= rf) & nz;
                if (b3) to = __ffs(b3) - 1;
                else {
                    lastA = 31 - __clz(nz);
                    sA = __shfl_sync(FULL, sc, 31);
                }
            }
        }
        if (to < 0) {  // electrode targets: istar -> electrode `lane`
            float rr = 0.0f;
            if (lane < P) rr = fmaxf(lds_f(c.a_elF_e + istar * 4) * boltz(ve_mine - e_star, kTt) - offE, 0.0f);
            const int e = pick_group<5>(rr, rf - sA);
            to = (e >= 0) ? N + e : lastA;
        }
        if (to < 0) return fallback;
    } else {  // empty acceptor: events electrode `lane` -> istar
        to = istar;
        float rr = 0.0f;
        if (lane < P) rr = fmaxf(lds_f(c.a_elR_e + istar * 4) * boltz(e_star - ve_mine, kTt) - offE, 0.0f);
        from = pick_group<5>(rr, rf);
        if (from < 0) return fallback;
        from += N;
    }
    if (from < N && to < N) return (uint32_t)from | ((uint32_t)to << 5);
    if (from < N) return (uint32_t)from | ((uint32_t)(32 + to - N) << 5);
    return (uint32_t)to | ((uint32_t)(64 + from - N) << 5);
}

// Everything that is not a way-0 hit in sector A.  st: 0 = nothing to do for this thread, 1 = way 0 holds another state,
// 2 = way-0 hit, but the event lies beyond sector A (f7 = its last word).  Returns {code, bits of -ln2/total}; bit 31 of
// the code: the state is dead (no transition possible); bit 30: the state was evaluated for this thread.
template <int PT, bool DBG, int NR>
__device__ __forceinline__ uint2 lanes_cold(const EnsembleDev &E, const LaneCtx &c, uint32_t occ, uint32_t xr, uint32_t st, uint32_t f1,
                                         uint32_t f7, uint32_t ri, uint32_t setofs, int leader) {
    const int lane = c.lane;
    const uint32_t X = xr | 0xfffu;
    uint32_t code = 0, rt = f1;
    // ---- read phase (per thread): sector B of way 0, or way 1
    bool need_eval = false, need_tail = false, want_swap = false;
    unsigned char *line = c.wtab + (size_t)(setofs + ((occ * 0x9E3779B1u) >> c.hshift)) * LSETB;
    if (st == 2) {
        const Sector b = ldg_sector(line + 32);
        if (X > b.w[7]) need_tail = true;
        else code = select8(X, b);
    } else if (st == 1) {
        if (c.use_table) {
            const Sector a = ldg_sector(line + 64);
            if (a.w[0] == occ) {
                want_swap = true;
                rt = a.w[1];
                if (X > a.w[7]) {
                    const Sector b = ldg_sector(line + 96);
                    if (X > b.w[7]) need_tail = true;
                    else code = select8(X, b);
                } else code = select6(X, a.w[2], a.w[3], a.w[4], a.w[5], a.w[6], a.w[7]);
            } else need_eval = true;
        } else need_eval = true;
    }
    __syncwarp();  // every read of this step is done before the table is written

    // ---- way swaps (LRU: the way just hit becomes way 0), one set at a time, all lanes move one word each
    uint32_t swaps = __ballot_sync(FULL, want_swap);
    while (swaps) {
        const int t = __ffs(swaps) - 1;
        unsigned char *ln = (unsigned char *)__shfl_sync(FULL, (unsigned long long)line, t);
        swaps &= ~__ballot_sync(FULL, want_swap && line == ln);
        const uint32_t v = ldg_u32(ln + lane * 4);
        __syncwarp();
        stg_u32(ln + ((lane * 4) ^ 64), v);
        __syncwarp();
    }

    // ---- evaluations (states that are not in the table) and tail picks (states that are, but x lies beyond the front)
    uint32_t need = __ballot_sync(FULL, need_eval || need_tail);
    bool died = false, evaluated = false;
    while (need) {
        const int t = __ffs(need) - 1;
        const uint32_t occu = __shfl_sync(FULL, occ, t);
        const int r = (int)__shfl_sync(FULL, ri, t);
        double E64;
        float ve_mine, kTt;
        run_params<PT>(E, c, r, __shfl_sync(FULL, leader, t), E64, ve_mine, kTt);
        Eval<NR> ev;
        const bool ok = evaluate_state<PT, NR>(c, occu, E64, ve_mine, kTt, ev);
        // threads of this run that wait on this very state (t among them) are all served by this evaluation
        const uint32_t same = __ballot_sync(FULL, ((need >> lane) & 1u) && occ == occu && ri == (uint32_t)r);
        need &= ~same;
        if ((same >> lane) & 1u) evaluated = true;
        if (!ok) {
            if ((same >> lane) & 1u) died = true;
            continue;
        }
        const uint32_t ins = __ballot_sync(FULL, need_eval) & same;
        if (ins && c.use_table) {  // insert: way 0 <- the new entry, way 1 <- the old way 0 (if it held a state)
            unsigned char *ln = (unsigned char *)__shfl_sync(FULL, (unsigned long long)line, t);
            const uint32_t v = ldg_u32(ln + (lane & 15) * 4);
            const uint32_t oldkey = __shfl_sync(FULL, v, 0);
            __syncwarp();
            if (oldkey != 0xffffffffu && lane < 16) stg_u32(ln + 64 + lane * 4, v);
            if (lane < 6) stg_u32(ln + 8 + lane * 4, ev.word);
            else if (lane < LANES_K) stg_u32(ln + 32 + (lane - 6) * 4, ev.word);
            else if (lane == LANES_K) stg_u32(ln, occu);
            else if (lane == LANES_K + 1) stg_u32(ln + 4, __float_as_uint(ev.rtp));
            __syncwarp();
        }
        uint32_t todo = same;
        while (todo) {
            const int t2 = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t x2 = __shfl_sync(FULL, xr, t2);
            const uint32_t bb = __ballot_sync(FULL, (x2 | 0xfffu) <= ev.word);  // lanes >= K hold 0xffffffff
            const int k = __ffs(bb) - 1;
            uint32_t c2;
            if (k < LANES_K) c2 = __shfl_sync(FULL, ev.word, k) & 4095u;
            else c2 = tail_pick<NR>(c, ev, occu, kTt, ve_mine, x2);
            if (lane == t2) {
                code = c2;
                rt = __float_as_uint(ev.rtp);
            }
        }
    }
    return make_uint2((code & 4095u) | (died ? 0x80000000u : 0u) | (evaluated ? 0x40000000u : 0u), rt);
}

template <int PT, bool DBG, int NR>
__global__ void __launch_bounds__(32 * LANES_WARPS, LANES_MIN_CTAS) kmc_lanes_kernel(const LayoutDev L, const __grid_constant__ EnsembleDev E) {
    using G = LanesGeom<PT>;
    constexpr int PV = G::PV;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = L.N, S = L.S;
    const int P = PT > 0 ? PT : L.P;
    const int tid = threadIdx.x, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int warp = __shfl_sync(FULL, tid >> 5, 0);

    // ---- stage the layout: pair table for acceptor targets, two planes (i->e, e->i) for the electrodes
    {
        float2 *acc = reinterpret_cast<float2 *>(smem_raw);
        float *elF = reinterpret_cast<float *>(smem_raw + (size_t)N * ROWB);
        float *elR = elF + P * 33;
        // (trip counts that do not depend on the thread, and a __syncwarp() behind: with `idx = tid; idx < P * 33` ptxas 12.9
        // left the warp holding the last elements diverged up to the barrier and ran warp-uniform instructions of the code
        // below once per group, read-modify-write ones included -- compute-sanitizer: shared stores far out of bounds)
        for (int i0 = 0; i0 < N * 33; i0 += blockDim.x) {
            const int idx = i0 + tid;
            if (idx < N * 33) acc[idx] = L.tblf[idx];
        }
        for (int i0 = 0; i0 < P * 33; i0 += blockDim.x) {
            const int idx = i0 + tid;
            if (idx < P * 33) {
                const float2 v = L.tblf[N * 33 + idx];
                elF[idx] = v.x;
                elR[idx] = v.y;
            }
        }
    }
    __syncwarp();
    __syncthreads();

    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t a_elF = sb + (uint32_t)N * ROWB, a_elR = a_elF + (uint32_t)P * ELB;
    const uint32_t wb = sb + (((uint32_t)N * ROWB + 2u * (uint32_t)P * ELB + 15u) & ~15u) + (uint32_t)warp * G::WARP_BYTES;
    const uint32_t a_mir = wb, a_run = wb + G::MIRB, a_tal = a_run + LANES_RMAX * G::RUNB;
    LaneCtx ctx;
    ctx.a_row_me = sb + lane * 8u;    // + j*ROWB     : pair (source lane  -> target j)
    ctx.a_col_me = sb + lane * ROWB;  // + istar*8    : pair (source istar -> target lane)
    ctx.a_elF = a_elF;
    ctx.a_elR = a_elR;
    ctx.a_elF_e = a_elF + lane * ELB;  // + istar*4    : istar -> electrode lane
    ctx.a_elR_e = a_elR + lane * ELB;  // + istar*4    : electrode lane -> istar
    ctx.a_mir = a_mir;
    ctx.a_run = a_run;
    ctx.accm = (1u << N) - 1u;  // N <= 31
    ctx.lane = lane;
    ctx.N = N;
    ctx.P = P;

    const int tlog = E.gtab_log;  // log2(sets per warp slot), >= 6
    const int64_t wslot = (int64_t)blockIdx.x * nwarps + warp;
    unsigned char *const wtab = E.gtab + ((size_t)wslot << tlog) * LSETB;
    // (the table can only be switched off in the DBG instantiation: the hot loop of the production one carries no test for it)
    const bool use_table = DBG ? !(E.lanes_flags & 1) : true;
    ctx.wtab = wtab;
    ctx.use_table = use_table;
    const int total_hops = (int)(E.prehops + E.hops), prehops = (int)E.prehops;  // (the host routes runs of 2^31 hops elsewhere)
    const int64_t nblocks = (E.B + 31) >> 5;
    const int ns = E.lanes_ns;
    // halves: an ensemble that would leave more than half of the device's warp slots empty gives every block of 32 members
    // TWO warps; if a run of identical members starts at member 16 (two runs of 16 seeds: nothing is shared across the
    // halves), they run 16 members each -- a warp's hop costs the same whatever the number of live lanes --, otherwise
    // the first one runs the whole block and the second one returns (a run of 32 keeps ONE table).
    const bool halves = E.lanes_halves != 0;

    // ---- persistent: every warp pulls work items from the global queue
    //      (halves: every unit has a warp slot of its own -- the host checks it --, so a warp's ONE item is its slot number.
    //      Where no block splits, the live units are then the first halves in the first CTAs, which the hardware spreads evenly
    //      over the SMs; taken from a queue by 4096 warps arriving together, the live warps per SM came out binomial.)
    if (halves) {
        if (wslot >= 2 * nblocks) return;
        if (wslot >= nblocks) {  // a second half lives only if a run starts at member 16 of its block
            const int64_t m16 = ((wslot - nblocks) << 5) + 16;
            if (m16 >= E.B) return;
            bool same = __double_as_longlong(E.kT[m16]) == __double_as_longlong(E.kT[m16 - 1]);
            for (int p = lane; p < P; p += 32)
                same = same && __double_as_longlong(E.electrode_v[m16 * P + p]) == __double_as_longlong(E.electrode_v[(m16 - 1) * P + p]);
            if (E.E_constant)
                for (int i = lane; i < N; i += 32)
                    same = same && __double_as_longlong(E.E_constant[m16 * N + i]) == __double_as_longlong(E.E_constant[(m16 - 1) * N + i]);
            if (__all_sync(FULL, same)) return;
        }
    }
    // (what follows keeps nothing of this bookkeeping in registers across the hop loop: everything is re-derived from the
    // kernel parameters per item -- the hop loop spills at the slightest provocation)
    for (;;) {
        int64_t item;
        if (E.lanes_halves) item = wslot;
        else {
            unsigned long long mq = 0;
            if (lane == 0) mq = atomicAdd(E.queue, 1ULL);
            item = (int64_t)__shfl_sync(FULL, mq, 0);
        }
        const int64_t nb_full = E.lanes_nb_full, nb_sl = (nblocks << (E.lanes_halves != 0)) - nb_full;
        if (item >= nb_full + nb_sl * ns) break;
        int64_t unit;
        int hA, hB;
        int s0 = 0, s1 = ns;
        if (item < nb_full) {
            unit = item;
            hA = 0;
            hB = total_hops;
        } else {
            const int64_t q = item - nb_full;
            s0 = (int)(q / nb_sl);
            s1 = s0 + 1;
            unit = nb_full + q % nb_sl;
            hA = s0 * (int)E.lanes_slice_hops;
            hB = (s1 == ns) ? total_hops : s1 * (int)E.lanes_slice_hops;
        }
        const int prog_ix = (int)(unit - nb_full);  // (sliced units: < 2^31 of them)
        // (first halves first, see the queue above)
        const int half = (int)(unit >= nblocks);
        const int64_t blk = unit - (half ? nblocks : 0);
        const bool first = s0 == 0, last = s1 == ns;
        const int64_t base = blk << 5;
        const int64_t m = base + lane;
        const bool present = m < E.B;
        const int64_t mc = present ? m : E.B - 1;
        ctx.base = base;

        // ---- runs of identical members: bit t of sp = member t has exactly the parameters of member t-1
        uint32_t sp;
        {
            bool eq = lane > 0 && present;
            if (eq) {
                eq = __double_as_longlong(E.kT[mc]) == __double_as_longlong(E.kT[mc - 1]);
                for (int p = 0; p < P && eq; ++p)
                    eq = __double_as_longlong(E.electrode_v[mc * P + p]) == __double_as_longlong(E.electrode_v[(mc - 1) * P + p]);
                if (E.E_constant)
                    for (int i = 0; i < N && eq; ++i)
                        eq = __double_as_longlong(E.E_constant[mc * N + i]) == __double_as_longlong(E.E_constant[(mc - 1) * N + i]);
            }
            sp = __ballot_sync(FULL, eq || (!present && lane > 0));
        }
        // bit l = member l starts a run (bit 0 always does); a warp that runs one half of a split block sees the runs of its
        // half only (they share the whole of its table), and its idle lanes count as members of one of them
        const bool split = halves && ((~sp >> 16) & 1u);
        // (a second half whose block does not split never gets here: the taker of the first half runs the whole block)
        const uint32_t lead_mask = split ? (~sp & (half ? 0xffff0000u : 0x0000ffffu)) : ~sp;
        const bool active = present && (!split || (lane >> 4) == half);
        const int lane_r = (split && (lane >> 4) != half) ? (half ? 16 : 15) : lane;
        const int leader = 31 - __clz(lead_mask & (0xffffffffu >> (31 - lane_r)));
        const int nruns = __popc(lead_mask);
        const int slog = tlog - (nruns > 1 ? 32 - __clz(nruns - 1) : 0);  // log2(sets per run) >= 1
        const int hshift = 32 - slog;
        const uint32_t ri = (uint32_t)__popc(lead_mask & ((1u << leader) - 1u));  // index of this thread's run
        const uint32_t setofs = ri << slog;
        ctx.lead_mask = lead_mask;
        ctx.hshift = hshift;
#define LANES_SET(mask) (wtab + (size_t)(setofs + (((mask) * 0x9E3779B1u) >> hshift)) * LSETB)

        // ---- parameters of the first runs into shared memory; tallies; empty table
        __syncwarp();
        {
            uint32_t lm = lead_mask;
            for (int r = 0; r < LANES_RMAX && lm; ++r) {
                const int ld = __ffs(lm) - 1;
                lm &= lm - 1;
                const int64_t mt = base + ld;
                const uint32_t a = a_run + (uint32_t)r * G::RUNB;
                float ef = 0.0f;
                if (lane < N) {
                    double E64;
                    if (E.E_constant) E64 = E.E_constant[mt * N + lane];
                    else {
                        E64 = E.basis[(int64_t)P * N + lane];
                        for (int p = 0; p < P; ++p) E64 += E.electrode_v[mt * P + p] * E.basis[(int64_t)p * N + lane];
                    }
                    ef = (float)E64;
                }
                sts_f(a + lane * 4, ef);
                if (lane < P) sts_f(a + 128 + lane * 4, (float)E.electrode_v[mt * P + lane]);
                if (lane == 0) sts_f(a + 128 + 4 * PV, (float)E.kT[mt]);
            }
        }
        if (use_table)
            for (uint32_t s = lane; s < (2u << tlog); s += 32) stg_u32(wtab + (size_t)s * 64u, 0xffffffffu);

        // ---- state: from the inputs (first slice) or from the checkpoint the previous slice left in the outputs
        uint32_t occ = 0;
        bool alive = active, dead = false;
        double t_acc = 0.0;
        long long n_miss = 0;
        if (first) {
            for (int e = 0; e < P; ++e) sts_u(a_tal + (uint32_t)e * 128u + lane * 4u, 0u);
            if (E.occupation0 && active)
                for (int i = 0; i < N; ++i) occ |= (uint32_t)(E.occupation0[m * N + i] != 0) << i;
            if (DBG && E.avg_occupation && active)
                for (int i = 0; i < N; ++i) E.avg_occupation[m * N + i] = 0.0;
        } else {
            if (lane == 0) {
                const volatile uint32_t *pr = E.lanes_prog + prog_ix;
                while (*pr < (uint32_t)s0) __nanosleep(200);
            }
            __syncwarp();
            __threadfence();
            if (active) {
                occ = __ldcg(E.lanes_ck + m);
                t_acc = __ldcg(E.time + m);
                if (!(t_acc < __longlong_as_double(0x7ff0000000000000LL))) { alive = false; dead = true; }
                if (DBG && E.misses) n_miss = __ldcg(E.misses + m);
            }
            for (int e = 0; e < P; ++e)
                sts_u(a_tal + (uint32_t)e * 128u + lane * 4u, active ? (uint32_t)__ldcg((const long long *)E.electrode_occ + m * P + e) : 0u);
        }
        __syncwarp();

        const uint64_t gm = E.member_index0 + (uint64_t)mc;
        float t_part = 0.0f;
        // sector A of way 0 of the current state's set (fetched right after the previous hop was applied)
        uint32_t f0 = 0xffffffffu, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0, f6 = 0, f7 = 0;
#define LANES_FETCH(mask)                                                                                              \
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                            \
                 : "=r"(f0), "=r"(f1), "=r"(f2), "=r"(f3), "=r"(f4), "=r"(f5), "=r"(f6), "=r"(f7)                     \
                 : "l"(LANES_SET(mask))                                                                                \
                 : "memory")
        if (alive && use_table) LANES_FETCH(occ);
        // (a thread without a live trajectory looks like a sector-A hit -- f0 == occ, X <= f7 -- and is skipped at `apply`)
        if (!alive) { f0 = occ; f7 = 0xffffffffu; }
        // (opaque to ptxas: otherwise it re-derives these from %tid and the kernel parameters inside the hop loop)
        uint32_t a_tal_me = a_tal + lane * 4u, gm_lo = (uint32_t)gm, gm_hi = (uint32_t)(gm >> 32);
        asm volatile("" : "+r"(a_tal_me), "+r"(gm_lo), "+r"(gm_hi));

        // one hop of this thread's trajectory: er -> dwell time, xr -> event
        bool stop = false;
#define LANES_HOP(lg_, xr_, h_)                                                                                        \
    do {                                                                                                               \
        const uint32_t xr = (xr_);                                                                                     \
        const float lg = (lg_);                                                                                        \
        const uint32_t X = xr | 0xfffu;                                                                                \
        uint32_t code = select6(X, f2, f3, f4, f5, f6, f7);                                                            \
        uint32_t rt = f1;                                                                                              \
        const uint32_t st = (f0 != occ) ? 1u : (X > f7 ? 2u : 0u);                                                     \
        if (__any_sync(FULL, st != 0)) {                                                                               \
            /* hop-loop bookkeeping the evaluation does not need waits in shared memory meanwhile: ptxas otherwise */   \
            /* spills it for good and reloads it in EVERY hop (2-3 LDL per hop in front of the loop branch) */         \
            const uint32_t a_st = a_tal_me + G::TALB;                                                                  \
            sts_u(a_st, (uint32_t)q1); sts_u(a_st + 128u, blk0); sts_u(a_st + 256u, gm_lo);                            \
            sts_u(a_st + 384u, gm_hi); sts_u(a_st + 512u, (uint32_t)hend); sts_u(a_st + 640u, (uint32_t)h0);           \
            const uint2 r = lanes_cold<PT, DBG, NR>(E, ctx, occ, xr, st, f1, f7, ri, setofs, leader);                  \
            q1 = (int)lds_u(a_st); blk0 = lds_u(a_st + 128u); gm_lo = lds_u(a_st + 256u);                              \
            gm_hi = lds_u(a_st + 384u); hend = (int)lds_u(a_st + 512u); h0 = (int)lds_u(a_st + 640u);                  \
            if (st) {                                                                                                  \
                code = r.x;                                                                                            \
                rt = r.y;                                                                                              \
                if (r.x >> 31) { alive = false; dead = true; f0 = occ; f7 = 0xffffffffu; }                             \
                if (DBG && ((r.x >> 30) & 1u)) ++n_miss;                                                               \
            }                                                                                                          \
            if (!__any_sync(FULL, alive)) { stop = true; break; }                                                      \
        }                                                                                                              \
        /* apply (simulation.go:107-130, 306-319): the code's acceptor flips; an acceptor partner flips too; an */    \
        /* electrode partner gains (32+e) or loses (64+e) one hole */                                                  \
        if (alive) {                                                                                                   \
            const uint32_t evt = (code >> 5) & 127u;                                                                   \
            occ ^= bit_wrap(code) | bit_clamp(evt);                                                                    \
            if (evt >= 32u) {                                                                                          \
                const uint32_t a = a_tal_me + (evt & 31u) * 128u;                                                      \
                sts_u(a, lds_u(a) + 1u - ((evt >> 5) & 2u));                                                           \
            }                                                                                                          \
            t_part = fmaf(lg, __uint_as_float(rt), t_part);                                                            \
            if (DBG && (E.trace || E.traffic || E.avg_occupation) && (h_) >= prehops) {                                \
                const int site = (int)(code & 31u);                                                                    \
                int from, to;                                                                                          \
                if (evt < 32u) { from = site; to = (int)evt; }                                                         \
                else if (evt < 64u) { from = site; to = N + (int)evt - 32; }                                           \
                else { from = N + (int)evt - 64; to = site; }                                                          \
                if (E.trace) {                                                                                         \
                    int32_t *tp = E.trace + (m * E.hops + (int64_t)((h_) - prehops)) * 2;                              \
                    tp[0] = from;                                                                                      \
                    tp[1] = to;                                                                                        \
                }                                                                                                      \
                if (E.traffic) { /* simulation.go:310-311: antisymmetric hop counts */                                 \
                    double *tr = E.traffic + m * (int64_t)S * S;                                                       \
                    tr[from * S + to] += 1.0;                                                                          \
                    tr[to * S + from] -= 1.0;                                                                          \
                }                                                                                                      \
                if (E.avg_occupation) { /* simulation.go:312-316 as time stamps: a site collects t_off - t_on */      \
                    const double now = t_acc + (double)t_part;                                                         \
                    double *row = E.avg_occupation + m * N;                                                            \
                    if (from < N) row[from] += now;                                                                    \
                    if (to < N) row[to] -= now;                                                                        \
                }                                                                                                      \
            }                                                                                                          \
            if (use_table) LANES_FETCH(occ);                                                                           \
        }                                                                                                              \
    } while (0)

        // segments: a segment never straddles a 64-hop variate block or the prehops boundary
        for (int h0 = hA; h0 < hB && !stop;) {
            if (h0 == prehops && prehops > 0) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
                t_acc = 0.0;
                t_part = 0.0f;
                for (int e = 0; e < P; ++e) sts_u(a_tal + (uint32_t)e * 128u + lane * 4u, 0u);
                if (DBG && E.avg_occupation && active)
                    for (int i = 0; i < N; ++i) E.avg_occupation[m * N + i] = 0.0;
            } else {
                t_acc += (double)t_part;
                t_part = 0.0f;
            }
            int hend = (h0 | 63) + 1;
            if (hend > hB) hend = hB;
            if (h0 < prehops && hend > prehops) hend = prehops;
            const int q0 = h0 & 63;
            int q1 = q0 + (hend - h0);
            uint32_t blk0 = (uint32_t)(h0 >> 6) * 32u;  // (< 2^30: the counter's second word stays 0)
            uint4 r4 = make_uint4(0u, 0u, 0u, 0u);
            for (int q = q0; q < q1 && !stop; ++q) {
                uint32_t er, xq;
                if (!(q & 1) || q == q0) {  // (q0 odd: second half of a pair whose first hop belonged to the previous segment)
                    r4 = philox_rk(make_uint4(blk0 + (uint32_t)(q >> 1), 0u, gm_lo, gm_hi), E.rk);
                }
                if (q & 1) { er = r4.z; xq = r4.w; }
                else { er = r4.x; xq = r4.y; }
                // dwell time = lg * (-ln2 / total): lg = log2 of a uniform in (0, 1), or of exp(-e) for an injected variate
                float lgv;
                if (DBG && E.stream_e) {  // injected stream: e = Exp(1) variate (f64), u = uniform in [0, 1) (f32), simulation.go:297-299
                    const int64_t ix = m * (int64_t)total_hops + (h0 + (q - q0));
                    lgv = active ? (float)(-1.4426950408889634 * E.stream_e[ix]) : 0.0f;
                    xq = active ? __float2uint_rz(E.stream_u[ix] * 4294967296.0f) : 0u;
                } else lgv = lg2_approx(fmaf((float)er, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
                LANES_HOP(lgv, xq, h0 + (q - q0));
            }
            h0 = hend;
        }
#undef LANES_HOP
#undef LANES_FETCH

        // ---- results (last slice) or checkpoint
        t_acc += (double)t_part;
        if (dead) t_acc = __longlong_as_double(0x7ff0000000000000LL);  // +inf, as time_step = e/0 would give
        __syncwarp();
        if (active) {
            E.time[m] = t_acc;
            for (int e = 0; e < P; ++e) E.electrode_occ[m * P + e] = (int64_t)(int32_t)lds_u(a_tal + (uint32_t)e * 128u + lane * 4u);
            if (DBG && E.misses) E.misses[m] = n_miss;
            if (!last) E.lanes_ck[m] = occ;
            else {
                if (E.occupation_out)
                    for (int i = 0; i < N; ++i) E.occupation_out[m * N + i] = (occ >> i) & 1u;
                if (DBG && E.avg_occupation)  // sites still occupied collect t_end - t_on
                    for (int i = 0; i < N; ++i)
                        if ((occ >> i) & 1u) E.avg_occupation[m * N + i] += t_acc;
            }
        }
        if (last && E.site_energies_out) {
            const uint32_t actm = __ballot_sync(FULL, active);
            for (int t = 0; t < 32; ++t) {
                const int64_t mt = base + t;
                if (mt >= E.B) break;
                if (!((actm >> t) & 1u)) continue;  // (the other half of a split block)
                const uint32_t occu = __shfl_sync(FULL, occ, t);
                double E64;
                float ve_t, kTt;
                run_params<PT>(E, ctx, (int)__shfl_sync(FULL, ri, t), __shfl_sync(FULL, leader, t), E64, ve_t, kTt);
                if (lane < N) E.site_energies_out[mt * S + lane] = energy_of(occu, ctx.accm, E64, ctx.a_row_me);
                if (lane < P) E.site_energies_out[mt * S + N + lane] = (double)ve_t;
            }
        }
        if (!last) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(E.lanes_prog + prog_ix, (uint32_t)s1);
        }
        __syncwarp();
#undef LANES_SET
        if (E.lanes_halves) break;  // (one unit per warp slot)
    }  // work items
}


// =====================================================================================================================
// kmc_solo_kernel -- a FEW trajectories, each as fast as one serial chain can go (the drop-in's single go_simulation call).
//
// A lone Markov chain is bound by the latency of its dependent instructions -- a single warp issues a dependent instruction
// every 4-5 cycles -- and on the kernels above that chain runs through hashing the state, a table read, the select tree, the
// mask update, tallies and the next hash: 90 ns per hop on the warp-per-trajectory kernel against 28 ns on one CPU core with
// the reference's cache.  Here one warp owns one trajectory and keeps the states it has visited as a GRAPH in shared memory:
// an entry holds the event words of the kernel above (same evaluate_state, same tail_pick: results are bit-identical with
// kmc_lanes_kernel) and, for each of its 4 most likely events, the shared-memory ADDRESS of the successor's entry.  Lane 0
// WALKS the graph and does nothing else: per hop two 128-bit shared loads at a known address, four compares, three selects on
// the successor addresses, one store of the address it was at.  No hash, no mask, no tallies, no time on the chain.  A
// successor that is not known yet points to a trap entry whose thresholds send the walk to the exit it already has for the
// rare events (beyond the 4 most likely ones: 0.2-0.8 % of the hops on C3).  Everything else is done by all 32 lanes for 64 hops at a time: the variates
// before the walk (Philox numbering of the kernels above), and after it -- from the addresses the walk left behind -- event codes,
// electrode tallies (ballots) and the dwell times (one lane, in hop order: fp32 sums must not be re-associated).  Leaving
// the walk: the warp resolves the event (second sector, or evaluate + tail_pick), finds or evaluates the successor (lane =
// acceptor, as above) and links it.  A table that fills up is dropped and rebuilt (a pure cache).  Record outputs, traces and
// injected streams stay on the kernels above.
#define SOLO_ENT 128u    // bytes per entry: e0-e3 | s0-s3 | key rt e4 e5 | e6-e9 | e10-e13 | t4..t13 (u16 entry numbers) | -
#define SOLO_HS 4096u    // slots of the mask -> entry hash (u16 entry numbers, 0 = empty)

template <int PT>
struct SoloGeom {
    using G = LanesGeom<PT>;
    // per hop of a 64-hop block: x | 0xfff, lg2(u), entry address, -ln2/total, event code, x -- three blocks (the one being walked,
    // the one being accounted for, the one being filled with variates); then 8 control words
    static constexpr int RINGB = 64 * 4 * 6;
    static constexpr int FIXED = G::WARP_BYTES + 3 * RINGB + 32 + (int)SOLO_HS * 2;
};

// (not inlined: both warps meet at the SAME instruction, which is what compute-sanitizer's synccheck expects of a barrier that
// counts every thread of the CTA)
__device__ __noinline__ void solo_rendezvous() {
    __syncwarp();  // (bar.sync is the .aligned form: the warp must arrive as one)
    asm volatile("bar.sync 1, 64;" ::: "memory");
}
// warp 1 has read the entries the previous block points to (does not wait) / warp 0 waits for that
__device__ __forceinline__ void solo_entries_read() {
    __syncwarp();
    asm volatile("bar.arrive 2, 64;" ::: "memory");
}
__device__ __forceinline__ void solo_wait_entries_read() {
    __syncwarp();
    asm volatile("bar.sync 2, 64;" ::: "memory");
}
__device__ __forceinline__ uint32_t lds_u_volatile(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// Two warps per trajectory.  Warp 0 walks block b (and does everything that needs the layout: evaluations, tail picks, linking);
// warp 1 meanwhile turns the addresses block b-1 left behind into event codes, tallies and dwell times and draws the variates of
// block b+1.  They meet once per block (named barrier 1).  Before that warp 1 signals (arrives at barrier 2, without waiting) that
// it has read the entries block b-1 points to; warp 0 waits for that signal at the end of its walk, when it came long ago -- or
// earlier, if it has to drop a full table.
template <int PT, int NR>
__global__ void __launch_bounds__(64) kmc_solo_kernel(const LayoutDev L, const __grid_constant__ EnsembleDev E) {
    using G = LanesGeom<PT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = L.N, S = L.S;
    const int P = PT > 0 ? PT : L.P;
    const int tid = threadIdx.x, lane = tid & 31;
    const int role = __shfl_sync(FULL, tid >> 5, 0);  // 0: walks, 1: helps
    {
        float2 *acc = reinterpret_cast<float2 *>(smem_raw);
        float *elF = reinterpret_cast<float *>(smem_raw + (size_t)N * ROWB);
        float *elR = elF + P * 33;
        for (int i0 = 0; i0 < N * 33; i0 += 64)
            if (i0 + tid < N * 33) acc[i0 + tid] = L.tblf[i0 + tid];
        for (int i0 = 0; i0 < P * 33; i0 += 64)
            if (i0 + tid < P * 33) {
                const float2 v = L.tblf[N * 33 + i0 + tid];
                elF[i0 + tid] = v.x;
                elR[i0 + tid] = v.y;
            }
    }
    __syncwarp();
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t a_elF = sb + (uint32_t)N * ROWB, a_elR = a_elF + (uint32_t)P * ELB;
    const uint32_t wb = sb + (((uint32_t)N * ROWB + 2u * (uint32_t)P * ELB + 15u) & ~15u);
    const uint32_t a_mir = wb, a_run = wb + G::MIRB;
    const uint32_t a_ring = wb + G::WARP_BYTES;  // block b: a_ring + (b % 3) * RINGB + {0 X, 256 lg, 512 entry, 768 rt, 1024 code, 1280 x}
    const uint32_t a_ctl = a_ring + 3u * SoloGeom<PT>::RINGB;  // +0 / +4 / +8: hops done in the block (by b % 3), +12 dead
    const uint32_t a_hash = a_ctl + 32u;
    const uint32_t a_ent = (a_hash + SOLO_HS * 2u + 127u) & ~127u;  // entry e at a_ent + e * 128; entry 0 is the trap
    const int emax = E.solo_emax;
    const int total_hops = (int)(E.prehops + E.hops), prehops = (int)E.prehops;
    // the trap: thresholds 0, so that every x leaves the walk
    if (role == 0) sts_u(a_ent + lane * 4u, 0u);
    __syncthreads();

    for (int64_t m = blockIdx.x; m < E.B; m += gridDim.x) {
        const uint64_t gm = E.member_index0 + (uint64_t)m;
        const uint32_t gm_lo = (uint32_t)gm, gm_hi = (uint32_t)(gm >> 32);
        if (tid < 8) sts_u(a_ctl + tid * 4u, 0u);
        __syncthreads();

        if (role == 1) {
            // =========================================== warp 1: variates before, codes / tallies / time after ===========
            auto gen = [&](int h0, int par) {  // lane l draws hops 2l and 2l+1 of the 64-hop block h0 lies in
                const uint32_t rb = a_ring + (uint32_t)par * SoloGeom<PT>::RINGB;
                const uint32_t blk0 = (uint32_t)(h0 >> 6) * 32u;
                const uint4 r4 = philox_rk(make_uint4(blk0 + (uint32_t)lane, 0u, gm_lo, gm_hi), E.rk);
                sts_f(rb + 256u + (2 * lane) * 4u, lg2_approx(fmaf((float)r4.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f)));
                sts_f(rb + 256u + (2 * lane + 1) * 4u, lg2_approx(fmaf((float)r4.z, 2.3283064365386963e-10f, 1.1641532182693481e-10f)));
                sts_u(rb + (2 * lane) * 4u, r4.y | 0xfffu);
                sts_u(rb + (2 * lane + 1) * 4u, r4.w | 0xfffu);
                sts_u(rb + 1280u + (2 * lane) * 4u, r4.y);
                sts_u(rb + 1280u + (2 * lane + 1) * 4u, r4.w);
            };
            int tally = 0;        // lane e < P: net holes into electrode e
            double t_acc = 0.0;   // (lane 0's copy is the one that counts)
            float t_part = 0.0f;
            auto post = [&](int ri, int h0) {  // the block in ring slot ri started at hop h0; the walk did its hops [q0, qd)
                const uint32_t rb = a_ring + (uint32_t)ri * SoloGeom<PT>::RINGB;
                const int q0 = h0 & 63, qd = (int)lds_u(a_ctl + (uint32_t)ri * 4u);
                if (h0 == prehops && prehops > 0) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
                    t_acc = 0.0;
                    t_part = 0.0f;
                    tally = 0;
                } else {
                    t_acc += (double)t_part;
                    t_part = 0.0f;
                }
                for (int h = q0 + lane; h < qd; h += 32) {
                    const uint32_t a = lds_u(rb + 512u + h * 4u);
                    if (a) {
                        const uint32_t X = lds_u(rb + h * 4u);
                        const uint4 A = lds_u4(a);
                        sts_u(rb + 1024u + h * 4u, select4(X, A.x, A.y, A.z, A.w) & 4095u);
                        sts_u(rb + 768u + h * 4u, lds_u(a + 36u));
                    }
                }
                solo_entries_read();  // the entries that block points to are not needed any more: warp 0 may drop the table
                for (int hb = q0; hb < qd; hb += 32) {
                    const int h = hb + lane;
                    const uint32_t evt = (h < qd) ? (lds_u(rb + 1024u + h * 4u) >> 5) : 0u;
                    for (int e = 0; e < P; ++e) {
                        const int d = __popc(__ballot_sync(FULL, evt == 32u + (uint32_t)e)) - __popc(__ballot_sync(FULL, evt == 64u + (uint32_t)e));
                        if (lane == e) tally += d;
                    }
                }
                if (lane == 0) {  // four hops per pair of 128-bit loads, the sums strictly in hop order
                    const uint32_t a_lg = rb + 256u, a_rt = rb + 768u;
                    int h = q0;
                    for (; h < qd && (h & 3); ++h) t_part = fmaf(lds_f(a_lg + h * 4u), lds_f(a_rt + h * 4u), t_part);
#pragma unroll 4
                    for (; h + 4 <= qd; h += 4) {
                        const uint4 g = lds_u4(a_lg + h * 4u), r = lds_u4(a_rt + h * 4u);
                        t_part = fmaf(__uint_as_float(g.x), __uint_as_float(r.x), t_part);
                        t_part = fmaf(__uint_as_float(g.y), __uint_as_float(r.y), t_part);
                        t_part = fmaf(__uint_as_float(g.z), __uint_as_float(r.z), t_part);
                        t_part = fmaf(__uint_as_float(g.w), __uint_as_float(r.w), t_part);
                    }
                    for (; h < qd; ++h) t_part = fmaf(lds_f(a_lg + h * 4u), lds_f(a_rt + h * 4u), t_part);
                }
                __syncwarp();
            };
            auto end_of = [&](int h0) {
                int hend = (h0 | 63) + 1;
                if (hend > total_hops) hend = total_hops;
                if (h0 < prehops && hend > prehops) hend = prehops;
                return hend;
            };
            if (total_hops > 0) gen(0, 0);
            solo_rendezvous();  // the variates of block 0 are there
            int b = 0, h_prev = 0, ri = 0;  // ri = b % 3
            for (int h0 = 0; h0 < total_hops; ++b) {
                const int hend = end_of(h0);
                if (hend < total_hops) gen(hend, ri == 2 ? 0 : ri + 1);
                if (b >= 1) post(ri == 0 ? 2 : ri - 1, h_prev);  // (meets warp 0 once inside)
                h_prev = h0;
                h0 = hend;
                ri = ri == 2 ? 0 : ri + 1;
                solo_rendezvous();  // block b is walked; block b-1 is accounted for; the variates of block b+1 are there
            }
            if (b >= 1) post(ri == 0 ? 2 : ri - 1, h_prev);
            solo_rendezvous();
            t_acc += (double)t_part;
            if (lds_u(a_ctl + 12u)) t_acc = __longlong_as_double(0x7ff0000000000000LL);  // dead: +inf, as e/0 would give
            if (lane == 0) E.time[m] = t_acc;
            if (lane < P) E.electrode_occ[m * P + lane] = (int64_t)tally;
        } else {
            // =========================================== warp 0: the walk and everything that needs the layout ===========
            LaneCtx ctx;
            ctx.a_row_me = sb + lane * 8u;
            ctx.a_col_me = sb + lane * ROWB;
            ctx.a_elF = a_elF;
            ctx.a_elR = a_elR;
            ctx.a_elF_e = a_elF + lane * ELB;
            ctx.a_elR_e = a_elR + lane * ELB;
            ctx.a_mir = a_mir;
            ctx.a_run = a_run;
            ctx.accm = (1u << N) - 1u;
            ctx.lane = lane;
            ctx.N = N;
            ctx.P = P;
            ctx.lead_mask = 1u;
            ctx.hshift = 0;
            ctx.wtab = nullptr;
            ctx.base = 0;
            ctx.use_table = false;
            // ---- member parameters (as the kernel above narrows them)
            double E64 = 0.0;
            if (lane < N) {
                double e64;
                if (E.E_constant) e64 = E.E_constant[m * N + lane];
                else {
                    e64 = E.basis[(int64_t)P * N + lane];
                    for (int p = 0; p < P; ++p) e64 += E.electrode_v[m * P + p] * E.basis[(int64_t)p * N + lane];
                }
                E64 = (double)(float)e64;
            }
            const float ve_mine = (lane < P) ? (float)E.electrode_v[m * P + lane] : 0.0f;
            const float kTt = (float)E.kT[m];
            uint32_t occ0 = 0;
            if (E.occupation0 && lane == 0)
                for (int i = 0; i < N; ++i) occ0 |= (uint32_t)(E.occupation0[m * N + i] != 0) << i;
            occ0 = __shfl_sync(FULL, occ0, 0);

            int count = 0, generation = 0;  // entries in the table, times it was dropped (warp-uniform)
            int q_mat = 0;                   // hops of the current block below q_mat have their code / rt slots filled
            int ri = 0;                      // ring slot of the block being walked (block number % 3)
            bool met = false;                // this block's first meeting with warp 1 is behind us
            uint32_t rb = a_ring;            // its slots
            // fills code / rt of the hops [q_mat, qe) the walk went through (their entry slot holds the entry they left from)
            auto materialise = [&](int qe) {
                __syncwarp();
                for (int h = q_mat + lane; h < qe; h += 32) {
                    const uint32_t a = lds_u(rb + 512u + h * 4u);
                    if (a) {
                        const uint32_t X = lds_u(rb + h * 4u);
                        const uint4 A = lds_u4(a);
                        sts_u(rb + 1024u + h * 4u, select4(X, A.x, A.y, A.z, A.w) & 4095u);
                        sts_u(rb + 768u + h * 4u, lds_u(a + 36u));
                        sts_u(rb + 512u + h * 4u, 0u);
                    }
                }
                q_mat = qe;
                __syncwarp();
            };
            auto reset = [&]() {
                __syncwarp();
                for (uint32_t s = lane; s < SOLO_HS / 2u; s += 32) sts_u(a_hash + s * 4u, 0u);
                count = 0;
                ++generation;
                __syncwarp();
            };
            // entry number of state occu, 0 if it is not in the table (lane 0 probes; warp-uniform result)
            auto lookup = [&](uint32_t occu) {
                uint32_t found = 0;
                if (lane == 0) {
                    uint32_t slot = (occu * 0x9E3779B1u) >> 20;  // 12 bits
                    for (;;) {
                        const uint32_t w = lds_u(a_hash + (slot >> 1) * 4u);
                        const uint32_t e = (slot & 1u) ? (w >> 16) : (w & 0xffffu);
                        if (!e) break;
                        if (lds_u(a_ent + e * SOLO_ENT + 32u) == occu) { found = e; break; }
                        slot = (slot + 1u) & (SOLO_HS - 1u);
                    }
                }
                return __shfl_sync(FULL, found, 0);
            };
            // evaluates state occu and appends its entry (a state without any transition gets an entry without events: the walk
            // leaves it at once, and the re-evaluation in the tail finds it dead); returns the entry number.  qd: hops of the
            // current block done so far (what they point to must be read before a full table is dropped -- and warp 1 must
            // have read what the previous block points to)
            auto insert = [&](uint32_t occu, int qd) {
                if (count >= emax) {
                    materialise(qd);
                    if (!met) {
                        solo_wait_entries_read();
                        met = true;
                    }
                    reset();
                }
                Eval<NR> ev;
                const bool ok = evaluate_state<PT, NR>(ctx, occu, E64, ve_mine, kTt, ev);
                const uint32_t e = (uint32_t)++count;
                const uint32_t a = a_ent + e * SOLO_ENT;
                const uint32_t w = ok ? ev.word : 0u;
                __syncwarp();
                if (lane < 4) sts_u(a + lane * 4u, w);
                else if (lane < 6) sts_u(a + 40u + (lane - 4) * 4u, w);
                else if (lane < LANES_K) sts_u(a + 48u + (lane - 6) * 4u, w);
                else if (lane == LANES_K) sts_u(a + 32u, occu);
                else if (lane == LANES_K + 1) sts_u(a + 36u, ok ? __float_as_uint(ev.rtp) : 0u);
                else if (lane < LANES_K + 2 + 4) sts_u(a + 16u + (lane - LANES_K - 2) * 4u, a_ent);  // successors of events 0..3: the trap
                else if (lane < LANES_K + 2 + 4 + 5) sts_u(a + 80u + (lane - LANES_K - 6) * 4u, 0u);  // successors of events 4..13: unknown
                if (lane == 0) {
                    uint32_t slot = (occu * 0x9E3779B1u) >> 20;
                    for (;;) {
                        const uint32_t hw = lds_u(a_hash + (slot >> 1) * 4u);
                        const uint32_t c = (slot & 1u) ? (hw >> 16) : (hw & 0xffffu);
                        if (!c) {
                            sts_u(a_hash + (slot >> 1) * 4u, (slot & 1u) ? (hw | (e << 16)) : (hw | e));
                            break;
                        }
                        slot = (slot + 1u) & (SOLO_HS - 1u);
                    }
                }
                __syncwarp();
                return e;
            };

            reset();
            uint32_t cur = a_ent + insert(occ0, 0) * SOLO_ENT;  // address of the current state's entry (warp-uniform between walks)
            bool dead = false;
            solo_rendezvous();  // the variates of block 0 are there

            met = true;  // (nothing of an earlier block is left before block 0)
            for (int h0 = 0; h0 < total_hops; ri = ri == 2 ? 0 : ri + 1) {
                int hend = (h0 | 63) + 1;
                if (hend > total_hops) hend = total_hops;
                if (h0 < prehops && hend > prehops) hend = prehops;
                const int q0 = h0 & 63, q1 = q0 + (hend - h0);
                rb = a_ring + (uint32_t)ri * SoloGeom<PT>::RINGB;
                const uint32_t a_X = rb, a_tr = rb + 512u, a_rt = rb + 768u, a_cd = rb + 1024u, a_xr = rb + 1280u;
                int q = q0;
                q_mat = q0;
                while (q < q1 && !dead) {
                    // ---- the walk (lane 0): until the block is used up or a uniform lies beyond the entry's first 4 events
                    if (lane == 0) {
                        // ONE branch per hop, at the end: with an exit branch in the middle ptxas sinks two of the loads below
                        // it and the walk pays the shared-memory latency twice per hop.  A hop that leaves has stored its entry
                        // like the others (harmless) and moved on to a successor that is not used: the entry is read back.
                        uint32_t a = cur, pX = a_X + q * 4u;
                        const uint32_t pEnd = a_X + q1 * 4u;
                        bool out = false;
                        // one hop: entry loads, the store of where we are, exit test, successor; the exit branch comes last
#define SOLO_HOP(Xv)                                                                                                   \
    {                                                                                                                  \
        const uint32_t X = (Xv);                                                                                       \
        const uint4 A = lds_u4(a), Sx = lds_u4(a + 16u); /* e0-e3 | s0-s3 */                                           \
        sts_u(pX + 512u, a);                                                                                           \
        out = X > A.w;                                                                                                 \
        const uint32_t s01 = X > A.x ? Sx.y : Sx.x, s23 = X > A.z ? Sx.w : Sx.z;                                       \
        a = X > A.y ? s23 : s01;                                                                                       \
        pX += 4u;                                                                                                      \
    }
                        // (the variates four at a time where the block allows it: the lone warp is bound by the number of
                        // shared-memory instructions it can issue as much as by their latency)
                        while (!out && pX < pEnd && (pX & 15u)) SOLO_HOP(lds_u(pX));
                        while (!out && pX + 16u <= pEnd) {
                            const uint4 X4 = lds_u4(pX);
                            SOLO_HOP(X4.x);
                            if (out) break;
                            SOLO_HOP(X4.y);
                            if (out) break;
                            SOLO_HOP(X4.z);
                            if (out) break;
                            SOLO_HOP(X4.w);
                        }
                        while (!out && pX < pEnd) SOLO_HOP(lds_u(pX));
#undef SOLO_HOP
                        if (out) {
                            pX -= 4u;
                            a = lds_u(pX + 512u);
                        }
                        q = (int)((pX - a_X) >> 2);
                        cur = a;
                    }
                    q = __shfl_sync(FULL, q, 0);
                    cur = __shfl_sync(FULL, cur, 0);
                    if (cur == a_ent) {
                        // ---- the trap: hop q-1 left entry `from` by an event whose successor was not known
                        const int qp = q - 1;
                        const uint32_t from = lds_u(a_tr + qp * 4u);
                        const uint32_t Xp = lds_u(a_X + qp * 4u);
                        const uint4 A = lds_u4(from);
                        const uint32_t k = (Xp > A.x) + (Xp > A.y) + (Xp > A.z);
                        const uint32_t code = select4(Xp, A.x, A.y, A.z, A.w) & 4095u;
                        const uint32_t succ = lds_u(from + 32u) ^ (bit_wrap(code) | bit_clamp((code >> 5) & 127u));
                        uint32_t e = lookup(succ);
                        const int gen0 = generation;
                        if (!e) e = insert(succ, q);
                        cur = a_ent + e * SOLO_ENT;
                        if (generation == gen0 && lane == 0) sts_u(from + 16u + k * 4u, cur);  // (a dropped table took `from` with it)
                        __syncwarp();
                        continue;
                    }
                    if (q >= q1) break;
                    // ---- hop q of entry `cur` lies beyond its first 4 events: events 4 .. 13, or the tail
                    const uint32_t X = lds_u(a_X + q * 4u);
                    const uint4 W1 = lds_u4(cur + 32u), W2 = lds_u4(cur + 48u), W3 = lds_u4(cur + 64u);  // key rt e4 e5 | e6-e9 | e10-e13
                    const uint32_t occ = W1.x;
                    uint32_t code, e = 0, k = 0;
                    float rtv = __uint_as_float(W1.y);
                    const bool second = !(X > W3.w);
                    if (second) {
                        const uint32_t ws[10] = {W1.z, W1.w, W2.x, W2.y, W2.z, W2.w, W3.x, W3.y, W3.z, W3.w};
                        uint32_t word = ws[0];
#pragma unroll
                        for (int j = 0; j < 9; ++j)
                            if (X > ws[j]) {  // (the words ascend: the last j with X > ws[j] decides)
                                k = (uint32_t)j + 1u;
                                word = ws[j + 1];
                            }
                        code = word & 4095u;  // event 4 + k
                        const uint32_t tw = lds_u(cur + 80u + (k >> 1) * 4u);
                        e = (k & 1u) ? (tw >> 16) : (tw & 0xffffu);
                    } else {  // tail: evaluate again, exact pick
                        Eval<NR> ev;
                        if (!evaluate_state<PT, NR>(ctx, occ, E64, ve_mine, kTt, ev)) {
                            dead = true;
                            break;
                        }
                        code = tail_pick<NR>(ctx, ev, occ, kTt, ve_mine, lds_u(a_xr + q * 4u));
                        rtv = ev.rtp;
                    }
                    // this hop's code and rate go straight into its slots
                    if (lane == 0) {
                        sts_u(a_cd + q * 4u, code);
                        sts_f(a_rt + q * 4u, rtv);
                        sts_u(a_tr + q * 4u, 0u);
                    }
                    const uint32_t succ = occ ^ (bit_wrap(code) | bit_clamp((code >> 5) & 127u));
                    const int gen0 = generation;
                    const uint32_t from = cur;
                    if (!e) {
                        e = lookup(succ);
                        if (!e) e = insert(succ, q + 1);
                        if (second && generation == gen0 && lane == 0) {
                            const uint32_t ap = from + 80u + (k >> 1) * 4u;
                            const uint32_t tw = lds_u(ap);
                            sts_u(ap, (k & 1u) ? ((tw & 0xffffu) | (e << 16)) : ((tw & 0xffff0000u) | e));
                        }
                    }
                    __syncwarp();
                    cur = a_ent + e * SOLO_ENT;
                    ++q;
                }
                if (lane == 0) {
                    sts_u(a_ctl + (uint32_t)ri * 4u, (uint32_t)q);  // hops [q0, q) of this block are done
                    if (dead) sts_u(a_ctl + 12u, 1u);
                }
                h0 = hend;
                if (!met) solo_wait_entries_read();  // (warp 1 has read what the previous block points to: normally long ago)
                solo_rendezvous();            // (warp 1: the previous block accounted for, variates of the next block drawn)
                met = false;
            }
            if (!met) solo_wait_entries_read();
            solo_rendezvous();  // warp 1 has accounted for the last block
            const uint32_t occ = lds_u(cur + 32u);
            if (E.occupation_out && lane < N) E.occupation_out[m * N + lane] = (occ >> lane) & 1u;
            if (E.site_energies_out) {
                if (lane < P) sts_f(a_mir + 128 + lane * 4, ve_mine);
                if (lane < N) E.site_energies_out[m * S + lane] = energy_of(occ, ctx.accm, E64, ctx.a_row_me);
                if (lane < P) E.site_energies_out[m * S + N + lane] = (double)ve_mine;
            }
        }
        __syncthreads();
    }
}

template <int PT>
static cudaError_t launch_solo_t(const LayoutDev &L, EnsembleDev E, cudaStream_t st, int *launches) {
    const size_t fixed = (((size_t)L.N * ROWB + 2 * (size_t)L.P * ELB + 15) & ~size_t(15)) + SoloGeom<PT>::FIXED + 128;
    const int nr = L.N <= 12 ? 3 : 2;
    auto kern = nr == 3 ? kmc_solo_kernel<PT, 3> : kmc_solo_kernel<PT, 2>;
    int dev = 0, sms = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // one CTA per member; two per SM when there are more members than SMs (the table shrinks accordingly)
    const int per_sm = E.B > sms ? 2 : 1;
    size_t budget = per_sm == 1 ? (size_t)optin : (size_t)((optin + 1024) / 2 - 1024);
    if (budget <= fixed + 64 * SOLO_ENT) return cudaErrorInvalidValue;
    int emax = (int)((budget - fixed) / SOLO_ENT) - 1;
    if (emax > 2700) emax = 2700;  // (the hash has 4096 slots)
    if (const char *ev = getenv("KMCB200_SOLO_EMAX")) emax = std::max(2, std::min(emax, atoi(ev)));  // (tests: a table that keeps filling up)
    E.solo_emax = emax;
    const size_t smem = fixed + (size_t)(emax + 1) * SOLO_ENT;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int64_t cap = (int64_t)sms * per_sm;
    const unsigned grid = (unsigned)(E.B < cap ? E.B : cap);
    kern<<<grid, 64, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// A few trajectories of a layout with N <= 31 acceptors, fewer than 2^31 hops, no record outputs: one warp each, state graph in
// shared memory.  E.rk must hold the Philox key schedule (as for launch_lanes).
cudaError_t launch_solo(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches) {
    if (E.B <= 0) return cudaSuccess;
    if (L.N > 31 || L.P > 32 || E.prehops + E.hops >= (int64_t)1 << 31) return cudaErrorInvalidValue;
    if (E.trace || E.misses || E.traffic || E.avg_occupation || E.stream_e) return cudaErrorInvalidValue;
    switch (L.P) {
        case 2: return launch_solo_t<2>(L, E, st, launches);
        case 8: return launch_solo_t<8>(L, E, st, launches);
        default: return launch_solo_t<0>(L, E, st, launches);
    }
}

template <int PT>
static cudaError_t launch_lanes_t(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches, MemoPlan *plan_only) {
    using G = LanesGeom<PT>;
    // (record outputs, injected streams and the table-off switch exist only in the DBG instantiation)
    const bool dbg = E.trace || E.misses || (E.lanes_flags & 1) || E.traffic || E.avg_occupation || E.stream_e;
    const int warps = LANES_WARPS;
    const size_t smem = (((size_t)L.N * ROWB + 2 * (size_t)L.P * ELB + 15) & ~size_t(15)) + (size_t)warps * G::WARP_BYTES;
    // candidates per acceptor for the K largest events of a state
    const int nr = L.N <= 12 ? 3 : 2;
    auto kern = nr == 3 ? (dbg ? kmc_lanes_kernel<PT, true, 3> : kmc_lanes_kernel<PT, false, 3>)
                        : (dbg ? kmc_lanes_kernel<PT, true, 2> : kmc_lanes_kernel<PT, false, 2>);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    // persistent CTAs: as many as stay resident; every warp loops over work items
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    const int64_t want = ((E.lanes_halves ? 2 : 1) * ((E.B + 31) / 32) + warps - 1) / warps;
    if (E.lanes_halves && want > (int64_t)sms * per_sm) return cudaErrorInvalidValue;  // (halves: one warp slot per unit, no queue)
    const unsigned grid = (unsigned)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
    if (plan_only) {
        plan_only->warp_slots = (int64_t)grid * warps;
        plan_only->max_slots = (int64_t)sms * per_sm * warps;
        return cudaSuccess;
    }
    kern<<<grid, warps * 32, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// plan != nullptr: only report the launch geometry (number of persistent warp slots) -- the caller sizes the table
// E.gtab = warp_slots * 2^E.gtab_log sets of 128 bytes and the slicing of the queue's tail from it.
cudaError_t launch_lanes(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches, MemoPlan *plan) {
    if (E.B <= 0) {
        if (plan) { plan->warp_slots = 0; plan->max_slots = 0; }
        return cudaSuccess;
    }
    if (L.N > 31 || L.P > 32 || L.pitchf != 33) return cudaErrorInvalidValue;
    if (!plan && !(E.lanes_flags & 1) && (!E.gtab || E.gtab_log < 6)) return cudaErrorInvalidValue;
    if (L.P == 8) return launch_lanes_t<8>(L, E, st, launches, plan);
    if (L.P == 2) return launch_lanes_t<2>(L, E, st, launches, plan);
    return launch_lanes_t<0>(L, E, st, launches, plan);
}

}  // namespace kmcb200
