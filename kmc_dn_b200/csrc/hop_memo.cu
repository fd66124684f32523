// hop_memo.cu -- production KMC hop loop with STATE MEMOISATION (KMCB200_MODE_FAST, N <= 31 acceptors).
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
// One warp = one trajectory; lane l owns acceptor l.  Physics, arithmetic and event order are those of
// hop_fast.cu (read its header first): every allowed pair is evaluated exactly once per sweep -- lane i
// evaluates i->j for the empty sites j (warp-uniform bit-loop) and, per electrode e, i->e if it is occupied or
// e->i if it is empty; fp32 rates with MUFU.EX2, fp64 incremental energies (exact, drift-free), fp64 prefix
// over the lane sums, two-level pick, Philox4x32-10 variates pooled in shared memory once per 64 hops.
//
// What this kernel adds is the GPU-native counterpart of the reference's state cache.  A trajectory revisits
// a few dozen occupation states over and over (C3: 16 states cover 80-99 % of 1e5 hops), and here -- exactly as
// in simulateRecordPlus, which recomputes the energies from scratch on a miss (simulation.go:378-386) -- the
// cumulative rate structure is a PURE function of the occupation bit-mask, because the fp64 incremental
// energies are exact.  Every warp keeps a direct-mapped cache in shared memory:
//
//     key   = 32-bit occupation mask              (slot = multiplicative hash, 2^LOGK slots: 16, or 64 for small ensembles)
//     value = 31 event SLOTS as a NORMALISED fp64 exclusive prefix (with the event -- partner site and direction, for
//             NR > 1 also the acceptor -- in the 7 / 12 mantissa LSBs), lane 31 = the normalised mass of the slot events
//             (sentinel), the fp32 reciprocal of the total rate and the fp64 total (272 B).
//             Slot s holds the (s / N + 1)-th largest rate of acceptor s % N: for N >= 25 simply every acceptor's
//             largest event (NR = 1); smaller layouts leave lanes free for the 2nd / 3rd largest (NR = 2 / 3) -- on a
//             regular grid all downhill hops of an acceptor have the SAME rate, and one event per acceptor would
//             leave 10 % of the mass to the slow path.
//     A second level of the same entries (+ a (launch, member) tag, so it is never reset) lives in global memory / L2.
//
// Hit:  the sweep and the fp64 scan are skipped; the hop costs ONE ballot of (key matches && prefix < uniform) on the
//       speculatively prefetched line -- an empty ballot means another state's line, the sentinel lane (the 0.2 % of
//       the hops that fall outside the top events take the exact slow path) or a dead state --, one shuffle for the
//       winning lane's event code and a branch-free update of the mask and the electrode tallies: 27 SASS
//       instructions per hop.
// Miss: sweep + scan, then the prefix is parked in the slot.
// Memoising a pure function cannot change a result: with the cache disabled (LOGK = -1 instantiation,
// KMCB200_FLAG_NO_MEMO) the kernel produces bit-identical trajectories (tests/test_gpu_parity.py).
// The reference's cache stores the full per-pair list per state (up to 150e6 floats per trajectory); here the
// rest of the list is recomputed when needed instead of stored, so 272 B per state keep 16 states per warp on chip.
//
// Shared-memory accesses in the hop loop go through explicit ld/st.shared on 32-bit shared addresses: every
// branch condition is then provably warp-uniform for the compiler (votes), and no generic-address arithmetic
// is left in the loop.
#include "memo_common.cuh"

namespace kmcb200 {

template <int LOGK, int PT>
struct MemoGeom {
    static constexpr int K = LOGK >= 0 ? (1 << LOGK) : 0;
    static constexpr int ENTRY = (int)ENTB;
    // mirror (acceptor energies 128 B + electrode energies) | variates 64 x 16 B | entries
    static constexpr int MIRB = PT > 0 ? 128 + ((PT * 4 + 15) & ~15) : 256;
    static constexpr int WARP_BYTES = MIRB + 1024 + (K > 0 ? K : 1) * ENTRY;  // K = 0: one scratch entry
};


// The common case of a hop: one of the cached top events.  Slot istar = the highest one whose interval starts below
// the uniform; branch-free update (simulation.go:107-130): the slot's acceptor flips; an acceptor partner flips too;
// an electrode partner (code >= 32: the clamped bit mask is 0) gains (32+e) or loses (64+e) one hole.
// NR == 1: slot = acceptor.  NR > 1: the acceptor sits in bits 7..11 of the event code.
template <bool DBG, int NR>
__device__ __forceinline__ void apply_top(uint32_t bal, double pre, int lane, int N, uint32_t &occ, int &eoc, int &from, int &to) {
    const int istar = bfind_u(bal);
    const uint32_t c = (uint32_t)__shfl_sync(FULL, __double2loint(pre), istar);
    const uint32_t code = c & 127u;
    const uint32_t site = NR == 1 ? (uint32_t)istar : ((c >> 7) & 31u);
    uint32_t sbit;
    if (NR == 1) sbit = bit_clamp((uint32_t)istar);
    else asm("bmsk.wrap.b32 %0, %1, 1;" : "=r"(sbit) : "r"(c >> 7));  // (wrap: position = low 5 bits)
    occ ^= sbit | bit_clamp(code);
    asm("{ .reg .pred p, q;\n"
        "  setp.eq.u32 p, %1, %2;\n"
        "  setp.eq.u32 q, %1, %3;\n"
        "  @p add.s32 %0, %0, 1;\n"
        "  @q add.s32 %0, %0, -1; }"
        : "+r"(eoc)
        : "r"(code), "r"(lane + 32), "r"(lane + 64));
    if (DBG) {
        if (code < 32u) { from = (int)site; to = (int)code; }
        else if (code < 64u) { from = (int)site; to = N + (int)code - 32; }
        else { from = N + (int)code - 64; to = (int)site; }
    }
}

// Speculative fetch of the cache line of the state just entered (multiplicative hash -> first-level slot).
#define PREFETCH_LINE()                                                                                          \
    do {                                                                                                         \
        const uint32_t slot_ = LOGK > 0 ? ((occ * 0x9E3779B1u) >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u;            \
        uint32_t a_ent_, a_pre_; /* (opaque to the optimiser: one IMAD each, no rematerialised bases) */         \
        asm volatile("mad.lo.u32 %0, %1, 272, %2;" : "=r"(a_ent_) : "r"(slot_), "r"(a_cache_op));                \
        asm volatile("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a_pre_) : "r"(lane), "r"(a_ent_));                       \
        pre = lds_d(a_pre_);                                                                                     \
        tailv = lds_u2(a_ent_ + 256); /* rtot | key */                                                           \
    } while (0)

template <int PT, int LOGK, bool DBG, int NR>
__global__ void __launch_bounds__(256, 4) kmc_memo_kernel(const LayoutDev L, const EnsembleDev E) {
    using G = MemoGeom<LOGK, PT>;
    constexpr int K = G::K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = L.N, S = L.S;
    const int P = PT > 0 ? PT : L.P;
    const int tid = threadIdx.x, lane = tid & 31, nwarps = blockDim.x >> 5;
    // broadcast from lane 0, so that the compiler knows the warp index is warp-uniform (it drives the member loop)
    const int warp = __shfl_sync(FULL, tid >> 5, 0);

    // ---- stage the layout: pair table for acceptor targets, two planes (i->e, e->i) for the electrodes
    {
        float2 *acc = reinterpret_cast<float2 *>(smem_raw);
        float *elF = reinterpret_cast<float *>(smem_raw + (size_t)N * ROWB);
        float *elR = elF + P * 33;
        // (thread-independent trip counts: see the staging loop of hop_lanes.cu)
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
    const uint32_t a_mir = wb, a_rng = wb + G::MIRB, a_cache = wb + G::MIRB + 1024;
    const uint32_t a_row_me = sb + lane * 8u;         // + j*ROWB     : pair (source lane  -> target j)
    const uint32_t a_col_me = sb + lane * ROWB;       // + istar*8    : pair (source istar -> target lane)
    const uint32_t a_elF_e = a_elF + lane * ELB;      // + istar*4    : istar -> electrode lane
    const uint32_t a_elR_e = a_elR + lane * ELB;      // + istar*4    : electrode lane -> istar
    const uint32_t accm = (1u << N) - 1u;             // N <= 31: lane 31 is never an acceptor (it holds the sentinel)
    // the cache base as an OPAQUE register value (round trip through shared memory): otherwise ptxas re-derives it
    // from the warp base inside the hop loop and spends an extra IMAD per hop on a second entry address
    sts_u(a_mir + 124, a_cache);
    __syncwarp();
    const uint32_t a_cache_op = lds_u(a_mir + 124);
    __syncwarp();
    // event slots (NR > 1): slot `lane` serves rank sl_r of acceptor sl_a; acceptor `lane` owns n_slots of its ranks
    const int sl_a = (NR > 1 && lane < 31) ? lane % N : lane;
    const int sl_r = (NR > 1) ? (lane < 31 ? lane / N : NR) : 0;
    int n_slots = 1;
    if (NR > 1) {
        n_slots = 0;
        if (lane < N)
            for (int r = 0; r < NR; ++r) n_slots += (lane + r * N <= 30);
    }
    constexpr int CODEMASK = NR > 1 ? 4095 : 127;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);

    // Second-level cache of this warp slot in global memory (L2-resident): 2^GLOG entries of GENTB bytes, the
    // first level's layout + a tag.  Direct-mapped with an independent hash; looked up on a first-level miss,
    // filled together with the first level.  An entry is valid only with this launch's and this member's tag, so
    // the table is never reset.
    const int GLOG = (K > 0) ? E.gtab_log : 0;
    const int64_t wslot = (int64_t)blockIdx.x * nwarps + warp;
    unsigned char *gtab = (GLOG > 0) ? E.gtab + ((size_t)wslot << GLOG) * GENTB : nullptr;

    // ---- persistent: every warp slot pulls members from a global queue until it is empty (members differ in cost --
    //      the miss rate depends on the voltages -- so a static split would leave slots idle at the end)
    for (;;) {
        unsigned long long mq = 0;
        if (lane == 0) mq = atomicAdd(E.queue, 1ULL);
        const int64_t m = (int64_t)__shfl_sync(FULL, mq, 0);
        if (m >= E.B) break;
    // ---- member parameters
    const float nb = (float)E.kT[m];  // kT, narrowed as the cgo wrappers do (simulationWrapper.go:92)
    const float ve_mine = (lane < P) ? (float)E.electrode_v[m * P + lane] : 0.0f;  // electrode `lane`
    sts_f(a_mir + 128 + lane * 4, ve_mine);
    __syncwarp();
    if (K > 0) {  // empty caches: a key that hashes to another slot can never hit
        for (int e_ = lane; e_ < K; e_ += 32) sts_u(a_cache + e_ * ENTB + 260, e_ == 0 ? 1u : 0u);
    }
    const uint2 tag = make_uint2(E.launch_id, (uint32_t)m + 1u);

    // ---- initial state
    bool o0 = false;
    double E64 = 0.0;  // E_constant of this lane's acceptor, narrowed to float32 (simulationWrapper.go:50-56)
    if (lane < N) {
        if (E.occupation0) o0 = E.occupation0[m * N + lane] != 0;
        if (E.E_constant) E64 = E.E_constant[m * N + lane];
        else {
            E64 = E.basis[(int64_t)P * N + lane];
            for (int p = 0; p < P; ++p) E64 += E.electrode_v[m * P + p] * E.basis[(int64_t)p * N + lane];
        }
        E64 = (double)(float)E64;
    }
    uint32_t occ = __ballot_sync(FULL, o0);

    const uint64_t gm = E.member_index0 + (uint64_t)m;
    const uint2 key = make_uint2((uint32_t)E.seed, (uint32_t)(E.seed >> 32));
    const bool inject = DBG && E.stream_e != nullptr;
    const int64_t total_hops = E.prehops + E.hops, prehops = E.prehops;

    double t_acc = 0.0;
    float t_part = 0.0f;
    int eoc = 0;
    double occtime = 0.0;
    bool dead = false;
    long long n_miss = 0;

    float e_me = 0.0f, rest = 0.0f;
    float tk[NR];
    int pk[NR];

    // Loop-carried cache line of the CURRENT state, fetched speculatively when the previous hop was applied:
    //   pre   NORMALISED exclusive fp64 prefix over the lanes' TOP rates (prefix / total rate of the state); its 7
    //         mantissa LSBs carry the lane's top event: partner acceptor j | 32+e (hole into electrode e) | 64+e
    //         (hole out of electrode e).  +inf for a lane without a positive rate (it can never win the ballot);
    //         lane 31 holds the SENTINEL: the normalised mass of all top events -- a uniform above it (0.2 % of the
    //         hops on C3) falls into the rest of the list.
    //   rtot  1/total (fp32) for the dwell time
    uint2 tailv = make_uint2(0u, ~occ);  // {1/total, key}; first hop: miss
    double pre = 0.0;
    __syncwarp();

    int64_t h = 0;
    while (h < total_hops && !dead) {
        // a chunk never straddles a 64-hop variate block or the prehops boundary
        const int q0 = (int)(h & 63);
        int64_t hend = h - q0 + 64;
        if (hend > total_hops) hend = total_hops;
        if (h < prehops && hend > prehops) hend = prehops;
        const int q1 = q0 + (int)(hend - h);
        if (!inject) {
            // 64 hops' worth of variates: lane l serves hops 2l and 2l+1 of this block
            t_acc += (double)t_part;
            t_part = 0.0f;
            const uint64_t blk = (uint64_t)(h >> 6) * 32u + (uint64_t)lane;
            const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)gm, (uint32_t)(gm >> 32)), key);
            // per hop: a unit exponential (float) for the dwell time and a uniform in (0,1) (double, (x+0.5)/2^32)
            const float e0 = -0.6931471805599453f * lg2_approx(fmaf((float)r.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
            const float e1 = -0.6931471805599453f * lg2_approx(fmaf((float)r.z, 2.3283064365386963e-10f, 1.1641532182693481e-10f));
            const double u0 = ((double)r.y + 0.5) * 2.3283064365386963e-10;
            const double u1 = ((double)r.w + 0.5) * 2.3283064365386963e-10;
            __syncwarp();
            sts_u4(a_rng + lane * 32, make_uint4((uint32_t)__double2loint(u0), (uint32_t)__double2hiint(u0), __float_as_uint(e0), 0u));
            sts_u4(a_rng + lane * 32 + 16, make_uint4((uint32_t)__double2loint(u1), (uint32_t)__double2hiint(u1), __float_as_uint(e1), 0u));
            __syncwarp();
        }
        uint32_t a_rq = a_rng + (uint32_t)q0 * 16u;
        const uint32_t a_rq1 = a_rng + (uint32_t)q1 * 16u;
        for (; a_rq != a_rq1; a_rq += 16u) {
            // ---- random variates: unit exponential for the dwell time (simulation.go:297), uniform for the pick (:164)
            double u, dtd = 0.0;
            float ek = 0.0f;
            if (!inject) {
                const uint4 rv = lds_u4(a_rq);
                ek = __uint_as_float(rv.z);
                u = __hiloint2double((int)rv.y, (int)rv.x);
            } else {
                const int64_t hh = h + (int64_t)((a_rq - a_rng) >> 4) - q0;
                u = (double)E.stream_u[m * total_hops + hh];
            }

            // ---- speculative pick on the prefetched line: the ballot is empty if the line belongs to another state
            //      (all lanes see the same key), if the sentinel fired, or if nothing can happen
            int from = 0, to = 0;
            const uint32_t occ_old = occ;
            uint32_t bal = __ballot_sync(FULL, (K > 0 && !inject && tailv.y == occ) && pre < u);
            if (__builtin_expect((int)bal > 0, 1)) {
                t_part = fmaf(ek, __uint_as_float(tailv.x), t_part);
                if (DBG) dtd = (double)(ek * __uint_as_float(tailv.x));
                apply_top<DBG, NR>(bal, pre, lane, N, occ, eoc, from, to);
            } else {
                // ---- event structure of this state: cached after all, or computed and parked
                // (the variates are re-read here rather than kept alive across the hot path's registers)
                double u2 = u;
                float ek2 = ek;
                if (!inject) {
                    const uint4 rv = lds_u4_again(a_rq);
                    ek2 = __uint_as_float(rv.z);
                    u2 = __hiloint2double((int)rv.y, (int)rv.x);
                }
                bool hit = false;
                if (K > 0) hit = __all_sync(FULL, tailv.y == occ);
                if (!hit) {
                    double total;
                    float rtot;
                    // (the fast path moves the mask through a shuffle; REDUX tells the compiler it is warp-uniform again)
                    const uint32_t occu = __reduce_or_sync(FULL, occ);
                    const uint32_t a_ent = a_cache + (LOGK > 0 ? ((occu * 0x9E3779B1u) >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u) * ENTB;
                    bool hit2 = false;
                    unsigned char *gent = nullptr;
                    if (GLOG > 0) {  // second level (global memory, L2): all loads in flight at once, one latency
                        gent = gtab + (size_t)((occu * 0x85EBCA6Bu) >> ((32 - GLOG) & 31)) * GENTB;
                        const double g_pre = __ldcg(reinterpret_cast<const double *>(gent + lane * 8));
                        const uint4 g1 = __ldcg(reinterpret_cast<const uint4 *>(gent + 256));  // rtot | key | total
                        const uint2 g2 = __ldcg(reinterpret_cast<const uint2 *>(gent + 272));  // tag
                        hit2 = __all_sync(FULL, g1.y == occu && g2.x == tag.x && g2.y == tag.y);
                        if (hit2) {
                            pre = g_pre;
                            rtot = __uint_as_float(g1.x);
                            total = __hiloint2double((int)g1.w, (int)g1.z);
                        }
                    }
                    if (!hit2) {
                        if (DBG) ++n_miss;
                        sweep_state<PT, NR>(occu, accm, E64, lane, N, P, nb, a_row_me, a_mir, a_elF, a_elR, e_me, tk, pk, rest);
                        // event slots: slot s <-> rank s / N of acceptor s % N (s = 0..30; lane 31 is the sentinel)
                        float sv = tk[0];       // this slot's rate
                        int sp = pk[0];         // ... its partner site
                        float rest_tot = rest;  // this ACCEPTOR's mass outside the slots
                        if (NR > 1) {
#pragma unroll
                            for (int r = 1; r < NR; ++r) {
                                const float tv = __shfl_sync(FULL, tk[r], sl_a);
                                const int pv = __shfl_sync(FULL, pk[r], sl_a);
                                if (sl_r == r) { sv = tv; sp = pv; }
                                if (r >= n_slots) rest_tot += tk[r];
                            }
                            if (sl_r >= NR) sv = 0.0f;
                        }
                        const double incl = scan_d((double)sv);
                        const double mtop = __shfl_sync(FULL, incl, 31);
                        double ex = __shfl_up_sync(FULL, incl, 1);  // exact exclusive prefix
                        if (lane == 0) ex = 0.0;
                        double rsum = (double)rest_tot;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) rsum += __shfl_xor_sync(FULL, rsum, d);
                        total = mtop + rsum;
                        if (__all_sync(FULL, !(total > 0.0))) {  // no transition possible (simulation.go:297 would divide by zero)
                            dead = true;
                            break;
                        }
                        const double inv = 1.0 / total;
                        rtot = (float)inv;
                        const double pn = ex * inv;
                        const bool occ_a = (occu >> sl_a) & 1u;
                        uint32_t code = ((uint32_t)sp < (uint32_t)N) ? (uint32_t)sp : ((uint32_t)sp - (uint32_t)N + (occ_a ? 32u : 64u));
                        if (NR > 1) code |= (uint32_t)sl_a << 7;
                        pre = (sv > 0.0f) ? __hiloint2double(__double2hiint(pn), (__double2loint(pn) & ~CODEMASK) | (int)code) : INF;
                        if (lane == 31) pre = mtop * inv;
                        if (GLOG > 0) {
                            __stcg(reinterpret_cast<double *>(gent + lane * 8), pre);
                            if (lane == 0) {
                                __stcg(reinterpret_cast<uint4 *>(gent + 256),
                                       make_uint4(__float_as_uint(rtot), occu, (uint32_t)__double2loint(total), (uint32_t)__double2hiint(total)));
                                __stcg(reinterpret_cast<uint2 *>(gent + 272), tag);
                            }
                        }
                    }
                    // install in the first level (without memoisation: a scratch entry that never hits; the slow path and
                    // the replay read the total from it)
                    __syncwarp();  // (the other lanes' prefetch reads of this slot are ordered before lane 0's writes)
                    sts_d(a_ent + lane * 8, pre);
                    if (lane == 0) {
                        sts_u2(a_ent + 256, make_uint2(__float_as_uint(rtot), K > 0 ? occu : ~occu));
                        sts_d(a_ent + 264, total);
                    }
                    __syncwarp();
                    // (read back through the same loads as the prefetch below: one register assignment for both paths)
                    pre = lds_d(a_ent + lane * 8);
                    tailv = lds_u2(a_ent + 256);
                }

                if (!inject) {
                    t_part = fmaf(ek2, __uint_as_float(tailv.x), t_part);
                    if (DBG) dtd = (double)(ek2 * __uint_as_float(tailv.x));
                } else {
                    const int64_t hh = h + (int64_t)((a_rq - a_rng) >> 4) - q0;
                    const uint32_t a_ent = a_cache + (LOGK > 0 ? ((occ * 0x9E3779B1u) >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u) * ENTB;
                    dtd = E.stream_e[m * total_hops + hh] / lds_d(a_ent + 264);
                    t_acc += dtd;
                }
                bal = __ballot_sync(FULL, pre < u2);
                if ((int)bal > 0) {
                    apply_top<DBG, NR>(bal, pre, lane, N, occ, eoc, from, to);
                } else {
                    // ---- the rest of the list: exact two-level pick over all events EXCEPT the ones that own a slot
                    const uint32_t occu = __reduce_or_sync(FULL, occ);
                    // (rare enough that the sweep is simply repeated, even when this very hop already missed)
                    if (DBG && hit) ++n_miss;
                    sweep_state<PT, NR>(occu, accm, E64, lane, N, P, nb, a_row_me, a_mir, a_elF, a_elR, e_me, tk, pk, rest);
                    float rest_tot = rest;  // this acceptor's mass outside the slots; sk[]: its partners that own a slot
                    int sk[NR];
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        sk[r] = (r < n_slots && tk[r] > 0.0f) ? pk[r] : -1;
                        if (r >= n_slots) rest_tot += tk[r];
                    }
                    const uint32_t a_ent = a_cache + (LOGK > 0 ? ((occu * 0x9E3779B1u) >> (32 - (LOGK > 0 ? LOGK : 1))) : 0u) * ENTB;
                    const double total = lds_d(a_ent + 264);
                    int istar = -1;
                    float rf = BIGE;
                    if (bal >> 31) {
                        const double mtopn = __shfl_sync(FULL, pre, 31);
                        const double rres = (u2 - mtopn) * total;
                        const double incl = scan_d((double)rest_tot);
                        double ex = __shfl_up_sync(FULL, incl, 1);
                        if (lane == 0) ex = 0.0;
                        const uint32_t rpos = __ballot_sync(FULL, rest_tot > 0.0f);
                        uint32_t b2 = __ballot_sync(FULL, ex < rres) & rpos;
                        if (!b2) b2 = rpos & (0u - rpos);
                        if (b2) {
                            istar = 31 - __clz(b2);
                            rf = __shfl_sync(FULL, (float)(rres - ex), istar);
                        }
                    }
                    if (istar < 0) {
                        // no mass outside the slots (rounding), or a uniform of exactly 0 (injected stream): take the
                        // last (first) slot's event instead
                        const uint32_t posu = __ballot_sync(FULL, pre < INF) & 0x7fffffffu;
                        if (!posu) {
                            dead = true;
                            break;
                        }
                        const int slot = (bal >> 31) ? 31 - __clz(posu) : __ffs(posu) - 1;
                        apply_top<DBG, NR>(1u << slot, pre, lane, N, occ, eoc, from, to);
                    } else {
                        uint32_t occ_new = occu;
                        int deo = 0;
                        int skip[NR];  // (offset by one through the unsigned OR-broadcast)
#pragma unroll
                        for (int r = 0; r < NR; ++r) skip[r] = (int)bcast_u((uint32_t)(sk[r] + 1), istar, lane) - 1;
                        const int ptop = (int)bcast_u((uint32_t)pk[0], istar, lane);  // rounding fallback: the acceptor's largest event
                        const bool rowocc = (occu >> istar) & 1u;
                        const float e_star = lds_f(a_mir + istar * 4);
                        bool keepA = true, keepE = true;  // target `lane` / electrode `lane` does not own a slot
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            keepA = keepA && lane != skip[r];
                            keepE = keepE && N + lane != skip[r];
                        }
                        if (rowocc) {
                            from = istar;
                            to = -1;
                            int lastA = -1;
                            float sA = 0.0f;
                            const uint32_t emp = ~occu & accm;
                            if (emp) {  // acceptor targets: istar -> empty `lane`
                                float rr = 0.0f;
                                if (((emp >> lane) & 1u) && keepA) {
                                    const float2 v = lds_f2(a_col_me + istar * 8);
                                    rr = ma(v.x, v.y, e_me, e_star, nb);
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
                                if (lane < P && keepE)
                                    rr = lds_f(a_elF_e + istar * 4) * boltz(ve_mine - e_star, nb);
                                const int e = pick_group<5>(rr, rf - sA);
                                to = (e >= 0) ? N + e : lastA;
                            }
                            if (to < 0) to = ptop;
                        } else {  // empty acceptor: events electrode `lane` -> istar
                            to = istar;
                            float rr = 0.0f;
                            if (lane < P && keepE)
                                rr = lds_f(a_elR_e + istar * 4) * boltz(e_star - ve_mine, nb);
                            from = pick_group<5>(rr, rf);
                            from = (from >= 0) ? from + N : ptop;
                        }
                        if (from < N) occ_new &= ~(1u << from);
                        else deo -= (int)(lane == from - N);
                        if (to < N) occ_new |= (1u << to);
                        else deo += (int)(lane == to - N);
                        eoc += deo;
                        occ = occ_new;
                    }
                }
            }

            // ---- tallies (simulation.go:309-317: pre-hop occupation, antisymmetric traffic)
            if (DBG && h >= prehops) {
                if ((occ_old >> lane) & 1u) occtime += dtd;
                if (lane == 0) {
                    const int64_t hh = h + (int64_t)((a_rq - a_rng) >> 4) - q0;
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
            }

            // ---- the new state (the energies follow from the new mask when next needed)
            if (K > 0) PREFETCH_LINE();
        }
        h = hend;
        if (h == prehops && prehops > 0 && !dead) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
            t_acc = 0.0;
            t_part = 0.0f;
            eoc = 0;
            occtime = 0.0;
        }
    }

    // ---- results
    occ = __reduce_or_sync(FULL, occ);
    t_acc += (double)t_part;
    if (dead) t_acc = INF;  // +inf, as time_step = e/0 would give
    if (lane == 0) E.time[m] = t_acc;
    if (lane < P) E.electrode_occ[m * P + lane] = (int64_t)eoc;
    if (lane < N) {
        if (E.occupation_out) E.occupation_out[m * N + lane] = (occ >> lane) & 1u;
        if (DBG && E.avg_occupation) E.avg_occupation[m * N + lane] = occtime;
        if (E.site_energies_out) E.site_energies_out[m * S + lane] = energy_of(occ, accm, E64, a_row_me);
    }
    if (E.site_energies_out && lane < P) E.site_energies_out[m * S + N + lane] = (double)ve_mine;
    if (DBG && E.misses && lane == 0) E.misses[m] = n_miss;
    __syncwarp();
    }  // members of this warp slot
}

// ---- parity probe: energies + dense rate matrix of one state with the production arithmetic
__global__ void kmc_probe_kernel(const LayoutDev L, const double *E_constant, const double *electrode_v, float kT,
                                 const uint8_t *occ, float *se_io, int se_given, float *rates) {
    const int N = L.N, S = L.S, pitch = L.pitchf;
    const int lane = threadIdx.x;
    if (!se_given) {
        for (int i = lane; i < S; i += 32) {
            double e = (i < N) ? (double)(float)E_constant[i] : (double)(float)electrode_v[i - N];
            if (i < N)
                for (int j = 0; j < N; ++j)
                    if (!occ[j]) e -= (double)L.tblf[j * pitch + i].y;
            se_io[i] = (float)e;
        }
    }
    __syncwarp();
    const float nb = kT;
    for (int idx = lane; idx < S * S; idx += 32) {
        const int i = idx / S, j = idx % S;
        bool ok = (i != j) && !(i >= N && j >= N);
        if (ok && i < N) ok = occ[i] != 0;
        if (ok && j < N) ok = occ[j] == 0;
        float r = 0.0f;
        if (ok) {
            if (i < N && j < N) {
                const float2 v = L.tblf[j * pitch + i];
                r = ma(v.x, v.y, se_io[j], se_io[i], nb);
            } else if (i < N) {  // i -> electrode j
                r = L.tblf[j * pitch + i].x * boltz(se_io[j] - se_io[i], nb);
            } else {             // electrode i -> acceptor j
                r = L.tblf[i * pitch + j].y * boltz(se_io[j] - se_io[i], nb);
            }
        }
        rates[idx] = r;
    }
}

template <int PT, int LOGK>
static cudaError_t launch_memo_t(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches, MemoPlan *plan_only) {
    using G = MemoGeom<LOGK, PT>;
    const bool dbg = E.avg_occupation || E.traffic || E.trace || E.stream_e || E.misses;
    int warps = 8;
    while (warps > 1 && (E.B + warps - 1) / warps < 2 * 148) warps >>= 1;
    const size_t smem = (((size_t)L.N * ROWB + 2 * (size_t)L.P * ELB + 15) & ~size_t(15)) + (size_t)warps * G::WARP_BYTES;
    // ranked events per acceptor: as many as the 31 slots hold (3 for N <= 10, 2 up to N = 24, else 1)
    const int nr = L.N <= 10 ? 3 : (L.N <= 24 ? 2 : 1);
    auto kern = nr == 3 ? (dbg ? kmc_memo_kernel<PT, LOGK, true, 3> : kmc_memo_kernel<PT, LOGK, false, 3>)
              : nr == 2 ? (dbg ? kmc_memo_kernel<PT, LOGK, true, 2> : kmc_memo_kernel<PT, LOGK, false, 2>)
                        : (dbg ? kmc_memo_kernel<PT, LOGK, true, 1> : kmc_memo_kernel<PT, LOGK, false, 1>);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    // persistent CTAs: as many as stay resident; every warp slot loops over its members
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

template <int PT>
static cudaError_t launch_memo_p(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches, MemoPlan *plan) {
    switch (logk) {
        case -1: return launch_memo_t<PT, -1>(L, E, st, launches, plan);
        case 6: return launch_memo_t<PT, 6>(L, E, st, launches, plan);
        default: return launch_memo_t<PT, 4>(L, E, st, launches, plan);
    }
}

// logk: log2(first-level cache slots per warp); -1 disables the memoisation (same code path, every hop a miss).
// plan != nullptr: only report the launch geometry (number of persistent warp slots) -- the caller sizes the
// second-level table E.gtab = warp_slots * 2^E.gtab_log * 288 bytes from it.
cudaError_t launch_memo(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches, MemoPlan *plan) {
    if (E.B <= 0) {
        if (plan) plan->warp_slots = 0;
        return cudaSuccess;
    }
    if (L.N > 31 || L.P > 32 || L.pitchf != 33) return cudaErrorInvalidValue;
    if (L.P == 8) return launch_memo_p<8>(L, E, logk, st, launches, plan);
    if (L.P == 2) return launch_memo_p<2>(L, E, logk, st, launches, plan);
    return launch_memo_p<0>(L, E, logk, st, launches, plan);
}

cudaError_t launch_probe(const LayoutDev &L, const double *E_constant, const double *electrode_v, double kT,
                         const uint8_t *occ, float *se_io, int se_given, float *rates, cudaStream_t st, int *launches) {
    kmc_probe_kernel<<<1, 32, 0, st>>>(L, E_constant, electrode_v, (float)kT, occ, se_io, se_given, rates);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace kmcb200
