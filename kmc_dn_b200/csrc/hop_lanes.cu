// hop_lanes.cu -- production KMC hop loop, ONE THREAD PER TRAJECTORY on the common path (KMCB200_MODE_FAST, N <= 31
// acceptors, large ensembles).
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
// hop_memo.cu (read its header first) runs one trajectory per WARP: a hop that finds its state in the cache still
// costs 27 warp-instructions.  Here a warp carries 32 trajectories in lock-step, one per thread, and a hop whose
// state is in the table is plain per-thread code: Philox, one hash, one 64-byte read of the entry's first chunk
// (header + the 8 most likely events), a branch-free count of thresholds, the mask update -- about 5.5 warp-
// instructions per hop and trajectory.  Only what is NOT in the table is warp-cooperative: the warp stops, evaluates
// the missing state of one of its trajectories with the sweep of hop_memo.cu (lane i = acceptor i), parks the result
// and goes on.  Measured on C3 (1 M members x 1e4 hops, one B200): 6.5e10 hops/s against 2.1e10 of hop_memo.cu;
// 8.5 warp-instructions per hop against 42.8 (profiles/ncu_r01_v8_lanes_kernel_final.txt).
//
// Table.  The cumulative rate structure is a PURE function of (layout, E_constant, electrode energies, kT, occupation
// mask) -- the fp64 energies are exact sums of fp32 terms (hop_memo.cu) -- so trajectories with identical parameters
// (the seeds of one voltage vector / temperature) SHARE one table: the warp detects the runs of consecutive identical
// members among its 32 and gives every run one direct-mapped table in global memory (hot entries live in L2; the
// hardware cache replaces the hand-managed first level of hop_memo.cu).  An entry (512 B) is keyed by the full
// occupation mask and tagged (launch, first member of the run), so the table is zeroed once and never reset:
//
//     0   u32 key | f32 1/total | u32 launch id | u32 first member of the run + 1
//     16 + 64c   8 x u16   codes of events 8c .. 8c+7: event (partner acceptor j | 32+e hole into electrode e | 64+e
//                          hole out of electrode e) | acceptor << 7 | (rate > 0) << 12
//     32 + 64c   8 x u32   their thresholds: inclusive cumulative rate / total in 0.32 fixed point   (c = 0 .. 3)
//     64  f64 total rate | f64 mass of the slot events
//     256 32 x f32  per acceptor: mass of its events outside the slots     384 32 x f32  per acceptor: site energy
//   (the first chunk is read with two 256-bit loads: header + codes, thresholds; the second half of the entry serves
//    the rest-of-list picks, so that they need no second evaluation of the state)
//
// The events are the 31 event slots of hop_memo.cu (every acceptor's largest rates), SORTED by decreasing rate: on C3
// the first chunk answers most hops.  The pick compares the raw 32-bit Philox output x against the thresholds (event k
// iff thr[k-1] <= x < thr[k]) -- the same partition of [0,1) that hop_memo.cu's fp64 compare against (x+0.5)/2^32
// realises, in integers.  x >= thr[31] (the mass of all slot events: 0.4 % of the hops on C3 on average, up to 8 % for
// some voltage vectors) takes the exact two-level pick over the rest of the list (slow_pick below), warp-cooperatively,
// from the entry's second half.
//
// Lock-step and misses.  All 32 trajectories of a warp execute hop h together.  Step 1: every thread probes its
// table (the entry was fetched right after the previous hop); threads that hit resolve their event on their own.
// Step 2: rest-of-list picks of threads that hit -- they READ entries, so they run before this step writes any.
// Step 3: for every thread that missed, the warp evaluates the state, writes the entry and resolves from the registers
// every waiting thread of the run that sits in this very state (an entry may be evicted by the next evaluation of the
// same step; nobody depends on re-reading it).  Step 4: every thread applies its event.
// With the table disabled (lanes_flags & 1) every hop takes step 3 -- same arithmetic, bit-identical results (tested).
//
// RNG: the same Philox4x32-10 numbering as hop_memo.cu (key = seed, counter = (64-hop block * 32 + pair, global member
// index), two hops per call), so streams do not depend on batching, on the number of GPUs or on the kernel's geometry.
#include "memo_common.cuh"

namespace kmcb200 {

#define LENTB 512u  // bytes per table entry
#ifndef LANES_MIN_CTAS
#define LANES_MIN_CTAS 6  // resident CTAs of 4 warps per SM the register budget is set for
#endif

template <int PT>
struct LanesGeom {
    static constexpr int PV = PT > 0 ? PT : 32;   // electrode slots per trajectory
    static constexpr int MIRB = 256;              // mirror: acceptor energies (128 B) | electrode energies (128 B)
    static constexpr int EFB = 32 * 32 * 4;       // E_constant (the narrowed fp32 values) of the 32 trajectories
    static constexpr int VEB = 32 * PV * 4;       // electrode energies of the 32 trajectories
    static constexpr int TALB = PV * 32 * 4;      // electrode tallies [electrode][trajectory]
    static constexpr int WARP_BYTES = MIRB + EFB + VEB + TALB;
};

__device__ __forceinline__ uint4 ldg_u4(const unsigned char *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// header + first chunk (codes, thresholds; 64 B) in ONE statement of two 256-bit loads, issued back to back
__device__ __forceinline__ void ldg_head(const unsigned char *p, uint4 &h, uint4 &c, uint4 &a, uint4 &b) {
    asm volatile(
        "ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n\t"
        "ld.global.v8.u32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%16+32];"
        : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w),
          "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
        : "l"(p)
        : "memory");
}
// codes + thresholds of a later chunk (48 B at p = entry + 64 c + 16)
__device__ __forceinline__ void ldg_chunk(const unsigned char *p, uint4 &c, uint4 &a, uint4 &b) {
    asm volatile(
        "ld.global.v4.u32 {%0, %1, %2, %3}, [%12];\n\t"
        "ld.global.v8.u32 {%4, %5, %6, %7, %8, %9, %10, %11}, [%12+16];"
        : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w), "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
        : "l"(p)
        : "memory");
}
__device__ __forceinline__ uint32_t ldg_u16(const unsigned char *p) {
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.global.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg_u4(unsigned char *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_u32(unsigned char *p, uint32_t v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ float ldg_f32(const unsigned char *p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ldg_f64(const unsigned char *p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg_f32(unsigned char *p, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void stg_f64(unsigned char *p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void stg_u16(unsigned char *p, uint32_t v) {
    asm volatile("{ .reg .u16 t; cvt.u16.u32 t, %1; st.global.u16 [%0], t; }" ::"l"(p), "r"(v) : "memory");
}

// The rest of the list: exact two-level pick over all events EXCEPT the ones that own a slot (x >= mass of the slot
// events; hop_memo.cu's slow path).  Works on the state's per-acceptor data -- rest_tot (mass outside the slots), e_me
// (fp32 site energy), code_l (ANY permutation of the slot codes: event | acceptor << 7 | valid << 12) -- either fresh
// from a sweep or read back from the table entry: the result does not depend on which, nor on the order of the slots.
// Warp-cooperative (lane = acceptor / target / electrode); returns the event code (event | acceptor << 7), warp-uniform.
__device__ __noinline__ uint32_t slow_pick(uint32_t occu, uint32_t accm, float nbt, uint32_t x, double total, double mtop,
                                              float rest_tot, float e_me, float ve_mine, uint32_t code_l, int lane, int N, int P,
                                              uint32_t a_col_me, uint32_t a_elF_e, uint32_t a_elR_e) {
    const double rres = ((double)x + 0.5) * 2.3283064365386963e-10 * total - mtop;
    const double incl = scan_d((double)rest_tot);
    double ex = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) ex = 0.0;
    const uint32_t rpos = __ballot_sync(FULL, rest_tot > 0.0f);
    uint32_t b2 = __ballot_sync(FULL, ex < rres) & rpos;
    if (!b2) b2 = rpos & (0u - rpos);
    const bool valid = (code_l >> 12) & 1u;
    const uint32_t c12 = code_l & 4095u;
    // an event that is certainly allowed, for the cases rounding leaves without a pick
    const uint32_t any_valid = __reduce_max_sync(FULL, valid ? c12 : 0u);
    if (!b2) return any_valid;  // no mass outside the slots (rounding)
    const int istar = 31 - __clz(b2);
    const float rf = __shfl_sync(FULL, (float)(rres - ex), istar);
    // the acceptor's events that own a slot are skipped; the smallest of their codes is the rounding fallback
    const bool match = valid && (int)((code_l >> 7) & 31u) == istar;
    const uint32_t evt = code_l & 127u;
    const uint32_t skipA = __reduce_or_sync(FULL, (match && evt < 32u) ? (1u << evt) : 0u);
    const uint32_t skipE = __reduce_or_sync(FULL, (match && evt >= 32u) ? (1u << (evt & 31u)) : 0u);
    uint32_t fallback = __reduce_min_sync(FULL, match ? c12 : 0xffffu);
    if (fallback == 0xffffu) fallback = any_valid;
    const bool keepA = !((skipA >> lane) & 1u), keepE = !((skipE >> lane) & 1u);
    const bool rowocc = (occu >> istar) & 1u;
    const float e_star = __shfl_sync(FULL, e_me, istar);
    int from, to;
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
                rr = ma(v.x, v.y, e_me, e_star, nbt);
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
            if (lane < P && keepE) rr = lds_f(a_elF_e + istar * 4) * ex2_approx(fminf((ve_mine - e_star) * nbt, 0.0f));
            const int e = pick_group<5>(rr, rf - sA);
            to = (e >= 0) ? N + e : lastA;
        }
        if (to < 0) return fallback;
    } else {  // empty acceptor: events electrode `lane` -> istar
        to = istar;
        float rr = 0.0f;
        if (lane < P && keepE) rr = lds_f(a_elR_e + istar * 4) * ex2_approx(fminf((e_star - ve_mine) * nbt, 0.0f));
        from = pick_group<5>(rr, rf);
        if (from < 0) return fallback;
        from += N;
    }
    if (from < N && to < N) return (uint32_t)to | ((uint32_t)from << 7);
    if (from < N) return (uint32_t)(32 + to - N) | ((uint32_t)from << 7);
    return (uint32_t)(64 + from - N) | ((uint32_t)to << 7);
}

// The entry of the NEXT state is fetched right after a hop is applied: the loads fly while the next hop's variates are
// generated.  (Measured on C3: 6 resident CTAs per SM beat 5, 7 and 8; two 256-bit loads -- served from L2 -- beat four
// 128-bit loads that allocate in L1, 6.3e10 against 5.7e10 hops/s.)
template <int PT, bool DBG, int NR>
__global__ void __launch_bounds__(128, LANES_MIN_CTAS) kmc_lanes_kernel(const LayoutDev L, const EnsembleDev E) {
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
        for (int idx = tid; idx < N * 33; idx += blockDim.x) acc[idx] = L.tblf[idx];
        for (int idx = tid; idx < P * 33; idx += blockDim.x) {
            const float2 v = L.tblf[N * 33 + idx];
            elF[idx] = v.x;
            elR[idx] = v.y;
        }
    }
    __syncthreads();

    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t a_elF = sb + (uint32_t)N * ROWB, a_elR = a_elF + (uint32_t)P * ELB;
    const uint32_t wb = sb + (((uint32_t)N * ROWB + 2u * (uint32_t)P * ELB + 15u) & ~15u) + (uint32_t)warp * G::WARP_BYTES;
    const uint32_t a_mir = wb, a_ef = wb + G::MIRB, a_ve = a_ef + G::EFB, a_tal = a_ve + G::VEB;
    const uint32_t a_row_me = sb + lane * 8u;         // + j*ROWB     : pair (source lane  -> target j)
    const uint32_t a_col_me = sb + lane * ROWB;       // + istar*8    : pair (source istar -> target lane)
    const uint32_t a_elF_e = a_elF + lane * ELB;      // + istar*4    : istar -> electrode lane
    const uint32_t a_elR_e = a_elR + lane * ELB;      // + istar*4    : electrode lane -> istar
    const uint32_t accm = (1u << N) - 1u;             // N <= 31
    // event slots (hop_memo.cu): slot `lane` serves rank sl_r of acceptor sl_a; acceptor `lane` owns n_slots of its ranks
    const int sl_a = (NR > 1 && lane < 31) ? lane % N : lane;
    const int sl_r = (NR > 1) ? (lane < 31 ? lane / N : NR) : 0;
    int n_slots = 1;
    if (NR > 1) {
        n_slots = 0;
        if (lane < N)
            for (int r = 0; r < NR; ++r) n_slots += (lane + r * N <= 30);
    }

    const int tlog = E.gtab_log;  // log2(table entries per warp slot), >= 6
    const int64_t wslot = (int64_t)blockIdx.x * nwarps + warp;
    unsigned char *const wtab = E.gtab + ((size_t)wslot << tlog) * LENTB;
    const bool use_table = !(E.lanes_flags & 1);
    const uint2 key = make_uint2((uint32_t)E.seed, (uint32_t)(E.seed >> 32));
    const int64_t total_hops = E.prehops + E.hops, prehops = E.prehops;
    const uint32_t tagx = E.launch_id;

    // ---- persistent: every warp pulls blocks of 32 members from the global queue
    for (;;) {
        unsigned long long mq = 0;
        if (lane == 0) mq = atomicAdd(E.queue, 32ULL);
        const int64_t base = (int64_t)__shfl_sync(FULL, mq, 0);
        if (base >= E.B) break;
        const int64_t m = base + lane;
        const bool active = m < E.B;
        const int64_t mc = active ? m : E.B - 1;

        // ---- member parameters: thread t = trajectory t
        const float nb = -1.4426950408889634f / (float)E.kT[mc];
        __syncwarp();
        for (int e = 0; e < P; ++e) {
            sts_f(a_ve + (uint32_t)(lane * PV + e) * 4u, active ? (float)E.electrode_v[mc * P + e] : 0.0f);
            sts_u(a_tal + (uint32_t)e * 128u + lane * 4u, 0u);
        }
        uint32_t occ = 0;
        if (E.occupation0 && active)
            for (int i = 0; i < N; ++i) occ |= (uint32_t)(E.occupation0[m * N + i] != 0) << i;
        // E_constant of every trajectory, narrowed to float32 (simulationWrapper.go:50-56): row t, lane = acceptor
        for (int t = 0; t < 32; ++t) {
            const int64_t mt = base + t;
            float ef = 0.0f;
            if (mt < E.B && lane < N) {
                double E64;
                if (E.E_constant) E64 = E.E_constant[mt * N + lane];
                else {
                    E64 = E.basis[(int64_t)P * N + lane];
                    for (int p = 0; p < P; ++p) E64 += E.electrode_v[mt * P + p] * E.basis[(int64_t)p * N + lane];
                }
                ef = (float)E64;
            }
            sts_f(a_ef + (uint32_t)(t * 32 + lane) * 4u, ef);
        }
        __syncwarp();

        // ---- runs of identical members: bit t of sp = member t has the parameters of member t-1
        uint32_t sp = 0;
        {
            const float nbp = __shfl_up_sync(FULL, nb, 1);
            const uint32_t spk = __ballot_sync(FULL, !active || (lane > 0 && __float_as_uint(nbp) == __float_as_uint(nb)));
            for (int t = 1; t < 32; ++t) {
                bool eq = true;
                if (lane < N) eq = lds_u(a_ef + (uint32_t)(t * 32 + lane) * 4u) == lds_u(a_ef + (uint32_t)((t - 1) * 32 + lane) * 4u);
                if (lane < P) eq = eq && lds_u(a_ve + (uint32_t)(t * PV + lane) * 4u) == lds_u(a_ve + (uint32_t)((t - 1) * PV + lane) * 4u);
                if (__all_sync(FULL, eq) || base + t >= E.B) sp |= 1u << t;
            }
            sp &= spk;
        }
        // runs of identical members (any length, any alignment): every run of the warp gets the same power-of-two share
        // of the warp slot's table entries
        const uint32_t lead_mask = ~sp;  // bit l = member l starts a run (bit 0 always does)
        const int leader = 31 - __clz(lead_mask & (0xffffffffu >> (31 - lane)));
        const int nruns = __popc(lead_mask);
        const int slog = tlog - (nruns > 1 ? 32 - __clz(nruns - 1) : 0);  // log2(table entries per run) >= 1
        const int hshift = 32 - slog;
        const uint32_t grp = (uint32_t)leader;
        const uint32_t gofs = ((uint32_t)__popc(lead_mask & ((1u << leader) - 1u)) << slog) * LENTB;  // (< 2^25: 32 bits)
        const uint32_t tagy = (uint32_t)(base + leader) + 1u;
#define LANES_HEAD(p_) ldg_head(p_, hd, tc, ta, tb)
#define LANES_ENT(mask) (wtab + (size_t)(gofs + (((mask) * 0x9E3779B1u) >> hshift) * LENTB))

        const uint64_t gm = E.member_index0 + (uint64_t)mc;
        bool alive = active, dead = false;
        double t_acc = 0.0;
        float t_part = 0.0f;
        long long n_miss = 0;
        uint4 r = make_uint4(0u, 0u, 0u, 0u);
        uint4 hd = make_uint4(0u, 0u, 0u, 0u), ta = hd, tb = hd, tc = hd;  // (launch ids start at 1: never a valid header)
        if (alive && use_table) LANES_HEAD(LANES_ENT(occ));

        // segments: a segment never straddles a 64-hop variate block or the prehops boundary
        bool stop = false;
        for (int64_t h0 = 0; h0 < total_hops && !stop;) {
            if (h0 == prehops && prehops > 0) {  // kmc_dopant_networks.py:580-585: tallies restart, occupation is kept
                t_acc = 0.0;
                t_part = 0.0f;
                for (int e = 0; e < P; ++e) sts_u(a_tal + (uint32_t)e * 128u + lane * 4u, 0u);
            } else {
                t_acc += (double)t_part;
                t_part = 0.0f;
            }
            int64_t hend = (h0 | 63) + 1;
            if (hend > total_hops) hend = total_hops;
            if (h0 < prehops && hend > prehops) hend = prehops;
            const int q0 = (int)(h0 & 63), q1 = q0 + (int)(hend - h0);
            const uint64_t blk0 = (uint64_t)(h0 >> 6) * 32u;
            for (int q = q0; q < q1; ++q) {
            // ---- random variates: unit exponential for the dwell time (simulation.go:297), 32 uniform bits for the pick (:164)
            uint32_t xr, er;
            if (!(q & 1)) {
                const uint64_t blk = blk0 + (uint64_t)(q >> 1);
                r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)gm, (uint32_t)(gm >> 32)), key);
                er = r.x;
                xr = r.y;
            } else {
                er = r.z;
                xr = r.w;
            }
            const float ek = -0.6931471805599453f * lg2_approx(fmaf((float)er, 2.3283064365386963e-10f, 1.1641532182693481e-10f));

            // ---- step 1: probe the table; a thread that finds its state resolves its event on its own
            uint32_t code = 0;
            float rt = 0.0f;
            bool hit = false, slow = false;
            // (thresholds never decrease: the last clause is always true for a valid entry -- it keeps ptxas from
            //  sinking the threshold loads below the branch, which would cost a second round trip)
            hit = alive && use_table && hd.x == occ && hd.z == tagx && hd.w == tagy && tb.w >= ta.x;
            if (hit) {
                rt = __uint_as_float(hd.y);
                int c = 0;
                for (;;) {
                    if (xr < tb.w) {
                        const int k = (int)(xr >= ta.x) + (int)(xr >= ta.y) + (int)(xr >= ta.z) + (int)(xr >= ta.w) +
                                      (int)(xr >= tb.x) + (int)(xr >= tb.y) + (int)(xr >= tb.z);
                        const uint32_t w2 = (k & 4) ? ((k & 2) ? tc.w : tc.z) : ((k & 2) ? tc.y : tc.x);
                        code = (k & 1) ? (w2 >> 16) : (w2 & 0xffffu);
                        break;
                    }
                    if (++c == 4) {
                        slow = true;
                        break;
                    }
                    ldg_chunk(LANES_ENT(occ) + (uint32_t)c * 64u + 16u, tc, ta, tb);
                }
            }

            // ---- step 2: rest-of-list picks of threads that hit (from their entries, before this step writes any)
            uint32_t need = __ballot_sync(FULL, alive && !hit);
            uint32_t slowm = __ballot_sync(FULL, slow);
            bool died = false;
            if (need | slowm) __syncwarp();  // (step 1's reads of the table are ordered before this step's writes)
            while (slowm) {
                const int t = __ffs(slowm) - 1;
                slowm &= slowm - 1;
                const uint32_t occu = __shfl_sync(FULL, occ, t);
                const unsigned char *ent = wtab + (size_t)(__shfl_sync(FULL, gofs, t) + ((occu * 0x9E3779B1u) >> hshift) * LENTB);
                const float rest_tot = ldg_f32(ent + 256 + lane * 4);
                const float e_me = ldg_f32(ent + 384 + lane * 4);
                const uint32_t code_l = ldg_u16(ent + 16 + (lane >> 3) * 64 + (lane & 7) * 2);
                const double total = ldg_f64(ent + 64), mtop = ldg_f64(ent + 72);
                const float ve_mine = (lane < P) ? lds_f(a_ve + (uint32_t)(t * PV + lane) * 4u) : 0.0f;
                const uint32_t rcode = slow_pick(occu, accm, __shfl_sync(FULL, nb, t), __shfl_sync(FULL, xr, t), total, mtop, rest_tot, e_me,
                                                 ve_mine, code_l, lane, N, P, a_col_me, a_elF_e, a_elR_e);
                if (lane == t) code = rcode;
            }

            // ---- step 3: warp-cooperative evaluation of the states that are not in the table
            while (need) {
                const int t = __ffs(need) - 1;
                const uint32_t occu = __shfl_sync(FULL, occ, t);
                const float nbt = __shfl_sync(FULL, nb, t);
                const double E64 = (double)lds_f(a_ef + (uint32_t)(t * 32 + lane) * 4u);
                float ve_mine = 0.0f;  // electrode `lane` of trajectory t
                __syncwarp();
                if (lane < P) {
                    ve_mine = lds_f(a_ve + (uint32_t)(t * PV + lane) * 4u);
                    sts_f(a_mir + 128 + lane * 4, ve_mine);
                }
                if (DBG && lane == t) ++n_miss;
                float e_me, rest;
                float tk[NR];
                int pk[NR];
                sweep_state<PT, NR>(occu, accm, E64, lane, N, P, nbt, a_row_me, a_mir, a_elF, a_elR, e_me, tk, pk, rest);
                // event slots: slot s <-> rank s / N of acceptor s % N (s = 0..30)
                float sv = tk[0];       // this slot's rate
                int spn = pk[0];        // ... its partner site
                float rest_tot = rest;  // this ACCEPTOR's mass outside the slots
                if (NR > 1) {
#pragma unroll
                    for (int rr = 1; rr < NR; ++rr) {
                        const float tv = __shfl_sync(FULL, tk[rr], sl_a);
                        const int pv = __shfl_sync(FULL, pk[rr], sl_a);
                        if (sl_r == rr) { sv = tv; spn = pv; }
                        if (rr >= n_slots) rest_tot += tk[rr];
                    }
                    if (sl_r >= NR) sv = 0.0f;
                }
                if (lane == 31) sv = 0.0f;
                double mtop = (double)sv, total = (double)rest_tot;  // mass of the slot events | of everything
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    mtop += __shfl_xor_sync(FULL, mtop, d);
                    total += __shfl_xor_sync(FULL, total, d);
                }
                total += mtop;
                // threads of this run that wait on this very state (t among them) are all served by this evaluation
                const uint32_t grp_t = __shfl_sync(FULL, grp, t);
                uint32_t same = __ballot_sync(FULL, ((need >> lane) & 1u) && occ == occu && grp == grp_t);
                need &= ~same;
                if (!(total > 0.0)) {  // no transition possible (simulation.go:297 would divide by zero)
                    if ((same >> lane) & 1u) { alive = false; dead = true; }
                    died = true;
                    continue;
                }
                const bool occ_a = (occu >> sl_a) & 1u;
                // slot code: event | acceptor << 7 | (rate > 0) << 12
                const uint32_t mycode = (((uint32_t)spn < (uint32_t)N) ? (uint32_t)spn : ((uint32_t)spn - (uint32_t)N + (occ_a ? 32u : 64u))) |
                                        ((uint32_t)sl_a << 7) | (sv > 0.0f ? 4096u : 0u);
                const double inv = 1.0 / total;
                const float rtot = (float)inv;
                // slots by decreasing rate: bitonic network on (rate bits with the 5 low mantissa bits replaced by
                // 31 - slot) -- unique keys; the order only decides which events share the first chunk, not the result
                uint32_t skey = (__float_as_uint(sv) & ~31u) | (uint32_t)(31 - lane);
#pragma unroll
                for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                        const uint32_t other = __shfl_xor_sync(FULL, skey, jj);
                        const bool keepmax = ((lane & jj) == 0) == ((lane & kk) == 0);
                        skey = keepmax ? max(skey, other) : min(skey, other);
                    }
                }
                const int ssrc = 31 - (int)(skey & 31u);
                const float ssv = __shfl_sync(FULL, sv, ssrc);
                const uint32_t scode = __shfl_sync(FULL, mycode, ssrc);
                const double incl = scan_d((double)ssv);
                const uint32_t thr = __double2uint_rn(incl * inv * 4294967296.0);  // (saturates at 2^32 - 1)
                if (use_table) {
                    const uint32_t tagy_t = __shfl_sync(FULL, tagy, t);
                    unsigned char *ent = wtab + (size_t)(__shfl_sync(FULL, gofs, t) + ((occu * 0x9E3779B1u) >> hshift) * LENTB);
                    unsigned char *ch = ent + (uint32_t)(lane >> 3) * 64u;
                    stg_u32(ch + 32 + (lane & 7) * 4, thr);
                    stg_u16(ch + 16 + (lane & 7) * 2, scode);
                    stg_f32(ent + 256 + lane * 4, rest_tot);
                    stg_f32(ent + 384 + lane * 4, e_me);
                    if (lane == 0) {
                        stg_f64(ent + 64, total);
                        stg_f64(ent + 72, mtop);
                        stg_u4(ent, make_uint4(occu, __float_as_uint(rtot), tagx, tagy_t));
                    }
                    __syncwarp();
                }
                while (same) {
                    const int t2 = __ffs(same) - 1;
                    same &= same - 1;
                    const uint32_t x2 = __shfl_sync(FULL, xr, t2);
                    const uint32_t bb = __ballot_sync(FULL, x2 < thr);
                    uint32_t c2;
                    if (bb) c2 = __shfl_sync(FULL, scode, __ffs(bb) - 1);
                    else c2 = slow_pick(occu, accm, nbt, x2, total, mtop, rest_tot, e_me, ve_mine, mycode, lane, N, P, a_col_me, a_elF_e, a_elR_e);
                    if (lane == t2) {
                        code = c2;
                        rt = rtot;
                    }
                }
            }
            if (died && !__any_sync(FULL, alive)) {
                stop = true;
                break;
            }

            // ---- step 4: apply (simulation.go:107-130, 306-319): the slot's acceptor flips; an acceptor partner flips
            //      too; an electrode partner gains (32+e) or loses (64+e) one hole
            if (alive) {
                const uint32_t site = (code >> 7) & 31u, evt = code & 127u;
                occ ^= (1u << site) | bit_clamp(evt);
                if (evt >= 32u) {
                    const uint32_t a = a_tal + (evt & 31u) * 128u + lane * 4u;
                    sts_u(a, lds_u(a) + (evt < 64u ? 1u : 0xffffffffu));
                }
                t_part = fmaf(ek, rt, t_part);
                if (DBG && E.trace && h0 >= prehops) {
                    int from, to;
                    if (evt < 32u) { from = (int)site; to = (int)evt; }
                    else if (evt < 64u) { from = (int)site; to = N + (int)evt - 32; }
                    else { from = N + (int)evt - 64; to = (int)site; }
                    int32_t *tp = E.trace + (m * E.hops + (h0 + (q - q0) - prehops)) * 2;
                    tp[0] = from;
                    tp[1] = to;
                }
                if (use_table) LANES_HEAD(LANES_ENT(occ));
            }
            }  // hops of the segment
            h0 = hend;
        }

        // ---- results
        t_acc += (double)t_part;
        if (dead) t_acc = __longlong_as_double(0x7ff0000000000000LL);  // +inf, as time_step = e/0 would give
        __syncwarp();
        if (active) {
            E.time[m] = t_acc;
            for (int e = 0; e < P; ++e) E.electrode_occ[m * P + e] = (int64_t)(int32_t)lds_u(a_tal + (uint32_t)e * 128u + lane * 4u);
            if (E.occupation_out)
                for (int i = 0; i < N; ++i) E.occupation_out[m * N + i] = (occ >> i) & 1u;
            if (DBG && E.misses) E.misses[m] = n_miss;
        }
        if (E.site_energies_out) {
            for (int t = 0; t < 32; ++t) {
                const int64_t mt = base + t;
                const uint32_t occu = __shfl_sync(FULL, occ, t);
                if (mt >= E.B) continue;
                if (lane < N)
                    E.site_energies_out[mt * S + lane] = energy_of(occu, accm, (double)lds_f(a_ef + (uint32_t)(t * 32 + lane) * 4u), a_row_me);
                if (lane < P) E.site_energies_out[mt * S + N + lane] = (double)lds_f(a_ve + (uint32_t)(t * PV + lane) * 4u);
            }
        }
        __syncwarp();
    }  // blocks of members
}

template <int PT>
static cudaError_t launch_lanes_t(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches, MemoPlan *plan_only) {
    using G = LanesGeom<PT>;
    const bool dbg = E.trace || E.misses;
    const int warps = 4;
    const size_t smem = (((size_t)L.N * ROWB + 2 * (size_t)L.P * ELB + 15) & ~size_t(15)) + (size_t)warps * G::WARP_BYTES;
    // ranked events per acceptor: as many as the 31 slots hold (3 for N <= 10, 2 up to N = 24, else 1)
    const int nr = L.N <= 10 ? 3 : (L.N <= 24 ? 2 : 1);
    auto kern = nr == 3 ? (dbg ? kmc_lanes_kernel<PT, true, 3> : kmc_lanes_kernel<PT, false, 3>)
              : nr == 2 ? (dbg ? kmc_lanes_kernel<PT, true, 2> : kmc_lanes_kernel<PT, false, 2>)
                        : (dbg ? kmc_lanes_kernel<PT, true, 1> : kmc_lanes_kernel<PT, false, 1>);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    // persistent CTAs: as many as stay resident; every warp loops over blocks of 32 members
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (E.B + warps * 32 - 1) / (warps * 32);
    const unsigned grid = (unsigned)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
    if (plan_only) {
        plan_only->warp_slots = (int64_t)grid * warps;
        return cudaSuccess;
    }
    kern<<<grid, warps * 32, smem, st>>>(L, E);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// plan != nullptr: only report the launch geometry (number of persistent warp slots) -- the caller sizes the table
// E.gtab = warp_slots * 2^E.gtab_log * 512 bytes from it.
cudaError_t launch_lanes(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches, MemoPlan *plan) {
    if (E.B <= 0) {
        if (plan) plan->warp_slots = 0;
        return cudaSuccess;
    }
    if (L.N > 31 || L.P > 32 || L.pitchf != 33) return cudaErrorInvalidValue;
    if (!plan && !(E.lanes_flags & 1) && (!E.gtab || E.gtab_log < 6)) return cudaErrorInvalidValue;
    if (L.P == 8) return launch_lanes_t<8>(L, E, st, launches, plan);
    if (L.P == 2) return launch_lanes_t<2>(L, E, st, launches, plan);
    return launch_lanes_t<0>(L, E, st, launches, plan);
}

}  // namespace kmcb200
