// memo_common.cuh -- device helpers shared by the memoised narrow kernels (hop_memo.cu: one warp per trajectory;
// hop_lanes.cu: one thread per trajectory on the hit path, warp-cooperative state evaluation): explicit shared-memory
// accesses, warp scans, the Miller-Abrahams rate in the reference's operation order (goSimulation/simulation.go:58-80)
// and the SWEEP -- every allowed pair of a state exactly once (simulation.go:40-55), per-lane top events + rest.
#pragma once
#include "kmc_device.cuh"
#include "kmc_internal.cuh"

namespace kmcb200 {

#define BIGE 1.0e30f
#define ROWB 264u  // bytes per acceptor-target row of the pair table: 33 float2
#define ELB 132u   // bytes per electrode row of the electrode planes: 33 float
#define ENTB 272u  // bytes per first-level entry: 32 x f64 prefix | f32 1/total | u32 key | f64 total
#define GENTB 288u  // second-level entry: the same + tag {launch id, member + 1} at byte 272

__device__ __forceinline__ float lds_f(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// (volatile at the PTX level: ptxas may not merge it with an earlier identical load -- used where a value is re-read
//  on a cold path precisely so that it need not stay in a register across the hot path)
__device__ __forceinline__ uint4 lds_u4_again(uint32_t a) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_d(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void sts_u(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts_d(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

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

// Broadcast of a small unsigned value from one lane through REDUX.OR: unlike SHFL the result is warp-uniform
// for the compiler, so everything derived from it (the occupation mask, loop trip counts) stays uniform.
__device__ __forceinline__ uint32_t bcast_u(uint32_t v, int src, int lane) {
    return __reduce_or_sync(FULL, lane == src ? v : 0u);
}

// Miller-Abrahams factor in the reference's operation order (simulation.go:66-77): dE = e_to - e_from - kd,
// rate = tc * exp(-dE/kT) for dE > 0, else tc (boltz: kmc_device.cuh).
__device__ __forceinline__ float ma(float tc, float kd, float e_to, float e_from, float kT) {
    const float dE = (e_to - e_from) - kd;
    return tc * boltz(dE, kT);
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

// Site energy of acceptor `lane` for occupation mask o, from scratch: E_const - sum over EMPTY j of I0*R/d_ij
// (simulation.go:226-234).  fp64 sums of fp32 terms in ascending j: exact, hence a pure function of the mask.
__device__ __forceinline__ double energy_of(uint32_t o, uint32_t accm, double E64, uint32_t a_row_me) {
    double e = E64;
    uint32_t mm = ~o & accm;
    while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        e -= (double)lds_f2(a_row_me + j * ROWB).y;
    }
    return e;
}

// Running top-NR of a lane's rates: t[0] >= t[1] >= ... with their partner sites, everything else summed into rest.
// Exact (every rate ends up in exactly one of t[] / rest); ties keep the earlier event in the higher rank.
template <int NR>
__device__ __forceinline__ void rank_insert(float x, int site, float (&t)[NR], int (&p)[NR], float &rest) {
    if (NR == 1) {
        rest += fminf(x, t[0]);
        if (x > t[0]) p[0] = site;
        t[0] = fmaxf(x, t[0]);
    } else {
        rest += fminf(x, t[NR - 1]);  // whatever drops out of the last rank (x itself if it does not make it)
#pragma unroll
        for (int r = NR - 1; r >= 1; --r) {
            const bool in_above = x > t[r - 1];  // x belongs above rank r: rank r inherits rank r-1
            const bool in_here = x > t[r];
            p[r] = in_above ? p[r - 1] : (in_here ? site : p[r]);
            t[r] = in_above ? t[r - 1] : fmaxf(x, t[r]);
        }
        if (x > t[0]) p[0] = site;
        t[0] = fmaxf(x, t[0]);
    }
}

// Sweep: every allowed pair of the current state exactly once.  Per lane (= acceptor): its NR LARGEST rates t[] with
// the partner sites p[], and the sum of all its other rates (rest).  Publishes the fp32 energies to the warp's mirror.
template <int PT, int NR>
__device__ __forceinline__ void sweep_state(uint32_t occ, uint32_t accm, double E64, int lane, int N, int P, float kT,
                                            uint32_t a_row_me, uint32_t a_mir,
                                            uint32_t a_elF, uint32_t a_elR, float &e_me, float (&t)[NR], int (&p)[NR],
                                            float &rest) {
    e_me = (float)energy_of(occ, accm, E64, a_row_me);
    __syncwarp();
    sts_f(a_mir + lane * 4, e_me);
    __syncwarp();
    const bool o = (occ >> lane) & 1u;
    const float src = o ? e_me : -BIGE;         // only occupied acceptors emit to acceptors
    const float sg = o ? 1.0f : -1.0f;          // occupied: i->e, dE = V_e - e_i ; empty: e->i, dE = e_i - V_e
    const uint32_t a_el = (o ? a_elF : a_elR) + lane * 4u;
    rest = 0.0f;
#pragma unroll
    for (int r = 0; r < NR; ++r) { t[r] = 0.0f; p[r] = 0; }
    uint32_t mm = ~occ & accm;
    while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const float ej = lds_f(a_mir + j * 4);
        const float2 v = lds_f2(a_row_me + j * ROWB);
        rank_insert<NR>(ma(v.x, v.y, ej, src, kT), j, t, p, rest);
    }
    if (PT > 0) {
#pragma unroll
        for (int e = 0; e < (PT > 0 ? PT : 1); ++e) {
            const float x = lds_f(a_el + e * ELB) * boltz((lds_f(a_mir + 128 + e * 4) - e_me) * sg, kT);
            rank_insert<NR>(x, N + e, t, p, rest);
        }
    } else {
        for (int e = 0; e < P; ++e) {
            const float x = lds_f(a_el + e * ELB) * boltz((lds_f(a_mir + 128 + e * 4) - e_me) * sg, kT);
            rank_insert<NR>(x, N + e, t, p, rest);
        }
    }
}

// highest set bit (bfind: one FLO, no 31-clz round trip); shift that clamps (shl.b32 gives 0 for amounts > 31)
__device__ __forceinline__ int bfind_u(uint32_t v) {
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
// single-bit mask 1 << n, 0 for n > 31 (BMSK: no constant operand to materialise)
__device__ __forceinline__ uint32_t bit_clamp(uint32_t n) {
    uint32_t r;
    asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(r) : "r"(n));
    return r;
}
__device__ __forceinline__ void sts_u2(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y)); }

}  // namespace kmcb200
