// kmc_device.cuh -- device helpers shared by the warp-per-trajectory hop kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmcb200 {

#define FULL 0xffffffffu

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Boltzmann factor of the Miller-Abrahams rate: exp(-dE/kT) for dE > 0, else 1 (simulation.go:72-76), with the
// reference's roundings: q = float32(-dE/kT) (one correctly rounded fp32 division -- exact, and skipped, for kT = 1), then
// exp(q) = 2^(q * log2e) with the product carried in two floats, so that the argument of ex2.approx is exact to 2^-48 and
// the result is within ex2.approx's 2^-22 of float32(math.Exp(float64(q))) for EVERY dE, not only for the large rates.
// kT is warp-uniform wherever this is called.
__device__ __forceinline__ float boltz(float dE, float kT) {
#ifdef KMCB200_FAST_BOLTZ  // (A/B measurements only: round 1's three-instruction form, 1e-5 on the smallest rates)
    return ex2_approx(fminf(dE * (-1.4426950408889634f / kT), 0.0f));
#endif
    const float d = fmaxf(dE, 0.0f);
    const float q = (kT == 1.0f) ? -d : __fdiv_rn(-d, kT);
    const float hi = q * 1.4426950216293335f;                 // float32(log2 e)
    float lo = fmaf(q, 1.4426950216293335f, -hi);
    lo = fmaf(q, 1.9259629911e-8f, lo);                        // log2 e - float32(log2 e)
    const float t = ex2_approx(hi);
    return fmaf(t, lo * 0.6931471805599453f, t);
}

// Miller-Abrahams rate of one pair: v = {nu*tc, I0*R/d (0 unless acceptor-acceptor)}.
__device__ __forceinline__ float ma_rate(float2 v, float e_to, float e_from, float kT) {
    const float dE = (e_to - e_from) - v.y;
    return v.x * boltz(dE, kT);
}


}  // namespace kmcb200
