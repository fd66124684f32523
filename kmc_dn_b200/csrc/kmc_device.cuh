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

// Miller-Abrahams rate of one pair: v = {nu*tc, I0*R/d (0 unless acceptor-acceptor)}.
// dE>0 -> exp(-dE/kT), else 1   ==   exp2(min(-dE*log2e/kT, 0)).
__device__ __forceinline__ float ma_rate(float2 v, float e_to, float e_from, float negbeta) {
    const float dE = (e_to - e_from) - v.y;
    return v.x * ex2_approx(fminf(dE * negbeta, 0.0f));
}


}  // namespace kmcb200
