// peaks.cu -- micro-kernels that measure, on the device the bench runs on, the pipe peaks the hop loop is
// bounded by (SURVEY.md section 8d asks for measured denominators: MEASURED_PEAKS.json only has HBM and bf16).
//   what = 0: MUFU.EX2 throughput (ex2.approx per second, all SMs)
//   what = 1: FP32 FFMA throughput (fma per second)
//   what = 2: warp-instruction issue rate (FFMA + LOP3 mix on both FP/INT pipes; warp-instructions per second)
//   what = 3: scattered table lookups: every thread reads 32-byte sectors at independent pseudo-random places of an
//             L2-resident 32 MiB table with 256-bit loads, 8 in flight (sectors per second) -- the access pattern of the
//             thread-per-trajectory kernel's hit path (hop_lanes.cu: one 32-byte entry sector per hop and thread), whose
//             hardware limit is the L1TEX data pipe: one wavefront (here = one thread's sector) per clock and SM
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmcb200 {

template <int WHAT>
__global__ void __launch_bounds__(256) peak_kernel(float *sink, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
    float a4 = a0 + 0.4f, a5 = a0 + 0.5f, a6 = a0 + 0.6f, a7 = a0 + 0.7f;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    for (int it = 0; it < iters; ++it) {
        if (WHAT == 0) {
#define EX2(x) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x))
            EX2(a0); EX2(a1); EX2(a2); EX2(a3); EX2(a4); EX2(a5); EX2(a6); EX2(a7);
#undef EX2
        } else if (WHAT == 1) {
#define FMA(x) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x))
            FMA(a0); FMA(a1); FMA(a2); FMA(a3); FMA(a4); FMA(a5); FMA(a6); FMA(a7);
#undef FMA
        } else {
            // 4 FFMA (fma pipe) + 4 XOR in a rotating dependency ring (alu pipe): nothing ptxas can fold
#define FMA(x) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x))
#define XR(x, y) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y))
            FMA(a0); XR(i0, i1); FMA(a1); XR(i1, i2); FMA(a2); XR(i2, i3); FMA(a3); XR(i3, i0);
#undef FMA
#undef XR
        }
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7);
    if (s == 1.2345e-30f) sink[0] = s;
}

__global__ void __launch_bounds__(256) lookup_peak_kernel(const unsigned char *table, uint32_t mask, uint32_t *sink, int iters) {
    uint32_t x = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t a[8], v[8][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x = x * 1664525u + 1013904223u;
            a[k] = (x >> 5) & mask;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[k][0]), "=r"(v[k][1]), "=r"(v[k][2]), "=r"(v[k][3]), "=r"(v[k][4]), "=r"(v[k][5]), "=r"(v[k][6]), "=r"(v[k][7])
                         : "l"(table + (size_t)a[k] * 32u));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc ^= v[k][0] ^ v[k][7];
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

static double measure_lookup_peak(int sms, int *launches) {
    const size_t bytes = (size_t)32 << 20;
    unsigned char *table = nullptr;
    uint32_t *sink = nullptr;
    if (cudaMalloc(&table, bytes) != cudaSuccess) return -1.0;
    if (cudaMalloc(&sink, 4) != cudaSuccess) { cudaFree(table); return -1.0; }
    cudaMemset(table, 1, bytes);
    const int blocks = sms * 8, threads = 256, iters = 1 << 10;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        lookup_peak_kernel<<<blocks, threads>>>(table, (uint32_t)(bytes / 32 - 1), sink, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        if (launches) ++*launches;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)blocks * threads * iters * 8.0 / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(table); cudaFree(sink);
    return best;
}

// returns operations per second (thread-level ops for 0/1/3, warp-instructions for 2); <0 on error
double measure_peak(int what, int *launches) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (what == 3) return measure_lookup_peak(sms, launches);
    float *sink = nullptr;
    if (cudaMalloc(&sink, 4) != cudaSuccess) return -1.0;
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (what == 0) peak_kernel<0><<<blocks, threads>>>(sink, iters);
        else if (what == 1) peak_kernel<1><<<blocks, threads>>>(sink, iters);
        else peak_kernel<2><<<blocks, threads>>>(sink, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        if (launches) ++*launches;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * threads * iters * 8.0;
        if (what == 2) ops /= 32.0;
        const double rate = ops / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

}  // namespace kmcb200
