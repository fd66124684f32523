// reduce.cu -- per-voltage-vector statistics of the electrode currents on the device ("optional
// ncclAllReduce of (sum x, sum x^2, n)").  The consumers of the hop loop (voltage_search.evaluate_error_corr_parallel,
// voltage_search.py:160-185; thesis_indrek/validate_tests.py:80-135) use the MEAN and SPREAD of the current
// electrode_occupation / time (kmc_dopant_networks.py:618) over the repeats of one voltage vector; with the seeds of a
// voltage vector stored as `group` consecutive members, this kernel turns [B] times and [B,P] tallies into
// [B/group, P] sums, sums of squares and [B/group] counts -- 8 MB instead of 75 MB for the C3 ensemble.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmc_internal.cuh"

namespace kmcb200 {

__global__ void reduce_currents_kernel(const double *__restrict__ time, const int64_t *__restrict__ eo, int64_t n_groups, int P, int group,
                                       double *__restrict__ sum, double *__restrict__ sumsq, double *__restrict__ count) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_groups * P) return;
    const int64_t g = idx / P;
    const int e = (int)(idx % P);
    double s = 0.0, q = 0.0, n = 0.0;
    for (int k = 0; k < group; ++k) {
        const int64_t m = g * group + k;
        const double t = time[m];
        if (t > 0.0 && t < __longlong_as_double(0x7ff0000000000000LL)) {  // (a dead trajectory has time = +inf)
            const double c = (double)eo[m * P + e] / t;
            s += c;
            q += c * c;
            n += 1.0;
        }
    }
    sum[idx] = s;
    sumsq[idx] = q;
    if (e == 0 && count) count[g] = n;
}

cudaError_t launch_reduce_currents(const double *time, const int64_t *eo, int64_t B, int P, int group, double *sum, double *sumsq,
                                   double *count, cudaStream_t st, int *launches) {
    if (B <= 0 || P <= 0) return cudaSuccess;
    const int64_t n_groups = B / group, total = n_groups * P;
    if (total <= 0) return cudaSuccess;
    reduce_currents_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(time, eo, n_groups, P, group, sum, sumsq, count);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace kmcb200
