// stats_kernels.cu — K5 standalone: reduce per-env-step result arrays into the STG_NSTATS episode-statistics vector
// (include/stg.h STG_STAT_*). The step kernels fuse this reduction into their epilogue; this entry point serves arrays that
// were produced without it (collect_stats off, stored rollouts, solver batches). HBM-bound: <= 30 bytes per element read once;
// grid-stride loop over a grid sized to the machine, per-thread partial sums, warp shuffle + shared memory, one atomic per CTA
// and statistic. stats_fold_kernel sums the replicated buffer of the step kernels (include/stg.h, STG_STAT_REPLICAS).
// Reference analogue: the per-env rolling sums of EnvironmentMonitor (utils/monitoring.py:89-116,180-229).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stg.h"

namespace stg {

__device__ __forceinline__ double stats_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) stats_reduce_kernel(const double* reward, const double* energy, const uint8_t* terminated,
                                                           const uint8_t* truncated, const int32_t* n_sub,
                                                           const int32_t* status, const int32_t* step_count, int64_t n,
                                                           double* stats) {
    double v[STG_NSTATS];
#pragma unroll
    for (int q = 0; q < STG_NSTATS; ++q) v[q] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool term = terminated && terminated[i];
        const bool trunc = truncated && truncated[i];
        v[STG_STAT_STEPS] += 1.0;
        if (n_sub) v[STG_STAT_SUBSTEPS] += (double)n_sub[i];
        if (term) v[STG_STAT_TERMINATED] += 1.0;
        if (!term && trunc) v[STG_STAT_TRUNCATED] += 1.0;
        if (energy) v[STG_STAT_ENERGY] += energy[i];
        if (reward) v[STG_STAT_REWARD] += reward[i];
        if (status && (status[i] & 1)) v[STG_STAT_GUARD] += 1.0;
        if (step_count && (term || trunc)) v[STG_STAT_EPLEN] += (double)step_count[i];
    }
    // warp shuffle, then the 8 warps of the CTA through shared memory: one atomic per statistic and CTA
    __shared__ double part[8][STG_NSTATS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < STG_NSTATS; ++q) {
        const double s = stats_warp_sum(v[q]);
        if (lane == 0) part[warp][q] = s;
    }
    __syncthreads();
    if (threadIdx.x < STG_NSTATS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
        if (s != 0.0) atomicAdd(stats + threadIdx.x, s);
    }
}

// column sums of the replicated buffer: thread (r, q) layout [STG_STAT_REPLICAS][STG_NSTATS], 256 threads, q = tid % 8
__global__ void __launch_bounds__(256) stats_fold_kernel(const double* rep, double* out, int accumulate) {
    __shared__ double part[256];
    const int q = threadIdx.x % STG_NSTATS, r0 = threadIdx.x / STG_NSTATS;      // 32 row groups
    double s = 0.0;
    for (int r = r0; r < STG_STAT_REPLICAS; r += 256 / STG_NSTATS) s += rep[r * STG_NSTATS + q];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < STG_NSTATS) {
        double t = 0.0;
        for (int g = 0; g < 256 / STG_NSTATS; ++g) t += part[g * STG_NSTATS + threadIdx.x];
        out[threadIdx.x] = accumulate ? out[threadIdx.x] + t : t;
    }
}

}  // namespace stg

extern "C" int stg_stats_reduce_f64(const double* d_reward, const double* d_step_energy, const uint8_t* d_terminated,
                                    const uint8_t* d_truncated, const int32_t* d_n_sub, const int32_t* d_status,
                                    const int32_t* d_step_count, int64_t n, double* d_stats, void* stream) {
    if (!d_stats) return STG_E_NULL;
    if (n < 0) return STG_E_SIZE;
    if (n == 0) return STG_OK;
    const int64_t want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);        // 8 CTAs of 256 threads per SM, 148 SMs
    stg::stats_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_reward, d_step_energy, d_terminated, d_truncated, d_n_sub,
                                                                     d_status, d_step_count, n, d_stats);
    return (int)cudaGetLastError();
}

extern "C" int stg_stats_fold_f64(const double* d_replicas, double* d_out, int32_t accumulate, void* stream) {
    if (!d_replicas || !d_out) return STG_E_NULL;
    stg::stats_fold_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_replicas, d_out, accumulate);
    return (int)cudaGetLastError();
}

// Device-side address of a pinned (page-locked, mapped) host allocation: what a kernel must be given to write its outputs
// straight into host memory (under unified addressing it equals the host pointer, but that is the runtime's call to make).
extern "C" int stg_host_device_pointer(void* host_ptr, void** device_ptr) {
    if (!host_ptr || !device_ptr) return STG_E_NULL;
    return (int)cudaHostGetDevicePointer(device_ptr, host_ptr, 0);
}
