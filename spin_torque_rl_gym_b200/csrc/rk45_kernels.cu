// rk45_kernels.cu — K2 launch + C-ABI entry (see rk45_core.cuh). One trajectory per thread; lanes whose trajectory has finished
// idle until the warp's slowest lane is done, so the caller passes a permutation that sorts the batch by (parameter set, estimated
// attempt count: stg_llgs_rk45_cost_f64) and warps hold trajectories of similar cost. A resident grid whose lanes pick up the next trajectory when theirs ends was
// measured instead (262,144 trajectories): 8 % slower on a uniform batch, 10 % faster on t_end ~ U(0.05, 1) - with ~3.5
// trajectories per resident thread the tail of every CTA costs a quarter of its lifetime - and dropped in favour of the sort.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stg.h"
#include "rk45_core.cuh"
#include "sort_utils.cuh"

namespace stg {

#ifndef STG_RK45_MINBLOCKS
// Occupancy sweep on the configs[2] mix (262,144 trajectories x 114 attempts, B200, end of round 2, solve incl. the sort): min
// blocks 6 (164 registers, no spill) 1.91 ms; 7: 1.82; 8 (128 registers, 16 warps/SM): 1.80; 9 / 10 (96 registers, 208 B stack):
// 1.88 / 1.87; 12 (80 registers): 1.95. (Round 1, before the kernel was trimmed: 1: 3.99, 8: 3.04, 16: 4.02 ms.)
// The kernel waits on dependent FP64 chains (ncu: `wait` 2.7 stalls per issue at 3.6 warps per scheduler), so warps beat registers.
#define STG_RK45_MINBLOCKS 8
#endif
template <bool SEG>      // SEG: piecewise-constant control tables (StgRk45Args.d_seg_*) instead of pulse + constant field
__global__ void __launch_bounds__(64, STG_RK45_MINBLOCKS) llgs_rk45_kernel(const __grid_constant__ StgRk45Args a) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_envs) return;
    // d_perm (optional): thread `slot` integrates trajectory perm[slot]. Inputs, outputs and the Philox id stay indexed by the
    // trajectory, so a permutation that sorts by (parameter set, t_end) makes the lanes of a warp homogeneous without moving data.
    rk45_body<SEG>(a, a.d_perm ? (int64_t)a.d_perm[slot] : slot);
}

template <bool SEG>
__global__ void __launch_bounds__(256) llgs_rk45_cost_kernel(const __grid_constant__ StgRk45Args a, double* cost) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < a.n_envs) cost[e] = rk45_cost_estimate<SEG>(a, e);
}

// Counting sort by (parameter set, estimated cost descending): bin = (set mod 8) * 1024 + (1023 - cost bin), cost bins of 1.5 %
// (48 per octave over 2^0 .. 2^21 attempted steps). The order inside a bin is arbitrary; results are per trajectory.
template <bool SEG>
__device__ __forceinline__ int rk45_sort_bin(const StgRk45Args& a, int64_t e) {
    const double c = rk45_cost_estimate<SEG>(a, e);
    int b = c > 1.0 ? (int)(48.0f * log2f((float)c)) : 0;
    b = b > 1023 ? 1023 : b;
    const int set = a.d_param_index ? (a.d_param_index[e] & 7) : 0;
    return set * 1024 + (1023 - b);
}
template <bool SEG>
__global__ void __launch_bounds__(256) llgs_rk45_sort_hist_kernel(const __grid_constant__ StgRk45Args a, int32_t* hist) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < a.n_envs;
    warp_aggregated_inc(hist, valid ? rk45_sort_bin<SEG>(a, e) : 0, valid);
}
template <bool SEG>
__global__ void __launch_bounds__(256) llgs_rk45_sort_scatter_kernel(const __grid_constant__ StgRk45Args a, int32_t* hist,
                                                                     int32_t* perm) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < a.n_envs;
    const int pos = warp_aggregated_inc(hist, valid ? rk45_sort_bin<SEG>(a, e) : 0, valid);
    if (valid) perm[pos] = (int32_t)e;
}

}  // namespace stg

extern "C" int stg_llgs_rk45_f64(const StgRk45Args* args, void* stream) {
    if (!args) return STG_E_NULL;
    const StgRk45Args& a = *args;
    if (a.n_envs < 0 || a.n_sets <= 0) return STG_E_SIZE;
    if (!a.d_table || !a.d_m0 || !a.d_t_end || !a.d_y_out) return STG_E_NULL;
    if (!(a.rtol > 0.0) || !(a.atol > 0.0) || !(a.max_step > 0.0)) return STG_E_SIZE;
    if ((a.flags & STG_F_THERMAL_PHILOX) && (a.flags & STG_F_THERMAL_INJECT)) return STG_E_ENUM;
    if ((a.flags & STG_F_THERMAL_INJECT) && (!a.d_noise || a.noise_stride <= 0)) return STG_E_NULL;
    if (a.d_traj && a.traj_stride <= 0) return STG_E_SIZE;
    if (a.n_seg < 0 || (a.n_seg > 0 && (!a.d_seg_t || !a.d_seg_current || (a.seg_rows != 1 && a.seg_rows != a.n_envs))))
        return a.n_seg < 0 ? STG_E_SIZE : (!a.d_seg_t || !a.d_seg_current ? STG_E_NULL : STG_E_SIZE);
    if (a.n_envs == 0) return STG_OK;
    const int64_t ctas = (a.n_envs + 63) / 64;
    if (a.n_seg > 0) stg::llgs_rk45_kernel<true><<<(unsigned)ctas, 64, 0, (cudaStream_t)stream>>>(a);
    else stg::llgs_rk45_kernel<false><<<(unsigned)ctas, 64, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int stg_llgs_rk45_cost_f64(const StgRk45Args* args, double* d_cost, void* stream) {
    if (!args || !d_cost) return STG_E_NULL;
    const StgRk45Args& a = *args;
    if (a.n_envs < 0 || a.n_sets <= 0) return STG_E_SIZE;
    if (!a.d_table || !a.d_m0 || !a.d_t_end || !(a.max_step > 0.0)) return a.max_step > 0.0 ? STG_E_NULL : STG_E_SIZE;
    if (a.n_seg > 0 && (!a.d_seg_t || !a.d_seg_current)) return STG_E_NULL;
    if (a.n_envs == 0) return STG_OK;
    const unsigned grid = (unsigned)((a.n_envs + 255) / 256);
    if (a.n_seg > 0) stg::llgs_rk45_cost_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a, d_cost);
    else stg::llgs_rk45_cost_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a, d_cost);
    return (int)cudaGetLastError();
}

extern "C" int stg_llgs_rk45_sort_f64(const StgRk45Args* args, int32_t* d_perm, int32_t* d_work, void* stream) {
    if (!args || !d_perm || !d_work) return STG_E_NULL;
    const StgRk45Args& a = *args;
    if (a.n_envs < 0 || a.n_sets <= 0 || a.n_envs > 2147483647LL) return STG_E_SIZE;
    if (!a.d_table || !a.d_m0 || !a.d_t_end || !(a.max_step > 0.0)) return a.max_step > 0.0 ? STG_E_NULL : STG_E_SIZE;
    if (a.n_seg > 0 && (!a.d_seg_t || !a.d_seg_current)) return STG_E_NULL;
    if (a.n_envs == 0) return STG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaMemsetAsync(d_work, 0, sizeof(int32_t) * STG_SORT_WORK_INTS, s);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)((a.n_envs + 255) / 256);
    if (a.n_seg > 0) {
        stg::llgs_rk45_sort_hist_kernel<true><<<grid, 256, 0, s>>>(a, d_work);
        stg::sort_scan_kernel<<<1, 1024, 0, s>>>(d_work);
        stg::llgs_rk45_sort_scatter_kernel<true><<<grid, 256, 0, s>>>(a, d_work, d_perm);
    } else {
        stg::llgs_rk45_sort_hist_kernel<false><<<grid, 256, 0, s>>>(a, d_work);
        stg::sort_scan_kernel<<<1, 1024, 0, s>>>(d_work);
        stg::llgs_rk45_sort_scatter_kernel<false><<<grid, 256, 0, s>>>(a, d_work, d_perm);
    }
    return (int)cudaGetLastError();
}
