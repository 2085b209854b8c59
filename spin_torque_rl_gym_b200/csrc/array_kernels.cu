// array_kernels.cu — K3 launch + C-ABI: one CTA per crossbar array. The pattern and the target live in shared memory for
// the whole step (coalesced FP64 loads, coalesced f32 observation stores); the sequential Gauss-Seidel device update is
// executed by one thread (its order is part of the reference's semantics), the reductions by NumPy-ordered sums.
// HBM-bound: ~ (24+24) B read + (24+24) B written per device and step.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stg.h"
#include "array_core.cuh"

namespace stg {

constexpr int kArrayBlock = 32;     // one warp per crossbar array; 24 CTAs (= arrays) resident per SM hide the FP64 latency
constexpr int kArrayMinBlocks = 24;

__device__ __forceinline__ void array_draw_pattern(const StgArrayStepArgs& a, int64_t arr, uint32_t episode, int nd,
                                                   double* pattern) {
    Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    const uint64_t gid = a.array_offset + (uint64_t)arr;
    for (int i = threadIdx.x; i < nd; i += blockDim.x) {
        uint32_t o[4];
        float z0, z1, z2, z3;
        uint32_t attempt = 0;
        do {
            ph((uint32_t)gid, (uint32_t)(gid >> 32), episode, ((uint32_t)i << 4) + attempt, o);
            box_muller(o[0], o[1], z0, z1);
            box_muller(o[2], o[3], z2, z3);
            ++attempt;
        } while (z0 * z0 + z1 * z1 + z2 * z2 < 1e-12f && attempt < 8);
        const double n = sqrt((double)z0 * z0 + (double)z1 * z1 + (double)z2 * z2);
        pattern[3 * i] = z0 / n; pattern[3 * i + 1] = z1 / n; pattern[3 * i + 2] = z2 / n;
    }
}

__device__ __forceinline__ void array_store_obs(float* obs, const double* pattern, const double* target, int nd) {
    // [D][6] f32: consecutive threads write consecutive floats
    for (int q = threadIdx.x; q < nd * 6; q += blockDim.x) {
        const int d = q / 6, c = q % 6;
        obs[q] = (float)(c < 3 ? pattern[3 * d + c] : target[3 * d + c - 3]);
    }
}

__global__ void __launch_bounds__(kArrayBlock, kArrayMinBlocks) array_step_kernel(const __grid_constant__ StgArrayStepArgs a) {
    extern __shared__ double smem[];
    const int nd = a.params.n_rows * a.params.n_cols;
    double* pattern = smem;                 // [nd][3]
    double* target = smem + 3 * nd;         // [nd][3]
    double* scratch = smem + 6 * nd;        // [nd]
    __shared__ double s_out[6];
    __shared__ int s_flags[3];
    const int64_t arr = blockIdx.x;
    double* gp = a.d_pattern + arr * nd * 3;
    const double* gt = a.d_target + arr * nd * 3;
    for (int q = threadIdx.x; q < nd * 3; q += blockDim.x) { pattern[q] = gp[q]; target[q] = gt[q]; }
    __syncthreads();
    const StgArrayParams& p = a.params;
    __shared__ double s_prev, s_energy, s_sim, s_mean;
    __shared__ ArrayAction s_act;
    // per-device dot products / norms in parallel, NumPy-ordered sums by lane 0 (bit-identical to the sequential form)
    for (int i = threadIdx.x; i < nd; i += blockDim.x) scratch[i] = dot_u(pattern + 3 * i, target + 3 * i);
    __syncthreads();
    if (threadIdx.x == 0) {
        s_prev = ddiv(numpy_sum(scratch, nd), (double)nd);
        s_act = array_parse_action(p, a.d_action + arr * a.action_stride);
        s_energy = array_apply_action(p, a.d_coupling, pattern, s_act);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nd; i += blockDim.x) scratch[i] = dot_u(pattern + 3 * i, target + 3 * i);
    __syncthreads();
    if (threadIdx.x == 0) s_sim = ddiv(numpy_sum(scratch, nd), (double)nd);
    __syncthreads();
    for (int i = threadIdx.x; i < nd; i += blockDim.x) scratch[i] = norm_u(pattern + 3 * i);
    __syncthreads();
    if (threadIdx.x == 0) s_mean = ddiv(numpy_sum(scratch, nd), (double)nd);
    __syncthreads();
    for (int i = threadIdx.x; i < nd; i += blockDim.x) { const double d = dadd(scratch[i], -s_mean); scratch[i] = dmul(d, d); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double prev = s_prev, energy = s_energy, sim = s_sim;
        const ArrayAction act = s_act;
        const bool success = sim >= p.success_threshold;
        const double sd = sqrt(ddiv(numpy_sum(scratch, nd), (double)nd));
        const double reward = array_reward(p, success, sim, energy, dadd(sim, -prev), sd);
        const int step = a.d_step_count[arr] + 1;
        const double tot = dadd(a.d_total_energy[arr], energy);
        const bool trunc = step >= p.max_steps;
        a.d_reward[arr] = reward;
        a.d_terminated[arr] = success ? 1 : 0;
        a.d_truncated[arr] = trunc ? 1 : 0;
        if (a.d_step_energy) a.d_step_energy[arr] = energy;
        if (a.d_similarity) a.d_similarity[arr] = sim;
        const bool reset = (a.flags & STG_F_AUTORESET) && (success || trunc);
        a.d_step_count[arr] = reset ? 0 : step;
        a.d_total_energy[arr] = reset ? 0.0 : tot;
        s_flags[0] = reset ? 1 : 0;
        if (a.d_stats) {
            atomicAdd(a.d_stats + STG_STAT_STEPS, 1.0);
            atomicAdd(a.d_stats + STG_STAT_SUBSTEPS, 10.0 * act.count);
            if (success) atomicAdd(a.d_stats + STG_STAT_TERMINATED, 1.0);
            else if (trunc) atomicAdd(a.d_stats + STG_STAT_TRUNCATED, 1.0);
            atomicAdd(a.d_stats + STG_STAT_ENERGY, energy);
            atomicAdd(a.d_stats + STG_STAT_REWARD, reward);
            if (success || trunc) atomicAdd(a.d_stats + STG_STAT_EPLEN, (double)step);
        }
    }
    __syncthreads();
    const bool reset = s_flags[0] != 0;
    if (reset) {
        if (a.d_final_obs) array_store_obs(a.d_final_obs + arr * nd * 6, pattern, target, nd);
        __syncthreads();
        const uint32_t ep = (uint32_t)a.d_episode[arr] + 1u;
        __syncthreads();
        if (threadIdx.x == 0) a.d_episode[arr] = (int32_t)ep;
        array_draw_pattern(a, arr, ep, nd, pattern);
        __syncthreads();
    } else if ((a.flags & STG_F_AUTORESET) && a.d_final_obs) {
        for (int q = threadIdx.x; q < nd * 6; q += blockDim.x) a.d_final_obs[arr * nd * 6 + q] = 0.0f;
    }
    for (int q = threadIdx.x; q < nd * 3; q += blockDim.x) gp[q] = pattern[q];
    array_store_obs(a.d_obs + arr * nd * 6, pattern, target, nd);
}

__global__ void __launch_bounds__(kArrayBlock) array_reset_kernel(const __grid_constant__ StgArrayStepArgs a,
                                                                   const uint8_t* mask, const double* pattern0) {
    extern __shared__ double smem[];
    const int nd = a.params.n_rows * a.params.n_cols;
    double* pattern = smem;
    double* target = smem + 3 * nd;
    const int64_t arr = blockIdx.x;
    if (mask && !mask[arr]) return;
    double* gp = a.d_pattern + arr * nd * 3;
    const double* gt = a.d_target + arr * nd * 3;
    const uint32_t ep = (uint32_t)a.d_episode[arr] + 1u;
    if (pattern0) {
        for (int q = threadIdx.x; q < nd * 3; q += blockDim.x) pattern[q] = pattern0[arr * nd * 3 + q];
    } else {
        array_draw_pattern(a, arr, ep, nd, pattern);
    }
    for (int q = threadIdx.x; q < nd * 3; q += blockDim.x) target[q] = gt[q];
    __syncthreads();
    for (int q = threadIdx.x; q < nd * 3; q += blockDim.x) gp[q] = pattern[q];
    if (a.d_obs) array_store_obs(a.d_obs + arr * nd * 6, pattern, target, nd);
    if (threadIdx.x == 0) {
        a.d_episode[arr] = (int32_t)ep;
        a.d_step_count[arr] = 0;
        a.d_total_energy[arr] = 0.0;
    }
}

}  // namespace stg

static int check_array_args(const StgArrayStepArgs& a) {
    const int nd = a.params.n_rows * a.params.n_cols;
    if (a.n_arrays < 0 || a.params.n_rows <= 0 || a.params.n_cols <= 0 || nd > STG_ARRAY_MAX_DEVICES) return STG_E_SIZE;
    if (a.params.action_mode < STG_ARRAY_INDIVIDUAL || a.params.action_mode > STG_ARRAY_GLOBAL) return STG_E_ENUM;
    if (a.params.device_kind < STG_DEV_STT || a.params.device_kind > STG_DEV_VCMA) return STG_E_ENUM;
    if (!a.d_pattern || !a.d_target || !a.d_total_energy || !a.d_step_count || !a.d_episode) return STG_E_NULL;
    return STG_OK;
}

extern "C" int stg_array_step_f64(const StgArrayStepArgs* args, void* stream) {
    if (!args) return STG_E_NULL;
    const StgArrayStepArgs& a = *args;
    int rc = check_array_args(a);
    if (rc != STG_OK) return rc;
    if (!a.d_action || !a.d_obs || !a.d_reward || !a.d_terminated || !a.d_truncated) return STG_E_NULL;
    if (a.action_stride < (a.params.action_mode == STG_ARRAY_GLOBAL ? 2 : 3)) return STG_E_SIZE;
    if (a.n_arrays == 0) return STG_OK;
    const int nd = a.params.n_rows * a.params.n_cols;
    const size_t smem = sizeof(double) * 7 * (size_t)nd;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(stg::array_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    stg::array_step_kernel<<<(unsigned)a.n_arrays, stg::kArrayBlock, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int stg_array_reset(const StgArrayStepArgs* args, const uint8_t* d_mask, const double* d_pattern0, void* stream) {
    if (!args) return STG_E_NULL;
    const StgArrayStepArgs& a = *args;
    int rc = check_array_args(a);
    if (rc != STG_OK) return rc;
    if (a.n_arrays == 0) return STG_OK;
    const int nd = a.params.n_rows * a.params.n_cols;
    const size_t smem = sizeof(double) * 6 * (size_t)nd;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(stg::array_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    stg::array_reset_kernel<<<(unsigned)a.n_arrays, stg::kArrayBlock, smem, (cudaStream_t)stream>>>(a, d_mask, d_pattern0);
    return (int)cudaGetLastError();
}
