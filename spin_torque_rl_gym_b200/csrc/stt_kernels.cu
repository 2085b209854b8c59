// stt_kernels.cu — K1: fused SpinTorque-v0 env step / reset / solver kernels + their C-ABI entry points (sm_100a).
//
// Thread mapping: one env per thread. Per env-step a thread (stt_env_core.cuh: env_step_body)
//   1. loads its state (FP64 SoA planes, coalesced) and its action,
//   2. sanitises the action (SafetyWrapper.validate_action + _parse_action), derives the substep plan in FP64,
//   3. integrates n_sub fixed RK4/Euler substeps entirely in registers (stage arithmetic in R = float/double, state FP64),
//   4. evaluates Joule energy, alignment, reward, flags, observation in FP64, optionally resets finished episodes (Philox),
// then the block
//   5. stages the 12-float observation rows through shared memory so [n][12] f32 is written as coalesced float4s,
//   6. accumulates episode statistics (warp shuffle + one atomic per warp and statistic).
// Reference lines are cited next to each block; the CPU restatement lives in oracle/stt_oracle.py.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/stg.h"
#include "llgs_core.cuh"
#include "stt_env_core.cuh"
#include "sort_utils.cuh"

namespace stg {

#ifndef STG_BLOCK
#define STG_BLOCK 64
#endif
#ifndef STG_MINBLOCKS_DET
#define STG_MINBLOCKS_DET 12    // no-noise variants: 80 registers, 24 warps/SM measured best (profiles/)
#endif
#ifndef STG_MINBLOCKS_NOISE
#define STG_MINBLOCKS_NOISE 1   // FP64-stage thermal variants: let ptxas keep the 12 Gaussians + generator state in registers
#endif
#ifndef STG_MINBLOCKS_NOISE_F32
// FP32 thermal variants (one env per thread): 106 registers at 8 CTAs (16 warps) per SM, no spills; 10 / 12 CTAs per SM (93 / 80
// registers, the latter spilling) measured 9.50 / 9.35 ms against 9.31 per 1M-env x 999-substep step.
#define STG_MINBLOCKS_NOISE_F32 8
#endif
constexpr int kBlock = STG_BLOCK;   // 64: 65,536 envs -> 1024 CTAs = 6.9 per SM, balanced to 1.2 % on 148 SMs

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

using StepArgs = StgSttStepArgs;
using ResetArgs = StgSttResetArgs;
using SolveArgs = StgSttSolveArgs;

// An env the FP32 stages declined (env_step_body returned false, nothing written): append it to the list of the second pass.
// This launch writes NOTHING for it - no observation row, no flags, no statistics; the second pass writes all of them.
__device__ __forceinline__ void redo_push(const StgSttStepArgs& a, int64_t e) {
    const int pos = atomicAdd(a.d_redo, 1);
    a.d_redo[STG_REDO_HEADER + pos] = (int32_t)e;
}

__device__ __forceinline__ void store_row(float* dst, const float* o) {
    float4* d = reinterpret_cast<float4*>(dst);
    d[0] = make_float4(o[0], o[1], o[2], o[3]);
    d[1] = make_float4(o[4], o[5], o[6], o[7]);
    d[2] = make_float4(o[8], o[9], o[10], o[11]);
}

template <typename R, bool AXIS_Z, int NOISE, bool EULER>
__global__ void __launch_bounds__(kBlock, NOISE == 0 ? STG_MINBLOCKS_DET : (sizeof(R) == 4 ? STG_MINBLOCKS_NOISE_F32 : STG_MINBLOCKS_NOISE))
stt_env_step_kernel(const __grid_constant__ StepArgs a) {
    __shared__ __align__(16) float s_obs[kBlock * kObs];
    __shared__ uint8_t s_skip[kBlock];
    const int64_t slot = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool active = slot < a.n_envs;
    const bool sorted = (a.flags & STG_F_SORTED) != 0;
    const bool want_fin = (a.flags & STG_F_AUTORESET) != 0 && a.out.final_obs != nullptr;
    const int64_t e = active ? (sorted ? (int64_t)a.d_perm[slot] : slot) : 0;

    EnvStepResult r;
    r.did_reset = false;
    bool stepped = active;
    if (active) {
        stepped = env_step_body<R, AXIS_Z, NOISE, EULER>(a, e, r);
        if (!stepped) redo_push(a, e);      // FP32 stages only: repeated with FP64 stages by stt_env_redo_kernel
    }

    // ---- observation rows ------------------------------------------------------------------------------------------------
    if (!sorted) {
        // consecutive slots are consecutive envs: stage rows in shared memory, write the block's rows as coalesced float4s
        // (rows of envs left to the second pass are skipped)
        const int64_t base = (int64_t)blockIdx.x * kBlock;
        const int64_t rows = (a.n_envs - base) < kBlock ? (a.n_envs - base) : kBlock;
        const int n4 = (int)(rows * kObs / 4);   // rows*12 floats is always a multiple of 4
        if (stepped) {
#pragma unroll
            for (int q = 0; q < kObs; ++q) s_obs[threadIdx.x * kObs + q] = r.obs[q];
        }
        s_skip[threadIdx.x] = stepped ? 0 : 1;
        __syncthreads();
        float4* dst = reinterpret_cast<float4*>(a.out.obs + base * kObs);
        const float4* src = reinterpret_cast<const float4*>(s_obs);
        for (int q = threadIdx.x; q < n4; q += kBlock)
            if (!s_skip[q / (kObs / 4)]) dst[q] = src[q];
    } else if (stepped) {
        store_row(a.out.obs + e * kObs, r.obs);
    }
    // last observation of an episode that ended in this step: written for those envs only (rows of the other envs keep what
    // they held; the caller selects rows with terminated | truncated)
    if (want_fin && stepped && r.did_reset) store_row(a.out.final_obs + e * kObs, r.final_obs);

    // ---- episode statistics: warp shuffle reduction, one atomic per warp and statistic (K5 input) ---------------------
    if (a.out.stats) {
        double v[STG_NSTATS];
#pragma unroll
        for (int q = 0; q < STG_NSTATS; ++q) v[q] = 0.0;
        if (stepped) {
            const bool ended = r.terminated || r.truncated;
            v[STG_STAT_STEPS] = 1.0;
            v[STG_STAT_SUBSTEPS] = r.valid ? (double)r.n_sub : 0.0;
            v[STG_STAT_TERMINATED] = r.terminated ? 1.0 : 0.0;
            v[STG_STAT_TRUNCATED] = (!r.terminated && r.truncated) ? 1.0 : 0.0;
            v[STG_STAT_ENERGY] = r.energy;
            v[STG_STAT_REWARD] = r.reward;
            v[STG_STAT_GUARD] = (r.status & 1) ? 1.0 : 0.0;
            v[STG_STAT_EPLEN] = ended ? (double)r.step_after : 0.0;
        }
#pragma unroll
        for (int q = 0; q < STG_NSTATS; ++q) {
            const double s = warp_sum(v[q]);
            if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(a.out.stats + (blockIdx.x % STG_STAT_REPLICAS) * STG_NSTATS + q, s);
        }
    }
}

// ---- two envs per thread: packed FP32x2 (FFMA2) variant of the fast path ------------------------------------------------
// R = float, e = z^, RK4, NOISE in {0 (none), 1 (in-kernel stream)}. Thread t of a CTA owns the adjacent slots 2t, 2t+1, so the FP64
// state planes are read as 16-byte pairs and one CTA covers 2*kBlock observation rows.
__device__ __forceinline__ void accumulate_stats(double* v, const EnvStepResult& r) {
    const bool ended = r.terminated || r.truncated;
    v[STG_STAT_STEPS] += 1.0;
    v[STG_STAT_SUBSTEPS] += r.valid ? (double)r.n_sub : 0.0;
    v[STG_STAT_TERMINATED] += r.terminated ? 1.0 : 0.0;
    v[STG_STAT_TRUNCATED] += (!r.terminated && r.truncated) ? 1.0 : 0.0;
    v[STG_STAT_ENERGY] += r.energy;
    v[STG_STAT_REWARD] += r.reward;
    v[STG_STAT_GUARD] += (r.status & 1) ? 1.0 : 0.0;
    v[STG_STAT_EPLEN] += ended ? (double)r.step_after : 0.0;
}

#ifndef STG_PAIR_MINBLOCKS
#define STG_PAIR_MINBLOCKS 8
#endif
#ifndef STG_PAIR_THERMAL_MIN_ENVS
#define STG_PAIR_THERMAL_MIN_ENVS (1 << 18)
#endif
#ifndef STG_PAIR_MINBLOCKS_TH
// thermal pair kernel: 154 registers without spills at 6 CTAs (12 warps) per SM: 8.83 ms per 1M-env x 999-substep step; held at
// 128 registers (8 CTAs) it spills 40 bytes: 9.21 ms; 5 CTAs (166 registers): 9.26 ms
#define STG_PAIR_MINBLOCKS_TH 6
#endif
template <int NOISE>
__global__ void __launch_bounds__(kBlock, NOISE == 0 ? STG_PAIR_MINBLOCKS : (NOISE == 3 ? 8 : STG_PAIR_MINBLOCKS_TH)) stt_env_step_pair_kernel(const __grid_constant__ StepArgs a) {
    __shared__ __align__(16) float s_obs[2 * kBlock * kObs];
    __shared__ uint8_t s_skip[2 * kBlock];
    const int64_t base = (int64_t)blockIdx.x * (2 * kBlock);
    const int64_t slotA = base + 2 * threadIdx.x, slotB = slotA + 1;
    const bool actA = slotA < a.n_envs, actB = slotB < a.n_envs;
    const bool sorted = (a.flags & STG_F_SORTED) != 0;
    const bool want_fin = (a.flags & STG_F_AUTORESET) != 0 && a.out.final_obs != nullptr;
    const int64_t eA = actA ? (sorted ? (int64_t)a.d_perm[slotA] : slotA) : 0;
    const int64_t eB = actB ? (sorted ? (int64_t)a.d_perm[slotB] : slotB) : 0;

    EnvStepResult rA, rB;
    rA.did_reset = rB.did_reset = false;
    int redo = 0;
    if (actB) redo = env_step_pair_body<NOISE>(a, eA, eB, rA, rB);
    else if (actA) redo = env_step_body<float, true, NOISE, false>(a, eA, rA) ? 0 : 1;
    if (redo & 1) redo_push(a, eA);
    if (redo & 2) redo_push(a, eB);
    const bool doneA = actA && !(redo & 1), doneB = actB && !(redo & 2);

    if (!sorted) {
        const int64_t rows = (a.n_envs - base) < 2 * kBlock ? (a.n_envs - base) : 2 * kBlock;
        const int n4 = (int)(rows * kObs / 4);
#pragma unroll
        for (int q = 0; q < kObs; ++q) {
            if (doneA) s_obs[(2 * threadIdx.x) * kObs + q] = rA.obs[q];
            if (doneB) s_obs[(2 * threadIdx.x + 1) * kObs + q] = rB.obs[q];
        }
        s_skip[2 * threadIdx.x] = doneA ? 0 : 1;
        s_skip[2 * threadIdx.x + 1] = doneB ? 0 : 1;
        __syncthreads();
        float4* dst = reinterpret_cast<float4*>(a.out.obs + base * kObs);
        const float4* src = reinterpret_cast<const float4*>(s_obs);
        for (int q = threadIdx.x; q < n4; q += kBlock)
            if (!s_skip[q / (kObs / 4)]) dst[q] = src[q];
    } else {
        if (doneA) store_row(a.out.obs + eA * kObs, rA.obs);
        if (doneB) store_row(a.out.obs + eB * kObs, rB.obs);
    }
    if (want_fin) {      // see stt_env_step_kernel
        if (doneA && rA.did_reset) store_row(a.out.final_obs + eA * kObs, rA.final_obs);
        if (doneB && rB.did_reset) store_row(a.out.final_obs + eB * kObs, rB.final_obs);
    }
    if (a.out.stats) {
        double v[STG_NSTATS];
#pragma unroll
        for (int q = 0; q < STG_NSTATS; ++q) v[q] = 0.0;
        if (doneA) accumulate_stats(v, rA);
        if (doneB) accumulate_stats(v, rB);
#pragma unroll
        for (int q = 0; q < STG_NSTATS; ++q) {
            const double sum = warp_sum(v[q]);
            if ((threadIdx.x & 31) == 0 && sum != 0.0) atomicAdd(a.out.stats + (blockIdx.x % STG_STAT_REPLICAS) * STG_NSTATS + q, sum);
        }
    }
}

// (A warp-specialised variant of the thermal step - three producer warps evaluating the noise into a shared-memory ring for one
// consumer warp - was measured in round 2 and removed: 11.5 ms against 10.4 ms for the kernel above; profiles/README.md.)

// ---- second pass of stg_stt_step_f32: the envs the FP32 stages declined, compacted, with FP64 stages ---------------------
// Grid-stride over the list d_redo[STG_REDO_HEADER ..] (count in d_redo[0], written by the first pass on the same stream).
// Typically empty or a fraction of a per cent of the batch, so rows are stored per thread and statistics added per thread.
// It is the only writer of these envs' outputs: the first pass leaves their rows untouched.
template <int NOISE>
__global__ void __launch_bounds__(kBlock) stt_env_redo_kernel(const __grid_constant__ StepArgs a) {
    const int64_t count = a.d_redo[0] < a.n_envs ? (int64_t)a.d_redo[0] : a.n_envs;
    const bool want_fin = (a.flags & STG_F_AUTORESET) != 0 && a.out.final_obs != nullptr;
    for (int64_t slot = (int64_t)blockIdx.x * kBlock + threadIdx.x; slot < count; slot += (int64_t)gridDim.x * kBlock) {
        const int64_t e = a.d_redo[STG_REDO_HEADER + slot];
        EnvStepResult r;
        r.did_reset = false;
        env_step_body<double, true, NOISE, false>(a, e, r, STG_STATUS_REDONE_F64);
        store_row(a.out.obs + e * kObs, r.obs);
        if (want_fin && r.did_reset) store_row(a.out.final_obs + e * kObs, r.final_obs);
        if (a.out.stats) {
            double v[STG_NSTATS];
#pragma unroll
            for (int q = 0; q < STG_NSTATS; ++q) v[q] = 0.0;
            accumulate_stats(v, r);
#pragma unroll
            for (int q = 0; q < STG_NSTATS; ++q)
                if (v[q] != 0.0) atomicAdd(a.out.stats + (blockIdx.x % STG_STAT_REPLICAS) * STG_NSTATS + q, v[q]);
        }
    }
}

__global__ void __launch_bounds__(128) stt_env_reset_kernel(const __grid_constant__ ResetArgs a) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n_envs) return;
    if (a.d_mask && !a.d_mask[e]) return;
    env_reset_body(a, e);
}

// ---- counting sort of envs by substep count (descending); helpers in sort_utils.cuh ---------------------------------------
__global__ void sort_hist_kernel(const StgSttFolded* table, const int32_t* pidx, const float* action, int32_t* hist,
                                 int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < n;
    warp_aggregated_inc(hist, valid ? action_bin(table, pidx, action, e) : 0, valid);
}
__global__ void sort_scatter_kernel(const StgSttFolded* table, const int32_t* pidx, const float* action, int32_t* hist,
                                    int32_t* perm, int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < n;
    if (hist[STG_SORT_BINS] <= 1) {   // every env integrates the same number of substeps: keep the coalesced identity order
        if (valid) perm[e] = (int32_t)e;
        return;
    }
    const int pos = warp_aggregated_inc(hist, valid ? action_bin(table, pidx, action, e) : 0, valid);
    if (valid) perm[pos] = (int32_t)e;
}

template <typename R, bool AXIS_Z, int NOISE, bool EULER>
__global__ void __launch_bounds__(kBlock) stt_solve_kernel(const __grid_constant__ SolveArgs a) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= a.n_envs) return;
    solve_body<R, AXIS_Z, NOISE, EULER>(a, e);
}

template <int NOISE, bool EULER>
__global__ void __launch_bounds__(kBlock) stt_solve_grid_kernel(const __grid_constant__ SolveArgs a) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= a.n_envs) return;
    solve_grid_body<NOISE, EULER>(a, e);
}

// ---- dispatch --------------------------------------------------------------------------------------------------------
// (Measured and dropped in round 2: capping the residency with unused dynamic shared memory so that a grid of 1.73 waves - the
// strong-scaling shard of 131,072 envs - runs as two equal waves at 14 warps/SM instead of a full and a 73 % one: 1.415 ms
// against 1.373 ms as it is, the kernel's throughput still grows from 14 to 16 warps per SM. profiles/README.md.)
template <typename R, bool AXIS_Z, int NOISE>
static cudaError_t launch_step2(const StepArgs& a, cudaStream_t s) {
    const unsigned grid = (unsigned)((a.n_envs + kBlock - 1) / kBlock);
    if constexpr (sizeof(R) == 8 && NOISE != 3) {      // Euler always runs FP64 stages (launch_step), noise from Philox blocks
        if (a.flags & STG_F_EULER) {
            stt_env_step_kernel<R, AXIS_Z, NOISE, true><<<grid, kBlock, 0, s>>>(a);
            return cudaGetLastError();
        }
    }
    stt_env_step_kernel<R, AXIS_Z, NOISE, false><<<grid, kBlock, 0, s>>>(a);
    return cudaGetLastError();
}
// FP32 stages without the Philox stream: zero the redo counter, run the FP32 kernel, then the compacted FP64 pass
static bool step_uses_redo(const StepArgs& a, bool f32, bool axis_z) {
    return f32 && axis_z && !(a.flags & (STG_F_EULER | STG_F_THERMAL_PHILOX));
}
template <int NOISE>
static cudaError_t launch_redo(const StepArgs& a, cudaStream_t s) {
    const int64_t blocks = (a.n_envs + kBlock - 1) / kBlock;
    const unsigned grid = (unsigned)(blocks < 4 * 148 ? blocks : 4 * 148);
    stt_env_redo_kernel<NOISE><<<grid, kBlock, 0, s>>>(a);
    return cudaGetLastError();
}
// FP32 stage arithmetic exists for the axis-aligned geometry only (compensated constants + block scaling, llgs_core.cuh).
// A tilted easy axis / applied field always runs FP64 stages, also through the _f32 entry points: rounding the axis
// components to 24 bits is a systematic rate error the 1e-4 contract does not survive over thousands of substeps.
// The explicit Euler map amplifies rounding errors chaotically (0.35 rad per substep, renormalised): FP32 stages do not hold the
// 1e-4 contract there either, so Euler runs FP64 stages through both entry points as well.
// Does the thermal FP32 step (axis z, RK4, in-kernel stream) of n_envs envs take the two-envs-per-thread kernel? (exported as
// stg_stt_thermal_pair_dispatch for tests and tools)
//   * STG_F_NO_PAIR / STG_F_PAIR_ALWAYS decide by themselves;
//   * from STG_PAIR_THERMAL_MIN_ENVS envs (262,144; all-Philox stream: 524,288, where the packed kernel only draws level - 5.24 vs
//     5.23 ms, 2.70 vs 2.65 at 262,144) the GPU is several waves deep and the packed kernel's throughput wins (8.55 vs 9.22 ms
//     at 1,048,576 envs, 2.21 vs 2.35 at 262,144);
//   * batches of at most one wave of the packed kernel: every CTA is two warps and an SM has four schedulers, so with k CTAs per
//     SM the busiest scheduler runs ceil(k / 2) warps and the step time is a staircase in that count (measured, 999 substeps,
//     one env per thread: 0.25 / 0.375 / 0.56 / 0.72 / 0.90 / 1.05 / 1.22 ms for 1 .. 7 warps; two envs per thread: 0.415 / 0.70 /
//     0.99 / 1.30 ms for 1 .. 4). The packed kernel has half as many CTAs, each ~1.75x as long: it wins (by 3 - 6 %) exactly
//     when it halves the busiest scheduler's warp count and that count is at least 2 - e.g. 65,536 envs (BASELINE configs[1]):
//     7 CTAs per SM = 4 warps against 4 CTAs = 2 warps, 0.70 vs 0.72 ms; 95,000 .. 113,664 envs: 0.99 vs 1.05 - and loses
//     otherwise (76,000 .. 94,000 envs: 0.99 vs 0.90; 131,072: 1.30 vs 1.21). profiles/README.md.
static bool thermal_pair_dispatch(int64_t n_envs, uint32_t flags, int sms) {
    if (flags & STG_F_NO_PAIR) return false;
    if (flags & STG_F_PAIR_ALWAYS) return true;
    const bool all_philox = (flags & STG_F_STREAM_PHILOX10) != 0;
    if (n_envs >= (all_philox ? 2 : 1) * (int64_t)STG_PAIR_THERMAL_MIN_ENVS) return true;
    if (all_philox || kBlock != 64 || sms <= 0) return false;
    const int64_t k_s = ((n_envs + kBlock - 1) / kBlock + sms - 1) / sms;
    const int64_t k_p = ((n_envs + 2 * kBlock - 1) / (2 * kBlock) + sms - 1) / sms;
    const int64_t w_s = (k_s + 1) / 2, w_p = (k_p + 1) / 2;
    return k_p <= STG_PAIR_MINBLOCKS_TH && w_p >= 2 && w_s == 2 * w_p;
}
static int sm_count() {      // of the current device (every GPU of a node is the same part); 148 on B200
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            return 148;
    }
    return n;
}
template <typename R>
static cudaError_t launch_step(const StepArgs& a, bool axis_z, cudaStream_t s) {
    // 3: the in-kernel stream with every word from Philox4x32-10 (Euler draws from Philox blocks in either case)
    const int noise = (a.flags & STG_F_THERMAL_INJECT) ? 2 : ((a.flags & STG_F_THERMAL_PHILOX)
                      ? (((a.flags & STG_F_STREAM_PHILOX10) && !(a.flags & STG_F_EULER)) ? 3 : 1) : 0);
    if (sizeof(R) == 4 && (a.flags & STG_F_EULER)) return launch_step<double>(a, axis_z, s);
    if (step_uses_redo(a, sizeof(R) == 4, axis_z)) {
        cudaError_t err = cudaMemsetAsync(a.d_redo, 0, sizeof(int32_t) * STG_REDO_HEADER, s);
        if (err != cudaSuccess) return err;
        if (noise == 0 && !(a.flags & STG_F_NO_PAIR)) {
            // two envs per thread, Blackwell packed FP32x2 arithmetic (bit-identical to the one-env-per-thread kernels).
            // Measured (profiles/README.md): +12 % without thermal noise.
            const unsigned grid = (unsigned)((a.n_envs + 2 * kBlock - 1) / (2 * kBlock));
            stt_env_step_pair_kernel<0><<<grid, kBlock, 0, s>>>(a);
            err = cudaGetLastError();
        } else {
            err = noise == 0 ? launch_step2<float, true, 0>(a, s) : launch_step2<float, true, 2>(a, s);
        }
        if (err != cudaSuccess) return err;
        return noise == 0 ? launch_redo<0>(a, s) : launch_redo<2>(a, s);
    }
    if (axis_z) {
        if (sizeof(R) == 4 && (noise == 1 || noise == 3) && thermal_pair_dispatch(a.n_envs, a.flags, sm_count())) {
            // two envs per thread on packed FP32x2 arithmetic, one noise stream per lane (bit-identical to one env per thread).
            // Measured (profiles/README.md, 999 substeps): 8.78 vs 9.27 ms at 1,048,576 envs, 2.26 vs 2.36 at 262,144; below
            // that the one-env-per-thread kernel has twice as many threads to fill the GPU with (1.33 vs 1.21 ms at 131,072
            // envs, 0.42 vs 0.26 ms up to 16,384), so smaller batches take it.
            const unsigned grid = (unsigned)((a.n_envs + 2 * kBlock - 1) / (2 * kBlock));
            if (noise == 3) stt_env_step_pair_kernel<3><<<grid, kBlock, 0, s>>>(a);
            else stt_env_step_pair_kernel<1><<<grid, kBlock, 0, s>>>(a);
            return cudaGetLastError();
        }
        if (noise == 0) return launch_step2<R, true, 0>(a, s);
        if (noise == 1) return launch_step2<R, true, 1>(a, s);
        if (noise == 3) return launch_step2<R, true, 3>(a, s);
        return launch_step2<R, true, 2>(a, s);
    }
    if (noise == 0) return launch_step2<double, false, 0>(a, s);
    if (noise == 1) return launch_step2<double, false, 1>(a, s);
    if (noise == 3) return launch_step2<double, false, 3>(a, s);
    return launch_step2<double, false, 2>(a, s);
}

template <typename R, bool AXIS_Z, int NOISE>
static cudaError_t launch_solve2(const SolveArgs& a, uint32_t flags, cudaStream_t s) {
    const unsigned grid = (unsigned)((a.n_envs + kBlock - 1) / kBlock);
    if (flags & STG_F_EULER)
        stt_solve_kernel<R, AXIS_Z, NOISE, true><<<grid, kBlock, 0, s>>>(a);
    else
        stt_solve_kernel<R, AXIS_Z, NOISE, false><<<grid, kBlock, 0, s>>>(a);
    return cudaGetLastError();
}
template <int NOISE>
static cudaError_t launch_solve_grid(const SolveArgs& a, uint32_t flags, cudaStream_t s) {
    const unsigned grid = (unsigned)((a.n_envs + kBlock - 1) / kBlock);
    if (flags & STG_F_EULER)
        stt_solve_grid_kernel<NOISE, true><<<grid, kBlock, 0, s>>>(a);
    else
        stt_solve_grid_kernel<NOISE, false><<<grid, kBlock, 0, s>>>(a);
    return cudaGetLastError();
}
template <typename R>
static cudaError_t launch_solve(const SolveArgs& a, uint32_t flags, bool axis_z, cudaStream_t s) {
    const int noise = (flags & STG_F_THERMAL_INJECT) ? 2 : ((flags & STG_F_THERMAL_PHILOX) ? 1 : 0);
    if (a.d_current_grid || a.d_field_grid) {      // host-sampled callables: FP64 general stages
        if (noise == 0) return launch_solve_grid<0>(a, flags, s);
        if (noise == 1) return launch_solve_grid<1>(a, flags, s);
        return launch_solve_grid<2>(a, flags, s);
    }
    if (axis_z) {
        if (noise == 0) return launch_solve2<R, true, 0>(a, flags, s);
        if (noise == 1) return launch_solve2<R, true, 1>(a, flags, s);
        return launch_solve2<R, true, 2>(a, flags, s);
    }
    if (noise == 0) return launch_solve2<double, false, 0>(a, flags, s);
    if (noise == 1) return launch_solve2<double, false, 1>(a, flags, s);
    return launch_solve2<double, false, 2>(a, flags, s);
}

}  // namespace stg

// =====================================================================================================================
// C-ABI
// =====================================================================================================================
using namespace stg;

extern "C" int stg_abi_version(void) { return STG_ABI_VERSION; }
extern "C" int stg_stt_thermal_pair_dispatch(int64_t n_envs, uint32_t flags, int sm_count) {
    return stg::thermal_pair_dispatch(n_envs, flags, sm_count > 0 ? sm_count : stg::sm_count()) ? 1 : 0;
}

extern "C" const char* stg_error_string(int code) {
    switch (code) {
        case STG_OK: return "ok";
        case STG_E_NULL: return "required pointer is NULL";
        case STG_E_SIZE: return "negative or inconsistent size";
        case STG_E_ENUM: return "unknown enum value or flag combination";
        case STG_E_ALIGN: return "pointer not aligned as required";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown stg error";
    }
}

extern "C" int stg_stt_fold(const StgSttParams* params, int32_t n_sets, StgSttFolded* out) {
    if (!params || !out) return STG_E_NULL;
    if (n_sets <= 0) return STG_E_SIZE;
    for (int i = 0; i < n_sets; ++i) {
        const StgSttParams& p = params[i];
        if (p.device_kind < STG_DEV_STT || p.device_kind > STG_DEV_VCMA) return STG_E_ENUM;
        double* v = out[i].v;
        memset(v, 0, sizeof(out[i].v));
        const double alpha = p.damping, ms = p.saturation_magnetization, vol = p.volume;
        v[FI_ALPHA] = alpha;
        v[FI_GEFF] = kGamma / (1.0 + alpha * alpha);                               // physics/simple_solver.py:338
        v[FI_HK] = (2.0 * p.uniaxial_anisotropy) / (kMu0 * ms);                    // :368
        v[FI_MS] = ms;
        v[FI_AJ_PER_J] = p.polarization / (ms * vol);                              // :330
        v[FI_HTH] = (p.thermal && p.temperature > 0.0)
                        ? sqrt(2.0 * alpha * kKbSolver * p.temperature / (kMu0 * ms * vol * kGamma))   // :375-380
                        : 0.0;
        const double en = sqrt(p.easy_axis[0] * p.easy_axis[0] + p.easy_axis[1] * p.easy_axis[1] +
                               p.easy_axis[2] * p.easy_axis[2]);
        const double rn = sqrt(p.reference_magnetization[0] * p.reference_magnetization[0] +
                               p.reference_magnetization[1] * p.reference_magnetization[1] +
                               p.reference_magnetization[2] * p.reference_magnetization[2]);
        if (!(en > 0.0) || !(rn > 0.0)) return STG_E_SIZE;
        for (int k = 0; k < 3; ++k) {
            v[FI_EX + k] = p.easy_axis[k] / en;                                    // :319
            v[FI_REFX + k] = p.reference_magnetization[k] / rn;
            v[FI_HAX + k] = p.applied_field[k];
        }
        v[FI_RP] = p.resistance_parallel;
        v[FI_RAP] = p.resistance_antiparallel;
        v[FI_TMR] = (p.resistance_antiparallel - p.resistance_parallel) / p.resistance_parallel;
        v[FI_AREA] = p.area;
        v[FI_RSERIES] = p.series_resistance;
        v[FI_TEMP] = p.temperature;
        v[FI_MAXCUR] = p.max_current;
        v[FI_MAXDUR] = p.max_duration;
        v[FI_SUCC] = p.success_threshold;
        v[FI_WE] = p.energy_penalty_weight;
        v[FI_MAXSTEP_DT] = p.max_step > 0.0 ? p.max_step : 1e-12;
        v[FI_MAXSTEPS] = (double)p.max_steps;
        v[FI_KIND] = (double)p.device_kind;
        v[FI_THERMAL] = (double)p.thermal;
        v[FI_VALID] = (double)p.solver_valid;
        const bool axis_z = v[FI_EX] == 0.0 && v[FI_EY] == 0.0 && v[FI_EZ] == 1.0 && p.applied_field[0] == 0.0 &&
                            p.applied_field[1] == 0.0 && p.applied_field[2] == 0.0;
        v[FI_AXISZ] = axis_z ? 1.0 : 0.0;
    }
    return STG_OK;
}

extern "C" int stg_stt_all_axis_z(const StgSttFolded* folded_host, int32_t n_sets) {
    if (!folded_host || n_sets <= 0) return 0;
    for (int i = 0; i < n_sets; ++i)
        if (folded_host[i].v[FI_AXISZ] == 0.0) return 0;
    return 1;
}

static int check_step_args(const StgSttStepArgs& a) {
    if (a.n_envs < 0 || a.n_sets <= 0) return STG_E_SIZE;
    const StgSttState& st = a.state;
    if (!a.d_table || !st.m || !st.target || !st.total_energy || !st.last_action || !st.step_count || !st.episode ||
        !a.d_action || !a.out.obs || !a.out.reward || !a.out.terminated || !a.out.truncated)
        return STG_E_NULL;
    if ((a.flags & STG_F_THERMAL_PHILOX) && (a.flags & STG_F_THERMAL_INJECT)) return STG_E_ENUM;
    if ((a.flags & STG_F_THERMAL_INJECT) && (!a.d_noise || a.noise_stride <= 0)) return STG_E_NULL;
    if ((a.flags & STG_F_SORTED) && !a.d_perm) return STG_E_NULL;
    if ((a.flags & STG_F_AUTORESET) && (!a.d_target_table || a.n_targets <= 0)) return STG_E_NULL;
    if (((uintptr_t)a.d_action & 7u) || ((uintptr_t)a.out.obs & 15u) ||
        (a.out.final_obs && ((uintptr_t)a.out.final_obs & 15u)))
        return STG_E_ALIGN;
    return STG_OK;
}

template <typename R>
static int stt_step_impl(const StgSttStepArgs* args, void* stream) {
    if (!args) return STG_E_NULL;
    int rc = check_step_args(*args);
    if (rc != STG_OK) return rc;
    if (step_uses_redo(*args, sizeof(R) == 4, (args->flags & STG_F_AXIS_Z) != 0)) {
        if (!args->d_redo) return STG_E_NULL;
        if (args->n_envs > 2147483647LL - STG_REDO_HEADER) return STG_E_SIZE;
    }
    if (args->n_envs == 0) return STG_OK;
    return (int)launch_step<R>(*args, (args->flags & STG_F_AXIS_Z) != 0, (cudaStream_t)stream);
}
extern "C" int stg_stt_step_f32(const StgSttStepArgs* args, void* stream) { return stt_step_impl<float>(args, stream); }
extern "C" int stg_stt_step_f64(const StgSttStepArgs* args, void* stream) { return stt_step_impl<double>(args, stream); }

extern "C" int stg_stt_reset(const StgSttResetArgs* args, void* stream) {
    if (!args) return STG_E_NULL;
    const StgSttResetArgs& a = *args;
    if (a.n_envs < 0 || a.n_sets <= 0) return STG_E_SIZE;
    const StgSttState& st = a.state;
    if (!a.d_table || !st.m || !st.target || !st.total_energy || !st.last_action || !st.step_count || !st.episode)
        return STG_E_NULL;
    if (!a.d_target0 && (!a.d_target_table || a.n_targets <= 0)) return STG_E_NULL;
    if (a.d_obs && ((uintptr_t)a.d_obs & 15u)) return STG_E_ALIGN;
    if (a.n_envs == 0) return STG_OK;
    const unsigned grid = (unsigned)((a.n_envs + 127) / 128);
    stt_env_reset_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int stg_stt_sort_by_substeps(const StgSttFolded* d_table, int32_t n_sets, const int32_t* d_param_index,
                                        const float* d_action, int32_t* d_perm, int32_t* d_work, int64_t n_envs,
                                        void* stream) {
    if (!d_table || !d_action || !d_perm || !d_work) return STG_E_NULL;
    if (n_envs < 0 || n_sets <= 0 || n_envs > 2147483647LL) return STG_E_SIZE;
    if (n_envs == 0) return STG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaMemsetAsync(d_work, 0, sizeof(int32_t) * STG_SORT_WORK_INTS, s);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)((n_envs + 255) / 256);
    sort_hist_kernel<<<grid, 256, 0, s>>>(d_table, d_param_index, d_action, d_work, n_envs);
    sort_scan_kernel<<<1, 1024, 0, s>>>(d_work);
    sort_scatter_kernel<<<grid, 256, 0, s>>>(d_table, d_param_index, d_action, d_work, d_perm, n_envs);
    return (int)cudaGetLastError();
}

template <typename R>
static int stt_solve_impl(const StgSttSolveArgs* args, void* stream) {
    if (!args) return STG_E_NULL;
    const StgSttSolveArgs& a = *args;
    if (a.n_envs < 0 || a.n_sets <= 0) return STG_E_SIZE;
    if (!a.d_table || !a.d_m0 || !a.d_pulse || !a.d_m_out) return STG_E_NULL;
    if ((a.flags & STG_F_THERMAL_PHILOX) && (a.flags & STG_F_THERMAL_INJECT)) return STG_E_ENUM;
    if ((a.flags & STG_F_THERMAL_INJECT) && (!a.d_noise || a.noise_stride <= 0)) return STG_E_NULL;
    if (a.d_traj && a.traj_stride <= 0) return STG_E_SIZE;
    if ((a.d_current_grid || a.d_field_grid) &&
        (a.grid_stride <= 0 || (a.grid_envs != 1 && a.grid_envs != a.n_envs) || (a.flags & STG_F_VECTORIZED_PLAN)))
        return STG_E_SIZE;
    if (a.n_envs == 0) return STG_OK;
    return (int)launch_solve<R>(a, a.flags, (a.flags & STG_F_AXIS_Z) != 0, (cudaStream_t)stream);
}
extern "C" int stg_stt_solve_f32(const StgSttSolveArgs* args, void* stream) { return stt_solve_impl<float>(args, stream); }
extern "C" int stg_stt_solve_f64(const StgSttSolveArgs* args, void* stream) { return stt_solve_impl<double>(args, stream); }
