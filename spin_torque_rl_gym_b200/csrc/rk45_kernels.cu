// rk45_kernels.cu — K2 launch + C-ABI entry (see rk45_core.cuh). One trajectory per thread; lanes whose trajectory has finished
// idle until the warp's slowest lane is done, so the caller passes a permutation that sorts the batch by (parameter set, t_end)
// and warps hold trajectories of similar length. A resident grid whose lanes pick up the next trajectory when theirs ends was
// measured instead (262,144 trajectories): 8 % slower on a uniform batch, 10 % faster on t_end ~ U(0.05, 1) - with ~3.5
// trajectories per resident thread the tail of every CTA costs a quarter of its lifetime - and dropped in favour of the sort.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stg.h"
#include "rk45_core.cuh"

namespace stg {

#ifndef STG_RK45_MINBLOCKS
// Occupancy sweep on the configs[2] mix (262,144 trajectories x 114 attempts, B200): min blocks 1 (184 registers, 8 warps/SM)
// 3.99 ms; 5: 3.34; 6: 3.41; 7: 3.23; 8 (128 registers + 72 B spill, 16 warps/SM): 3.04; 10: 3.17; 12: 3.47; 16: 4.02 ms.
// The kernel waits on dependent FP64 chains (ncu: `wait` 2.1 stalls per issue at 1.9 warps per scheduler), so warps beat registers.
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
