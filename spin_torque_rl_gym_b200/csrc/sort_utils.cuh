// sort_utils.cuh — counting-sort helpers shared by the K1 substep sort (stt_kernels.cu) and the K2 cost sort (rk45_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stg.h"

namespace stg {

// Warp-aggregated atomics: lanes that fall into the same bin elect a leader which issues ONE atomicAdd for the group
// (with a fixed pulse duration all 1M envs share a bin; un-aggregated that is 1M serialised atomics on one address).
__device__ __forceinline__ int warp_aggregated_inc(int32_t* counters, int bin, bool valid) {
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    int pos = 0;
    if (valid) {
        const unsigned peers = __match_any_sync(active, bin);
        const int leader = __ffs(peers) - 1;
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(counters + bin, __popc(peers));
        base = __shfl_sync(peers, base, leader);
        pos = base + __popc(peers & ((1u << lane) - 1u));
    }
    return pos;
}

// one block of 1024 threads, 8 bins each (STG_SORT_BINS = 8192): exclusive scan of the histogram in place;
// hist[STG_SORT_BINS] receives the number of non-empty bins
static __global__ void sort_scan_kernel(int32_t* hist) {
    __shared__ int32_t s[1024];
    __shared__ int32_t s_nonempty;
    const int t = threadIdx.x;
    if (t == 0) s_nonempty = 0;
    __syncthreads();
    int32_t loc[8];
    int32_t sum = 0, ne = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { loc[q] = hist[t * 8 + q]; sum += loc[q]; ne += loc[q] != 0; }
    if (ne) atomicAdd(&s_nonempty, ne);
    s[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int32_t v = (t >= off) ? s[t - off] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    int32_t run = s[t] - sum;
#pragma unroll
    for (int q = 0; q < 8; ++q) { hist[t * 8 + q] = run; run += loc[q]; }
    if (t == 0) hist[STG_SORT_BINS] = s_nonempty;
}

}  // namespace stg
