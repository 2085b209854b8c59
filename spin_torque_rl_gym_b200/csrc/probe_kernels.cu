// probe_kernels.cu — FP32 / FP64 FMA-pipe throughput probes used by bench.py as the measured roofline denominator
// (MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the LLGS path is bound by the FP32 FMA pipe).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stg.h"

namespace stg {

template <typename T>
__global__ void __launch_bounds__(256) fma_probe_kernel(T* out, int iters, T a, T b) {
    // 8 independent dependent-FMA chains per thread: enough ILP to cover the 4-cycle pipe latency at any occupancy
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    T s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == (T)12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chains alive
}

// packed FP32x2 variant (Blackwell FFMA2): 8 independent chains of float2 per thread = 128 FMAs per loop iteration
__global__ void __launch_bounds__(256) fma2_probe_kernel(float2* out, int iters, float a, float b) {
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    float2 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = make_float2((float)threadIdx.x + u, (float)threadIdx.x - u);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __ffma2_rn(x[u], a2, b2);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += x[u].x + x[u].y;
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(s, s);
}

}  // namespace stg

// Launches blocks*256 threads, each executing iters*64 FMAs. flops = blocks*256*iters*64*2.
extern "C" int stg_probe_fma(void* d_out, int32_t blocks, int32_t iters, int32_t f64, void* stream) {
    if (!d_out) return STG_E_NULL;
    if (blocks <= 0 || iters <= 0) return STG_E_SIZE;
    if (f64 == 2)   // FP32x2 packed: blocks*256 threads x iters*64 FFMA2 = iters*128 FMAs per thread
        stg::fma2_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((float2*)d_out, iters, 0.999999f, 1e-7f);
    else if (f64)
        stg::fma_probe_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)d_out, iters, 0.999999, 1e-7);
    else
        stg::fma_probe_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)d_out, iters, 0.999999f, 1e-7f);
    return (int)cudaGetLastError();
}
