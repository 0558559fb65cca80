// FP32 peak of the device this context runs on, measured live: packed FFMA2 (fma.rn.f32x2), 8 independent
// chains per thread, no memory traffic.  bench.py divides k_mp_iter's achieved FLOP/s by this number instead of a
// datasheet figure (MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only).
#pragma once
#include <cuda_runtime.h>

#include "gnn_kernels.cuh"

namespace enero {

constexpr int UB_ITERS = 4096;
__global__ void __launch_bounds__(256)
k_ubench_ffma2(float* __restrict__ out, float x, float y) {
    u64 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pack2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const u64 X = pack2(x, x), Y = pack2(y, y);
    for (int it = 0; it < UB_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], X, Y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; unpack2(a[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace enero
