// Micro-benchmark: FP32 FMA issue rate on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ffma2 ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }

template <int ITERS>
__global__ void k_ffma(float* out, float x, float y) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ITERS>
__global__ void k_ffma2(float* out, float x, float y) {
    u64 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pack(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const u64 X = pack(x, x), Y = pack(y, y);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], X, Y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += lo(a[i]) + hi(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// broadcast-LDS.128 fed FMAs: 1 LDS.128 per 4 FFMA (R=1) and per 8 FFMA (R=2)
template <int ITERS, int R>
__global__ void k_lds_ffma(float* out, float x) {
    __shared__ __align__(16) float w[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) w[i] = 1e-6f * i;
    __syncthreads();
    float a[R][4], xs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { xs[r] = x + r; for (int i = 0; i < 4; ++i) a[r][i] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(&w[((it & 15) * 64 + k) * 4]);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                a[r][0] = fmaf(xs[r], v.x, a[r][0]); a[r][1] = fmaf(xs[r], v.y, a[r][1]);
                a[r][2] = fmaf(xs[r], v.z, a[r][2]); a[r][3] = fmaf(xs[r], v.w, a[r][3]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) for (int i = 0; i < 4; ++i) s += a[r][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, tpb = 256; constexpr int IT = 4096;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * tpb);
    float ms1 = time_ms([&] { k_ffma<IT><<<blocks, tpb>>>(out, 1.0001f, 0.5f); });
    float ms2 = time_ms([&] { k_ffma2<IT><<<blocks, tpb>>>(out, 1.0001f, 0.5f); });
    const double fl = 2.0 * 16 * IT * (double)blocks * tpb;
    printf("SMs %d clock %d kHz\n", p.multiProcessorCount, p.clockRate);
    printf("FFMA : %.3f ms  %.1f TFLOP/s\n", ms1, fl / ms1 / 1e9);
    printf("FFMA2: %.3f ms  %.1f TFLOP/s\n", ms2, fl / ms2 / 1e9);
    constexpr int IT2 = 256;
    float ms3 = time_ms([&] { k_lds_ffma<IT2, 1><<<blocks, tpb>>>(out, 1.0001f); });
    float ms4 = time_ms([&] { k_lds_ffma<IT2, 2><<<blocks, tpb>>>(out, 1.0001f); });
    const double fl3 = 2.0 * 4 * 64 * IT2 * (double)blocks * tpb;
    printf("LDS.128 + 4 FFMA : %.3f ms  %.1f TFLOP/s\n", ms3, fl3 / ms3 / 1e9);
    printf("LDS.128 + 8 FFMA : %.3f ms  %.1f TFLOP/s\n", ms4, 2 * fl3 / ms4 / 1e9);
    return 0;
}
