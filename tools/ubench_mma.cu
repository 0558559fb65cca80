// Micro-benchmark: legacy warp-level mma.sync TF32 (m16n8k8) issue rate on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma ubench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHAINS>
__global__ void __launch_bounds__(128) k(float* out, int iters, long long* cyc) {
    float d[CHAINS][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1e-3f * (threadIdx.x + i));
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(1e-3f * (threadIdx.x + 7 * i));
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma_tf32(d[c], a, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 4 * 128); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    for (int cps = 1; cps <= 4; ++cps) {
        k<8><<<p.multiProcessorCount * cps, 128>>>(out, iters, cyc); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<8><<<p.multiProcessorCount * cps, 128>>>(out, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double mmas_per_smsp = (double)iters * 8 * cps;       // one warp per SMSP per CTA
        const double tflops = 2.0 * 16 * 8 * 8 * iters * 8.0 * 4 * cps * p.multiProcessorCount / (ms * 1e-3) / 1e12;
        printf("CTAs/SM %d: %.2f cycles per MMA per SMSP, %.1f dense TF32 TFLOP/s\n", cps, h / mmas_per_smsp, tflops);
    }
    return 0;
}
