// Micro-benchmark: shared-memory load throughput per SM on sm_100a for warp-uniform (broadcast) and
// per-lane conflict-free addresses at 32/64/128-bit width.  Reports cycles per warp-level LDS per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_lds ubench_lds.cu
#include <cstdio>
#include <cuda_runtime.h>
// BCAST: 1 = one address per warp, 2 = two addresses (even / odd lanes), 0 = one address per lane
template <int W, int BCAST>
__global__ void __launch_bounds__(128) k(float* out, int iters, long long* cyc) {
    __shared__ __align__(16) float sm[128 * 20 + 4096];
    for (int i = threadIdx.x; i < 128 * 20 + 4096; i += 128) sm[i] = i * 1e-6f;
    __syncthreads();
    float acc = 0.f;
    const int base = BCAST == 1 ? 0 : (BCAST == 2 ? (threadIdx.x & 1) * 800 : threadIdx.x * (W / 4));   // per-lane: consecutive W-bit words (conflict-free)
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int off = (base + ((it + u) & 63) * 4);
            const unsigned a = (unsigned)__cvta_generic_to_shared(&sm[W == 16 ? (off & ~3) : (W == 8 ? (off & ~1) : off)]);
            if (W == 16) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); acc += x + y + z + w; }
            else if (W == 8) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); acc += x + y; }
            else { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); acc += x; }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int W, int B> void run(const char* name, int sms, int cps) {
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * sms * cps * 128); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<W, B><<<sms * cps, 128>>>(out, iters, cyc); cudaDeviceSynchronize();
    k<W, B><<<sms * cps, 128>>>(out, iters, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_lds = (double)iters * 16 * 4 * cps;      // warp-level LDS per SM
    printf("%-28s CTAs/SM %d : %.2f cycles per warp-LDS per SM\n", name, cps, h / warp_lds);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int cps = 1; cps <= 3; cps += 2) {
        run<4, 1>("LDS.32  broadcast", p.multiProcessorCount, cps);
        run<8, 1>("LDS.64  broadcast", p.multiProcessorCount, cps);
        run<16, 1>("LDS.128 broadcast", p.multiProcessorCount, cps);
        run<8, 2>("LDS.64  two addresses", p.multiProcessorCount, cps);
        run<16, 2>("LDS.128 two addresses", p.multiProcessorCount, cps);
        run<4, 0>("LDS.32  per-lane", p.multiProcessorCount, cps);
        run<8, 0>("LDS.64  per-lane", p.multiProcessorCount, cps);
        run<16, 0>("LDS.128 per-lane", p.multiProcessorCount, cps);
    }
    return 0;
}
