// Micro-benchmark of the per-row 20xN mat-vec inner loop: where should the (warp-uniform) weights live?
//   smem  : broadcast LDS.128
//   const : __constant__ operands (constant cache / uniform loads), rolled over output blocks
// R rows per thread, x in registers, outputs stored to shared memory.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int LD = 160;                 // 20 x 160 floats = 12.8 KB of weights, like the GNN's message+GRU part
__constant__ float cW[20 * LD];

template <int R, bool CONST>
__global__ void __launch_bounds__(128) k_mv(const float* __restrict__ gw, float* out, int iters) {
    __shared__ __align__(16) float sw[20 * LD];
    __shared__ float so[128 * 8];
    for (int i = threadIdx.x; i < 20 * LD; i += 128) sw[i] = gw[i];
    __syncthreads();
    float x[R][20];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < 20; ++k) x[r][k] = 1e-3f * (threadIdx.x + r + k);
    float s = 0.f;
    for (int it = 0; it < iters; ++it) {
        for (int jb = 0; jb < LD / 4; ++jb) {
            float acc[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                float4 w;
                if (CONST) w = *reinterpret_cast<const float4*>(&cW[k * LD + 4 * jb]);
                else w = *reinterpret_cast<const float4*>(&sw[k * LD + 4 * jb]);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r][0] = fmaf(x[r][k], w.x, acc[r][0]); acc[r][1] = fmaf(x[r][k], w.y, acc[r][1]);
                    acc[r][2] = fmaf(x[r][k], w.z, acc[r][2]); acc[r][3] = fmaf(x[r][k], w.w, acc[r][3]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) s += acc[r][0] + acc[r][1] + acc[r][2] + acc[r][3];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { if (it & 1) x[r][0] += s * 1e-9f; else x[r][1] += s * 1e-9f; }
    }
    so[threadIdx.x] = s;
    out[blockIdx.x * 128 + threadIdx.x] = so[threadIdx.x];
}
template <class F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float h[20 * LD]; for (int i = 0; i < 20 * LD; ++i) h[i] = 1e-4f * (i % 97);
    cudaMemcpyToSymbol(cW, h, sizeof(h));
    float* gw; cudaMalloc(&gw, sizeof(h)); cudaMemcpy(gw, h, sizeof(h), cudaMemcpyHostToDevice);
    const int iters = 64;
    for (int cps = 1; cps <= 4; cps += (cps == 1 ? 2 : 1)) {
        const int blocks = p.multiProcessorCount * cps;
        float* out; cudaMalloc(&out, sizeof(float) * blocks * 128);
        auto rep = [&](const char* name, int R, float ms) {
            double fl = 2.0 * 20 * LD * R * iters * (double)blocks * 128;
            printf("CTAs/SM %d  %-10s R=%d : %.3f ms  %.1f TFLOP/s\n", cps, name, R, ms, fl / ms / 1e9);
        };
        rep("smem", 1, time_ms([&] { k_mv<1, false><<<blocks, 128>>>(gw, out, iters); }));
        rep("smem", 2, time_ms([&] { k_mv<2, false><<<blocks, 128>>>(gw, out, iters); }));
        rep("smem", 4, time_ms([&] { k_mv<4, false><<<blocks, 128>>>(gw, out, iters); }));
        rep("const", 1, time_ms([&] { k_mv<1, true><<<blocks, 128>>>(gw, out, iters); }));
        rep("const", 2, time_ms([&] { k_mv<2, true><<<blocks, 128>>>(gw, out, iters); }));
        rep("const", 4, time_ms([&] { k_mv<4, true><<<blocks, 128>>>(gw, out, iters); }));
        cudaFree(out);
    }
    return 0;
}
