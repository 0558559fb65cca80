// Probe of the tcgen05 plumbing used by the tensor-core message-passing kernel: D[128 x N] = A[128 x 24] . W[24 x N]
// as 3xTF32 (A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, fp32-grade products) with K-major no-swizzle shared-memory operand
// tiles, TMEM accumulators and tcgen05.ld read-back, checked against fp64 on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu && ./tc_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int KC = 6;            // 16-byte chunks along K (24 tf32)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    return d;                                             // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // F32 acc, TF32 x TF32, K-major both
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D, int split_mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float4* a_hi = reinterpret_cast<float4*>(smem);                        // [KC][128] chunks of 16 B
    float4* a_lo = a_hi + KC * 128;
    float4* w_hi = a_lo + KC * 128;                                        // [KC][N]
    float4* w_lo = w_hi + KC * N;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto split = [&](float x, float& hi, float& lo) {
        if (split_mode == 0) { hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u); lo = x - hi; }
        else {
            uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x)); hi = __uint_as_float(h); lo = x - hi;
            if (split_mode >= 2) { uint32_t l; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(lo)); lo = __uint_as_float(l); }
        }
    };
    // A operand: thread = row
    for (int c = 0; c < KC; ++c) {
        float hi[4], lo[4];
        for (int i = 0; i < 4; ++i) split(A[tid * 24 + 4 * c + i], hi[i], lo[i]);
        a_hi[c * 128 + tid] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        a_lo[c * 128 + tid] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    // B operand: [n][k] chunk-major
    for (int i = tid; i < KC * N; i += 128) {
        const int c = i / N, n = i % N;
        float hi[4], lo[4];
        for (int q = 0; q < 4; ++q) split(W[(4 * c + q) * N + n], hi[q], lo[q]);
        w_hi[c * N + n] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        w_lo[c * N + n] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
        uint32_t acc = 0;
        const float4* aop[3] = {a_hi, a_lo, a_hi};
        const float4* bop[3] = {w_hi, w_hi, w_lo};
        if (split_mode == 3) { aop[0] = a_lo; aop[1] = a_hi; aop[2] = a_hi; bop[0] = w_hi; bop[1] = w_lo; bop[2] = w_hi; }   // small terms first
        for (int t = 0; t < 3; ++t)
            for (int j = 0; j < 3; ++j) {
                const uint64_t ad = make_desc(smem_u32(aop[t] + 2 * j * 128), 128 * 16, 128);
                const uint64_t bd = make_desc(smem_u32(bop[t] + 2 * j * N), N * 16, 128);
                mma_tf32(tm, ad, bd, idesc, acc);
                acc = 1;
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    {   // wait for the MMAs
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(128));
}

template <int N> int run(int split_mode) {
    std::vector<float> A(128 * 24, 0.f), W(24 * N, 0.f), D(128 * N);
    srand(7);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 20; ++k) A[r * 24 + k] = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (int k = 0; k < 20; ++k) for (int n = 0; n < N; ++n) W[k * N + n] = (rand() / (float)RAND_MAX - 0.5f);
    float *dA, *dW, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * KC * 128 + 2 * KC * N) * 16;
    cudaFuncSetAttribute(k_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<N><<<1, 128, smem>>>(dA, dW, dD, split_mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0, err_f32 = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0; float f = 0.f;
            for (int k = 0; k < 24; ++k) { ref += (double)A[r * 24 + k] * W[k * N + n]; f = fmaf(A[r * 24 + k], W[k * N + n], f); }
            max_err = fmax(max_err, fabs(D[r * N + n] - ref)); max_ref = fmax(max_ref, fabs(ref)); err_f32 = fmax(err_f32, fabs(f - ref));
        }
    printf("N=%3d split=%d: max |D - fp64| = %.3e (fp32 fma chain: %.3e), max |ref| = %.3f  D[5][3]=%f\n", N, split_mode, max_err, err_f32, max_ref, D[5 * N + 3]);
    return max_err < 1e-5 ? 0 : 1;
}
int main() {
    int bad = 0;
    bad += run<80>(0); bad += run<80>(1); bad += run<80>(2); bad += run<80>(3); bad += run<48>(2); bad += run<32>(2); bad += run<64>(3);
    printf(bad ? "FAILED\n" : "OK\n");
    return bad;
}
