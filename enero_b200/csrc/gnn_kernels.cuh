// Message-passing GNN kernels (actor and critic share them) for sm_100a.
//
// Reference semantics: actorPPOmiddR.py:42-69 / criticPPO.py:38-64
//   T x { gather(first), gather(second), concat, Dense(40->20, selu),
//         unsorted_segment_sum(by second), GRUCell(20, reset_after=True) }
//   -> per-graph segment_sum -> Dense20-selu, Dense20-selu, Dense1.
//
// Formulation used here (one thread per link-state row, fp32 FMA pipe):
//   [h_f || h_s] . Wm  ==  h_f . Wm[0:20] + h_s . Wm[20:40]
// so every row publishes A_r = h_r . Wm[0:20] once per iteration (20x20 instead of
// 40x20 per *pair*), and the receiving row s adds its own B_s = h_s . Wm[20:40] + bm:
//   agg_s = sum_{p : second_p = s, ascending p} selu(A[first_p] + B_s)
// The ascending-p order of the stable CSR reproduces TF-CPU's sequential
// unsorted_segment_sum accumulation order; no atomics anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "enero_internal.h"

namespace enero {

constexpr int H = 20;
constexpr int MP_TPB = 128;

// tf.nn.selu constants (tensorflow/core/kernels/relu_op_functor.h)
__device__ __forceinline__ float selu_f(float x) {
    const float scale = 1.0507009873554804934193349852946f;
    const float scale_alpha = 1.7580993408473768599402175208123f;
    return x < 0.f ? scale_alpha * (expf(x) - 1.f) : scale * x;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// acc[j] += sum_k x[k] * W[k*LD + j], W in shared memory (warp-uniform address ->
// one broadcast LDS.128 feeds four FMAs)
template <int LD>
__device__ __forceinline__ void mv20(float (&acc)[H], const float (&x)[H], const float* __restrict__ W) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + k * LD + j);
            acc[j + 0] = fmaf(x[k], w.x, acc[j + 0]);
            acc[j + 1] = fmaf(x[k], w.y, acc[j + 1]);
            acc[j + 2] = fmaf(x[k], w.z, acc[j + 2]);
            acc[j + 3] = fmaf(x[k], w.w, acc[j + 3]);
        }
    }
}

__device__ __forceinline__ void load_row(float (&v)[H], const float* __restrict__ p) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < H / 4; ++i) {
        const float4 t = q[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[H]) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < H / 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// ---- row-space indexers -------------------------------------------------------
struct RowRef {            // what k_mp_iter needs to know about one link-state row
    bool ok;               // live row
    int pb, pe;            // its incoming pairs: list[pb..pe)
    const int* list;       // source rows (before `add`)
    int64_t add;
    unsigned b, base;      // IdxTopo only: decision, and `add` within the decision (add = b * RPD + base)
};
// Explicit: arbitrary first/second of the signature-compatible path, as a CSR by
// destination row built on device (stable in pair order).
struct IdxExplicit {
    int64_t n;
    const int* seg_off;   // [n+1]
    const int* src;       // [p] source rows in CSR order
    __device__ __forceinline__ bool valid(int64_t r) const { return r < n; }
    __device__ __forceinline__ bool in_range(int64_t r) const { return r < n; }
    __host__ __device__ __forceinline__ int64_t rows() const { return n; }
    __device__ __forceinline__ bool tile_live(int64_t r0, int) const { return r0 < n; }
    __device__ __forceinline__ void seg(int64_t r, int& b, int& e, const int*& list, int64_t& add) const {
        b = seg_off[r]; e = seg_off[r + 1]; list = src; add = 0;
    }
    __device__ __forceinline__ void ref(int64_t r, RowRef& o) const {
        o.ok = r < n; o.pb = o.pe = 0; o.list = src; o.add = 0; o.b = o.base = 0;
        if (o.ok) { o.pb = seg_off[r]; o.pe = seg_off[r + 1]; }
    }
};
// Implicit: B decisions of one topology, decision b owns padded rows [b*RPD, (b+1)*RPD),
// of which nK[b]*E are live.  Row lr of a decision receives the pairs of local link
// x = lr % off_s from graph k = lr / off_s, whose sources sit at in_first + k*off_f:
// exactly the reference's old_cummax offsets (script_eval_on_single_topology.py:58-64,
// 110-117), including the case max(index)+1 < E.
struct IdxTopo {
    int64_t total;
    int E, off_f, off_s, RPD;
    const int* nK;        // [B]
    const int* in_off;    // [E+1]
    const int* in_first;  // [P]
    __device__ __forceinline__ bool in_range(int64_t r) const { return r < total; }
    __host__ __device__ __forceinline__ int64_t rows() const { return total; }
    // does the row range [r0, r0 + n) hold a live row of any decision?
    __device__ __forceinline__ bool tile_live(int64_t r0, int n) const {
        const int64_t r1 = r0 + n;                         // exclusive
        for (int64_t b = r0 / RPD; b * RPD < r1; ++b) {
            const int64_t lo = b * RPD, hi = lo + (int64_t)nK[b] * E;      // live rows of decision b
            if (hi > r0 && lo < r1 && hi > lo) return true;
        }
        return false;
    }
    __device__ __forceinline__ bool valid(int64_t r) const {
        if (r >= total) return false;
        const int b = (int)(r / RPD);
        return (int)(r - (int64_t)b * RPD) < nK[b] * E;
    }
    __device__ __forceinline__ void seg(int64_t r, int& bg, int& en, const int*& list, int64_t& add) const {
        const int b = (int)(r / RPD);
        const int lr = (int)(r - (int64_t)b * RPD);
        bg = en = 0; list = in_first; add = 0;
        if (off_s > 0) {
            const int k = lr / off_s, x = lr - k * off_s;
            if (k < nK[b]) { bg = in_off[x]; en = in_off[x + 1]; add = (int64_t)b * RPD + (int64_t)k * off_f; }
        }
    }
    // valid() + seg() in one go with 32-bit arithmetic (the host guarantees total < 2^31 for this path;
    // the 64-bit divisions of the separate calls cost ~7 % of the kernel, profiles/r01_v3_*)
    __device__ __forceinline__ void ref(int64_t r64, RowRef& o) const {
        o.ok = false; o.pb = o.pe = 0; o.list = in_first; o.add = 0; o.b = o.base = 0;
        if (r64 >= total) return;
        const unsigned r = (unsigned)r64;
        const unsigned b = r / (unsigned)RPD;
        const unsigned lr = r - b * (unsigned)RPD;
        const int nk = nK[b];
        if (lr >= (unsigned)(nk * E)) return;
        o.ok = true;
        if (off_s > 0) {
            const unsigned k = lr / (unsigned)off_s, x = lr - k * (unsigned)off_s;
            if ((int)k < nk) {
                o.pb = in_off[x]; o.pe = in_off[x + 1]; o.b = b; o.base = k * (unsigned)off_f;
                o.add = (int64_t)(b * (unsigned)RPD + k * (unsigned)off_f);
            }
        }
    }
};

// ---- packed fp32x2 helpers (Blackwell FFMA2: two FMAs per lane per issue slot) ---------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// c + x * (w.lo, w.hi): ptxas folds the {x, x} pack into FFMA2's scalar-broadcast operand form
__device__ __forceinline__ u64 fma2s(float x, u64 w, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pack2(x, x)), "l"(w), "l"(c));
    return d;
}

__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void prefetch_row_l1(const float* p) {
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p + H - 1));
}

// ---- 1-D TMA (cp.async.bulk) and mbarrier helpers ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// fast transcendental forms (MUFU.EX2 / MUFU.RCP); relative error ~1e-6, see DESIGN.md
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float selu_fast(float x) {
    const float scale = 1.0507009873554804934193349852946f;
    const float scale_alpha = 1.7580993408473768599402175208123f;
    const float e = ex2_approx(x * LOG2E);
    return x < 0.f ? fmaf(scale_alpha, e, -scale_alpha) : scale * x;
}
// selu_fast on two lanes: the packed add/mul/fma round each lane exactly like the scalar forms
__device__ __forceinline__ u64 selu_fast2(u64 x) {
    const float scale = 1.0507009873554804934193349852946f;
    const float scale_alpha = 1.7580993408473768599402175208123f;
    float xl, xh, ml, mh, nl, nh, pl, ph;
    unpack2(x, xl, xh);
    unpack2(mul2(x, pack2(LOG2E, LOG2E)), ml, mh);
    unpack2(fma2(pack2(ex2_approx(ml), ex2_approx(mh)), pack2(scale_alpha, scale_alpha), pack2(-scale_alpha, -scale_alpha)), nl, nh);
    unpack2(mul2(x, pack2(scale, scale)), pl, ph);
    return pack2(xl < 0.f ? nl : pl, xh < 0.f ? nh : ph);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.f + ex2_approx(-LOG2E * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(-2.f, rcp_approx(1.f + ex2_approx((2.f * LOG2E) * x)), 1.f); }

// ---- math policy of the message-passing kernels -----------------------------------------------------
// ACC = false ("fast", the default): MUFU.EX2 / MUFU.RCP forms above, ~1e-6 relative error per call.
// ACC = true  ("accurate"): libdevice expf / tanhf (<= 2 ulp) and IEEE division, the selu / sigmoid expressions
//              written exactly as TensorFlow's functors and the oracle write them (relu_op_functor.h:
//              scale_alpha * (exp(x) - 1); sigmoid = 1 / (1 + exp(-x))).
// Selected per context (enero_ctx_set_math / ENERO_MATH).  Against the fp64 arbiter both modes sit at the fp32
// rounding floor of the sums (profiles/r02_parity_sweep.json); accurate costs 46 % throughput.
template <bool ACC> struct Mth;
template <> struct Mth<false> {
    static __device__ __forceinline__ float selu(float x) { return selu_fast(x); }
    static __device__ __forceinline__ float sigmoid(float x) { return sigmoid_fast(x); }
    static __device__ __forceinline__ float tanh_(float x) { return tanh_fast(x); }
    static __device__ __forceinline__ float selu_grad(float u) {
        return u < 0.f ? 1.7580993408473768599402175208123f * ex2_approx(u * LOG2E) : 1.0507009873554804934193349852946f;
    }
    static __device__ __forceinline__ u64 selu2(u64 x) { return selu_fast2(x); }
};
template <> struct Mth<true> {
    static __device__ __forceinline__ float selu(float x) { return selu_f(x); }
    static __device__ __forceinline__ float sigmoid(float x) { return sigmoid_f(x); }
    static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float selu_grad(float u) {
        return u < 0.f ? 1.7580993408473768599402175208123f * expf(u) : 1.0507009873554804934193349852946f;
    }
    static __device__ __forceinline__ u64 selu2(u64 x) {
        float lo, hi;
        unpack2(x, lo, hi);
        return pack2(selu_f(lo), selu_f(hi));
    }
};

// ---- one message-passing iteration ------------------------------------------------
// Persistent CTAs of 128 threads; every thread owns MP_R = 2 link-state rows of a 256-row tile so
// that each broadcast LDS.128 of weights feeds 8 FMAs (the shared-memory return path, not the FMA
// pipe, is the limiter at 4 FMAs per load: tools/ubench_matvec.cu).  The row vectors (h, agg) stay in
// registers; mat-vecs run as rolled loops over blocks of 4 outputs (k fully unrolled) so the
// instruction footprint stays ~15 KB -- the first version, fully unrolled (~80 KB of straight-line
// code), spent 45 % of its issue slots waiting on instruction fetch (profiles/r01_*).
// Outputs of a block (dynamic index) go through per-thread shared-memory slots with an 80-byte
// stride, which is bank-conflict-free for LDS/STS.128.
constexpr int MP_R = 2;
constexpr int MP_ROWS = MP_TPB * MP_R;                       // rows per tile
constexpr int MP_SMEM_FLOATS = W_MP_FLOATS + 2 * MP_ROWS * H;  // weights + h slots + B/z/A' slots
constexpr size_t MP_SMEM_BYTES = sizeof(float) * MP_SMEM_FLOATS + 8 * (MP_TPB / 32);   // + one mbarrier per warp

// ---- where the (warp-uniform) message + GRU weights are read from -----------------------------------------
// WSmem: staged per CTA in shared memory, every read a broadcast LDS.128.  (The __constant__ bank was
// measured and dropped, DESIGN.md section 3: with a warp-uniform index ptxas emits LDCU.64 + uniform-register
// FFMA2 operands, which frees the LSU pipe but runs at half the speed -- 224 k vs 407 k decisions/s -- and
// per-thread LDC is 5x slower.)
struct WSmem {
    const float* sw;
    __device__ __forceinline__ ulonglong2 ld2(int off) const { return *reinterpret_cast<const ulonglong2*>(sw + off); }
    __device__ __forceinline__ float4 ld4(int off) const { return *reinterpret_cast<const float4*>(sw + off); }
};
// ---- the FMA phases of one message-passing iteration, shared by k_mp_iter and k_mp_tile ---------------
// B = bm + h . Wm[20:40]  ->  myb slots
template <int KH, class W>
__device__ __forceinline__ void mp_phase_B(const W sw, const float (&h)[MP_R][H], float* const (&myb)[MP_R]) {
#pragma unroll 1
    for (int jb = 0; jb < H / 4; ++jb) {
        // accumulators start at zero and the bias is added after the dot product (the op order of
        // Dense = matmul + BiasAdd); an accumulator initialised from a constant-bank value would also keep
        // ptxas from using uniform-register weight operands for the whole loop
        u64 acc[MP_R][2];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) { acc[r][0] = 0ull; acc[r][1] = 0ull; }
#pragma unroll
        for (int k = 0; k < KH; ++k) {
            const ulonglong2 w = sw.ld2(W_MSG_K + H * H + k * H + 4 * jb);
#pragma unroll
            for (int r = 0; r < MP_R; ++r) { acc[r][0] = fma2s(h[r][k], w.x, acc[r][0]); acc[r][1] = fma2s(h[r][k], w.y, acc[r][1]); }
        }
        const ulonglong2 b0 = sw.ld2(W_MSG_B + 4 * jb);
#pragma unroll
        for (int r = 0; r < MP_R; ++r) *reinterpret_cast<ulonglong2*>(myb[r] + 4 * jb) = make_ulonglong2(add2(acc[r][0], b0.x), add2(acc[r][1], b0.y));
    }
}

// agg += selu(a + B) for one gathered operand row (packed f32x2 lanes, rounding identical to the scalar form)
template <bool ACC>
__device__ __forceinline__ void mp_accumulate(u64 (&agg2)[H / 2], const float (&a)[H], const u64 (&t2)[H / 2]) {
#pragma unroll
    for (int j = 0; j < H / 2; ++j)
        agg2[j] = add2(agg2[j], Mth<ACC>::selu2(add2(pack2(a[2 * j], a[2 * j + 1]), t2[j])));
}
// the row's B vector as packed pairs, read once from its slot before the gather loop
__device__ __forceinline__ void mp_load_B(u64 (&t2)[H / 2], const float* __restrict__ myb) {
#pragma unroll
    for (int jb = 0; jb < H / 4; ++jb) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(myb + 4 * jb);
        t2[2 * jb] = t.x; t2[2 * jb + 1] = t.y;
    }
}

// z = sigmoid(agg.Kz + b0z + h.Rz + b1z)   (GRUCell, reset_after=True)  ->  myb slots
template <int KH, bool ACC, class W>
__device__ __forceinline__ void mp_phase_z(const W sw, const float (&h)[MP_R][H], const float (&agg)[MP_R][H],
                                           float* const (&myb)[MP_R]) {
#pragma unroll 1
    for (int jb = 0; jb < H / 4; ++jb) {
        u64 acc[MP_R][2];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) { acc[r][0] = 0ull; acc[r][1] = 0ull; }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const ulonglong2 wk = sw.ld2(W_GRU_K + k * 60 + 4 * jb);
#pragma unroll
            for (int r = 0; r < MP_R; ++r) { acc[r][0] = fma2s(agg[r][k], wk.x, acc[r][0]); acc[r][1] = fma2s(agg[r][k], wk.y, acc[r][1]); }
            if (k < KH) {
                const ulonglong2 wr = sw.ld2(W_GRU_R + k * 60 + 4 * jb);
#pragma unroll
                for (int r = 0; r < MP_R; ++r) { acc[r][0] = fma2s(h[r][k], wr.x, acc[r][0]); acc[r][1] = fma2s(h[r][k], wr.y, acc[r][1]); }
            }
        }
        const ulonglong2 b0 = sw.ld2(W_GRU_B + 4 * jb);
        const ulonglong2 b1 = sw.ld2(W_GRU_B + 60 + 4 * jb);
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            float v0, v1, v2, v3;
            unpack2(add2(add2(acc[r][0], b0.x), b1.x), v0, v1); unpack2(add2(add2(acc[r][1], b0.y), b1.y), v2, v3);
            *reinterpret_cast<float4*>(myb[r] + 4 * jb) = make_float4(Mth<ACC>::sigmoid(v0), Mth<ACC>::sigmoid(v1), Mth<ACC>::sigmoid(v2), Mth<ACC>::sigmoid(v3));
        }
    }
}

// r, candidate state and the new h, four outputs at a time: new h -> myh slots (and hout[r] + 4*jb if non-null)
template <int KH, bool ACC, class W>
__device__ __forceinline__ void mp_phase_rh(const W sw, const float (&h)[MP_R][H], const float (&agg)[MP_R][H],
                                            float* const (&myh)[MP_R], float* const (&myb)[MP_R], float* const (&hout)[MP_R]) {
#pragma unroll 1
    for (int jb = 0; jb < H / 4; ++jb) {
        u64 ar[MP_R][2], ax[MP_R][2], ay[MP_R][2];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) { ar[r][0] = ar[r][1] = ax[r][0] = ax[r][1] = ay[r][0] = ay[r][1] = 0ull; }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const ulonglong2 kr = sw.ld2(W_GRU_K + k * 60 + 20 + 4 * jb);
            const ulonglong2 kh = sw.ld2(W_GRU_K + k * 60 + 40 + 4 * jb);
#pragma unroll
            for (int r = 0; r < MP_R; ++r) {
                ar[r][0] = fma2s(agg[r][k], kr.x, ar[r][0]); ar[r][1] = fma2s(agg[r][k], kr.y, ar[r][1]);
                ax[r][0] = fma2s(agg[r][k], kh.x, ax[r][0]); ax[r][1] = fma2s(agg[r][k], kh.y, ax[r][1]);
            }
            if (k < KH) {
                const ulonglong2 rr = sw.ld2(W_GRU_R + k * 60 + 20 + 4 * jb);
                const ulonglong2 rh = sw.ld2(W_GRU_R + k * 60 + 40 + 4 * jb);
#pragma unroll
                for (int r = 0; r < MP_R; ++r) {
                    ar[r][0] = fma2s(h[r][k], rr.x, ar[r][0]); ar[r][1] = fma2s(h[r][k], rr.y, ar[r][1]);
                    ay[r][0] = fma2s(h[r][k], rh.x, ay[r][0]); ay[r][1] = fma2s(h[r][k], rh.y, ay[r][1]);
                }
            }
        }
        const ulonglong2 br = sw.ld2(W_GRU_B + 20 + 4 * jb);       // input biases of r and the candidate,
        const ulonglong2 bx = sw.ld2(W_GRU_B + 40 + 4 * jb);
        const ulonglong2 by = sw.ld2(W_GRU_B + 100 + 4 * jb);      // recurrent bias of the candidate,
        const float4 b1r = sw.ld4(W_GRU_B + 80 + 4 * jb);          // recurrent bias of r (added below)
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            float rv[4], xv[4], yv[4];
            unpack2(add2(ar[r][0], br.x), rv[0], rv[1]); unpack2(add2(ar[r][1], br.y), rv[2], rv[3]);
            unpack2(add2(ax[r][0], bx.x), xv[0], xv[1]); unpack2(add2(ax[r][1], bx.y), xv[2], xv[3]);
            unpack2(add2(ay[r][0], by.x), yv[0], yv[1]); unpack2(add2(ay[r][1], by.y), yv[2], yv[3]);
            const float4 z = *reinterpret_cast<const float4*>(myb[r] + 4 * jb);
            const float4 ho = *reinterpret_cast<const float4*>(myh[r] + 4 * jb);
            const float b1[4] = {b1r.x, b1r.y, b1r.z, b1r.w};
            const float zz[4] = {z.x, z.y, z.z, z.w};
            const float hh0[4] = {ho.x, ho.y, ho.z, ho.w};
            float hn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float rg = Mth<ACC>::sigmoid(rv[i] + b1[i]);
                const float cand = Mth<ACC>::tanh_(xv[i] + rg * yv[i]);
                hn[i] = zz[i] * hh0[i] + (1.f - zz[i]) * cand;
            }
            const float4 o = make_float4(hn[0], hn[1], hn[2], hn[3]);
            *reinterpret_cast<float4*>(myh[r] + 4 * jb) = o;
            if (hout[r]) *reinterpret_cast<float4*>(hout[r] + 4 * jb) = o;
        }
    }
}

// next iteration's message operand A' = h' . Wm[0:20]  ->  aout[r] + 4*jb (skipped if null)
template <class W>
__device__ __forceinline__ void mp_phase_A(const W sw, float* const (&myh)[MP_R], float* const (&aout)[MP_R]) {
    float h[MP_R][H];
#pragma unroll
    for (int r = 0; r < MP_R; ++r) load_row(h[r], myh[r]);
#pragma unroll 1
    for (int jb = 0; jb < H / 4; ++jb) {
        u64 acc[MP_R][2];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) { acc[r][0] = 0ull; acc[r][1] = 0ull; }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const ulonglong2 w = sw.ld2(W_MSG_K + k * H + 4 * jb);
#pragma unroll
            for (int r = 0; r < MP_R; ++r) { acc[r][0] = fma2s(h[r][k], w.x, acc[r][0]); acc[r][1] = fma2s(h[r][k], w.y, acc[r][1]); }
        }
#pragma unroll
        for (int r = 0; r < MP_R; ++r)
            if (aout[r]) *reinterpret_cast<ulonglong2*>(aout[r] + 4 * jb) = make_ulonglong2(acc[r][0], acc[r][1]);
    }
}

// KH = number of leading link-state columns that can be non-zero on entry: 20 in general, 3 for the
// first iteration of the fused decision path, whose input rows are [util, cap, mark, 0 x 17]
// (every h-side mat-vec then needs 3 instead of 20 k-steps; adding exact zeros changes no bit).
//
// Every WARP is its own pipeline over 64-row tiles handed out by a global counter: no block-wide barrier
// after the weights are staged.  Block barriers put the four warps of a CTA in the same phase at the same
// time; measured on B200, three barriers per tile cost 6-8 % because the FMA-dense phases of one warp
// then no longer overlap the gather / transcendental phases of its neighbours.
//
// TMA = true: the slot of row i of a warp's tile is wsh + 20*i, i.e. the slots ARE the tile's rows in
// global order, so the tile's link states arrive with one bulk asynchronous copy (cp.async.bulk, 1-D TMA,
// issued by lane 0, awaited on the warp's own mbarrier) and the new link states leave with one bulk
// store; A' goes through the B/z slots and leaves with a second bulk store.  No 80-byte-stride
// LDG/STG (~20 L1 wavefronts each) except the operand gathers, whose rows were prefetched into L1.
// TMA = false: per-thread LDG/STG for everything.
constexpr int MP_WROWS = 32 * MP_R;          // rows per warp tile
constexpr int MP_WARPS = MP_TPB / 32;

template <class IDX, bool LAST, int KH, bool ACC>
__global__ void __launch_bounds__(MP_TPB, 3)
k_mp_iter(const IDX idx, const float* __restrict__ wflat, const float* __restrict__ h_in,
          const float* __restrict__ A_in, float* __restrict__ h_out, float* __restrict__ A_out,
          int64_t ntiles, int* __restrict__ tile_counter) {
    constexpr bool TMA = true;                 // per-thread LDG/STG variant: measured within 1 %, kept for reference
    constexpr int NW = MP_WARPS;               // warps per CTA
    extern __shared__ __align__(16) float mp_smem[];
    float* sw_s = mp_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* slots = mp_smem + W_MP_FLOATS;
    float* wsh = slots + warp * (MP_WROWS * H);                                      // h (old, then new) of the warp's tile
    float* wsb = slots + NW * (MP_WROWS * H) + warp * (MP_WROWS * H);                // B, then z, then A'
    uint64_t* bar = reinterpret_cast<uint64_t*>(slots + 2 * NW * (MP_WROWS * H)) + warp;
    for (int i = threadIdx.x; i < W_MP_FLOATS / 4; i += MP_TPB)
        reinterpret_cast<float4*>(sw_s)[i] = reinterpret_cast<const float4*>(wflat)[i];
    WSmem sw; sw.sw = sw_s;
    if (TMA && lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t total = idx.rows();
    float* myh[MP_R]; float* myb[MP_R];
#pragma unroll
    for (int r = 0; r < MP_R; ++r) { myh[r] = wsh + (r * 32 + lane) * H; myb[r] = wsb + (r * 32 + lane) * H; }
    uint32_t phase = 0u;

    int64_t tile = (int64_t)blockIdx.x * NW + warp;
    while (tile < ntiles) {
        int nxt = 0;
        if (lane == 0) nxt = (int)gridDim.x * NW + atomicAdd(tile_counter, 1);
        const int64_t next = __shfl_sync(0xffffffffu, nxt, 0);
        const int64_t row0 = tile * MP_WROWS;
        const int nrows = (int)min((int64_t)MP_WROWS, total - row0);
        int64_t row[MP_R]; bool ok[MP_R]; RowRef ref[MP_R];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) { row[r] = row0 + r * 32 + lane; idx.ref(row[r], ref[r]); ok[r] = ref[r].ok; }
        bool any_ok = false;
#pragma unroll
        for (int r = 0; r < MP_R; ++r) any_ok |= ok[r];
        if (!__any_sync(0xffffffffu, any_ok)) { tile = next; continue; }
        if (TMA && lane == 0) {
            bulk_wait_read0();                               // this lane's previous bulk store has left the slots
            mbar_expect_tx(bar, (uint32_t)nrows * H * 4);
            bulk_g2s(wsh, h_in + row0 * H, (uint32_t)nrows * H * 4, bar);
        }
        // pull exactly the message operands this thread will gather into L1 while the copy lands and the B
        // mat-vec runs (prefetch.global.L1 takes no register and never blocks)
#pragma unroll
        for (int r = 0; r < MP_R; ++r)
            for (int p = ref[r].pb; p < ref[r].pe; ++p) prefetch_row_l1(A_in + ((int64_t)ref[r].list[p] + ref[r].add) * H);
        float h[MP_R][H];
        u64 agg2[MP_R][H / 2];
        if (TMA) { mbar_wait(bar, phase); phase ^= 1u; }
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            if (ok[r]) load_row(h[r], TMA ? myh[r] : h_in + row[r] * H);
            else {
#pragma unroll
                for (int j = 0; j < H; ++j) h[r][j] = 0.f;
            }
            if (!TMA || !ok[r]) store_row(myh[r], h[r]);
#pragma unroll
            for (int j = 0; j < H / 2; ++j) agg2[r][j] = 0ull;
        }
        // ---- B = bm + h . Wm[20:40]
        mp_phase_B<KH>(sw, h, myb);
        // ---- agg = sum over incoming pairs (ascending pair index) of selu(A[first] + B)
        // The two rows of the thread walk their pair lists in lock step so that two independent gathers
        // are in flight per step (each row still adds its pairs in ascending order); the SELU arithmetic
        // runs on packed f32x2 lanes (half the issue slots of the scalar form).
        {
            int steps = 0;
#pragma unroll
            for (int r = 0; r < MP_R; ++r) steps = max(steps, ref[r].pe - ref[r].pb);
            u64 t2[MP_R][H / 2];
#pragma unroll
            for (int r = 0; r < MP_R; ++r) mp_load_B(t2[r], myb[r]);
            for (int s = 0; s < steps; ++s) {
                float a[MP_R][H];
                bool on[MP_R];
#pragma unroll
                for (int r = 0; r < MP_R; ++r) {
                    on[r] = ref[r].pb + s < ref[r].pe;
                    if (on[r]) load_row(a[r], A_in + ((int64_t)ref[r].list[ref[r].pb + s] + ref[r].add) * H);
                }
#pragma unroll
                for (int r = 0; r < MP_R; ++r) {
                    if (!on[r]) continue;
                    mp_accumulate<ACC>(agg2[r], a[r], t2[r]);
                }
            }
        }
        // h was held in registers for the B mat-vec only; reload it for the GRU so the gather loop above
        // does not have to keep 40 more registers alive
        float agg[MP_R][H];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            load_row(h[r], myh[r]);
#pragma unroll
            for (int j = 0; j < H / 2; ++j) unpack2(agg2[r][j], agg[r][2 * j], agg[r][2 * j + 1]);
        }
        // ---- z = sigmoid(agg.Kz + b0z + h.Rz + b1z)   (GRUCell, reset_after=True)
        mp_phase_z<KH, ACC>(sw, h, agg, myb);
        // ---- r, candidate state and the new h; then A' = h' . Wm[0:20]
        float* hout[MP_R]; float* aout[MP_R];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            hout[r] = (!TMA && ok[r]) ? h_out + row[r] * H : nullptr;
            aout[r] = TMA ? myb[r] : ((ok[r] && !LAST) ? A_out + row[r] * H : nullptr);
        }
        mp_phase_rh<KH, ACC>(sw, h, agg, myh, myb, hout);
        if (!LAST) mp_phase_A(sw, myh, aout);
        if (TMA) {
            fence_proxy_async();                             // slot writes visible to the bulk store
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(h_out + row0 * H, wsh, (uint32_t)nrows * H * 4);
                if (!LAST) bulk_s2g(A_out + row0 * H, wsb, (uint32_t)nrows * H * 4);
                bulk_commit();
            }
            __syncwarp();
        }
        tile = next;
    }
    if (TMA && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
}

// A_0 = h_0 . Wm[0:20] for the signature-compatible path (arbitrary link_state)
__global__ void __launch_bounds__(MP_TPB)
k_transform0(const float* __restrict__ wflat, const float* __restrict__ h_in, float* __restrict__ A_out, int64_t n) {
    __shared__ __align__(16) float sw[H * H];
    for (int i = threadIdx.x; i < H * H / 4; i += MP_TPB)
        reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wflat + W_MSG_K)[i];
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * MP_TPB + threadIdx.x; r < n; r += (int64_t)gridDim.x * MP_TPB) {
        asm volatile("" ::: "memory");
        float h[H], t[H];
        load_row(h, h_in + r * H);
#pragma unroll
        for (int j = 0; j < H; ++j) t[j] = 0.f;
        mv20<H>(t, h, sw);
        store_row(A_out + r * H, t);
    }
}

// ---- readout: per-graph sum of link states + MLP 20-20-1 (one warp per graph) ----
struct GraphsExplicit {            // graph g = rows [g_off[g], g_off[g+1])
    int G; const int* g_off;
    __device__ __forceinline__ bool range(int g, int64_t& r0, int64_t& r1, int64_t& out) const {
        if (g >= G) return false;
        r0 = g_off[g]; r1 = g_off[g + 1]; out = g; return true;
    }
};
struct GraphsTopo {                // graph g = candidate k of decision b (padded to Kmax)
    int B, Kmax, E, RPD; const int* nK;
    __device__ __forceinline__ bool range(int g, int64_t& r0, int64_t& r1, int64_t& out) const {
        const int b = g / Kmax, k = g - b * Kmax;
        if (b >= B || k >= nK[b]) return false;
        r0 = (int64_t)b * RPD + (int64_t)k * E; r1 = r0 + E; out = g; return true;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-graph sum of the link states, rows in ASCENDING order per column -- the order of tf.math.segment_sum /
// reduce_sum(axis=0) on the CPU (Eigen reduces the outer dimension sequentially per output packet) and of the
// oracle's np.add.at: lane j < 20 owns column j and adds h[r][j] for r = r0, r0+1, ...; one row is one coalesced
// 80-byte read of the warp, eight rows are in flight at a time.  Every lane gets all 20 sums.
__device__ __forceinline__ void graph_row_sum(float (&acc)[H], const float* __restrict__ h, int64_t r0, int64_t r1, int lane) {
    const int j = lane < H ? lane : 0;
    float s = 0.f;
    int64_t r = r0;
    for (; r + 8 <= r1; r += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = h[(r + i) * H + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
    }
    for (; r < r1; ++r) s += h[r * H + j];
#pragma unroll
    for (int k = 0; k < H; ++k) acc[k] = __shfl_sync(0xffffffffu, s, k);
}

// sr: the readout weights W_R1_K..W_R3_B staged in shared memory; every lane returns the graph's scalar
__device__ __forceinline__ float readout_graph(const float* sr, const float* __restrict__ h, int64_t r0, int64_t r1, int lane) {
    float acc[H];
    graph_row_sum(acc, h, r0, r1, lane);
    const float* R1 = sr; const float* B1 = sr + (W_R1_B - W_R1_K);
    const float* R2 = sr + (W_R2_K - W_R1_K); const float* B2 = sr + (W_R2_B - W_R1_K);
    const float* R3 = sr + (W_R3_K - W_R1_K); const float* B3 = sr + (W_R3_B - W_R1_K);
    const int j = lane < H ? lane : 0;
    float a1 = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) a1 = fmaf(acc[k], R1[k * H + j], a1);
    a1 = selu_f(a1 + B1[j]);
    float a2 = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) a2 = fmaf(__shfl_sync(0xffffffffu, a1, k), R2[k * H + j], a2);
    a2 = selu_f(a2 + B2[j]);
    float y = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) y = fmaf(__shfl_sync(0xffffffffu, a2, k), R3[k], y);
    return y + B3[0];
}

template <class GR>
__global__ void __launch_bounds__(128)
k_readout(const GR gr, const float* __restrict__ wflat, const float* __restrict__ h, float* __restrict__ out) {
    __shared__ float sr[W_R3_B + 1 - W_R1_K];
    for (int i = threadIdx.x; i <= W_R3_B - W_R1_K; i += blockDim.x) sr[i] = wflat[W_R1_K + i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t r0, r1, o;
    if (!gr.range(g, r0, r1, o)) return;
    const float y = readout_graph(sr, h, r0, r1, lane);
    if (lane == 0) out[o] = y;
}

// ---- soft-max over the K' logits of each decision + greedy action --------------------
// tf.nn.softmax (script_eval_on_single_topology.py:125-126): exp(x - max) * (1 / sum);
// np.argmax (script :177): first maximum of the fp32 probabilities.  One warp; every lane returns the action.
__device__ __forceinline__ int policy_warp(const float* x, int n, int Kmax, float* probs, int lane) {
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += expf(x[i] - m);
    s = warp_sum(s);
    const float inv = 1.f / s;
    float best = -1.f; int bi = 0x7fffffff;
    for (int i = lane; i < Kmax; i += 32) {
        const float p = i < n ? expf(x[i] - m) * inv : 0.f;
        if (probs) probs[i] = p;
        if (i < n && p > best) { best = p; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    return bi;
}

__global__ void __launch_bounds__(128)
k_policy(int B, int Kmax, const int* __restrict__ nK, const float* __restrict__ logits,
         float* __restrict__ probs, int* __restrict__ action) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int bi = policy_warp(logits + (int64_t)b * Kmax, nK[b], Kmax, probs ? probs + (int64_t)b * Kmax : nullptr, lane);
    if (lane == 0 && action) action[b] = bi;
}

// ---- device CSR build for the signature-compatible path ---------------------------
__global__ void k_csr_count(const int* __restrict__ second, int64_t p, int64_t n, int* __restrict__ cnt, int* __restrict__ err) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = second[i];
        if (s < 0 || s >= n) { *err = 1; continue; }
        atomicAdd(&cnt[s], 1);
    }
}
// single-CTA exclusive scan (the compat path is not the throughput path)
__global__ void __launch_bounds__(1024) k_scan_excl(const int* __restrict__ cnt, int64_t n, int* __restrict__ off) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int v = i < n ? cnt[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const int prefix = carry + (wid ? warp_tot[wid - 1] : 0) + x - v;
        if (i < n) off[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = carry;
}
__global__ void k_csr_fill(const int* __restrict__ second, int64_t p, int64_t n, const int* __restrict__ off,
                           int* __restrict__ cursor, int* __restrict__ slot) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = second[i];
        if (s < 0 || s >= n) continue;
        slot[off[s] + atomicAdd(&cursor[s], 1)] = (int)i;
    }
}
// per-destination insertion sort by pair index (restores the sequential order of
// unsorted_segment_sum whatever order the atomics above landed in), then map to sources
__global__ void k_csr_sort(const int* __restrict__ first, int64_t n, const int* __restrict__ off, int* __restrict__ slot,
                           int* __restrict__ src, int* __restrict__ err) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int b = off[r], e = off[r + 1];
        for (int i = b + 1; i < e; ++i) {
            const int v = slot[i];
            int j = i - 1;
            while (j >= b && slot[j] > v) { slot[j + 1] = slot[j]; --j; }
            slot[j + 1] = v;
        }
        for (int i = b; i < e; ++i) {
            const int f = first[slot[i]];
            if (f < 0 || f >= n) { *err = 1; src[i] = 0; } else src[i] = f;
        }
    }
}
// graph_id (ascending) -> row offsets per graph
__global__ void k_graph_offsets(const int* __restrict__ gid, int64_t n, int G, int* __restrict__ g_off, int* __restrict__ err) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        const int cur = r < n ? gid[r] : G;
        const int prev = r > 0 ? gid[r - 1] : -1;
        if (cur < prev || (r < n && (cur < 0 || cur >= G))) { *err = 2; continue; }
        for (int g = prev + 1; g <= cur; ++g) g_off[g] = (int)r;
    }
}

}  // namespace enero
