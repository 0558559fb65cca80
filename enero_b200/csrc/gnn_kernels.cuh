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
// Explicit: arbitrary first/second of the signature-compatible path, as a CSR by
// destination row built on device (stable in pair order).
struct IdxExplicit {
    int64_t n;
    const int* seg_off;   // [n+1]
    const int* src;       // [p] source rows in CSR order
    __device__ __forceinline__ bool valid(int64_t r) const { return r < n; }
    __device__ __forceinline__ void seg(int64_t r, int& b, int& e, const int*& list, int64_t& add) const {
        b = seg_off[r]; e = seg_off[r + 1]; list = src; add = 0;
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
};

// ---- one message-passing iteration ------------------------------------------------
// Persistent CTAs: weights staged once in shared memory, then a grid-stride loop over
// 128-row tiles.  h_in/A_in are read, h_out/A_out written (ping-pong across launches).
template <class IDX, bool LAST>
__global__ void __launch_bounds__(MP_TPB)
k_mp_iter(const IDX idx, const float* __restrict__ wflat, const float* __restrict__ h_in,
          const float* __restrict__ A_in, float* __restrict__ h_out, float* __restrict__ A_out,
          int64_t ntiles) {
    __shared__ __align__(16) float sw[W_MP_FLOATS];
    for (int i = threadIdx.x; i < W_MP_FLOATS / 4; i += MP_TPB)
        reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wflat)[i];
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // keep ptxas from hoisting the (loop-invariant) weight LDS out of the tile loop into registers
        asm volatile("" ::: "memory");
        const int64_t r = tile * MP_TPB + threadIdx.x;
        if (!idx.valid(r)) continue;
        float h[H], t[H], agg[H];
        load_row(h, h_in + r * H);
        // B_r = bm + h . Wm[20:40]
#pragma unroll
        for (int j = 0; j < H; ++j) { t[j] = sw[W_MSG_B + j]; agg[j] = 0.f; }
        mv20<H>(t, h, sw + W_MSG_K + H * H);
        int pb, pe; const int* list; int64_t add;
        idx.seg(r, pb, pe, list, add);
        for (int p = pb; p < pe; ++p) {
            float a[H];
            load_row(a, A_in + ((int64_t)list[p] + add) * H);
#pragma unroll
            for (int j = 0; j < H; ++j) agg[j] += selu_f(a[j] + t[j]);
        }
        // GRUCell, reset_after=True (Keras recurrent_v2 GRUCell.call, implementation 2)
        float z[H];
#pragma unroll
        for (int j = 0; j < H; ++j) z[j] = sw[W_GRU_B + j];
        mv20<60>(z, agg, sw + W_GRU_K);
        mv20<60>(z, h, sw + W_GRU_R);
#pragma unroll
        for (int j = 0; j < H; ++j) z[j] = sigmoid_f(z[j] + sw[W_GRU_B + 60 + j]);
#pragma unroll
        for (int j = 0; j < H; ++j) t[j] = sw[W_GRU_B + 20 + j];
        mv20<60>(t, agg, sw + W_GRU_K + 20);
        mv20<60>(t, h, sw + W_GRU_R + 20);
#pragma unroll
        for (int j = 0; j < H; ++j) t[j] = sigmoid_f(t[j] + sw[W_GRU_B + 80 + j]);   // r
        float y[H];
#pragma unroll
        for (int j = 0; j < H; ++j) y[j] = sw[W_GRU_B + 100 + j];
        mv20<60>(y, h, sw + W_GRU_R + 40);
#pragma unroll
        for (int j = 0; j < H; ++j) { y[j] *= t[j]; t[j] = sw[W_GRU_B + 40 + j]; }
        mv20<60>(t, agg, sw + W_GRU_K + 40);
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float hh = tanhf(t[j] + y[j]);
            h[j] = z[j] * h[j] + (1.f - z[j]) * hh;
        }
        store_row(h_out + r * H, h);
        if (!LAST) {
#pragma unroll
            for (int j = 0; j < H; ++j) t[j] = 0.f;
            mv20<H>(t, h, sw + W_MSG_K);
            store_row(A_out + r * H, t);
        }
    }
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
    float acc[H];
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = 0.f;
    for (int64_t r = r0 + lane; r < r1; r += 32) {
        float v[H];
        load_row(v, h + r * H);
#pragma unroll
        for (int j = 0; j < H; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = warp_sum(acc[j]);
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
    if (lane == 0) out[o] = y + B3[0];
}

// ---- soft-max over the K' logits of each decision + greedy action --------------------
// tf.nn.softmax (script_eval_on_single_topology.py:125-126): exp(x - max) * (1 / sum);
// np.argmax (script :177): first maximum of the fp32 probabilities.
__global__ void __launch_bounds__(128)
k_policy(int B, int Kmax, const int* __restrict__ nK, const float* __restrict__ logits,
         float* __restrict__ probs, int* __restrict__ action) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int n = nK[b];
    const float* x = logits + (int64_t)b * Kmax;
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
        if (probs) probs[(int64_t)b * Kmax + i] = p;
        if (i < n && p > best) { best = p; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
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
