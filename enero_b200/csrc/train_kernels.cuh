// Backward kernels of the message-passing GNN and the fused PPO loss / clip / Adam step.
//
// Reference: train_Enero_3top_script.py:264-324 (_critic_step, _actor_step, _train_step_combined:
// GradientTape over the actor and critic, clip_by_global_norm(0.5), Adam.apply_gradients).
//
// Forward of a training minibatch keeps h_t and A_t = h_t.Wm[0:20] of every iteration (11 row
// buffers).  Backward runs t = T-1 .. 0 with two row-parallel kernels per iteration:
//   k_bwd_gates : recompute B, agg, the GRU gates of iteration t; back-propagate dh_{t+1} through the
//                 GRU; emit dagg, B, the partial dh_t and the gate-gradient rows used for weight grads
//   k_bwd_msgs  : back-propagate through selu and the two gathers.  dB_i sums over the pairs that
//                 arrive at row i (CSR by `second`), dA_i over the pairs that leave row i (CSR by `first`) --
//                 the transposes of the forward gathers, again without atomics.
// Weight gradients are tall-skinny reductions X^T.D over all rows and iterations: k_wgrad stages
// 128-row chunks in shared memory, every thread owns ~13 of the 3340 outputs, persistent CTAs write
// one partial vector each and k_reduce_partials adds them in CTA order (deterministic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "enero_internal.h"
#include "env_kernels.cuh"
#include "gnn_kernels.cuh"

namespace enero {

constexpr int BW_TPB = 128;
constexpr float SELU_S = 1.0507009873554804934193349852946f;
constexpr float SELU_SA = 1.7580993408473768599402175208123f;

// out-pairs of a row (CSR by `first`), mirroring IdxTopo::seg
struct OutTopo {
    int E, off_f, off_s, RPD;
    const int* nK; const int* out_off; const int* out_second;
    __device__ __forceinline__ void seg(int64_t r, int& bg, int& en, int64_t& add) const {
        const int b = (int)(r / RPD);
        const int lr = (int)(r - (int64_t)b * RPD);
        bg = en = 0; add = 0;
        if (off_f > 0) {
            const int k = lr / off_f, x = lr - k * off_f;
            if (k < nK[b]) { bg = out_off[x]; en = out_off[x + 1]; add = (int64_t)b * RPD + (int64_t)k * off_s; }
        }
    }
};

// y[j] (+)= sum_k x[k] * W[k*LD + j]   (forward orientation, 20 outputs)
template <int LD>
__device__ __forceinline__ void mv_fwd(float (&acc)[H], const float (&x)[H], const float* __restrict__ W) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + k * LD + j);
            acc[j] = fmaf(x[k], w.x, acc[j]); acc[j + 1] = fmaf(x[k], w.y, acc[j + 1]);
            acc[j + 2] = fmaf(x[k], w.z, acc[j + 2]); acc[j + 3] = fmaf(x[k], w.w, acc[j + 3]);
        }
    }
}
// y[k] += sum_j d[j] * W[k*LD + j]     (transposed orientation)
template <int LD>
__device__ __forceinline__ void mv_bwd(float (&acc)[H], const float (&d)[H], const float* __restrict__ W) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + k * LD + j);
            s0 = fmaf(d[j], w.x, s0); s1 = fmaf(d[j + 1], w.y, s1);
            s0 = fmaf(d[j + 2], w.z, s0); s1 = fmaf(d[j + 3], w.w, s1);
        }
        acc[k] += s0 + s1;
    }
}

// ---- GRU backward of one iteration ------------------------------------------------------------------
// ht = h_t, At = A_t (forward saves); dhn = dL/dh_{t+1}.  Writes agg_t, gate-gradient rows
// G = [dpz | dpr | dx | dy], B_t, dagg and the GRU part of dL/dh_t.
// Mat-vecs run as rolled loops over blocks of four outputs with the reduction index fully unrolled
// (same code-size discipline as k_mp_iter: a fully unrolled version is ~5000 FMAs of straight-line
// code that ptxas spills to a 16 KB stack frame); vectors whose block index is dynamic go through
// per-thread shared-memory slots (80-byte stride: conflict-free LDS/STS.128).
constexpr int BWG_SMEM_FLOATS = W_MP_FLOATS + 4 * BW_TPB * H;
constexpr size_t BWG_SMEM_BYTES = sizeof(float) * BWG_SMEM_FLOATS;

__device__ __forceinline__ float4 f4_fma(float x, const float4 w, float4 a) {
    a.x = fmaf(x, w.x, a.x); a.y = fmaf(x, w.y, a.y); a.z = fmaf(x, w.z, a.z); a.w = fmaf(x, w.w, a.w);
    return a;
}
__device__ __forceinline__ float f4_dot(const float (&d)[H], int j, const float4 w, float s) {
    s = fmaf(d[j], w.x, s); s = fmaf(d[j + 1], w.y, s); s = fmaf(d[j + 2], w.z, s); s = fmaf(d[j + 3], w.w, s);
    return s;
}

template <bool ACC>
__global__ void __launch_bounds__(BW_TPB)
k_bwd_gates(const IdxTopo idx, const float* __restrict__ wflat, const float* __restrict__ ht,
            const float* __restrict__ At, const float* __restrict__ dhn, float* __restrict__ agg_out,
            float* __restrict__ G_out, float* __restrict__ B_out, float* __restrict__ dagg_out,
            float* __restrict__ dh_part) {
    extern __shared__ __align__(16) float bwg_smem[];
    float* sw = bwg_smem;
    for (int i = threadIdx.x; i < W_MP_FLOATS / 4; i += BW_TPB)
        reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wflat)[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * BW_TPB + threadIdx.x;
    if (!idx.valid(r)) return;
    float* slot[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) slot[q] = sw + W_MP_FLOATS + (q * BW_TPB + threadIdx.x) * H;
    float h[H], agg[H];
    load_row(h, ht + r * H);
    // ---- B = bm + h . Wm[20:40]
#pragma unroll 1
    for (int jb = 0; jb < H / 4; ++jb) {
        float4 acc = *reinterpret_cast<const float4*>(sw + W_MSG_B + 4 * jb);
#pragma unroll
        for (int k = 0; k < H; ++k) acc = f4_fma(h[k], *reinterpret_cast<const float4*>(sw + W_MSG_K + H * H + k * H + 4 * jb), acc);
        *reinterpret_cast<float4*>(slot[0] + 4 * jb) = acc;
        *reinterpret_cast<float4*>(B_out + r * H + 4 * jb) = acc;
    }
    {
        float t[H];
        load_row(t, slot[0]);
#pragma unroll
        for (int j = 0; j < H; ++j) agg[j] = 0.f;
        int pb, pe; const int* list; int64_t add;
        idx.seg(r, pb, pe, list, add);
        for (int p = pb; p < pe; ++p) {
            float a[H];
            load_row(a, At + ((int64_t)list[p] + add) * H);
#pragma unroll
            for (int j = 0; j < H; ++j) agg[j] += Mth<ACC>::selu(a[j] + t[j]);
        }
    }
    store_row(agg_out + r * H, agg);
    // ---- gates of four outputs at a time and their local derivatives
#pragma unroll 1
    for (int jb = 0; jb < H / 4; ++jb) {
        float4 z = *reinterpret_cast<const float4*>(sw + W_GRU_B + 4 * jb);
        float4 rg = *reinterpret_cast<const float4*>(sw + W_GRU_B + 20 + 4 * jb);
        float4 c = *reinterpret_cast<const float4*>(sw + W_GRU_B + 40 + 4 * jb);
        float4 y = *reinterpret_cast<const float4*>(sw + W_GRU_B + 100 + 4 * jb);
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const float* K = sw + W_GRU_K + k * 60 + 4 * jb;
            const float* R = sw + W_GRU_R + k * 60 + 4 * jb;
            z = f4_fma(agg[k], *reinterpret_cast<const float4*>(K), z);
            z = f4_fma(h[k], *reinterpret_cast<const float4*>(R), z);
            rg = f4_fma(agg[k], *reinterpret_cast<const float4*>(K + 20), rg);
            rg = f4_fma(h[k], *reinterpret_cast<const float4*>(R + 20), rg);
            c = f4_fma(agg[k], *reinterpret_cast<const float4*>(K + 40), c);
            y = f4_fma(h[k], *reinterpret_cast<const float4*>(R + 40), y);
        }
        const float4 b1z = *reinterpret_cast<const float4*>(sw + W_GRU_B + 60 + 4 * jb);
        const float4 b1r = *reinterpret_cast<const float4*>(sw + W_GRU_B + 80 + 4 * jb);
        const float4 g4 = *reinterpret_cast<const float4*>(dhn + r * H + 4 * jb);
        const float4 h4 = *reinterpret_cast<const float4*>(ht + r * H + 4 * jb);
        const float zv[4] = {z.x + b1z.x, z.y + b1z.y, z.z + b1z.z, z.w + b1z.w};
        const float rv[4] = {rg.x + b1r.x, rg.y + b1r.y, rg.z + b1r.z, rg.w + b1r.w};
        const float cv[4] = {c.x, c.y, c.z, c.w}, yv[4] = {y.x, y.y, y.z, y.w};
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, hv[4] = {h4.x, h4.y, h4.z, h4.w};
        float dpz[4], dpr[4], dx[4], dy[4], dhz[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float zz = Mth<ACC>::sigmoid(zv[i]);
            const float rr = Mth<ACC>::sigmoid(rv[i]);
            const float cc = Mth<ACC>::tanh_(cv[i] + rr * yv[i]);
            const float dz = gv[i] * (hv[i] - cc);
            const float dpc = gv[i] * (1.f - zz) * (1.f - cc * cc);
            dx[i] = dpc;
            dy[i] = dpc * rr;
            dpr[i] = dpc * yv[i] * rr * (1.f - rr);
            dpz[i] = dz * zz * (1.f - zz);
            dhz[i] = gv[i] * zz;
        }
        const float4 o0 = make_float4(dpz[0], dpz[1], dpz[2], dpz[3]), o1 = make_float4(dpr[0], dpr[1], dpr[2], dpr[3]);
        const float4 o2 = make_float4(dx[0], dx[1], dx[2], dx[3]), o3 = make_float4(dy[0], dy[1], dy[2], dy[3]);
        *reinterpret_cast<float4*>(slot[0] + 4 * jb) = o0; *reinterpret_cast<float4*>(slot[1] + 4 * jb) = o1;
        *reinterpret_cast<float4*>(slot[2] + 4 * jb) = o2; *reinterpret_cast<float4*>(slot[3] + 4 * jb) = o3;
        float* G = G_out + r * 4 * H + 4 * jb;
        *reinterpret_cast<float4*>(G) = o0; *reinterpret_cast<float4*>(G + H) = o1;
        *reinterpret_cast<float4*>(G + 2 * H) = o2; *reinterpret_cast<float4*>(G + 3 * H) = o3;
        *reinterpret_cast<float4*>(dh_part + r * H + 4 * jb) = make_float4(dhz[0], dhz[1], dhz[2], dhz[3]);
    }
    // ---- dagg = [dpz dpr dx] . K^T ; dh += [dpz dpr dy] . R^T, four rows of K / R at a time
    float dpz[H], dpr[H], dx[H], dy[H];
    load_row(dpz, slot[0]); load_row(dpr, slot[1]); load_row(dx, slot[2]); load_row(dy, slot[3]);
#pragma unroll 1
    for (int kb = 0; kb < H / 4; ++kb) {
        float da[4], dhh[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float* K = sw + W_GRU_K + (4 * kb + kk) * 60;
            const float* R = sw + W_GRU_R + (4 * kb + kk) * 60;
            float s0 = 0.f, s1 = 0.f, u0 = 0.f, u1 = 0.f;
#pragma unroll
            for (int j = 0; j < H; j += 4) {
                s0 = f4_dot(dpz, j, *reinterpret_cast<const float4*>(K + j), s0);
                s1 = f4_dot(dpr, j, *reinterpret_cast<const float4*>(K + 20 + j), s1);
                s0 = f4_dot(dx, j, *reinterpret_cast<const float4*>(K + 40 + j), s0);
                u0 = f4_dot(dpz, j, *reinterpret_cast<const float4*>(R + j), u0);
                u1 = f4_dot(dpr, j, *reinterpret_cast<const float4*>(R + 20 + j), u1);
                u0 = f4_dot(dy, j, *reinterpret_cast<const float4*>(R + 40 + j), u0);
            }
            da[kk] = s0 + s1; dhh[kk] = u0 + u1;
        }
        *reinterpret_cast<float4*>(dagg_out + r * H + 4 * kb) = make_float4(da[0], da[1], da[2], da[3]);
        float4 prev = *reinterpret_cast<float4*>(dh_part + r * H + 4 * kb);
        prev.x += dhh[0]; prev.y += dhh[1]; prev.z += dhh[2]; prev.w += dhh[3];
        *reinterpret_cast<float4*>(dh_part + r * H + 4 * kb) = prev;
    }
}

// ---- message backward of one iteration -----------------------------------------------------------------
template <bool NEED_DH, bool ACC>
__global__ void __launch_bounds__(BW_TPB)
k_bwd_msgs(const IdxTopo idx, const OutTopo out, const float* __restrict__ wflat, const float* __restrict__ At,
           const float* __restrict__ B_in, const float* __restrict__ dagg_in, const float* __restrict__ dh_part,
           float* __restrict__ dB_out, float* __restrict__ dA_out, float* __restrict__ dh_prev) {
    __shared__ __align__(16) float sw[2 * H * H];
    for (int i = threadIdx.x; i < 2 * H * H / 4; i += BW_TPB)
        reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wflat + W_MSG_K)[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * BW_TPB + threadIdx.x;
    if (!idx.valid(r)) return;
    float B[H], dagg[H], dB[H], dA[H];
    load_row(B, B_in + r * H);
    load_row(dagg, dagg_in + r * H);
#pragma unroll
    for (int j = 0; j < H; ++j) { dB[j] = 0.f; dA[j] = 0.f; }
    {
        int pb, pe; const int* list; int64_t add;
        idx.seg(r, pb, pe, list, add);
        for (int p = pb; p < pe; ++p) {
            float a[H];
            load_row(a, At + ((int64_t)list[p] + add) * H);
#pragma unroll
            for (int j = 0; j < H; ++j) dB[j] += dagg[j] * Mth<ACC>::selu_grad(a[j] + B[j]);
        }
    }
    {
        float Ai[H];
        load_row(Ai, At + r * H);
        int pb, pe; int64_t add;
        out.seg(r, pb, pe, add);
        for (int p = pb; p < pe; ++p) {
            const int64_t dst = (int64_t)out.out_second[p] + add;
            float bs[H], ds[H];
            load_row(bs, B_in + dst * H);
            load_row(ds, dagg_in + dst * H);
#pragma unroll
            for (int j = 0; j < H; ++j) dA[j] += ds[j] * Mth<ACC>::selu_grad(Ai[j] + bs[j]);
        }
    }
    store_row(dB_out + r * H, dB);
    store_row(dA_out + r * H, dA);
    if (NEED_DH) {
        float dh[H];
        load_row(dh, dh_part + r * H);
        mv_bwd<H>(dh, dB, sw + H * H);     // B = bm + h.Wm[20:40]
        mv_bwd<H>(dh, dA, sw);             // A = h.Wm[0:20]
        store_row(dh_prev + r * H, dh);
    }
}

// ---- readout backward (one warp per graph) ------------------------------------------------------------------
// dlogit[graph] -> dL/dh_T rows, and per-graph vectors V = [g | a1 | a2 | dp1 | dp2 | dy] for the weight grads
__global__ void __launch_bounds__(128)
k_readout_bwd(const GraphsTopo gr, const float* __restrict__ wflat, const float* __restrict__ hT,
              const float* __restrict__ dlogit, float* __restrict__ dhT, float* __restrict__ V) {
    __shared__ float sr[W_R3_B + 1 - W_R1_K];
    for (int i = threadIdx.x; i <= W_R3_B - W_R1_K; i += blockDim.x) sr[i] = wflat[W_R1_K + i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t r0, r1, o;
    if (!gr.range(g, r0, r1, o)) return;
    float acc[H];
    graph_row_sum(acc, hT, r0, r1, lane);
    const float* R1 = sr; const float* B1 = sr + (W_R1_B - W_R1_K);
    const float* R2 = sr + (W_R2_K - W_R1_K); const float* B2 = sr + (W_R2_B - W_R1_K);
    const float* R3 = sr + (W_R3_K - W_R1_K);
    const int j = lane < H ? lane : 0;
    float p1 = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) p1 = fmaf(acc[k], R1[k * H + j], p1);
    p1 += B1[j];
    const float a1 = selu_f(p1);
    float p2 = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) p2 = fmaf(__shfl_sync(0xffffffffu, a1, k), R2[k * H + j], p2);
    p2 += B2[j];
    const float a2 = selu_f(p2);
    const float dy = dlogit[o];
    const float sg2 = p2 < 0.f ? SELU_SA * expf(p2) : SELU_S;
    const float dp2 = R3[j] * dy * sg2;                                  // lane j: da2[j] * selu'(p2[j])
    float da1 = 0.f;                                                     // lane j: sum_m R2[j][m] * dp2[m]
#pragma unroll
    for (int m = 0; m < H; ++m) da1 = fmaf(R2[j * H + m], __shfl_sync(0xffffffffu, dp2, m), da1);
    const float sg1 = p1 < 0.f ? SELU_SA * expf(p1) : SELU_S;
    const float dp1 = da1 * sg1;
    float dg = 0.f;
#pragma unroll
    for (int m = 0; m < H; ++m) dg = fmaf(R1[j * H + m], __shfl_sync(0xffffffffu, dp1, m), dg);
    // dL/dh_T of every row of the graph = dg
    float dgv[H];
#pragma unroll
    for (int k = 0; k < H; ++k) dgv[k] = __shfl_sync(0xffffffffu, dg, k);
    for (int64_t r = r0 + lane; r < r1; r += 32) store_row(dhT + r * H, dgv);
    if (lane < H) {
        float* v = V + o * 6 * H;
        v[lane] = acc[lane]; v[H + lane] = a1; v[2 * H + lane] = a2; v[3 * H + lane] = dp1; v[4 * H + lane] = dp2;
        v[5 * H + lane] = dy;
    }
}

// readout weight gradients: eight threads share output o of the 861 readout parameters; thread s of the eight walks
// the graphs g = s, s + 8, ... in ascending order (four in flight) reading the per-graph vectors V = [g | a1 | a2 | dp1 |
// dp2 | dy] straight from global memory, and the eight partial sums meet in a fixed shuffle tree (deterministic).  A
// padded graph (k >= nK) contributes an exact zero.  History: 32 graphs per CTA with two block barriers each, partial
// vectors and a second reduction kernel took 107 us per call; one thread per output still 160 us for an actor lane
// (400 dependent iterations on seven CTAs).
constexpr int RW_NOUT = W_R3_B + 1 - W_R1_K;     // 861
constexpr int RW_TPB = 128, RW_SPLIT = 8, RW_OUT_PER_CTA = RW_TPB / RW_SPLIT;
__global__ void __launch_bounds__(RW_TPB)
k_readout_wgrad(int n_graphs, int Kmax, const int* __restrict__ nK, const float* __restrict__ V, float* __restrict__ grad_out) {
    const int o = blockIdx.x * RW_OUT_PER_CTA + (threadIdx.x / RW_SPLIT), sl = threadIdx.x % RW_SPLIT;
    const bool live = o < RW_NOUT;
    int ia = 0, ib = -1;                          // value = V[ia] * V[ib]  (ib < 0: V[ia] alone)
    if (o < 400) { ia = o / H; ib = 3 * H + o % H; }                              // dR1[k][j] = g[k] * dp1[j]
    else if (o < 420) { ia = 3 * H + o - 400; ib = -1; }                          // db1
    else if (o < 820) { ia = H + (o - 420) / H; ib = 4 * H + (o - 420) % H; }     // dR2 = a1[k] * dp2[j]
    else if (o < 840) { ia = 4 * H + o - 820; ib = -1; }                          // db2
    else if (o < 860) { ia = 2 * H + o - 840; ib = 5 * H; }                       // dR3 = a2[k] * dy
    else if (live) { ia = 5 * H; ib = -1; }                                       // db3
    const int ib2 = ib >= 0 ? ib : ia;
    float acc = 0.f;
    for (int g0 = sl; g0 < n_graphs && live; g0 += 4 * RW_SPLIT) {
        float term[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int g = g0 + i * RW_SPLIT;
            term[i] = 0.f;
            if (g < n_graphs) {
                const int b = g / Kmax, k = g - b * Kmax;
                const float* v = V + (int64_t)g * 6 * H;
                const float x = v[ia], y = v[ib2];
                if (k < nK[b]) term[i] = ib >= 0 ? x * y : x;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc += term[i];
    }
#pragma unroll
    for (int d = RW_SPLIT / 2; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d, RW_SPLIT);
    if (live && sl == 0) grad_out[o] += acc;
}

// ---- message / GRU weight gradients -----------------------------------------------------------------------------
// For every row and iteration: X = [h(20) | agg(20) | 1], D = [dpz dpr dx dy (80) | dB (20) | dA (20)].
// Output o of the 3340 message+GRU parameters is sum over rows of X[xc(o)] * D[dc(o)].
struct WgradIter { const float* h; const float* agg; const float* G; const float* dB; const float* dA; };
constexpr int WG_TPB = 256, WG_ROWS = 64;
constexpr int WG_XP = 48, WG_DP = 128;             // X (41 columns) and D (120 columns) padded to 8x6 and 32x4 thread tiles
constexpr int WG_PART = WG_XP * WG_DP;             // floats of one CTA's partial outer product

__device__ __forceinline__ void wgrad_cols(int o, int& xc, int& dc) {
    if (o < 400) { xc = o / H; dc = 100 + o % H; }                         // dWm[0:20]   = h^T dA
    else if (o < 800) { xc = (o - 400) / H; dc = 80 + (o - 400) % H; }     // dWm[20:40]  = h^T dB
    else if (o < 820) { xc = 40; dc = 80 + (o - 800); }                    // dbm         = sum dB
    else if (o < 2020) { const int q = o - 820; xc = 20 + q / 60; dc = q % 60; }           // dK = agg^T [dpz dpr dx]
    else if (o < 3220) { const int q = o - 2020; const int j = q % 60; xc = q / 60; dc = j < 40 ? j : j + 20; }  // dR = h^T [dpz dpr dy]
    else if (o < 3280) { xc = 40; dc = o - 3220; }                         // db0 = sum [dpz dpr dx]
    else { const int j = o - 3280; xc = 40; dc = j < 40 ? j : j + 20; }    // db1 = sum [dpz dpr dy]
}

// Per CTA: the full outer product sum_rows X^T D (48 x 128, of which the 3340 parameter gradients are a
// subset) over its row tiles, register-tiled 6 x 4 per thread: per row a thread reads 6 X values (warp-uniform
// broadcast) and one float4 of D and issues 24 FMAs.  The previous version looked its 14 scattered
// (x, d) pairs up in shared memory row by row (2 LDS per FMA) and took 36 % of the PPO step.
__global__ void __launch_bounds__(WG_TPB)
k_wgrad(const IdxTopo idx, const WgradIter it, int64_t ntiles, float* __restrict__ partials, int accumulate) {
    __shared__ __align__(16) float sx[WG_ROWS * WG_XP];
    __shared__ __align__(16) float sd[WG_ROWS * WG_DP];
    __shared__ unsigned char s_ok[WG_ROWS];
    const int xg = threadIdx.x >> 5, dg = threadIdx.x & 31;
    float acc[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * WG_ROWS;
        __syncthreads();
        if (threadIdx.x < WG_ROWS) s_ok[threadIdx.x] = idx.valid(r0 + threadIdx.x) ? 1 : 0;
        __syncthreads();
        // stage the tile: 40 float4 per row (h 5 | agg 5 | G 20 | dB 5 | dA 5), ten per thread, all requested before the
        // first is stored (a scalar element loop exposed one load latency per element: 105 us per launch, half of a
        // backward pass)
        constexpr int WG_V4 = 40, WG_LD = WG_ROWS * WG_V4 / WG_TPB;
        float4 v[WG_LD];
#pragma unroll
        for (int u = 0; u < WG_LD; ++u) {
            const int i = threadIdx.x + u * WG_TPB;
            const int row = i / WG_V4, q = i - row * WG_V4;
            const int64_t r = r0 + row;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s_ok[row]) {
                const float* src = q < 5 ? it.h + r * H + 4 * q : q < 10 ? it.agg + r * H + 4 * (q - 5)
                                 : q < 30 ? it.G + r * 4 * H + 4 * (q - 10) : q < 35 ? it.dB + r * H + 4 * (q - 30)
                                 : it.dA + r * H + 4 * (q - 35);
                v[u] = *reinterpret_cast<const float4*>(src);
            }
        }
#pragma unroll
        for (int u = 0; u < WG_LD; ++u) {
            const int i = threadIdx.x + u * WG_TPB;
            const int row = i / WG_V4, q = i - row * WG_V4;
            float* dst = q < 10 ? sx + row * WG_XP + 4 * q : sd + row * WG_DP + 4 * (q - 10);
            *reinterpret_cast<float4*>(dst) = v[u];
        }
        if (threadIdx.x < WG_ROWS) {                      // the constant column of X, the padding of X and D
            const int row = threadIdx.x;
            *reinterpret_cast<float4*>(sx + row * WG_XP + 40) = make_float4(s_ok[row] ? 1.f : 0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(sx + row * WG_XP + 44) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(sd + row * WG_DP + 120) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(sd + row * WG_DP + 124) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int row = 0; row < WG_ROWS; ++row) {
            const float2* px = reinterpret_cast<const float2*>(sx + row * WG_XP + 6 * xg);
            const float2 x01 = px[0], x23 = px[1], x45 = px[2];
            const float4 d = *reinterpret_cast<const float4*>(sd + row * WG_DP + 4 * dg);
            const float x[6] = {x01.x, x01.y, x23.x, x23.y, x45.x, x45.y};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                acc[i][0] = fmaf(x[i], d.x, acc[i][0]); acc[i][1] = fmaf(x[i], d.y, acc[i][1]);
                acc[i][2] = fmaf(x[i], d.z, acc[i][2]); acc[i][3] = fmaf(x[i], d.w, acc[i][3]);
            }
        }
    }
    // the CTA's slot collects the iterations of one backward pass (same grid every iteration, newest iteration first);
    // k_reduce_partials runs once per pass
    float* out = partials + (int64_t)blockIdx.x * WG_PART;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float4* o4 = reinterpret_cast<float4*>(out + (6 * xg + i) * WG_DP + 4 * dg);
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (accumulate) { const float4 p = *o4; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
        *o4 = v;
    }
}

// grad[o] += sum over CTAs (ascending) of the (x, d) entry of parameter o
__global__ void k_reduce_partials(int n_part, const float* __restrict__ partials, float* __restrict__ grad) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= W_MP_FLOATS) return;
    int xc, dc;
    wgrad_cols(o, xc, dc);
    const float* p0 = partials + xc * WG_DP + dc;
    float s = 0.f;
    for (int p = 0; p < n_part; ++p) s += p0[(int64_t)p * WG_PART];
    grad[o] += s;
}

// ---- PPO loss head (one warp per sample) ------------------------------------------------------------------------
// _actor_step (train script :273-291): ratio, clipped surrogate, entropy; gradient wrt the logits of the
// minibatch-mean loss  mean(loss) - beta * mean(entropy).  inv_m = 1 / minibatch size.
__global__ void __launch_bounds__(128)
k_ppo_actor_head(int n, int Kmax, const int* __restrict__ nK, const float* __restrict__ logits,
                 const int* __restrict__ action, const float* __restrict__ old_prob, const float* __restrict__ adv,
                 float inv_m, float beta, float clip, float* __restrict__ dlogit, float* __restrict__ loss_out,
                 float* __restrict__ ent_out) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n) return;
    const int k = nK[b];
    const float* x = logits + (int64_t)b * Kmax;
    float m = -INFINITY;
    for (int i = lane; i < k; i += 32) m = fmaxf(m, x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int i = lane; i < k; i += 32) s += expf(x[i] - m);
    s = warp_sum(s);
    const float inv = 1.f / s, logs = logf(s);
    float ent = 0.f;
    for (int i = lane; i < k; i += 32) {
        const float p = expf(x[i] - m) * inv;
        const float lp = (x[i] - m) - logs;
        ent -= p * lp;
    }
    ent = warp_sum(ent);
    const int a = action[b];
    const float pa = expf(x[a] - m) * inv;
    const float ratio = expf(logf(pa) - logf(old_prob[b]));
    const float A = adv[b];
    const float surr1 = -ratio * A;
    const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
    const float surr2 = -rc * A;
    // tf.maximum sends the gradient to surr1 when surr1 >= surr2; clip_by_value passes it inside [lo, hi]
    float dr;
    if (surr1 >= surr2) dr = -A;
    else dr = (ratio >= 1.f - clip && ratio <= 1.f + clip) ? -A : 0.f;
    for (int i = lane; i < Kmax; i += 32) {
        float g = 0.f;
        if (i < k) {
            const float p = expf(x[i] - m) * inv;
            const float lp = (x[i] - m) - logs;
            g = dr * ratio * ((i == a ? 1.f : 0.f) - p);          // d loss / d logit_i
            g += beta * p * (lp + ent);                           // - beta * d entropy / d logit_i
            g *= inv_m;
        }
        dlogit[(int64_t)b * Kmax + i] = g;
    }
    if (lane == 0) { loss_out[b] = fmaxf(surr1, surr2); ent_out[b] = ent; }
}

// _critic_step (train script :264-271): (ret - V)^2
__global__ void k_ppo_critic_head(int n, const float* __restrict__ value, const float* __restrict__ ret, float inv_m,
                                  float* __restrict__ dvalue, float* __restrict__ loss_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const float d = ret[b] - value[b];
    loss_out[b] = d * d;
    dvalue[b] = -2.f * d * inv_m;
}

// dst += sum(x[0..n)) in a fixed order (one warp: lane-strided partial sums, then the shuffle tree)
// (warp w of the CTA: dst[w] += sum(x_w); x1 may be null with a one-warp launch)
__global__ void k_sum_into(int n, const float* __restrict__ x0, const float* __restrict__ x1, float* __restrict__ dst) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* x = w ? x1 : x0;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += x[i];
    s = warp_sum(s);
    if (lane == 0) dst[w] += s;
}

// ---- clip_by_global_norm + Adam for both models (single CTA) ----------------------------------------------------
// tf.clip_by_global_norm (train script :319): scale = clip * min(1/norm, 1/clip) over all 22 tensors;
// Keras Adam as TF's ApplyAdam kernel: m += (g-m)(1-b1); v += (g^2-v)(1-b2); theta -= m*alpha/(sqrt(v)+eps)
__global__ void __launch_bounds__(1024)
k_clip_adam(float* __restrict__ wa, float* __restrict__ wc, float* __restrict__ ga, float* __restrict__ gc,
            float* __restrict__ ma, float* __restrict__ va, float* __restrict__ mc, float* __restrict__ vc,
            float alpha, float beta1, float beta2, float eps, float clip_norm, float* __restrict__ norm_out,
            const float* __restrict__ sums, float* __restrict__ hist) {
    __shared__ float red[32];
    __shared__ float s_scale;
    float s = 0.f;
    for (int i = threadIdx.x; i < ENERO_NPARAMS; i += 1024) { s = fmaf(ga[i], ga[i], s); s = fmaf(gc[i], gc[i], s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = red[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            const float norm = sqrtf(v);
            s_scale = clip_norm * fminf(1.f / norm, 1.f / clip_norm);
            if (norm_out) *norm_out = norm;
            // per-step record read back once per ppo_update: global norm before clipping and the three loss sums
            if (hist) { hist[0] = norm; hist[1] = sums[0]; hist[2] = sums[1]; hist[3] = sums[2]; }
        }
    }
    __syncthreads();
    const float scale = s_scale;
    for (int i = threadIdx.x; i < 2 * ENERO_NPARAMS; i += 1024) {
        const bool actor = i < ENERO_NPARAMS;
        const int k = actor ? i : i - ENERO_NPARAMS;
        float* w = actor ? wa : wc; float* g = actor ? ga : gc; float* m = actor ? ma : mc; float* v = actor ? va : vc;
        const float gg = g[k] * scale;
        const float mm = m[k] + (gg - m[k]) * (1.f - beta1);
        const float vv = v[k] + (gg * gg - v[k]) * (1.f - beta2);
        m[k] = mm; v[k] = vv;
        w[k] = w[k] - (mm * alpha) / (sqrtf(vv) + eps);
        g[k] = 0.f;
    }
}

// ---- GAE (train script :367-377), fp64 like the reference's python floats; one thread, 835 steps -------------------
__global__ void k_gae(int n, const float* __restrict__ values /*[n+1]*/, const unsigned char* __restrict__ masks,
                      const double* __restrict__ rewards, double gamma, double lmbda,
                      double* __restrict__ returns, double* __restrict__ adv) {
    if (blockIdx.x || threadIdx.x) return;
    double gae = 0.0;
    for (int i = n - 1; i >= 0; --i) {
        const double mk = masks[i] ? 1.0 : 0.0;
        const double delta = rewards[i] + gamma * (double)values[i + 1] * mk - (double)values[i];
        gae = delta + gamma * lmbda * mk * gae;
        returns[i] = gae + (double)values[i];
    }
    double mean = 0.0;
    for (int i = 0; i < n; ++i) { adv[i] = returns[i] - (double)values[i]; mean += adv[i]; }
    mean /= n;
    double var = 0.0;
    for (int i = 0; i < n; ++i) { const double d = adv[i] - mean; var += d * d; }
    const double sd = sqrt(var / n);
    for (int i = 0; i < n; ++i) adv[i] = (adv[i] - mean) / (sd + 1e-10);
}

}  // namespace enero
