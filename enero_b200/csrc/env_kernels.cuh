// Environment-side kernels: link loads from a routing, per-candidate link features,
// the middlepoint env step, and the Local Search.  fp64 link loads are integer-valued
// (demands are floored on ingest, defo_process_results.py:224) so sums are exact and
// order-independent; every utilisation is a correctly-rounded fp64 division, i.e. the
// very numbers the reference's Python produces.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "enero_internal.h"
#include "gnn_kernels.cuh"

namespace enero {

struct TopoView {   // by-value kernel argument
    int N, E, Kmax;
    const double* cap; const float* capf;
    const int* sp_off; const int* sp_edge; const int* mid_off; const int* mids;
};

// ---- per-decision prologue: candidate count + fp32 utilisation feature -----------------
// utilization feature = fp32(load / capacity): fp64 divide then cast
// (script_eval_on_single_topology.py:144).
__device__ __forceinline__ int candidate_count(const TopoView& t, const int* __restrict__ sd, const uint8_t* __restrict__ skip, int b) {
    const int s = sd[2 * b], d = sd[2 * b + 1];
    if ((skip && skip[b]) || s < 0 || d < 0 || s >= t.N || d >= t.N || s == d) return 0;
    return t.mid_off[s * t.N + d + 1] - t.mid_off[s * t.N + d];
}

__global__ void k_decision_prep(TopoView t, int B, const double* __restrict__ loads, const int* __restrict__ sd,
                                const uint8_t* __restrict__ skip, int* __restrict__ nK, float* __restrict__ util) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * t.E) return;
    const int b = (int)(i / t.E), e = (int)(i - (int64_t)b * t.E);
    util[i] = (float)(loads[i] / t.cap[e]);
    if (e == 0) nK[b] = candidate_count(t, sd, skip, b);
}

// Tables of the fused first iteration (k_mp_tc<FUSED0>): bwcap[b][e] = fp32(bw / cap) (the double-rounded mark value of
// mark_action_sp, environment16.py:723) and one bit per (decision, candidate, link): link on path(s,mid) U path(mid,d).
// ptab[b][e][0 / 1][20]: the first message operand A_0 = h_0 . Wm[0:20] of link e of decision b, unmarked / marked, with
// the fma order of k_features (a zero mark feature adds nothing: fma(0, w, x) = x).
// The prologue (util, nK) is part of the same launch.
__global__ void k_decision_tables(TopoView t, int B, int W, const double* __restrict__ loads, const int* __restrict__ sd,
                                  const double* __restrict__ bw, const uint8_t* __restrict__ skip, const float* __restrict__ wflat,
                                  int* __restrict__ nK, float* __restrict__ util, float* __restrict__ bwcap,
                                  uint32_t* __restrict__ mask, float* __restrict__ ptab) {
    __shared__ __align__(16) float sw[3 * H];
    if (threadIdx.x < 3 * H) sw[threadIdx.x] = wflat[W_MSG_K + threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)B * t.E) {
        const int b = (int)(i / t.E), e = (int)(i - (int64_t)b * t.E);
        const double cap = t.cap[e];
        const float f0 = (float)(loads[i] / cap), f1 = t.capf[e], f2 = (float)(bw[b] / cap);
        util[i] = f0;
        bwcap[i] = f2;
        float p0[H], p1[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            p0[j] = fmaf(f1, sw[H + j], fmaf(f0, sw[j], 0.f));
            p1[j] = fmaf(f2, sw[2 * H + j], p0[j]);
        }
        store_row(ptab + i * (2 * H), p0);
        store_row(ptab + i * (2 * H) + H, p1);
        if (e == 0) nK[b] = candidate_count(t, sd, skip, b);
    }
    if (i < (int64_t)B * t.Kmax) {
        const int b = (int)(i / t.Kmax), k = (int)(i - (int64_t)b * t.Kmax);
        uint32_t* m = mask + i * W;
        for (int w = 0; w < W; ++w) m[w] = 0u;
        if (k < candidate_count(t, sd, skip, b)) {
            const int s = sd[2 * b], d = sd[2 * b + 1];
            const int mid = t.mids[t.mid_off[s * t.N + d] + k];
            for (int q = t.sp_off[s * t.N + mid]; q < t.sp_off[s * t.N + mid + 1]; ++q) m[t.sp_edge[q] >> 5] |= 1u << (t.sp_edge[q] & 31);
            if (mid != d)
                for (int q = t.sp_off[mid * t.N + d]; q < t.sp_off[mid * t.N + d + 1]; ++q) m[t.sp_edge[q] >> 5] |= 1u << (t.sp_edge[q] & 31);
        }
    }
}

// ---- candidate features (kernel 6, feature half) ----------------------------------------
// Row (b, k, e) of the batched link_state = [util, cap/maxCap, mark, 0 x 17]
// (mark_action_sp, environment16.py:711-725: assignment bw/cap on every link of
// path(s,mid) U path(mid,d); get_graph_features, script :131-160) and its first
// message operand A_0 = h_0 . Wm[0:20].
__global__ void __launch_bounds__(MP_TPB)
k_features(TopoView t, int B, int RPD, const int* __restrict__ sd, const double* __restrict__ bw,
           const int* __restrict__ nK, const float* __restrict__ util, const float* __restrict__ wflat,
           float* __restrict__ h0, float* __restrict__ A0) {
    __shared__ __align__(16) float sw[3 * H];
    if (threadIdx.x < 3 * H) sw[threadIdx.x] = wflat[W_MSG_K + threadIdx.x];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * MP_TPB + threadIdx.x;
    if (r >= (int64_t)B * RPD) return;
    const int b = (int)(r / RPD);
    const int lr = (int)(r - (int64_t)b * RPD);
    if (lr >= nK[b] * t.E) return;
    const int k = lr / t.E, e = lr - k * t.E;
    const int s = sd[2 * b], d = sd[2 * b + 1];
    const int mid = t.mids[t.mid_off[s * t.N + d] + k];
    bool on = false;
    for (int i = t.sp_off[s * t.N + mid]; i < t.sp_off[s * t.N + mid + 1]; ++i) on |= (t.sp_edge[i] == e);
    if (mid != d)
        for (int i = t.sp_off[mid * t.N + d]; i < t.sp_off[mid * t.N + d + 1]; ++i) on |= (t.sp_edge[i] == e);
    float h[H];
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = 0.f;
    h[0] = util[(int64_t)b * t.E + e];
    h[1] = t.capf[e];
    h[2] = on ? (float)(bw[b] / t.cap[e]) : 0.f;
    store_row(h0 + r * H, h);
    float a[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float v = 0.f;
        v = fmaf(h[0], sw[j], v);
        v = fmaf(h[1], sw[H + j], v);
        v = fmaf(h[2], sw[2 * H + j], v);
        a[j] = v;
    }
    store_row(A0 + r * H, a);
}

// critic features: [util, cap/maxCap, 0 x 18] on one graph per state
// (critic_get_graph_features, train_Enero_3top_script.py:204-230)
__global__ void __launch_bounds__(MP_TPB)
k_features_critic(TopoView t, int B, const float* __restrict__ util, const float* __restrict__ wflat,
                  float* __restrict__ h0, float* __restrict__ A0) {
    __shared__ __align__(16) float sw[2 * H];
    if (threadIdx.x < 2 * H) sw[threadIdx.x] = wflat[W_MSG_K + threadIdx.x];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * MP_TPB + threadIdx.x;
    if (r >= (int64_t)B * t.E) return;
    const int e = (int)(r % t.E);
    float h[H], a[H];
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = 0.f;
    h[0] = util[r];
    h[1] = t.capf[e];
    store_row(h0 + r * H, h);
#pragma unroll
    for (int j = 0; j < H; ++j) a[j] = fmaf(h[1], sw[H + j], fmaf(h[0], sw[j], 0.f));
    store_row(A0 + r * H, a);
}

// ---- warp helpers ------------------------------------------------------------------------
// argmax of load/cap over links with the reference's scan semantics: start from (0,0,0),
// strict '>', links in id order  ->  largest value, lowest id on ties, id -1 if nothing > 0
// (environment16.py:604-612).
__device__ __forceinline__ void warp_max_uti(const double* loads, const double* cap, int E, int lane,
                                             double& best, int& best_e) {
    best = 0.0; best_e = -1;
    for (int e = lane; e < E; e += 32) {
        const double u = loads[e] / cap[e];
        if (u > best) { best = u; best_e = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oe = __shfl_xor_sync(0xffffffffu, best_e, o);
        if (ob > best || (ob == best && oe >= 0 && (best_e < 0 || oe < best_e))) { best = ob; best_e = oe; }
    }
}

// add sign*bw on every link of the shortest path a->b (links of one SP are distinct)
__device__ __forceinline__ void warp_route_add(const TopoView& t, double* loads, int a, int b, double v, int lane) {
    const int i0 = t.sp_off[a * t.N + b], i1 = t.sp_off[a * t.N + b + 1];
    for (int i = i0 + lane; i < i1; i += 32) loads[t.sp_edge[i]] += v;
    __syncwarp();
}
__device__ __forceinline__ void warp_demand_add(const TopoView& t, double* loads, int s, int d, int mid, double v, int lane) {
    if (mid < 0 || mid == d) warp_route_add(t, loads, s, d, v, lane);
    else { warp_route_add(t, loads, s, mid, v, lane); warp_route_add(t, loads, mid, d, v, lane); }
}

// ---- link loads of a routing (kernel 6, load half) -----------------------------------------
// compute_link_utilization_reset (environment16.py:240-245) / reset_DRL_hill_sp
// (environment15.py:504-518): one CTA per episode, loads accumulated in shared memory.
// mid_cur[b][s*N+d] = middlepoint or -1 (direct shortest path).
__global__ void __launch_bounds__(256)
k_route_loads(TopoView t, const double* __restrict__ tm, const int* __restrict__ mid_cur,
              double* __restrict__ loads, double* __restrict__ max_uti, int* __restrict__ max_edge) {
    extern __shared__ double sl[];
    const int b = blockIdx.x, NN = t.N * t.N;
    for (int e = threadIdx.x; e < t.E; e += blockDim.x) sl[e] = 0.0;
    __syncthreads();
    for (int sdi = threadIdx.x; sdi < NN; sdi += blockDim.x) {
        const int s = sdi / t.N, d = sdi - s * t.N;
        if (s == d) continue;
        const double v = tm[(int64_t)b * NN + sdi];
        const int mid = mid_cur ? mid_cur[(int64_t)b * NN + sdi] : -1;
        const int m1 = (mid < 0 || mid == d) ? d : mid;
        for (int i = t.sp_off[s * t.N + m1]; i < t.sp_off[s * t.N + m1 + 1]; ++i) atomicAdd(&sl[t.sp_edge[i]], v);
        if (m1 != d)
            for (int i = t.sp_off[m1 * t.N + d]; i < t.sp_off[m1 * t.N + d + 1]; ++i) atomicAdd(&sl[t.sp_edge[i]], v);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < t.E; e += blockDim.x) loads[(int64_t)b * t.E + e] = sl[e];
    if (threadIdx.x < 32) {
        double best; int be;
        warp_max_uti(sl, t.cap, t.E, threadIdx.x, best, be);
        if (threadIdx.x == 0) { max_uti[b] = best; max_edge[b] = be; }
    }
}

// ---- critical-demand selection (environment16.py:672-674,188-200; environment15.py:537-538,471-483) -----------
// One CTA per episode.  The five links with the largest RAW load (stable: ties keep link-id order); walking them in
// that order, the demands crossing each -- first the un-rerouted ones in row-major (s,d) order (the direct-SP crossing
// list of the link), then the re-routed ones whose route crosses it, in snapshot (insertion) order -- appended if not
// seen yet; stable sort by bandwidth descending; truncation to ceil(N(N-1) * pct).  mid_cur / routing == NULL: no demand
// is re-routed (Env16.reset).  Order-preserving appends go through a warp ballot; the sort is a rank sort (keys
// (bandwidth, position) are distinct).
struct CritView {
    const int* cross_off; const int* cross_sd;      // direct-SP crossing lists per link
    const int* mid_cur;                             // [B][N*N] re-route middlepoint or -1, or NULL
    const int* routing; const int* n_routing; int rstride;   // [B][rstride][3] snapshot order, or NULL
    int lim, cap;                                   // truncation, row stride of elig
    int* scratch;                                   // [B][N*N] candidate list before sorting
};
__global__ void __launch_bounds__(256)
k_critical_demands(TopoView t, CritView cv, const double* __restrict__ loads, const double* __restrict__ tm,
                   int* __restrict__ elig, int* __restrict__ elig_len) {
    extern __shared__ unsigned char crit_smem[];
    double* sl = reinterpret_cast<double*>(crit_smem);                 // [E] loads (consumed by the top-5 selection)
    uint32_t* seen = reinterpret_cast<uint32_t*>(sl + t.E);            // [ceil(N*N/32)]
    __shared__ int s_top[5]; __shared__ int s_n;
    const int b = blockIdx.x, NN = t.N * t.N, tid = threadIdx.x, lane = tid & 31;
    const double* ld = loads + (int64_t)b * t.E;
    const double* tmb = tm + (int64_t)b * NN;
    const int* mid = cv.mid_cur ? cv.mid_cur + (int64_t)b * NN : nullptr;
    int* lst = cv.scratch + (int64_t)b * NN;
    for (int e = tid; e < t.E; e += blockDim.x) sl[e] = ld[e];
    for (int i = tid; i < (NN + 31) / 32; i += blockDim.x) seen[i] = 0u;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int ncrit = min(5, t.E);
    if (tid < 32) {                                                    // sorted(..., key=load, reverse=True)[:5], stable
        for (int c = 0; c < ncrit; ++c) {
            double bv = -1.0; int be = 0x7fffffff;
            for (int e = lane; e < t.E; e += 32) { const double v = sl[e]; if (v > bv) { bv = v; be = e; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oe = __shfl_xor_sync(0xffffffffu, be, o);
                if (ov > bv || (ov == bv && oe < be)) { bv = ov; be = oe; }
            }
            if (lane == 0) { s_top[c] = be; sl[be] = -2.0; }
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid < 32) {                                                    // one warp builds the list in reference order
        int n = 0;
        for (int c = 0; c < ncrit; ++c) {
            const int e = s_top[c];
            const int i0 = cv.cross_off[e], i1 = cv.cross_off[e + 1];
            for (int base = i0; base < i1; base += 32) {               // entries of one crossing list are distinct
                const int i = base + lane;
                int sd = -1; bool keep = false;
                if (i < i1) {
                    sd = cv.cross_sd[i];
                    keep = !(mid && mid[sd] >= 0) && !((seen[sd >> 5] >> (sd & 31)) & 1u);
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) { lst[n + __popc(m & ((1u << lane) - 1u))] = sd; atomicOr(&seen[sd >> 5], 1u << (sd & 31)); }
                n += __popc(m);
                __syncwarp();
            }
            if (cv.routing) {                                          // re-routed demands crossing e, snapshot order
                const int nr = cv.n_routing[b];
                const int* rt = cv.routing + (int64_t)b * cv.rstride * 3;
                for (int base = 0; base < nr; base += 32) {
                    const int i = base + lane;
                    int sd = -1; bool keep = false;
                    if (i < nr) {
                        const int s = rt[3 * i], d = rt[3 * i + 1], mm = rt[3 * i + 2];
                        sd = s * t.N + d;
                        if (mm != d && mid[sd] == mm && !((seen[sd >> 5] >> (sd & 31)) & 1u)) {
                            bool on = false;
                            for (int q = t.sp_off[s * t.N + mm]; q < t.sp_off[s * t.N + mm + 1]; ++q) on |= t.sp_edge[q] == e;
                            for (int q = t.sp_off[mm * t.N + d]; q < t.sp_off[mm * t.N + d + 1]; ++q) on |= t.sp_edge[q] == e;
                            keep = on;
                        }
                    }
                    // the same (s,d) may appear twice in a snapshot: the first occurrence within this batch wins
                    for (int o = 0; o < 32; ++o) {
                        const int osd = __shfl_sync(0xffffffffu, sd, o);
                        const bool ok = __shfl_sync(0xffffffffu, keep, o);
                        if (o < lane && ok && osd == sd) keep = false;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    if (keep) { lst[n + __popc(m & ((1u << lane) - 1u))] = sd; atomicOr(&seen[sd >> 5], 1u << (sd & 31)); }
                    n += __popc(m);
                    __syncwarp();
                }
            }
        }
        if (lane == 0) s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    __threadfence_block();
    // sorted(key=bw, reverse=True), stable: rank = #{j : bw_j > bw_i or (bw_j == bw_i and j < i)}
    for (int i = tid; i < n; i += blockDim.x) {
        const int sdi = lst[i];
        const double v = tmb[sdi];
        int rank = 0;
        for (int j = 0; j < n; ++j) { const double o = tmb[lst[j]]; rank += (o > v || (o == v && j < i)) ? 1 : 0; }
        if (rank < cv.lim) elig[(int64_t)b * cv.cap + rank] = sdi;
    }
    if (tid == 0) elig_len[b] = min(n, cv.lim);
}

// ---- episode state ------------------------------------------------------------------------
struct EpisodeView {
    int B, Dmax;
    const double* tm;     // [B][N*N]
    double* loads;        // [B][E]
    int* mid_cur;         // [B][N*N]   sp_middlepoints (-1 = none)
    const int* elig;      // [B][Dmax]  list_eligible_demands as s*N+d
    const int* elig_len;  // [B]
    int* it;              // [B] iter_list_elig_demn
    int* cur;             // [B][2] current demand (s,d)
    double* cur_bw;       // [B]
    uint8_t* done;        // [B]
    int* steps;           // [B] decisions taken
    double* max_uti;      // [B] edgeMaxUti[2]
    int* max_edge;        // [B]
    double* best; int* best_step; double* ospf;      // [B]
    int* h_action; double* h_reward; double* h_max;  // [B][Dmax]
    double* last_reward;  // [B]
};

// first demand of the episode: _obtain_demand + decrease_links_utilization_sp
// (environment16.py:682-688)
__global__ void k_episode_begin(TopoView t, EpisodeView ep) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= ep.B) return;
    const int NN = t.N * t.N;
    const int sdi = ep.elig[(int64_t)b * ep.Dmax];
    const int s = sdi / t.N, d = sdi - s * t.N;
    const double v = ep.tm[(int64_t)b * NN + sdi];
    warp_route_add(t, ep.loads + (int64_t)b * t.E, s, d, -v, lane);
    if (lane == 0) {
        ep.it[b] = 1; ep.cur[2 * b] = s; ep.cur[2 * b + 1] = d; ep.cur_bw[b] = v;
        ep.done[b] = 0; ep.steps[b] = 0;
        ep.best[b] = ep.max_uti[b]; ep.ospf[b] = ep.max_uti[b]; ep.best_step[b] = -1;
    }
}

// Env16.step (environment16.py:582-640), one warp per episode.
__device__ __forceinline__ void env_step_warp(const TopoView& t, const EpisodeView& ep, int b, int a, int lane) {
    const int NN = t.N * t.N;
    double* loads = ep.loads + (int64_t)b * t.E;
    int* mid_cur = ep.mid_cur + (int64_t)b * NN;
    const int s = ep.cur[2 * b], d = ep.cur[2 * b + 1];
    const double bw = ep.cur_bw[b];
    const int mid = t.mids[t.mid_off[s * t.N + d] + a];
    warp_demand_add(t, loads, s, d, mid, bw, lane);
    if (mid != d && lane == 0) mid_cur[s * t.N + d] = mid;
    __syncwarp();
    double nu; int ne;
    warp_max_uti(loads, t.cap, t.E, lane, nu, ne);
    const double old = ep.max_uti[b];
    const int step = ep.steps[b];
    // np.around(x, 2) == rint(x * 100) / 100 in fp64
    const double reward = rint((old - nu) * 10.0 * 100.0) / 100.0;
    int it = ep.it[b];
    int ns, nd; bool fin = false;
    if (it < ep.elig_len[b]) {
        const int sdi = ep.elig[(int64_t)b * ep.Dmax + it];
        ns = sdi / t.N; nd = sdi - ns * t.N; ++it;
    } else { ns = 1; nd = 2; fin = true; }       // dummy demand of the finished episode (:622-626)
    const double nbw = ep.tm[(int64_t)b * NN + ns * t.N + nd];
    const int m2 = mid_cur[ns * t.N + nd];
    __syncwarp();
    warp_demand_add(t, loads, ns, nd, m2, -nbw, lane);
    if (lane == 0) {
        // the snapshot the eval driver copies after step() no longer holds the next demand's entry
        if (m2 >= 0) mid_cur[ns * t.N + nd] = -1;
        ep.max_uti[b] = nu; ep.max_edge[b] = ne;
        ep.h_action[(int64_t)b * ep.Dmax + step] = a;
        ep.h_reward[(int64_t)b * ep.Dmax + step] = reward;
        ep.h_max[(int64_t)b * ep.Dmax + step] = nu;
        ep.last_reward[b] = reward;
        if (nu < ep.best[b]) { ep.best[b] = nu; ep.best_step[b] = step; }
        ep.it[b] = it; ep.cur[2 * b] = ns; ep.cur[2 * b + 1] = nd; ep.cur_bw[b] = nbw;
        ep.steps[b] = step + 1; ep.done[b] = fin ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128)
k_env_step(TopoView t, EpisodeView ep, const int* __restrict__ action) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= ep.B || ep.done[b]) return;
    env_step_warp(t, ep, b, action[b], lane);
}

// Tail of one decision in one launch (one CTA per decision): read-out of the K' candidate graphs (a warp each),
// soft-max + arg-max over their logits, and -- STEP -- Env16.step with the greedy action.  The arithmetic is that of
// k_readout / k_policy / k_env_step, which stay for the callers that need the stages apart.
constexpr int FIN_TPB = 512;
template <bool STEP>
__global__ void __launch_bounds__(FIN_TPB)
k_decision_finish(GraphsTopo gr, const float* __restrict__ wflat, const float* __restrict__ h, float* logits,
                  float* __restrict__ probs, int* __restrict__ action, TopoView t, EpisodeView ep) {
    __shared__ float sr[W_R3_B + 1 - W_R1_K];
    const int b = blockIdx.x;
    const int n = gr.nK[b];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* x = logits + (int64_t)b * gr.Kmax;
    if (n > 0) {
        for (int i = threadIdx.x; i <= W_R3_B - W_R1_K; i += blockDim.x) sr[i] = wflat[W_R1_K + i];
        __syncthreads();
        for (int k = warp; k < gr.Kmax; k += blockDim.x >> 5) {
            float y = 0.f;
            if (k < n) {
                const int64_t r0 = (int64_t)b * gr.RPD + (int64_t)k * gr.E;
                y = readout_graph(sr, h, r0, r0 + gr.E, lane);
            }
            if (lane == 0) x[k] = y;
        }
        __syncthreads();                        // the CTA's logits are visible to warp 0
    } else {
        for (int k = threadIdx.x; k < gr.Kmax; k += blockDim.x) x[k] = 0.f;
    }
    if (warp != 0) return;
    const int bi = policy_warp(x, n, gr.Kmax, probs ? probs + (int64_t)b * gr.Kmax : nullptr, lane);
    if (lane == 0 && action) action[b] = bi;
    if (STEP && !ep.done[b]) env_step_warp(t, ep, b, bi, lane);
}

// ---- sampled-action rollouts (train_Enero_3top_script.py:495-506) ---------------------------------------------
// Philox4x32-10 counter-based generator: one uniform in [0, 1) per (seed, episode, step), independent of launch
// geometry and of how the episodes are spread over GPUs.
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0], p1 = (uint64_t)0xCD9E8D57u * ctr[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
        ctr[1] = (uint32_t)p1; ctr[3] = (uint32_t)p0; ctr[0] = n0; ctr[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct RolloutRec {      // per (episode, step) records of one sampled rollout, [B][Dmax] each (loads: [B][Dmax][E])
    double* loads; int* sd; double* bw; int* action; float* prob; float* value;
};

// action = np.random.choice(K', p = probs) (train script :502) with a device generator: inverse CDF of the fp32
// probabilities accumulated in fp64; also records what the PPO buffer keeps of the decision (the state the actor
// saw, the action, its probability, the critic's value).  One warp per episode.
__global__ void __launch_bounds__(128)
k_rollout_pick(TopoView t, EpisodeView ep, int Kmax, const float* __restrict__ probs, const int* __restrict__ nK,
               const float* __restrict__ values, uint64_t seed, RolloutRec rec, int* __restrict__ action) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= ep.B || ep.done[b]) return;
    const int step = ep.steps[b], k = nK[b];
    const float* p = probs + (int64_t)b * Kmax;
    int a = k - 1;
    if (lane == 0) {
        uint32_t ctr[4] = {(uint32_t)b, (uint32_t)step, 0u, 0u};
        philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u = ((double)ctr[0] * 4294967296.0 + (double)ctr[1]) * (1.0 / 18446744073709551616.0);
        double tot = 0.0;
        for (int i = 0; i < k; ++i) tot += (double)p[i];
        double c = 0.0;
        for (int i = 0; i < k; ++i) { c += (double)p[i] / tot; if (u < c) { a = i; break; } }
    }
    a = __shfl_sync(0xffffffffu, a, 0);
    const int64_t slot = (int64_t)b * ep.Dmax + step;
    for (int e = lane; e < t.E; e += 32) rec.loads[slot * t.E + e] = ep.loads[(int64_t)b * t.E + e];
    if (lane == 0) {
        action[b] = a;
        rec.sd[2 * slot] = ep.cur[2 * b]; rec.sd[2 * slot + 1] = ep.cur[2 * b + 1];
        rec.bw[slot] = ep.cur_bw[b]; rec.action[slot] = a; rec.prob[slot] = p[a]; rec.value[slot] = values[b];
    }
}

// one allocate_to_destination_sp / decrease_links_utilization_sp (environment15.py:546-563, 150-165) on the
// Local-Search loads of one episode, followed by the max-utilisation scan of step_hill_sp (:389-399)
__global__ void k_ls_route_add(TopoView t, double* __restrict__ loads, int src, int dst, double v,
                               double* __restrict__ max_uti, int* __restrict__ max_edge) {
    const int lane = threadIdx.x & 31;
    warp_route_add(t, loads, src, dst, v, lane);
    double nu; int ne;
    warp_max_uti(loads, t.cap, t.E, lane, nu, ne);
    if (lane == 0) { *max_uti = nu; *max_edge = ne; }
}

// ---- Local Search (kernel 7) -----------------------------------------------------------------
// One CTA per episode runs the whole hill climbing of play_DRL_GNN_sp_hill_climbing_games
// (script_eval_on_single_topology.py:482-509) on device: every (demand, middlepoint) move of
// explore_neighbourhood_DRL_sp (:395-435) is evaluated by one warp against the current loads
// (get_value_sp, :317-351) with a warp-level fp64 max-reduction; the CTA keeps the first
// strictly best move in (demand, action) order, applies it (step_hill_sp, environment15.py:
// 378-404) and repeats until the reference's stop rule.
struct LsView {
    int B, Dcap, max_moves;
    const double* tm; double* loads; int* mid_cur;
    const int* elig; const int* elig_len;          // [B][Dcap] frozen demand list
    double* cur_val;                               // [B] in: max uti at reset; out: final
    int* n_moves; int* moves; double* move_val;    // [B], [B][max_moves][3], [B][max_moves]
};

constexpr int LS_TPB = 1024;
constexpr int LS_TOP = 32;       // most utilised links kept per LS iteration (see the move evaluation below)

// dynamic shared memory of k_local_search for a topology with E links
__host__ __device__ inline size_t ls_smem_bytes(int E) {
    const int nw = LS_TPB / 32;
    return (size_t)(3 * E + nw + LS_TOP) * 8 + (size_t)(3 * nw + LS_TOP) * 4 + (size_t)nw * E + 16;
}

// CL = CTAs per episode (thread-block cluster): with few episodes in flight the demand list of one episode
// is split over the CTAs of a cluster; every CTA keeps its own copy of the loads, the per-CTA best moves
// meet in CTA 0's shared memory (DSMEM), CTA 0 decides and writes the decision into every CTA's shared
// memory, and all CTAs apply the same move -- two cluster barriers per hill-climbing iteration.
constexpr int LS_MAX_CLUSTER = 8;

template <bool CLUSTER>
__global__ void __launch_bounds__(LS_TPB)
k_local_search(TopoView t, LsView ls, int CL) {
    extern __shared__ double ls_smem[];
    double* sl = ls_smem;                                // [E] loads
    double* scap = sl + t.E;                             // [E]
    double* su = scap + t.E;                             // [E] scratch: base utilisations during the top-M selection
    double* wbest = su + t.E;                            // [nwarps]
    double* top_u = wbest + LS_TPB / 32;                 // [LS_TOP] largest base utilisations, descending
    int* wbest_i = reinterpret_cast<int*>(top_u + LS_TOP);   // [nwarps] move order key
    int* wbest_m = wbest_i + LS_TPB / 32;                // [nwarps][2] (demand idx, action)
    int* top_e = wbest_m + 2 * (LS_TPB / 32);            // [LS_TOP] their link ids
    signed char* cnt = reinterpret_cast<signed char*>(top_e + LS_TOP);   // [nwarps][E]
    __shared__ double s_cur; __shared__ int s_stop; __shared__ int s_bd, s_ba, s_ntop;
    __shared__ double cl_best[LS_MAX_CLUSTER]; __shared__ int cl_key[LS_MAX_CLUSTER], cl_d[LS_MAX_CLUSTER], cl_a[LS_MAX_CLUSTER];
    namespace cg = cooperative_groups;
    const int crank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;

    const int b = CLUSTER ? blockIdx.x / CL : blockIdx.x, NN = t.N * t.N;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = LS_TPB / 32;
    const double* tm = ls.tm + (int64_t)b * NN;
    int* mid_cur = ls.mid_cur + (int64_t)b * NN;
    const int* elig = ls.elig + (int64_t)b * ls.Dcap;
    const int nd = ls.elig_len[b];
    for (int e = threadIdx.x; e < t.E; e += LS_TPB) { sl[e] = ls.loads[(int64_t)b * t.E + e]; scap[e] = t.cap[e]; }
    signed char* mycnt = cnt + wid * t.E;
    for (int e = lane; e < t.E; e += 32) mycnt[e] = 0;
    if (threadIdx.x == 0) { s_cur = ls.cur_val[b]; s_stop = 0; }
    int nmoves = 0;
    __syncthreads();

    while (true) {
        // ---- the LS_TOP most utilised links of the current loads.  A move changes the load of the links
        // on the old and the new route only, so its value is max(first top link the move does not touch,
        // utilisation of the touched links): O(path) instead of O(E) per move, and the same fp64 numbers
        // as the reference's full rescan (get_value_sp, script :331-339) because max() is exact.
        for (int e = threadIdx.x; e < t.E; e += LS_TPB) su[e] = sl[e] / scap[e];
        __syncthreads();
        if (wid == 0) {
            const int ntop = min(LS_TOP, t.E);
            for (int m = 0; m < ntop; ++m) {
                double bu = -1.0; int be = 0x7fffffff;
                for (int e = lane; e < t.E; e += 32) { const double u = su[e]; if (u > bu) { bu = u; be = e; } }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ou = __shfl_xor_sync(0xffffffffu, bu, o);
                    const int oe = __shfl_xor_sync(0xffffffffu, be, o);
                    if (ou > bu || (ou == bu && oe < be)) { bu = ou; be = oe; }
                }
                if (lane == 0) { top_u[m] = bu; top_e[m] = be; su[be] = -2.0; }
                __syncwarp();
            }
            if (lane == 0) s_ntop = ntop;
        }
        __syncthreads();
        const int ntop = s_ntop;
        // ---- explore the neighbourhood
        double best = 1e300; int best_key = 0x7fffffff, best_d = -1, best_a = -1;
        for (int di = crank * nw + wid; di < nd; di += nw * CL) {
            const int sdi = elig[di];
            const int s = sdi / t.N, d = sdi - s * t.N;
            const double bw = tm[sdi];
            const int m0 = mid_cur[sdi];
            const int k0 = t.mid_off[sdi], k1 = t.mid_off[sdi + 1];
            const int a0 = (m0 < 0 || m0 == d) ? d : m0;
            // visit counts of the current route, negated (de-allocation, :404-416)
            {
                for (int i = t.sp_off[s * t.N + a0] + lane; i < t.sp_off[s * t.N + a0 + 1]; i += 32) mycnt[t.sp_edge[i]] -= 1;
                __syncwarp();
                if (a0 != d) { for (int i = t.sp_off[a0 * t.N + d] + lane; i < t.sp_off[a0 * t.N + d + 1]; i += 32) mycnt[t.sp_edge[i]] -= 1; __syncwarp(); }
            }
            for (int a = 0; a < k1 - k0; ++a) {
                const int m1 = t.mids[k0 + a];
                for (int i = t.sp_off[s * t.N + m1] + lane; i < t.sp_off[s * t.N + m1 + 1]; i += 32) mycnt[t.sp_edge[i]] += 1;
                __syncwarp();
                if (m1 != d) { for (int i = t.sp_off[m1 * t.N + d] + lane; i < t.sp_off[m1 * t.N + d + 1]; i += 32) mycnt[t.sp_edge[i]] += 1; __syncwarp(); }
                double mx = -1000000.0;
                // untouched links: the first entry of the top list whose visit count is zero
                const bool untouched = lane < ntop && mycnt[top_e[lane < ntop ? lane : 0]] == 0;
                const unsigned free_mask = __ballot_sync(0xffffffffu, untouched);
                if (free_mask != 0u || ntop == t.E) {
                    if (free_mask != 0u) mx = top_u[__ffs(free_mask) - 1];
                    // touched links: old route (count < 0) and new route (count > 0)
                    for (int i = t.sp_off[s * t.N + a0] + lane; i < t.sp_off[s * t.N + a0 + 1]; i += 32) {
                        const int e = t.sp_edge[i]; mx = fmax(mx, (sl[e] + bw * (double)mycnt[e]) / scap[e]);
                    }
                    if (a0 != d)
                        for (int i = t.sp_off[a0 * t.N + d] + lane; i < t.sp_off[a0 * t.N + d + 1]; i += 32) {
                            const int e = t.sp_edge[i]; mx = fmax(mx, (sl[e] + bw * (double)mycnt[e]) / scap[e]);
                        }
                    for (int i = t.sp_off[s * t.N + m1] + lane; i < t.sp_off[s * t.N + m1 + 1]; i += 32) {
                        const int e = t.sp_edge[i]; mx = fmax(mx, (sl[e] + bw * (double)mycnt[e]) / scap[e]);
                    }
                    if (m1 != d)
                        for (int i = t.sp_off[m1 * t.N + d] + lane; i < t.sp_off[m1 * t.N + d + 1]; i += 32) {
                            const int e = t.sp_edge[i]; mx = fmax(mx, (sl[e] + bw * (double)mycnt[e]) / scap[e]);
                        }
                } else {
                    // every top link is touched by this move (very long routes): full rescan
                    for (int e = lane; e < t.E; e += 32) {
                        const double u = (sl[e] + bw * (double)mycnt[e]) / scap[e];
                        mx = fmax(mx, u);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                // evalState = -mx ; "if evalState > nextVal" keeps the first strictly better move
                const int key = di * t.Kmax + a;
                if (mx < best || (mx == best && key < best_key)) { best = mx; best_key = key; best_d = di; best_a = a; }
                __syncwarp();
                for (int i = t.sp_off[s * t.N + m1] + lane; i < t.sp_off[s * t.N + m1 + 1]; i += 32) mycnt[t.sp_edge[i]] -= 1;
                __syncwarp();
                if (m1 != d) { for (int i = t.sp_off[m1 * t.N + d] + lane; i < t.sp_off[m1 * t.N + d + 1]; i += 32) mycnt[t.sp_edge[i]] -= 1; __syncwarp(); }
            }
            {
                for (int i = t.sp_off[s * t.N + a0] + lane; i < t.sp_off[s * t.N + a0 + 1]; i += 32) mycnt[t.sp_edge[i]] += 1;
                __syncwarp();
                if (a0 != d) { for (int i = t.sp_off[a0 * t.N + d] + lane; i < t.sp_off[a0 * t.N + d + 1]; i += 32) mycnt[t.sp_edge[i]] += 1; __syncwarp(); }
            }
        }
        if (lane == 0) { wbest[wid] = best; wbest_i[wid] = best_key; wbest_m[2 * wid] = best_d; wbest_m[2 * wid + 1] = best_a; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double bb = 1e300; int bk = 0x7fffffff, bd = -1, ba = -1;
            for (int w = 0; w < nw; ++w)
                if (wbest[w] < bb || (wbest[w] == bb && wbest_i[w] < bk)) { bb = wbest[w]; bk = wbest_i[w]; bd = wbest_m[2 * w]; ba = wbest_m[2 * w + 1]; }
            if (CLUSTER) {
                cg::cluster_group cl = cg::this_cluster();
                *cl.map_shared_rank(&cl_best[crank], 0) = bb; *cl.map_shared_rank(&cl_key[crank], 0) = bk;
                *cl.map_shared_rank(&cl_d[crank], 0) = bd; *cl.map_shared_rank(&cl_a[crank], 0) = ba;
            } else { cl_best[0] = bb; cl_key[0] = bk; cl_d[0] = bd; cl_a[0] = ba; }
        }
        if (CLUSTER) cg::this_cluster().sync();
        if (threadIdx.x == 0 && crank == 0) {
            double bb = 1e300; int bk = 0x7fffffff, bd = -1, ba = -1;
            for (int c = 0; c < CL; ++c)
                if (cl_best[c] < bb || (cl_best[c] == bb && cl_key[c] < bk)) { bb = cl_best[c]; bk = cl_key[c]; bd = cl_d[c]; ba = cl_a[c]; }
            // nextVal = -bb (or -1000000 with no move), currentVal = -s_cur (script :483-486)
            const double nextVal = bd >= 0 ? -bb : -1000000.0;
            const double currentVal = -s_cur;
            const bool stop = (nextVal <= currentVal) || (fabs((-1.0) * nextVal - (-1.0) * currentVal) < 1e-4) ||
                              nmoves >= ls.max_moves;
            if (CLUSTER) {
                cg::cluster_group cl = cg::this_cluster();
                for (int c = 0; c < CL; ++c) {
                    *cl.map_shared_rank(&s_stop, c) = stop ? 1 : 0; *cl.map_shared_rank(&s_bd, c) = bd; *cl.map_shared_rank(&s_ba, c) = ba;
                }
            } else { s_stop = stop ? 1 : 0; s_bd = bd; s_ba = ba; }
        }
        if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
        if (s_stop) break;
        // ---- apply the move: de-allocate (:494-502) + step_hill_sp
        if (wid == 0) {
            const int sdi = elig[s_bd];
            const int s = sdi / t.N, d = sdi - s * t.N;
            const double bw = tm[sdi];
            const int m0 = mid_cur[sdi];
            const int m1 = t.mids[t.mid_off[sdi] + s_ba];
            warp_demand_add(t, sl, s, d, m0, -bw, lane);
            warp_demand_add(t, sl, s, d, m1, bw, lane);
            double nu; int ne;
            warp_max_uti(sl, scap, t.E, lane, nu, ne);
            if (lane == 0) {
                mid_cur[sdi] = (m1 != d) ? m1 : -1;      // every CTA of a cluster stores the same value (its own later reads see it)
                s_cur = nu;
                if (ls.moves && crank == 0) {
                    int* mv = ls.moves + ((int64_t)b * ls.max_moves + nmoves) * 3;
                    mv[0] = s_ba; mv[1] = s; mv[2] = d;
                    ls.move_val[(int64_t)b * ls.max_moves + nmoves] = nu;
                }
            }
        }
        ++nmoves;
        __syncthreads();
    }
    if (crank == 0) {
        for (int e = threadIdx.x; e < t.E; e += LS_TPB) ls.loads[(int64_t)b * t.E + e] = sl[e];
        if (threadIdx.x == 0) { ls.cur_val[b] = s_cur; ls.n_moves[b] = nmoves; }
    }
    if (CLUSTER) cg::this_cluster().sync();              // no CTA leaves while its shared memory may still be written
}

}  // namespace enero
