// SAP ("shortest available path") baseline, results[8] of the evaluation records.
//
// Reference: script_eval_on_single_topology.py:511-557 (SAPAgent.act, play_sap_games) on
// gym-graph/gym_graph/envs/environment20.py:124-163 (_generate_tm: every demand, stable-sorted by bandwidth
// descending), :240-291 (step: allocate on the chosen path, or on random.randint(0, len(paths)-1) when no path
// has room) and :300-316 (reset).  The episode of one traffic matrix is sequential (every allocation sees the
// loads left by the previous ones); the batch is parallel over traffic matrices, one warp per episode.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "env_kernels.cuh"

namespace enero {

struct KspView { const int* pair_off; const int* edge_off; const int* edge; };

// ---- demand order: sorted(list_eligible_demands, key = bw, reverse = True) (environment20.py:149) -----------
// Python's sort is stable also with reverse=True, so the order is bw descending, then row-major (src, dst).
// Rank sort: thread i counts the demands that come before demand i; the keys (bw, index) are all distinct.
__global__ void __launch_bounds__(256)
k_sap_order(int N, const double* __restrict__ tm, int* __restrict__ order) {
    __shared__ double sbw[256];
    const int b = blockIdx.y, NN = N * N;
    const double* t = tm + (int64_t)b * NN;
    const int i = blockIdx.x * 256 + threadIdx.x;          // index over all N*N cells, diagonal skipped
    const bool live = i < NN && (i / N) != (i % N);
    const double mine = live ? t[i] : 0.0;
    int rank = 0;
    for (int j0 = 0; j0 < NN; j0 += 256) {
        __syncthreads();
        const int j = j0 + threadIdx.x;
        sbw[threadIdx.x] = (j < NN && (j / N) != (j % N)) ? t[j] : -1.0;      // diagonal / padding: never "before"
        __syncthreads();
        if (live) {
            const int jn = min(256, NN - j0);
            for (int q = 0; q < jn; ++q) {
                const double o = sbw[q];
                rank += (o > mine || (o == mine && j0 + q < i)) ? 1 : 0;
            }
        }
    }
    // the -1.0 sentinels never count (bandwidths are >= 0), so rank is the position among the N(N-1) demands
    if (live) order[(int64_t)b * (NN - N) + rank] = i;
}

// ---- Python's random module (Mersenne Twister) as play_sap_games leaves it after env.seed(SEED) -----------------
// random.seed(int) = init_by_array([abs(seed)]); randint(0, n-1) = _randbelow(n): k = n.bit_length();
// r = getrandbits(k) until r < n; getrandbits(k <= 32) = genrand_uint32() >> (32 - k).
struct PyMT {
    uint32_t* mt; int* idx;       // [624], [1] in global memory
    __device__ void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        *idx = 624;
    }
    __device__ void seed(uint32_t key) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = 624; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key + (uint32_t)j;      // one key word: j stays 0
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    __device__ uint32_t next() {
        if (*idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            *idx = 0;
        }
        uint32_t y = mt[(*idx)++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    __device__ int randbelow(int n) {
        const int k = 32 - __clz(n);
        uint32_t r = next() >> (32 - k);
        while ((int)r >= n) r = next() >> (32 - k);
        return (int)r;
    }
};

// ---- one SAP episode per warp ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_sap_run(TopoView t, KspView ks, int B, int K, uint32_t seed, const double* __restrict__ tm, const int* __restrict__ order,
          double* __restrict__ loads, uint32_t* __restrict__ mt_state, int* __restrict__ mt_idx,
          double* __restrict__ max_uti, int* __restrict__ actions) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int NN = t.N * t.N, D = NN - t.N;
    double* ld = loads + (int64_t)b * t.E;
    for (int e = lane; e < t.E; e += 32) ld[e] = 0.0;
    PyMT rng{mt_state + (int64_t)b * 624, mt_idx + b};
    bool seeded = false;
    __syncwarp();
    for (int q = 0; q < D; ++q) {
        const int sdi = order[(int64_t)b * D + q];
        const double bw = tm[(int64_t)b * NN + sdi];
        const int p0 = ks.pair_off[sdi], np = ks.pair_off[sdi + 1] - p0;
        int action = -1;
        // SAPAgent.act: the first of the (at most K) shortest paths on which every link keeps (load + bw) / cap <= 1
        for (int p = 0; p < np && p < K && action < 0; ++p) {
            bool ok = true;
            for (int i = ks.edge_off[p0 + p] + lane; i < ks.edge_off[p0 + p + 1]; i += 32) {
                const int e = ks.edge[i];
                ok &= !((ld[e] + bw) / t.cap[e] > 1.0);
            }
            if (__all_sync(0xffffffffu, ok)) action = p;
        }
        if (actions && lane == 0) actions[(int64_t)b * D + q] = action;
        if (action < 0) {                                   // environment20.py:251-254: load balancing on a random path
            int a = 0;
            if (lane == 0) {
                if (!seeded) rng.seed(seed);
                a = rng.randbelow(np);
            }
            seeded = true;
            action = __shfl_sync(0xffffffffu, a, 0);
        }
        __syncwarp();
        for (int i = ks.edge_off[p0 + action] + lane; i < ks.edge_off[p0 + action + 1]; i += 32) ld[ks.edge[i]] += bw;
        __syncwarp();
    }
    double best; int be;
    warp_max_uti(ld, t.cap, t.E, lane, best, be);           // environment20.py:267-279, value returned by the last step
    if (lane == 0) max_uti[b] = best;
}

}  // namespace enero
