// K shortest simple paths per node pair: the path table of the SAP baseline environment.
//
// Reference: gym-graph/gym_graph/envs/environment20.py:165-196 (Env20.compute_SPs):
//   allPaths[s:d] = sorted(nx.all_simple_paths(G, s, d, cutoff = 2 * diameter), key = (len, node list))[:K]
// i.e. the K simple paths with the fewest hops among those of at most 2*diameter links, ties broken by the
// lexicographic order of the node sequence.  The reference enumerates every simple path (exponential); here a
// depth-first search enumerates the simple paths of exactly L links in lexicographic order for L = dist(s,d),
// dist(s,d)+1, ... (pruned by the hop distance to d) and stops at the K-th path, which yields the same list.
#include <algorithm>
#include <string>

#include "enero_internal.h"

namespace {

struct Ksp {
    const enero_topo& t;
    int N, K;
    std::vector<std::vector<int>> nbr;       // out-neighbours, ascending node id
    std::vector<int> dist;                   // dist[v*N + d], hops
    std::vector<int> cur; std::vector<char> used;
    std::vector<std::vector<int>>* found = nullptr;
    int target = 0;

    Ksp(const enero_topo& topo, int k) : t(topo), N(topo.N), K(k), nbr(topo.N), dist((size_t)topo.N * topo.N, 1 << 29), used(topo.N, 0) {
        for (int e = 0; e < t.E; ++e) nbr[t.edge_src[e]].push_back(t.edge_dst[e]);
        for (auto& v : nbr) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
        std::vector<std::vector<int>> radj(N);
        for (int e = 0; e < t.E; ++e) radj[t.edge_dst[e]].push_back(t.edge_src[e]);
        std::vector<int> frontier, next;
        for (int d = 0; d < N; ++d) {
            dist[(size_t)d * N + d] = 0;
            frontier.assign(1, d);
            while (!frontier.empty()) {
                next.clear();
                for (int v : frontier)
                    for (int u : radj[v])
                        if (dist[(size_t)u * N + d] == (1 << 29)) { dist[(size_t)u * N + d] = dist[(size_t)v * N + d] + 1; next.push_back(u); }
                frontier.swap(next);
            }
        }
    }
    int diameter() const {
        int m = 0;
        for (int s = 0; s < N; ++s) for (int d = 0; d < N; ++d) m = std::max(m, dist[(size_t)s * N + d]);
        return m;
    }
    // simple paths of exactly `remaining` more links from cur.back() to target, lexicographic order
    void dfs(int remaining) {
        if ((int)found->size() >= K) return;
        const int v = cur.back();
        if (remaining == 0) { if (v == target) found->push_back(cur); return; }
        for (int w : nbr[v]) {
            if (used[w] || dist[(size_t)w * N + target] > remaining - 1) continue;
            if (w == target && remaining != 1) continue;       // a simple path ends at its first visit of the target
            used[w] = 1; cur.push_back(w);
            dfs(remaining - 1);
            cur.pop_back(); used[w] = 0;
            if ((int)found->size() >= K) return;
        }
    }
    void pair(int s, int d, int cutoff, std::vector<std::vector<int>>& out) {
        out.clear();
        found = &out; target = d;
        for (int L = dist[(size_t)s * N + d]; L <= cutoff && (int)out.size() < K; ++L) {
            cur.assign(1, s);
            std::fill(used.begin(), used.end(), 0);
            used[s] = 1;
            dfs(L);
        }
    }
};

}  // namespace

void topo_build_ksp(enero_topo& t, int K) {
    Ksp k(t, K);
    const int N = t.N;
    const int cutoff = 2 * k.diameter();
    t.ksp_K = K;
    t.ksp_pair_off.assign((size_t)N * N + 1, 0);
    t.ksp_path_off.assign(1, 0);
    t.ksp_node.clear(); t.ksp_edge_off.assign(1, 0); t.ksp_edge.clear();
    std::vector<std::vector<int>> paths;
    for (int s = 0; s < N; ++s)
        for (int d = 0; d < N; ++d) {
            if (s != d) {
                k.pair(s, d, cutoff, paths);
                for (auto& p : paths) {
                    for (int v : p) t.ksp_node.push_back(v);
                    t.ksp_path_off.push_back((int)t.ksp_node.size());
                    for (size_t i = 0; i + 1 < p.size(); ++i) t.ksp_edge.push_back(t.eid[(size_t)p[i] * N + p[i + 1]]);
                    t.ksp_edge_off.push_back((int)t.ksp_edge.size());
                }
            }
            t.ksp_pair_off[(size_t)s * N + d + 1] = (int)t.ksp_path_off.size() - 1;
        }
}

extern "C" int enero_topo_build_ksp(enero_topo* t, int K) {
    if (!t || K < 1) { enero_set_error("enero_topo_build_ksp: bad argument"); return ENERO_ERR_INVALID; }
    if (t->ksp_K == K) return ENERO_OK;
    topo_build_ksp(*t, K);
    t->ksp_dev_stale = true;
    return ENERO_OK;
}
extern "C" int enero_topo_ksp_sizes(const enero_topo* t, int64_t out[2]) {
    if (!t || !out || t->ksp_K == 0) { enero_set_error("enero_topo_ksp_sizes: no K-shortest-path table (enero_topo_build_ksp)"); return ENERO_ERR_STATE; }
    out[0] = (int64_t)t->ksp_path_off.size() - 1; out[1] = (int64_t)t->ksp_node.size();
    return ENERO_OK;
}
extern "C" int enero_topo_get_ksp(const enero_topo* t, int32_t* pair_off, int32_t* path_off, int32_t* nodes) {
    if (!t || !pair_off || !path_off || !nodes || t->ksp_K == 0) { enero_set_error("enero_topo_get_ksp: bad argument / no table"); return ENERO_ERR_INVALID; }
    std::copy(t->ksp_pair_off.begin(), t->ksp_pair_off.end(), pair_off);
    std::copy(t->ksp_path_off.begin(), t->ksp_path_off.end(), path_off);
    std::copy(t->ksp_node.begin(), t->ksp_node.end(), nodes);
    return ENERO_OK;
}
