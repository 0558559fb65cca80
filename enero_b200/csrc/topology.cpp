// Host-side topology precompute (pure C++, bit-exact integer work).
//
// Replaces Env16/Env15.generate_environment of the reference:
//   edge ids                       gym-graph/gym_graph/envs/environment16.py:557-569
//   message pairs first/second     environment16.py:507-524
//   shortest paths                 environment16.py:478-505 (hop-count shortest, lexicographically
//                                  smallest node sequence == BFS next-hop table)
//   candidate middlepoints         environment16.py:395-476
// plus the per-link crossing lists of the direct shortest paths that the device-side critical-demand selection
// (k_critical_demands; environment16.py:672-674,188-200 ; environment15.py:537-538,471-483) walks.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <unordered_map>

#include "enero_internal.h"

static thread_local std::string g_last_error;
void enero_set_error(const std::string& msg) { g_last_error = msg; }
extern "C" const char* enero_last_error(void) { return g_last_error.c_str(); }
extern "C" int enero_version(void) { return 100; }

namespace {

int fail(const std::string& m) { enero_set_error(m); return ENERO_ERR_INVALID; }

void bfs_paths(enero_topo& t) {
    const int N = t.N;
    const int INF = 1 << 29;
    std::vector<std::vector<int>> radj(N);
    for (int e = 0; e < t.E; ++e) radj[t.edge_dst[e]].push_back(t.edge_src[e]);
    std::vector<int> dist((size_t)N * N, INF);   // dist[v*N + d]
    std::vector<int> frontier, next;
    for (int d = 0; d < N; ++d) {
        dist[(size_t)d * N + d] = 0;
        frontier.assign(1, d);
        while (!frontier.empty()) {
            next.clear();
            for (int v : frontier)
                for (int u : radj[v])
                    if (dist[(size_t)u * N + d] == INF) {
                        dist[(size_t)u * N + d] = dist[(size_t)v * N + d] + 1;
                        next.push_back(u);
                    }
            frontier.swap(next);
        }
    }
    t.sp_node_off.assign((size_t)N * N + 1, 0);
    t.sp_node.clear();
    for (int s = 0; s < N; ++s)
        for (int d = 0; d < N; ++d) {
            if (s != d) {
                if (dist[(size_t)s * N + d] >= INF) throw std::string("graph is not strongly connected");
                int cur = s;
                t.sp_node.push_back(cur);
                while (cur != d) {
                    int best = -1;
                    for (int v : t.adj[cur])
                        if (dist[(size_t)v * N + d] == dist[(size_t)cur * N + d] - 1 && (best < 0 || v < best)) best = v;
                    cur = best;
                    t.sp_node.push_back(cur);
                }
            }
            t.sp_node_off[(size_t)s * N + d + 1] = (int)t.sp_node.size();
        }
}

uint64_t hash_sig(const std::vector<int>& v) {
    uint64_t h = 1469598103934665603ull;
    for (int x : v) { h ^= (uint64_t)(x + 1); h *= 1099511628211ull; }
    return h;
}

void middlepoints(enero_topo& t) {
    const int N = t.N;
    t.mid_off.assign((size_t)N * N + 1, 0);
    t.mids.clear();
    t.Kmax = 0;
    std::vector<int> sig;
    std::vector<std::vector<int>> acc;
    std::vector<uint64_t> acc_h;
    for (int s = 0; s < N; ++s)
        for (int d = 0; d < N; ++d) {
            if (s != d) {
                acc.clear(); acc_h.clear();
                int kept = 0;
                for (int m = 0; m < t.K; ++m) {
                    if (m == s) continue;
                    // visit-count vector of the candidate == sorted multiset of its edge ids
                    topo_route(t, s, d, m, sig);
                    std::sort(sig.begin(), sig.end());
                    uint64_t h = hash_sig(sig);
                    bool rep = false;
                    for (size_t i = 0; i < acc.size() && !rep; ++i) rep = (acc_h[i] == h && acc[i] == sig);
                    if (rep) continue;
                    if (m != d) {
                        // nodes of path(s,m)[:-1] + path(m,d) equal to s or d must be exactly two
                        int cnt = 0;
                        int a0 = t.sp_node_off[(size_t)s * N + m], a1 = t.sp_node_off[(size_t)s * N + m + 1];
                        for (int i = a0; i < a1 - 1; ++i) cnt += (t.sp_node[i] == s || t.sp_node[i] == d);
                        int b0 = t.sp_node_off[(size_t)m * N + d], b1 = t.sp_node_off[(size_t)m * N + d + 1];
                        for (int i = b0; i < b1; ++i) cnt += (t.sp_node[i] == s || t.sp_node[i] == d);
                        if (cnt != 2) continue;
                    }
                    t.mids.push_back(m);
                    acc.push_back(sig);
                    acc_h.push_back(h);
                    ++kept;
                }
                t.Kmax = std::max(t.Kmax, kept);
            }
            t.mid_off[(size_t)s * N + d + 1] = (int)t.mids.size();
        }
}

}  // namespace

extern "C" int enero_topo_build(int N, int E, const int32_t* edge_src, const int32_t* edge_dst,
                                const double* capacity, int K, const int32_t* sp_off,
                                const int32_t* sp_nodes, enero_topo** out) {
    if (!out || !edge_src || !edge_dst || !capacity) return fail("enero_topo_build: null argument");
    if (N < 2 || E < 1 || K < 1) return fail("enero_topo_build: need N >= 2, E >= 1, K >= 1");
    auto* t = new enero_topo();
    try {
        t->N = N; t->E = E; t->K = std::min(K, N);
        t->edge_src.assign(edge_src, edge_src + E);
        t->edge_dst.assign(edge_dst, edge_dst + E);
        t->cap.assign(capacity, capacity + E);
        t->eid.assign((size_t)N * N, -1);
        t->adj.assign(N, {});
        std::vector<char> seen_src(N, 0);
        for (int e = 0; e < E; ++e) {
            int u = edge_src[e], v = edge_dst[e];
            if (u < 0 || u >= N || v < 0 || v >= N || u == v) throw std::string("bad link endpoint");
            if (t->eid[(size_t)u * N + v] >= 0) throw std::string("duplicate link");
            if (e > 0 && edge_src[e - 1] != u && seen_src[u]) throw std::string("links are not in edge-id order (source groups must be contiguous)");
            if (!seen_src[u]) { seen_src[u] = 1; t->node_order.push_back(u); }
            t->eid[(size_t)u * N + v] = e;
            t->adj[u].push_back(v);
            if (!(capacity[e] > 0)) throw std::string("capacity must be positive");
        }
        t->max_cap = *std::max_element(t->cap.begin(), t->cap.end());
        t->capf.resize(E);
        for (int e = 0; e < E; ++e) t->capf[e] = (float)(t->cap[e] / t->max_cap);   // environment16.py:574
        // message pairs
        for (int e = 0; e < E; ++e) {
            int i = edge_src[e], j = edge_dst[e];
            for (int n : t->adj[j])
                if (n != i) { t->first.push_back(e); t->second.push_back(t->eid[(size_t)j * N + n]); }
        }
        t->P = (int)t->first.size();
        t->off_f = t->P ? *std::max_element(t->first.begin(), t->first.end()) + 1 : 0;
        t->off_s = t->P ? *std::max_element(t->second.begin(), t->second.end()) + 1 : 0;
        // stable CSR by `second`: same accumulation order as TF's CPU unsorted_segment_sum
        t->in_off.assign(E + 1, 0);
        for (int p = 0; p < t->P; ++p) t->in_off[t->second[p] + 1]++;
        for (int e = 0; e < E; ++e) t->in_off[e + 1] += t->in_off[e];
        t->in_first.resize(t->P);
        { std::vector<int> cur(t->in_off.begin(), t->in_off.end() - 1);
          for (int p = 0; p < t->P; ++p) t->in_first[cur[t->second[p]]++] = t->first[p]; }
        // CSR by `first` (the transpose, used by the backward pass)
        t->out_off.assign(E + 1, 0);
        for (int p = 0; p < t->P; ++p) t->out_off[t->first[p] + 1]++;
        for (int e = 0; e < E; ++e) t->out_off[e + 1] += t->out_off[e];
        t->out_second.resize(t->P);
        { std::vector<int> cur(t->out_off.begin(), t->out_off.end() - 1);
          for (int p = 0; p < t->P; ++p) t->out_second[cur[t->first[p]]++] = t->second[p]; }
        // shortest paths
        if (sp_off && sp_nodes) {
            t->sp_node_off.assign(sp_off, sp_off + (size_t)N * N + 1);
            t->sp_node.assign(sp_nodes, sp_nodes + sp_off[(size_t)N * N]);
        } else {
            bfs_paths(*t);
        }
        t->sp_off.assign((size_t)N * N + 1, 0);
        for (int s = 0; s < N; ++s)
            for (int d = 0; d < N; ++d) {
                size_t sd = (size_t)s * N + d;
                int a = t->sp_node_off[sd], b = t->sp_node_off[sd + 1];
                if (s != d) {
                    if (b - a < 2 || t->sp_node[a] != s || t->sp_node[b - 1] != d) throw std::string("bad shortest path");
                    for (int i = a; i + 1 < b; ++i) {
                        int id = t->eid[(size_t)t->sp_node[i] * N + t->sp_node[i + 1]];
                        if (id < 0) throw std::string("shortest path uses a missing link");
                        t->sp_edge.push_back(id);
                    }
                }
                t->sp_off[sd + 1] = (int)t->sp_edge.size();
            }
        middlepoints(*t);
        // direct-SP crossing lists, row-major (s,d) order
        std::vector<std::vector<int>> cross(E);
        for (int s = 0; s < N; ++s)
            for (int d = 0; d < N; ++d)
                if (s != d)
                    for (int i = t->sp_off[(size_t)s * N + d]; i < t->sp_off[(size_t)s * N + d + 1]; ++i)
                        cross[t->sp_edge[i]].push_back(s * N + d);
        t->cross_off.assign(E + 1, 0);
        for (int e = 0; e < E; ++e) {
            t->cross_off[e + 1] = t->cross_off[e] + (int)cross[e].size();
            t->cross_sd.insert(t->cross_sd.end(), cross[e].begin(), cross[e].end());
        }
    } catch (const std::string& m) {
        delete t;
        return fail("enero_topo_build: " + m);
    }
    *out = t;
    return ENERO_OK;
}

extern "C" int enero_topo_sizes(const enero_topo* t, int64_t out[9]) {
    if (!t || !out) return fail("enero_topo_sizes: null argument");
    out[0] = t->N; out[1] = t->E; out[2] = t->P; out[3] = t->K; out[4] = t->off_f; out[5] = t->off_s;
    out[6] = (int64_t)t->mids.size(); out[7] = t->Kmax; out[8] = (int64_t)t->sp_edge.size();
    return ENERO_OK;
}

extern "C" int enero_topo_get_pairs(const enero_topo* t, int32_t* first, int32_t* second) {
    if (!t || !first || !second) return fail("enero_topo_get_pairs: null argument");
    std::copy(t->first.begin(), t->first.end(), first);
    std::copy(t->second.begin(), t->second.end(), second);
    return ENERO_OK;
}

extern "C" int enero_topo_get_sp_edges(const enero_topo* t, int32_t* off, int32_t* edges) {
    if (!t || !off || !edges) return fail("enero_topo_get_sp_edges: null argument");
    std::copy(t->sp_off.begin(), t->sp_off.end(), off);
    std::copy(t->sp_edge.begin(), t->sp_edge.end(), edges);
    return ENERO_OK;
}

extern "C" int enero_topo_get_middlepoints(const enero_topo* t, int32_t* off, int32_t* mids) {
    if (!t || !off || !mids) return fail("enero_topo_get_middlepoints: null argument");
    std::copy(t->mid_off.begin(), t->mid_off.end(), off);
    std::copy(t->mids.begin(), t->mids.end(), mids);
    return ENERO_OK;
}

extern "C" int enero_topo_get_capacity_feature(const enero_topo* t, float* capf) {
    if (!t || !capf) return fail("enero_topo_get_capacity_feature: null argument");
    std::copy(t->capf.begin(), t->capf.end(), capf);
    return ENERO_OK;
}
