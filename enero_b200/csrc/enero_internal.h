// Internal structures shared by the host (C++) and device (CUDA) translation units.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/enero_b200.h"

void enero_set_error(const std::string& msg);

// Flat weight offsets (floats) inside the Keras-ordered parameter vector.
enum : int {
    W_MSG_K = 0,       // [40][20]  rows 0..19 act on h[first], rows 20..39 on h[second]
    W_MSG_B = 800,     // [20]
    W_GRU_K = 820,     // [20][60]  gate order z, r, h
    W_GRU_R = 2020,    // [20][60]
    W_GRU_B = 3220,    // [2][60]   input bias, recurrent bias (reset_after=True)
    W_R1_K = 3340, W_R1_B = 3740, W_R2_K = 3760, W_R2_B = 4160, W_R3_K = 4180, W_R3_B = 4200,
    W_MP_FLOATS = 3340,  // message + update part, staged in shared memory by the MP kernel
};

struct TopoDev {               // device mirrors (all int32 unless noted)
    int N = 0, E = 0, P = 0, K = 0, Kmax = 0, off_f = 0, off_s = 0;
    double* cap = nullptr;        // [E]
    float* capf = nullptr;        // [E]  fp32(cap / maxCap)
    int* in_off = nullptr;        // [E+1] CSR by `second` (stable in pair order)
    int* in_first = nullptr;      // [P]
    int* sp_off = nullptr;        // [N*N+1]
    int* sp_edge = nullptr;       // [sp_total]
    int* mid_off = nullptr;       // [N*N+1]
    int* mids = nullptr;          // [mid_total]
    int* out_off = nullptr;       // [E+1] CSR by `first` (pairs leaving a link), pair order
    int* out_second = nullptr;    // [P]
    int* cross_off = nullptr;     // [E+1] direct-SP crossing lists per link: (s*N+d), row-major
    int* cross_sd = nullptr;
};

struct enero_topo {
    int N = 0, E = 0, P = 0, K = 0, Kmax = 0, off_f = 0, off_s = 0;
    std::vector<int> edge_src, edge_dst;       // [E], edge-id order
    std::vector<int> eid;                       // [N*N] -> edge id or -1
    std::vector<std::vector<int>> adj;          // adjacency in insertion order
    std::vector<int> node_order;
    std::vector<double> cap;
    std::vector<float> capf;
    double max_cap = 0;
    std::vector<int> first, second;             // [P]
    std::vector<int> in_off, in_first;          // CSR by second
    std::vector<int> out_off, out_second;       // CSR by first
    std::vector<int> sp_off, sp_edge;           // SP as edge lists
    std::vector<int> sp_node_off, sp_node;      // SP as node lists
    std::vector<int> mid_off, mids;
    std::vector<int> cross_off, cross_sd;       // per link: (s*N+d) of direct SPs crossing it, row-major
    // SAP baseline (Env20): K shortest simple paths per pair, built on demand (ksp.cpp)
    int ksp_K = 0;
    std::vector<int> ksp_pair_off;              // [N*N+1] -> path index
    std::vector<int> ksp_path_off, ksp_node;    // node lists (for allPaths)
    std::vector<int> ksp_edge_off, ksp_edge;    // edge-id lists (for the kernel)
    bool ksp_dev_stale = true;
    int* d_ksp_pair_off = nullptr; int* d_ksp_edge_off = nullptr; int* d_ksp_edge = nullptr;
    TopoDev dev;
    int dev_id = -1;                            // device the mirrors live on (-1 = none)
};

// route (with multiplicity) of demand (s,d) via mid (mid < 0 or mid == d: direct)
inline void topo_route(const enero_topo& t, int s, int d, int mid, std::vector<int>& out) {
    out.clear();
    auto seg = [&](int a, int b) {
        for (int i = t.sp_off[a * t.N + b]; i < t.sp_off[a * t.N + b + 1]; ++i) out.push_back(t.sp_edge[i]);
    };
    if (mid < 0 || mid == d) seg(s, d);
    else { seg(s, mid); seg(mid, d); }
}
