// k_mp_tile: one message-passing iteration over graph-aligned tiles with TMA-staged operands.
//
// Why: profiles/r01_v4_* showed k_mp_iter bound by the LSU / L1 data pipe, not by the FMA pipe: its
// per-thread 80-byte-stride global accesses (h row load, A gathers, h'/A' stores) cost ~20 L1 wavefronts
// per LDG/STG.128 and made up half of the LSU traffic.  Here every global access of the hot loop is a
// bulk asynchronous copy (cp.async.bulk, the 1-D TMA path) issued by one thread:
//   * the tile's link-state rows land directly in the per-thread shared-memory slots (slot i = row i,
//     so a contiguous copy IS the transposition),
//   * the message operands A of the graphs the tile holds land in a staging buffer and are gathered
//     from shared memory,
//   * h' and A' are written to the slots and leave with one bulk store each.
//
// Tiling (reference batching quirk Q1 included, script_eval_on_single_topology.py:58-64,110-117): rows of a
// decision are grouped by *message-passing graph* k = lr / off_s; a tile holds G consecutive graphs of one
// decision, i.e. destination rows [k0*off_s, (k0+Gt)*off_s), whose sources are exactly rows
// [k0*off_f, (k0+Gt)*off_f) of the same decision -- both contiguous.  The rows past the last graph,
// [nK*off_s, nK*E) (non-empty iff off_s < E; they receive no message), ride in the last group's tile when
// they fit in its free slots, else in the decision's extra tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gnn_kernels.cuh"

namespace enero {

struct TileTopo {
    int B, E, off_f, off_s, RPD, G, TPD;   // TPD = tiles per decision = ceil(Kmax / G) + 1
    const int* nK;                         // [B] live candidate graphs per decision
    const int* in_off;                     // [E+1] CSR by `second`
    const int* in_first;                   // [P]
};

constexpr int MP_TILE_SMEM_MAX = 227 * 1024;
// dynamic shared memory: weights | h slots [2*TPB][20] | B/z/A' slots [2*TPB][20] | stage [G*off_f][20] | mbarrier
__host__ __device__ inline size_t mp_tile_smem_bytes(int tpb, int stage_rows) {
    return sizeof(float) * ((size_t)W_MP_FLOATS + (size_t)2 * MP_R * tpb * H + (size_t)stage_rows * H) + 16;
}

// what one CTA needs to know about a tile (uniform across the CTA)
struct TileDesc {
    int nrows;        // live rows in the tile (0 = nothing to do)
    int ngroup;       // of which rows that belong to message-passing graphs (the rest receive no message)
    int nsrc;         // staged source rows
    int64_t row0;     // first global row of the tile
    int64_t src0;     // first global source row
};

__device__ __forceinline__ TileDesc mp_tile_describe(const TileTopo& tt, int tile, int slots) {
    TileDesc d; d.nrows = d.ngroup = d.nsrc = 0; d.row0 = d.src0 = 0;
    const int b = tile / tt.TPD, j = tile - b * tt.TPD;
    const int nk = tt.nK[b];
    if (nk <= 0) return d;
    const int rem = nk * (tt.E - tt.off_s);                 // rows past the last graph
    const int ng = (nk + tt.G - 1) / tt.G;                  // group tiles of this decision
    const int64_t base = (int64_t)b * tt.RPD;
    if (j < ng) {
        const int k0 = j * tt.G;
        const int gt = min(tt.G, nk - k0);
        d.ngroup = gt * tt.off_s;
        d.nrows = d.ngroup;
        if (j == ng - 1 && d.ngroup + rem <= slots) d.nrows += rem;
        d.nsrc = gt * tt.off_f;
        d.row0 = base + (int64_t)k0 * tt.off_s;
        d.src0 = base + (int64_t)k0 * tt.off_f;
    } else if (j == tt.TPD - 1 && rem > 0) {
        const int last_gt = nk - (ng - 1) * tt.G;
        if (last_gt * tt.off_s + rem > slots) {             // did not fit in the last group's tile
            d.nrows = rem;                                   // host guarantees rem <= slots
            d.row0 = base + (int64_t)nk * tt.off_s;
        }
    }
    return d;
}

template <bool LAST, int KH>
__global__ void __launch_bounds__(384, 1)
k_mp_tile(const TileTopo tt, const float* __restrict__ wflat, const float* __restrict__ h_in,
          const float* __restrict__ A_in, float* __restrict__ h_out, float* __restrict__ A_out, int ntiles) {
    extern __shared__ __align__(16) float mpt_smem[];
    const int TPB = blockDim.x;
    const int slots = MP_R * TPB;
    float* sw_s = mpt_smem;
    WSmem sw; sw.sw = sw_s;
    float* sh = sw_s + W_MP_FLOATS;            // slot i = tile row i: h (old, then new)
    float* sb = sh + slots * H;                // B, then z, then A'
    float* stage = sb + slots * H;             // message operands of the tile's graphs
    uint64_t* bar = reinterpret_cast<uint64_t*>(stage + tt.G * tt.off_f * H);
    for (int i = threadIdx.x; i < W_MP_FLOATS / 4; i += TPB)
        reinterpret_cast<float4*>(sw_s)[i] = reinterpret_cast<const float4*>(wflat)[i];
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    float* myh[MP_R]; float* myb[MP_R];
#pragma unroll
    for (int r = 0; r < MP_R; ++r) { myh[r] = sh + (r * TPB + threadIdx.x) * H; myb[r] = sb + (r * TPB + threadIdx.x) * H; }
    uint32_t parity = 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileDesc td = mp_tile_describe(tt, tile, slots);
        if (td.nrows == 0) continue;
        // ---- bulk loads: link-state rows -> slots, message operands -> stage
        if (threadIdx.x == 0) {
            bulk_wait_read0();                 // the previous tile's stores have finished reading the slots
            mbar_expect_tx(bar, (uint32_t)(td.nrows + td.nsrc) * H * 4);
            bulk_g2s(sh, h_in + td.row0 * H, (uint32_t)td.nrows * H * 4, bar);
            if (td.nsrc) bulk_g2s(stage, A_in + td.src0 * H, (uint32_t)td.nsrc * H * 4, bar);
        }
        // per-row bookkeeping while the copies fly
        bool ok[MP_R]; int pb[MP_R], pe[MP_R], soff[MP_R];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            const int i = r * TPB + threadIdx.x;
            ok[r] = i < td.nrows; pb[r] = pe[r] = 0; soff[r] = 0;
            if (i < td.ngroup) {
                const int kk = i / tt.off_s, x = i - kk * tt.off_s;
                pb[r] = tt.in_off[x]; pe[r] = tt.in_off[x + 1]; soff[r] = kk * tt.off_f;
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        float h[MP_R][H];
        u64 agg2[MP_R][H / 2];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            if (ok[r]) load_row(h[r], myh[r]);
            else {
#pragma unroll
                for (int j = 0; j < H; ++j) h[r][j] = 0.f;
                store_row(myh[r], h[r]);
            }
#pragma unroll
            for (int j = 0; j < H / 2; ++j) agg2[r][j] = 0ull;
        }
        // ---- B = bm + h . Wm[20:40]
        mp_phase_B<KH>(sw, h, myb);
        // ---- agg = sum over incoming pairs (ascending pair index) of selu(A[first] + B), operands from stage
        {
            const int steps = max(pe[0] - pb[0], pe[1] - pb[1]);
            u64 t2[MP_R][H / 2];
#pragma unroll
            for (int r = 0; r < MP_R; ++r) mp_load_B(t2[r], myb[r]);
            for (int s = 0; s < steps; ++s) {
#pragma unroll
                for (int r = 0; r < MP_R; ++r) {
                    if (pb[r] + s < pe[r]) {
                        float a[H];
                        load_row(a, stage + (tt.in_first[pb[r] + s] + soff[r]) * H);
                        mp_accumulate(agg2[r], a, t2[r]);
                    }
                }
            }
        }
        float agg[MP_R][H];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) {
            load_row(h[r], myh[r]);
#pragma unroll
            for (int j = 0; j < H / 2; ++j) unpack2(agg2[r][j], agg[r][2 * j], agg[r][2 * j + 1]);
        }
        mp_phase_z<KH>(sw, h, agg, myb);
        float* none[MP_R];
#pragma unroll
        for (int r = 0; r < MP_R; ++r) none[r] = nullptr;
        mp_phase_rh<KH>(sw, h, agg, myh, myb, none);          // new h stays in the slots
        if (!LAST) mp_phase_A(sw, myh, myb);                  // A' replaces z in the B slots
        // ---- bulk stores: slots -> h_out / A_out
        fence_proxy_async();                                   // generic-proxy writes visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_s2g(h_out + td.row0 * H, sh, (uint32_t)td.nrows * H * 4);
            if (!LAST) bulk_s2g(A_out + td.row0 * H, sb, (uint32_t)td.nrows * H * 4);
            bulk_commit();
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
}

}  // namespace enero
