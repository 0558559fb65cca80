// k_mp_tc: one message-passing iteration with the five mat-muls on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), everything else on the CUDA cores.
//
// Reference semantics: actorPPOmiddR.py:45-63 (same formulation as k_mp_iter, gnn_kernels.cuh).
//
// Why: k_mp_iter spends its time streaming the 3 340 weights through broadcast LDS.128 into FFMA2 (DESIGN.md
// section 3: the shared-memory pipe, not the FMA pipe, is its ceiling -- 47 % of the FP32 peak).  Here a CTA owns a
// tile of 128 link-state rows (thread t <-> row t <-> TMEM lane t) and the tensor core does
//     phase 1   [Y | B]  = h   . [R | Wm_bot]        128 x 24 . 24 x 80     (h-dependent products)
//     phase 2    X       = agg . K                   128 x 24 . 24 x 64
//     phase 3    A'      = h'  . Wm_top              128 x 24 . 24 x 32
// from K-major, un-swizzled shared-memory operand tiles; the threads read the accumulators back with tcgen05.ld
// (32 lanes x 32 bit: thread t gets columns of ITS row) and do the gather / SELU / segment sum, the GRU gates and
// the blend exactly as k_mp_iter does.
//
// fp32-grade products from TF32 tensor cores (3xTF32): x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi);
// A.W ~= A_lo.W_hi + A_hi.W_lo + A_hi.W_hi, accumulated in THAT order: the tensor core's accumulator truncates, so
// the small correction terms go in before the large one.  tools/tc_probe.cu measures this against fp64: max error
// 4.9e-7 on outputs of magnitude 3, the same as an fp32 FMA chain (4.1e-7); large-term-first is 3x worse.
// X and Y are kept in separate accumulators and added on the CUDA cores, (X + b0) + (Y + b1), the association of
// Keras' GRUCell (reset_after=True).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gnn_kernels.cuh"

namespace enero {

#ifdef ENERO_TC_TIMING
__device__ unsigned long long tc_dbg[8];
#define TC_TICK(i) do { if (t == 0) { const long long _n = clock64(); dbg[i] += _n - tick; tick = _n; } } while (0)
#else
#define TC_TICK(i) do {} while (0)
#endif

constexpr int TC_ROWS = 128;                    // rows per tile = threads per CTA = TMEM lanes
constexpr int TC_KC = 6;                        // 16-byte chunks along K: 24 tf32 (20 used)
constexpr int TC_N1 = 80, TC_N2 = 64, TC_N3 = 32;
constexpr int TC_COL_Y = 0;                     // phase 1: Yz 0..19, Yr 20..39, Yh 40..59
constexpr int TC_COL_B = 60;                    //          B 60..79
constexpr int TC_COL_X = 60;                    // phase 2 (B is dead by then): Xz 60..79, Xr 80..99, Xh 100..119
constexpr int TC_COL_A = 0;                     // phase 3 (Y is dead by then): A' 0..19
constexpr int TC_TMEM_COLS = 128;                // per group
constexpr int TC_W_F4 = (TC_N1 + TC_N2 + TC_N3) * TC_KC;          // float4 per hi (or lo) weight image
constexpr int TC_GROUPS = 2;                     // independent 128-thread tile pipelines per CTA (they share the weight images)
constexpr int TC_TPB = TC_GROUPS * TC_ROWS;
constexpr size_t TC_GROUP_BYTES = (size_t)2 * TC_KC * TC_ROWS * 16 + (size_t)TC_ROWS * H * 4 + 32;   // operands hi/lo, stage, 2 mbarriers
constexpr size_t TC_SMEM_BYTES = (size_t)2 * TC_W_F4 * 16 + 832 + 16 + TC_GROUPS * TC_GROUP_BYTES;   // 208 floats: biases (140 used)

__device__ __forceinline__ uint64_t tc_desc(const void* p, uint32_t lbo_bytes) {
    // K-major, no swizzle: 8 rows x 16 B core matrices, rows 16 B apart; SBO (next 8 rows) = 128 B, LBO (next
    // 16-byte chunk along K) = lbo_bytes; version 1 (Blackwell)
    return (uint64_t)((smem_u32(p) >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t tc_idesc(int N) {
    // F32 accumulate, A and B TF32, both K-major, M = 128
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
// 20 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld20(uint32_t taddr, float (&v)[H]) {
    uint32_t r[H];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]) : "r"(taddr + 16));
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < H; ++i) v[i] = __uint_as_float(r[i]);
}
// round to nearest, ties away from zero, onto the 10-bit TF32 mantissa -- cvt.rna.tf32.f32 written as two integer
// ALU operations (add half an ulp to the magnitude, clear the 13 low bits): the conversion instruction runs on the XU
// pipe, which the SELU / sigmoid / tanh MUFU operations already keep busy (ncu: 56 % with cvt, profiles/r02_tc_*)
__device__ __forceinline__ float tc_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
    hi = tc_rna(x);
    lo = tc_rna(x - hi);
}
// this thread's row as 3xTF32 operand chunks: chunk c of row t sits at [c][t] (512 contiguous bytes per warp)
template <int NCH>
__device__ __forceinline__ void tc_store_operand(float4* a_hi, float4* a_lo, int t, const float (&x)[H]) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) tc_split(x[4 * c + q], hi[q], lo[q]);
        a_hi[c * TC_ROWS + t] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        a_lo[c * TC_ROWS + t] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}
// A_lo.W_hi + A_hi.W_lo + A_hi.W_hi over `ksteps` k-steps of 8, small terms first; one thread issues
template <int N>
__device__ __forceinline__ void tc_issue(uint32_t d_tmem, const float4* a_hi, const float4* a_lo, const float4* w_hi, const float4* w_lo,
                                         int ksteps, uint64_t* bar) {
    constexpr uint32_t idesc = tc_idesc(N);
    uint32_t acc = 0;
    const float4* aop[3] = {a_lo, a_hi, a_hi};
    const float4* bop[3] = {w_hi, w_lo, w_hi};
#pragma unroll
    for (int t = 0; t < 3; ++t)
        for (int j = 0; j < ksteps; ++j) {
            tc_mma(d_tmem, tc_desc(aop[t] + 2 * j * TC_ROWS, TC_ROWS * 16), tc_desc(bop[t] + 2 * j * N, N * 16), idesc, acc);
            acc = 1;
        }
    tc_commit(bar);
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// weights of one phase as a K-major B operand: row n of the image = column n of the (k x n) weight block
template <class F>
__device__ __forceinline__ void tc_stage_weights(float4* w_hi, float4* w_lo, int N, F weight /* (k, n) -> float */) {
    for (int i = threadIdx.x; i < TC_KC * N; i += TC_TPB) {
        const int c = i / N, n = i - c * N;
        float hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) tc_split(weight(4 * c + q, n), hi[q], lo[q]);
        w_hi[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        w_lo[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// Iteration 0 of the fused decision path without a materialised h_0 / A_0: a link-state row of candidate graph k of
// decision b is [util(b,e), cap(e)/maxCap, on_path(b,k,e) ? bw(b)/cap(e) : 0, 0 x 17] (mark_action_sp + get_graph_features,
// environment16.py:711-725, script_eval_on_single_topology.py:131-160), so both the tile's own rows and the operand
// rows its gather needs are rebuilt from three small L1/L2-resident tables and one bit mask per candidate instead
// of being written by a feature kernel (160 B per row) and read back (80 B per row + 80 B per pair).  The message
// operand A_0 of a row takes one of two values per (decision, link) -- marked or not -- so those are tabulated
// (`ptab`, 160 B per link and decision instead of 80 B per link, candidate and decision) and the gather reads the one
// the mask bit selects.
struct FeatSrc {
    const float* util;        // [B][E]   fp32(load / cap)
    const float* bwcap;       // [B][E]   fp32(bw / cap)
    const float* capf;        // [E]      fp32(cap / maxCap)
    const uint32_t* mask;     // [B][Kmax][W] bit e = link e is on path(s, mid_k) U path(mid_k, d)
    const float* ptab;        // [B][E][2][20] A_0 of link e of decision b, unmarked / marked
    int W, Kmax;
};
__device__ __forceinline__ void tc_feature_row(const FeatSrc& f, const IdxTopo& idx, int64_t r, float& f0, float& f1, float& f2) {
    const unsigned b = (unsigned)r / (unsigned)idx.RPD;
    const unsigned lr = (unsigned)r - b * (unsigned)idx.RPD;
    const unsigned k = lr / (unsigned)idx.E, e = lr - k * (unsigned)idx.E;
    const unsigned be = b * (unsigned)idx.E + e;
    f0 = f.util[be];
    f1 = f.capf[e];
    const uint32_t m = f.mask[((size_t)b * f.Kmax + k) * f.W + (e >> 5)];
    f2 = ((m >> (e & 31)) & 1u) ? f.bwcap[be] : 0.f;
}

// KH: leading link-state columns that can be non-zero on entry (3 for iteration 0 of the fused path: phase 1 then
// needs one k-step instead of three).  LAST: no A' (phase 3 skipped).
//
// A CTA holds TC_GROUPS independent groups of 128 threads; a group owns one 128-row tile at a time (thread <-> row <->
// TMEM lane), has its own operand tiles, TMEM columns, mbarriers and named barrier, and walks its tiles
// (group index + k * number of groups) on its own -- the groups of an SM drift apart, so one group's memory phase
// overlaps another's tensor-core or gate phase.  Rows in flight per SM are bounded by TMEM: 124 accumulator columns
// per row tile -> 4 tiles of 128 rows in the 512 columns.
__device__ __forceinline__ void tc_group_sync(int g) { asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "r"(TC_ROWS) : "memory"); }

// FUSED0 (with KH = 3): iteration 0 rebuilt from the feature tables (`feat`), h_in / A_in unused.
template <class IDX, bool LAST, int KH, bool ACC, bool FUSED0 = false>
__global__ void __launch_bounds__(TC_TPB, 2)
k_mp_tc(const IDX idx, const float* __restrict__ wflat, const float* __restrict__ h_in, const float* __restrict__ A_in,
        float* __restrict__ h_out, float* __restrict__ A_out, int ntiles, const FeatSrc feat = FeatSrc()) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    float4* w1_hi = reinterpret_cast<float4*>(tc_smem);
    float4* w1_lo = w1_hi + TC_KC * TC_N1;
    float4* w2_hi = w1_lo + TC_KC * TC_N1;
    float4* w2_lo = w2_hi + TC_KC * TC_N2;
    float4* w3_hi = w2_lo + TC_KC * TC_N2;
    float4* w3_lo = w3_hi + TC_KC * TC_N3;
    float* sbias = reinterpret_cast<float*>(w3_lo + TC_KC * TC_N3);            // msg bias [20] | gru bias [2][60]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sbias + 208);
    const int g = threadIdx.x / TC_ROWS, t = threadIdx.x - g * TC_ROWS, wg = t >> 5;      // group, row in tile, warp in group
    unsigned char* gbase = reinterpret_cast<unsigned char*>(s_tmem + 4) + (size_t)g * TC_GROUP_BYTES;
    float4* a_hi = reinterpret_cast<float4*>(gbase);
    float4* a_lo = a_hi + TC_KC * TC_ROWS;
    float* stage = reinterpret_cast<float*>(a_lo + TC_KC * TC_ROWS);          // the tile's link states, landed by TMA one tile ahead
    uint64_t* bar = reinterpret_cast<uint64_t*>(stage + TC_ROWS * H);          // [0] MMA completion, [1] TMA arrival

    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(TC_GROUPS * TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_stage_weights(w1_hi, w1_lo, TC_N1, [&](int k, int n) {
        return k >= H ? 0.f : (n < 60 ? wflat[W_GRU_R + k * 60 + n] : wflat[W_MSG_K + (H + k) * H + (n - 60)]); });
    tc_stage_weights(w2_hi, w2_lo, TC_N2, [&](int k, int n) { return (k >= H || n >= 60) ? 0.f : wflat[W_GRU_K + k * 60 + n]; });
    tc_stage_weights(w3_hi, w3_lo, TC_N3, [&](int k, int n) { return (k >= H || n >= H) ? 0.f : wflat[W_MSG_K + k * H + n]; });
    for (int i = threadIdx.x; i < 140; i += TC_TPB) sbias[i] = i < H ? wflat[W_MSG_B + i] : wflat[W_GRU_B + i - H];
    a_hi[5 * TC_ROWS + t] = make_float4(0.f, 0.f, 0.f, 0.f);                   // k = 20..23: never written again
    a_lo[5 * TC_ROWS + t] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm0 = *s_tmem + (uint32_t)(g * TC_TMEM_COLS);               // this group's columns
    const uint32_t tm = tm0 + ((uint32_t)(wg * 32) << 16);                     // ... and this warp's 32 TMEM lanes
    uint32_t parity = 0, parity_tma = 0;
    constexpr int KS1 = KH <= 8 ? 1 : 3;                                       // k-steps of phase 1

#ifdef ENERO_TC_TIMING
    long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tick = clock64();
#endif
    // Every tile is prepared one tile ahead (the memory phases are DRAM-latency bound at 16 warps per SM): the tile's
    // own 128 link states arrive by one 1-D TMA bulk copy into `stage`, and the message operands its rows will gather
    // are pulled into L2.  Tiles are dealt round-robin to the groups; a tile without a live row (padding past nK*E
    // of a decision) is skipped by every thread of the group on its own (the test is uniform).
    const int64_t total = idx.rows();
    const int stride = (int)gridDim.x * TC_GROUPS;
    auto live_from = [&](int cand) {
        while (cand < ntiles && !idx.tile_live((int64_t)cand * TC_ROWS, TC_ROWS)) cand += stride;
        return cand;
    };
    auto prepare = [&](int tl, RowRef& r) {          // all threads: prefetches of tile tl, its TMA by thread 0 of the group
        r.ok = false; r.pb = r.pe = 0;
        if (tl >= ntiles) return;
        const int64_t r0 = (int64_t)tl * TC_ROWS;
        idx.ref(r0 + t, r);
        if (!FUSED0) {
            // one prefetch per operand row (its first sector; prefetching all three sectors of the 80-byte row was
            // measured 2.5 % slower: the extra instructions cost more than the remaining misses)
            for (int p = r.pb; p < r.pe; ++p)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(A_in + ((int64_t)r.list[p] + r.add) * H));
            if (t == 0) {
                const uint32_t bytes = (uint32_t)min((int64_t)TC_ROWS, total - r0) * H * 4;
                mbar_expect_tx(bar + 1, bytes);
                bulk_g2s(stage, h_in + r0 * H, bytes, bar + 1);
            }
        }
    };
    int tile = live_from((int)blockIdx.x * TC_GROUPS + g);
    RowRef ref;
    prepare(tile, ref);
    while (tile < ntiles) {
        const int64_t row = (int64_t)tile * TC_ROWS + t;
        const bool ok = ref.ok;
        const int next = live_from(tile + stride);
        float h[H];
#pragma unroll
        for (int j = 0; j < H; ++j) h[j] = 0.f;
        if (FUSED0) {
            if constexpr (FUSED0) { if (ok) tc_feature_row(feat, idx, row, h[0], h[1], h[2]); }
        } else {
            mbar_wait(bar + 1, parity_tma); parity_tma ^= 1u;
            if (ok) load_row(h, stage + t * H);
        }
        tc_store_operand<KH <= 8 ? 2 : 5>(a_hi, a_lo, t, h);
        fence_proxy_async();
        tc_fence_before();
        tc_group_sync(g);
        TC_TICK(0);
        if (t == 0) {
            tc_fence_after();
            tc_issue<TC_N1>(tm0 + TC_COL_Y, a_hi, a_lo, w1_hi, w1_lo, KS1, bar);
        }
        const RowRef cur = ref;                                     // `stage` has been read by everyone: the next tile may land
        prepare(next, ref);
        // ---- B = h . Wm[20:40] + bm (bias after the dot product, the op order of Dense)
        tc_mbar_wait(bar, parity); parity ^= 1u;
        tc_fence_after();
        TC_TICK(1);
        float B[H];
        tc_ld20(tm + TC_COL_B, B);
        u64 t2[H / 2], agg2[H / 2];
#pragma unroll
        for (int j = 0; j < H / 2; ++j) { t2[j] = pack2(B[2 * j] + sbias[2 * j], B[2 * j + 1] + sbias[2 * j + 1]); agg2[j] = 0ull; }
        // ---- agg = sum over incoming pairs (ascending pair index) of selu(A[first] + B); two operand rows in flight
        if (FUSED0) {
            if constexpr (FUSED0) {
                // source row = link e of candidate k of the row's own decision: A_0 from the decision's table
                const float* ptab = feat.ptab + (size_t)cur.b * idx.E * (2 * H);
                const uint32_t* mrow = feat.mask + (size_t)cur.b * feat.Kmax * feat.W;
                const bool aligned = idx.off_f == idx.E;            // (else the old_cummax quirk moves sources across graphs)
                const unsigned k0 = aligned ? cur.base / (unsigned)idx.E : 0u;
                auto src = [&](int p) {
                    unsigned k = k0, e = (unsigned)cur.list[p];
                    if (!aligned) { const unsigned lr = cur.base + e; k = lr / (unsigned)idx.E; e = lr - k * (unsigned)idx.E; }
                    const uint32_t bit = (mrow[k * feat.W + (e >> 5)] >> (e & 31)) & 1u;
                    return ptab + (e * 2u + bit) * H;
                };
                int p = cur.pb;
                for (; p + 1 < cur.pe; p += 2) {
                    float a0[H], a1[H];
                    const float* s0 = src(p); const float* s1 = src(p + 1);
                    load_row(a0, s0); load_row(a1, s1);
                    mp_accumulate<ACC>(agg2, a0, t2);
                    mp_accumulate<ACC>(agg2, a1, t2);
                }
                if (p < cur.pe) {
                    float a0[H];
                    load_row(a0, src(p));
                    mp_accumulate<ACC>(agg2, a0, t2);
                }
            }
        } else {
            // two operand rows in flight.  Tried and dropped at the 128-register budget of two CTAs per SM: resolving the
            // source rows a tile ahead (four more registers live across the tile) 1.74 ms per launch, and requesting
            // rows before the wait on the phase-1 MMAs / four at once 1.93 ms, against 1.41 ms -- both spill in the hot loop.
            int p = cur.pb;
            for (; p + 1 < cur.pe; p += 2) {
                float a0[H], a1[H];
                load_row(a0, A_in + ((int64_t)cur.list[p] + cur.add) * H);
                load_row(a1, A_in + ((int64_t)cur.list[p + 1] + cur.add) * H);
                mp_accumulate<ACC>(agg2, a0, t2);
                mp_accumulate<ACC>(agg2, a1, t2);
            }
            if (p < cur.pe) {
                float a0[H];
                load_row(a0, A_in + ((int64_t)cur.list[p] + cur.add) * H);
                mp_accumulate<ACC>(agg2, a0, t2);
            }
        }
        {
            float agg[H];
#pragma unroll
            for (int j = 0; j < H / 2; ++j) unpack2(agg2[j], agg[2 * j], agg[2 * j + 1]);
            tc_store_operand<5>(a_hi, a_lo, t, agg);                // phase-1 MMAs have completed (the barrier fired)
        }
        fence_proxy_async();
        tc_fence_before();
        tc_group_sync(g);
        TC_TICK(2);
        if (t == 0) {
            tc_fence_after();
            tc_issue<TC_N2>(tm0 + TC_COL_X, a_hi, a_lo, w2_hi, w2_lo, 3, bar);
        }
        // ---- GRUCell (reset_after=True): z, r, candidate, blend -- four outputs at a time; h' replaces h in place
        tc_mbar_wait(bar, parity); parity ^= 1u;
        tc_fence_after();
        TC_TICK(3);
#pragma unroll
        for (int jb = 0; jb < H / 4; ++jb) {
            float yz[4], yr[4], yh[4], xz[4], xr[4], xh[4];
            tc_ld4(tm + TC_COL_Y + 4 * jb, yz); tc_ld4(tm + TC_COL_Y + 20 + 4 * jb, yr); tc_ld4(tm + TC_COL_Y + 40 + 4 * jb, yh);
            tc_ld4(tm + TC_COL_X + 4 * jb, xz); tc_ld4(tm + TC_COL_X + 20 + 4 * jb, xr); tc_ld4(tm + TC_COL_X + 40 + 4 * jb, xh);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 4 * jb + i;
                const float* b0 = sbias + H; const float* b1 = sbias + H + 60;
                const float z = Mth<ACC>::sigmoid((xz[i] + b0[j]) + (yz[i] + b1[j]));
                const float r = Mth<ACC>::sigmoid((xr[i] + b0[20 + j]) + (yr[i] + b1[20 + j]));
                const float cand = Mth<ACC>::tanh_((xh[i] + b0[40 + j]) + r * (yh[i] + b1[40 + j]));
                h[j] = z * h[j] + (1.f - z) * cand;
            }
        }
        if (ok) store_row(h_out + row * H, h);
        if (!LAST) {
            tc_store_operand<5>(a_hi, a_lo, t, h);
            fence_proxy_async();
            tc_fence_before();
            tc_group_sync(g);
            TC_TICK(4);
            if (t == 0) {
                tc_fence_after();
                tc_issue<TC_N3>(tm0 + TC_COL_A, a_hi, a_lo, w3_hi, w3_lo, 3, bar);
            }
            tc_mbar_wait(bar, parity); parity ^= 1u;
            tc_fence_after();
            TC_TICK(5);
            float an[H];
            tc_ld20(tm + TC_COL_A, an);
            if (ok) store_row(A_out + row * H, an);
        }
        // the next tile's phase 1 overwrites TMEM columns and operand chunks this tile has finished reading:
        // its leading group barrier (after every thread's wait::ld + fence::before_thread_sync) orders that
        TC_TICK(6);
        tile = next;
    }
#ifdef ENERO_TC_TIMING
    if (t == 0) for (int i = 0; i < 8; ++i) atomicAdd(&tc_dbg[i], (unsigned long long)dbg[i]);
#endif
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*s_tmem), "r"(TC_GROUPS * TC_TMEM_COLS));
}

}  // namespace enero
