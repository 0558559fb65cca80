// C-ABI glue (include/enero_b200.h): contexts, weights, device mirrors of the topology,
// the signature-compatible GNN path, the fused decision batch and the device-resident
// episode / Local-Search batch.  No CPU fallback: every entry point below that computes
// launches the kernels in gnn_kernels.cuh / env_kernels.cuh or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "enero_internal.h"
#include "env_kernels.cuh"
#include "gnn_kernels.cuh"
#include "mp_tc_kernel.cuh"
#include "sap_kernels.cuh"
#include "train_kernels.cuh"
#include "ubench.cuh"

using namespace enero;

#include "ctx.cuh"

void train_release_stores(enero_ctx* c);
void train_release_lanes(enero_ctx* c);

// ---------------------------------------------------------------------------------------
extern "C" int enero_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int enero_ctx_create(int device, enero_ctx** out) {
    if (!out) return fail(ENERO_ERR_INVALID, "enero_ctx_create: null out");
    int n = 0;
    CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(ENERO_ERR_INVALID, "enero_ctx_create: no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    auto* c = new enero_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (int i = 0; i < 2; ++i) {
        for (float** p : {&c->w[i], &c->adam_m[i], &c->adam_v[i]}) {
            CUDA_TRY(cudaMalloc(p, sizeof(float) * ENERO_GRAD_STRIDE));
            CUDA_TRY(cudaMemset(*p, 0, sizeof(float) * ENERO_GRAD_STRIDE));
        }
    }
    CUDA_TRY(cudaMalloc(&c->gradbuf, sizeof(float) * ENERO_GRAD_FLOATS));
    CUDA_TRY(cudaMemset(c->gradbuf, 0, sizeof(float) * ENERO_GRAD_FLOATS));
    c->grad[0] = c->gradbuf; c->grad[1] = c->gradbuf + ENERO_GRAD_STRIDE;
    CUDA_TRY(cudaMalloc(&c->tile_ctr, 16 * sizeof(int)));
    {
        int r = ENERO_OK;
        auto setup = [&](auto kern, int& occ) {
            if (r != ENERO_OK) return;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MP_SMEM_BYTES) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MP_TPB, MP_SMEM_BYTES) != cudaSuccess) {
                enero_set_error(std::string("k_mp_iter setup: ") + cudaGetErrorString(cudaGetLastError()));
                r = ENERO_ERR_CUDA;
            }
        };
        int tmp = 0;
        // [math][indexer][last]
        setup(k_mp_iter<IdxExplicit, false, 20, false>, c->occ_iter[0][0][0]);
        setup(k_mp_iter<IdxExplicit, true, 20, false>, c->occ_iter[0][0][1]);
        setup(k_mp_iter<IdxTopo, false, 20, false>, c->occ_iter[0][1][0]);
        setup(k_mp_iter<IdxTopo, true, 20, false>, c->occ_iter[0][1][1]);
        setup(k_mp_iter<IdxTopo, false, 3, false>, tmp);
        setup(k_mp_iter<IdxExplicit, false, 20, true>, c->occ_iter[1][0][0]);
        setup(k_mp_iter<IdxExplicit, true, 20, true>, c->occ_iter[1][0][1]);
        setup(k_mp_iter<IdxTopo, false, 20, true>, c->occ_iter[1][1][0]);
        setup(k_mp_iter<IdxTopo, true, 20, true>, c->occ_iter[1][1][1]);
        setup(k_mp_iter<IdxTopo, false, 3, true>, tmp);
        if (r != ENERO_OK) return r;
    }
    CUDA_TRY(cudaFuncSetAttribute(k_bwd_gates<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWG_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_bwd_gates<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWG_SMEM_BYTES));
    {
        int r = ENERO_OK;
        auto setup_tc = [&](auto kern) {
            if (r == ENERO_OK && (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess ||
                                  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)) {
                enero_set_error(std::string("k_mp_tc setup: ") + cudaGetErrorString(cudaGetLastError()));
                r = ENERO_ERR_CUDA;
            }
        };
        setup_tc(k_mp_tc<IdxTopo, false, 20, false>); setup_tc(k_mp_tc<IdxTopo, true, 20, false>); setup_tc(k_mp_tc<IdxTopo, false, 3, false>);
        setup_tc(k_mp_tc<IdxTopo, false, 20, true>); setup_tc(k_mp_tc<IdxTopo, true, 20, true>); setup_tc(k_mp_tc<IdxTopo, false, 3, true>);
        setup_tc(k_mp_tc<IdxTopo, false, 3, false, true>); setup_tc(k_mp_tc<IdxTopo, false, 3, true, true>);
        setup_tc(k_mp_tc<IdxExplicit, false, 20, false>); setup_tc(k_mp_tc<IdxExplicit, true, 20, false>);
        setup_tc(k_mp_tc<IdxExplicit, false, 20, true>); setup_tc(k_mp_tc<IdxExplicit, true, 20, true>);
        if (r != ENERO_OK) return r;
        // resident CTAs per SM from the kernel's own resources (registers, shared memory, 128 of the 512 TMEM columns each):
        // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for this kernel on CUDA 12.9, which is not what the SM does
        cudaFuncAttributes fa;
        CUDA_TRY(cudaFuncGetAttributes(&fa, k_mp_tc<IdxTopo, false, 20, false>));
        const int by_regs = 65536 / (std::max(1, fa.numRegs) * TC_TPB);
        const int by_smem = (int)((size_t)prop.sharedMemPerMultiprocessor / (TC_SMEM_BYTES + 1024));
        c->occ_tc = std::max(1, std::min({by_regs, by_smem, 512 / (TC_GROUPS * TC_TMEM_COLS)}));
    }
    c->mp_engine = ENERO_MP_DEFAULT;
    if (const char* e = std::getenv("ENERO_MP")) {
        if (!std::strcmp(e, "ffma")) c->mp_engine = 0;
        else if (!std::strcmp(e, "tc")) c->mp_engine = 1;
        else return fail(ENERO_ERR_INVALID, "ENERO_MP must be 'ffma' or 'tc'");
    }
    if (const char* e = std::getenv("ENERO_NO_GRAPHS")) c->no_graphs = std::atoi(e) != 0;
    if (const char* e = std::getenv("ENERO_UNFUSED_FEATURES")) c->prof_unfused = std::atoi(e) != 0;
    if (const char* e = std::getenv("ENERO_FUSED_TAIL")) c->tail_mode = std::atoi(e) != 0;
    c->math_acc = ENERO_MATH_DEFAULT;
    if (const char* e = std::getenv("ENERO_MATH")) {
        if (!std::strcmp(e, "fast")) c->math_acc = 0;
        else if (!std::strcmp(e, "accurate")) c->math_acc = 1;
        else return fail(ENERO_ERR_INVALID, "ENERO_MATH must be 'fast' or 'accurate'");
    }
    *out = c;
    return ENERO_OK;
}

extern "C" int enero_ctx_destroy(enero_ctx* c) {
    if (!c) return ENERO_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 2; ++i) { cudaFree(c->w[i]); cudaFree(c->adam_m[i]); cudaFree(c->adam_v[i]); }
    cudaFree(c->gradbuf); cudaFree(c->tile_ctr);
    train_release_stores(c);
    train_release_lanes(c);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return ENERO_OK;
}

extern "C" int enero_ctx_sync(enero_ctx* c) {
    if (!c) return fail(ENERO_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}
extern "C" int64_t enero_ctx_launch_count(enero_ctx* c) { return c ? c->launches.load() : (int64_t)-1; }
extern "C" void* enero_ctx_stream(enero_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int enero_ctx_set_stream(enero_ctx* c, void* stream) {
    if (!c) return fail(ENERO_ERR_INVALID, "enero_ctx_set_stream: null ctx");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));                  // nothing of ours is left in flight on the old stream
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    c->tile_ctr_next = 16;
    return ENERO_OK;
}

extern "C" int enero_ctx_set_math(enero_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 1) return fail(ENERO_ERR_INVALID, "enero_ctx_set_math: mode must be ENERO_MATH_FAST or ENERO_MATH_ACCURATE");
    c->math_acc = mode;
    return ENERO_OK;
}
extern "C" int enero_ctx_get_math(enero_ctx* c) { return c ? c->math_acc : -1; }
extern "C" int enero_ctx_set_engine(enero_ctx* c, int engine) {
    if (!c || engine < 0 || engine > 1) return fail(ENERO_ERR_INVALID, "enero_ctx_set_engine: engine must be ENERO_MP_FFMA or ENERO_MP_TC");
    c->mp_engine = engine;
    return ENERO_OK;
}
extern "C" int enero_ctx_get_engine(enero_ctx* c) { return c ? c->mp_engine : -1; }
#ifdef ENERO_TC_TIMING
// debugging aid (build with -DENERO_TC_TIMING): cycles thread 0 of every k_mp_tc CTA spent per segment, summed; reading resets
extern "C" int enero_debug_tc_counters(enero_ctx* c, unsigned long long* out) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(out, tc_dbg, 64));
    out[7] = (unsigned long long)c->occ_tc;
    unsigned long long z[8] = {};
    CUDA_TRY(cudaMemcpyToSymbol(tc_dbg, z, 64));
    return ENERO_OK;
}
#endif

extern "C" int enero_ctx_profile_begin(enero_ctx* c) {
    if (!c) return fail(ENERO_ERR_INVALID, "null ctx");
    for (auto& e : c->prof_ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    c->prof_ev.clear();
    c->prof = true;
    return ENERO_OK;
}
extern "C" int enero_ctx_profile_end(enero_ctx* c, double* mp_iter_ms, int64_t* mp_iter_launches) {
    if (!c) return fail(ENERO_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double tot = 0;
    for (auto& e : c->prof_ev) {
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e.first, e.second));
        tot += ms;
        cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    if (mp_iter_ms) *mp_iter_ms = tot;
    if (mp_iter_launches) *mp_iter_launches = (int64_t)c->prof_ev.size();
    c->prof_ev.clear();
    c->prof = false;
    return ENERO_OK;
}

extern "C" int enero_ctx_fp32_peak(enero_ctx* c, double* tflops) {
    if (!c || !tflops) return fail(ENERO_ERR_INVALID, "enero_ctx_fp32_peak: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const int blocks = c->num_sms * 8, tpb = 256;
    DevBuf out;
    TRY(out.ensure((size_t)blocks * tpb * 4));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {                       // first repetition warms up; best of the rest
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        k_ubench_ffma2<<<blocks, tpb, 0, c->stream>>>(out.as<float>(), 1.0001f, 0.5f);
        LAUNCH_CHECK(c);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = 2.0 * 16 * UB_ITERS * (double)blocks * tpb / (best * 1e-3) / 1e12;
    return ENERO_OK;
}

extern "C" int enero_weights_upload(enero_ctx* c, int which, const float* flat) {
    if (!c || !flat || which < 0 || which > 1) return fail(ENERO_ERR_INVALID, "enero_weights_upload: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(c->w[which], flat, sizeof(float) * ENERO_NPARAMS, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}
extern "C" int enero_weights_download(enero_ctx* c, int which, float* flat) {
    if (!c || !flat || which < 0 || which > 1) return fail(ENERO_ERR_INVALID, "enero_weights_download: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(flat, c->w[which], sizeof(float) * ENERO_NPARAMS, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

// ---- topology mirrors -------------------------------------------------------------------
template <class T>
static int upload_vec(T** dst, const std::vector<T>& v) {
    CUDA_TRY(cudaMalloc(dst, std::max<size_t>(1, v.size()) * sizeof(T)));
    if (!v.empty()) CUDA_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ENERO_OK;
}
static void topo_free_dev(enero_topo* t) {
    if (t->dev_id < 0) return;
    cudaSetDevice(t->dev_id);
    TopoDev& d = t->dev;
    cudaFree(d.cap); cudaFree(d.capf); cudaFree(d.in_off); cudaFree(d.in_first);
    cudaFree(d.sp_off); cudaFree(d.sp_edge); cudaFree(d.mid_off); cudaFree(d.mids); cudaFree(d.out_off); cudaFree(d.out_second);
    cudaFree(d.cross_off); cudaFree(d.cross_sd);
    d = TopoDev();
    cudaFree(t->d_ksp_pair_off); cudaFree(t->d_ksp_edge_off); cudaFree(t->d_ksp_edge);
    t->d_ksp_pair_off = t->d_ksp_edge_off = t->d_ksp_edge = nullptr;
    t->ksp_dev_stale = true;
    t->dev_id = -1;
}
extern "C" int enero_topo_destroy(enero_topo* t) {
    if (!t) return ENERO_OK;
    topo_free_dev(t);
    delete t;
    return ENERO_OK;
}
extern "C" int enero_topo_upload(enero_ctx* c, enero_topo* t) {
    if (!c || !t) return fail(ENERO_ERR_INVALID, "enero_topo_upload: null argument");
    if (t->dev_id == c->device) return ENERO_OK;
    topo_free_dev(t);
    CUDA_TRY(cudaSetDevice(c->device));
    TopoDev& d = t->dev;
    d.N = t->N; d.E = t->E; d.P = t->P; d.K = t->K; d.Kmax = t->Kmax; d.off_f = t->off_f; d.off_s = t->off_s;
    TRY(upload_vec(&d.cap, t->cap)); TRY(upload_vec(&d.capf, t->capf));
    TRY(upload_vec(&d.in_off, t->in_off)); TRY(upload_vec(&d.in_first, t->in_first));
    TRY(upload_vec(&d.sp_off, t->sp_off)); TRY(upload_vec(&d.sp_edge, t->sp_edge));
    TRY(upload_vec(&d.mid_off, t->mid_off)); TRY(upload_vec(&d.mids, t->mids));
    TRY(upload_vec(&d.out_off, t->out_off)); TRY(upload_vec(&d.out_second, t->out_second));
    TRY(upload_vec(&d.cross_off, t->cross_off)); TRY(upload_vec(&d.cross_sd, t->cross_sd));
    t->dev_id = c->device;
    return ENERO_OK;
}
static TopoView view_of(const enero_topo* t) {
    const TopoDev& d = t->dev;
    return TopoView{d.N, d.E, d.Kmax, d.cap, d.capf, d.sp_off, d.sp_edge, d.mid_off, d.mids};
}

// ---- message passing driver ---------------------------------------------------------------

// a zeroed tile counter for the next k_mp_iter launch (re-zeroed in batches of 16 on the stream)
static int next_tile_counter(enero_ctx* c, int** out, TrainLane* lane = nullptr) {
    int*& base = lane ? lane->tile_ctr : c->tile_ctr;
    int& next = lane ? lane->tile_ctr_next : c->tile_ctr_next;
    if (next >= 16) {
        CUDA_TRY(cudaMemsetAsync(base, 0, 16 * sizeof(int), lane ? lane->st : c->stream));
        next = 0;
    }
    *out = base + next++;
    return ENERO_OK;
}

// one message-passing iteration: picks the instantiation (last / 3-column first iteration / math mode),
// sizes the persistent grid and, when profiling is on, brackets the launch with CUDA events
template <class IDX>
static int launch_iteration(enero_ctx* c, const IDX& idx, int idx_kind, int which, bool last, bool sparse, const float* hin,
                            const float* Ain, float* hout, float* Aout, TrainLane* lane = nullptr) {
    const int64_t ntiles = (idx.rows() + MP_WROWS - 1) / MP_WROWS;          // 64-row warp tiles
    const cudaStream_t stream = lane ? lane->st : c->stream;
    const bool prof = c->prof && !lane;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) {
        CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, stream));
    }
    int* ctr = nullptr;
    if (c->mp_engine != 1) TRY(next_tile_counter(c, &ctr, lane));
    const int occ = std::max(1, c->occ_iter[c->math_acc][idx_kind][last ? 1 : 0]);
    const int grid = (int)std::min<int64_t>((ntiles + MP_WARPS - 1) / MP_WARPS, (int64_t)c->num_sms * occ);
    auto go = [&](auto kern) { kern<<<grid, MP_TPB, MP_SMEM_BYTES, stream>>>(idx, c->w[which], hin, Ain, hout, Aout, ntiles, ctr); };
    constexpr bool TOPO = std::is_same<IDX, IdxTopo>::value;     // the 3-column first iteration exists for the fused path only
    if (c->mp_engine == 1) {                                     // tensor-core engine: 128-row tiles, two tile pipelines per CTA
        const int64_t nt = (idx.rows() + TC_ROWS - 1) / TC_ROWS;
        const int g = (int)std::min<int64_t>((nt + TC_GROUPS - 1) / TC_GROUPS, (int64_t)c->num_sms * std::max(1, c->occ_tc));
        auto gtc = [&](auto kern) { kern<<<g, TC_TPB, TC_SMEM_BYTES, stream>>>(idx, c->w[which], hin, Ain, hout, Aout, (int)nt, FeatSrc()); };
        if (c->math_acc) {
            if (last) gtc(k_mp_tc<IDX, true, 20, true>); else if (TOPO && sparse) gtc(k_mp_tc<IDX, false, TOPO ? 3 : 20, true>); else gtc(k_mp_tc<IDX, false, 20, true>);
        } else {
            if (last) gtc(k_mp_tc<IDX, true, 20, false>); else if (TOPO && sparse) gtc(k_mp_tc<IDX, false, TOPO ? 3 : 20, false>); else gtc(k_mp_tc<IDX, false, 20, false>);
        }
        LAUNCH_CHECK(c);
        if (prof) { CUDA_TRY(cudaEventRecord(e1, stream)); c->prof_ev.emplace_back(e0, e1); }
        return ENERO_OK;
    }
    if (c->math_acc) {
        if (last) go(k_mp_iter<IDX, true, 20, true>);
        else if (TOPO && sparse) go(k_mp_iter<IDX, false, TOPO ? 3 : 20, true>);
        else go(k_mp_iter<IDX, false, 20, true>);
    } else {
        if (last) go(k_mp_iter<IDX, true, 20, false>);
        else if (TOPO && sparse) go(k_mp_iter<IDX, false, TOPO ? 3 : 20, false>);
        else go(k_mp_iter<IDX, false, 20, false>);
    }
    LAUNCH_CHECK(c);
    if (prof) { CUDA_TRY(cudaEventRecord(e1, stream)); c->prof_ev.emplace_back(e0, e1); }
    return ENERO_OK;
}
static int launch_topo_iteration(enero_ctx* c, const IdxTopo& idx, int which, bool last, bool sparse, const float* hin,
                                 const float* Ain, float* hout, float* Aout, TrainLane* lane = nullptr) {
    return launch_iteration(c, idx, 1, which, last, sparse, hin, Ain, hout, Aout, lane);
}

// sparse_first: the rows of iteration 0 have at most 3 non-zero leading columns (fused feature path)
template <class IDX>
static int run_iterations(enero_ctx* c, const IDX& idx, int which, int T, int64_t rows, int idx_kind, bool sparse_first = false) {
    if (rows >= (int64_t)1 << 31) return fail(ENERO_ERR_INVALID, "more than 2^31 link-state rows in one launch: split the batch");
    for (int it = 0; it < T; ++it) {
        const bool last = (it == T - 1);
        TRY(launch_iteration(c, idx, idx_kind, which, last, it == 0 && sparse_first && !last && idx_kind == 1,
                             c->h[it & 1].as<float>(), c->A[it & 1].as<float>(), c->h[(it + 1) & 1].as<float>(),
                             c->A[(it + 1) & 1].as<float>()));
    }
    return ENERO_OK;
}

static int ensure_rows(enero_ctx* c, int64_t rows) {
    const size_t bytes = (size_t)rows * H * sizeof(float);
    for (int i = 0; i < 2; ++i) { TRY(c->h[i].ensure(bytes)); TRY(c->A[i].ensure(bytes)); }
    return ENERO_OK;
}

// ---- signature-compatible GNN forward ---------------------------------------------------------
extern "C" int enero_gnn_forward(enero_ctx* c, int which, const float* link_state, const int32_t* graph_id,
                                 const int32_t* first, const int32_t* second, int64_t n, int64_t p, int G,
                                 int T, float* out, int ptr_kind) {
    ENERO_TRACE("enero_gnn_forward");
    if (!c || !link_state || !out || (p > 0 && (!first || !second))) return fail(ENERO_ERR_INVALID, "enero_gnn_forward: null argument");
    if (which < 0 || which > 1 || n < 1 || p < 0 || G < 1 || T < 1 || n > 0x7fffffff || p > 0x7fffffff)
        return fail(ENERO_ERR_INVALID, "enero_gnn_forward: bad sizes");
    if (!graph_id && G != 1) return fail(ENERO_ERR_INVALID, "enero_gnn_forward: graph_id == NULL means a single graph");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    TRY(ensure_rows(c, n));
    const int* d_first = first; const int* d_second = second; const int* d_gid = graph_id;
    float* d_out = out;
    if (ptr_kind == ENERO_HOST) {
        CUDA_TRY(cudaMemcpyAsync(c->h[0].p, link_state, (size_t)n * H * 4, cudaMemcpyHostToDevice, st));
        TRY(c->x_first.ensure((size_t)std::max<int64_t>(1, p) * 4)); TRY(c->x_second.ensure((size_t)std::max<int64_t>(1, p) * 4));
        if (p) {
            CUDA_TRY(cudaMemcpyAsync(c->x_first.p, first, (size_t)p * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(c->x_second.p, second, (size_t)p * 4, cudaMemcpyHostToDevice, st));
        }
        d_first = c->x_first.as<int>(); d_second = c->x_second.as<int>();
        if (graph_id) {
            TRY(c->x_gid.ensure((size_t)n * 4));
            CUDA_TRY(cudaMemcpyAsync(c->x_gid.p, graph_id, (size_t)n * 4, cudaMemcpyHostToDevice, st));
            d_gid = c->x_gid.as<int>();
        }
        TRY(c->x_out.ensure((size_t)G * 4));
        d_out = c->x_out.as<float>();
    } else {
        CUDA_TRY(cudaMemcpyAsync(c->h[0].p, link_state, (size_t)n * H * 4, cudaMemcpyDeviceToDevice, st));
    }
    // CSR by destination row, stable in pair order
    TRY(c->x_cnt.ensure((size_t)(n + 1) * 4)); TRY(c->x_off.ensure((size_t)(n + 1) * 4)); TRY(c->x_cur.ensure((size_t)(n + 1) * 4));
    TRY(c->x_slot.ensure((size_t)std::max<int64_t>(1, p) * 4)); TRY(c->x_src.ensure((size_t)std::max<int64_t>(1, p) * 4));
    TRY(c->x_goff.ensure((size_t)(G + 1) * 4)); TRY(c->x_err.ensure(4));
    CUDA_TRY(cudaMemsetAsync(c->x_cnt.p, 0, (size_t)(n + 1) * 4, st));
    CUDA_TRY(cudaMemsetAsync(c->x_cur.p, 0, (size_t)(n + 1) * 4, st));
    CUDA_TRY(cudaMemsetAsync(c->x_err.p, 0, 4, st));
    const int gp = (int)std::min<int64_t>(4096, (std::max<int64_t>(p, 1) + 255) / 256);
    const int gn = (int)std::min<int64_t>(4096, (n + 256) / 256);
    if (p) { k_csr_count<<<gp, 256, 0, st>>>(d_second, p, n, c->x_cnt.as<int>(), c->x_err.as<int>()); LAUNCH_CHECK(c); }
    k_scan_excl<<<1, 1024, 0, st>>>(c->x_cnt.as<int>(), n, c->x_off.as<int>()); LAUNCH_CHECK(c);
    if (p) {
        k_csr_fill<<<gp, 256, 0, st>>>(d_second, p, n, c->x_off.as<int>(), c->x_cur.as<int>(), c->x_slot.as<int>()); LAUNCH_CHECK(c);
        k_csr_sort<<<gn, 256, 0, st>>>(d_first, n, c->x_off.as<int>(), c->x_slot.as<int>(), c->x_src.as<int>(), c->x_err.as<int>()); LAUNCH_CHECK(c);
    }
    if (d_gid) { k_graph_offsets<<<gn, 256, 0, st>>>(d_gid, n, G, c->x_goff.as<int>(), c->x_err.as<int>()); LAUNCH_CHECK(c); }
    else { const int go[2] = {0, (int)n}; CUDA_TRY(cudaMemcpyAsync(c->x_goff.p, go, 8, cudaMemcpyHostToDevice, st)); CUDA_TRY(cudaStreamSynchronize(st)); }
    k_transform0<<<(int)std::min<int64_t>((n + MP_TPB - 1) / MP_TPB, 8 * c->num_sms), MP_TPB, 0, st>>>(c->w[which], c->h[0].as<float>(), c->A[0].as<float>(), n);
    LAUNCH_CHECK(c);
    IdxExplicit idx{n, c->x_off.as<int>(), c->x_src.as<int>()};
    TRY(run_iterations(c, idx, which, T, n, 0));
    GraphsExplicit gr{G, c->x_goff.as<int>()};
    k_readout<GraphsExplicit><<<(G + 3) / 4, 128, 0, st>>>(gr, c->w[which], c->h[T & 1].as<float>(), d_out); LAUNCH_CHECK(c);
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, c->x_err.p, 4, cudaMemcpyDeviceToHost, st));
    if (ptr_kind == ENERO_HOST) CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)G * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (err == 1) return fail(ENERO_ERR_INVALID, "enero_gnn_forward: first/second index out of range [0, n)");
    if (err == 2) return fail(ENERO_ERR_INVALID, "enero_gnn_forward: graph_id must be ascending in [0, n_graphs)");
    return ENERO_OK;
}

// ---- fused decision batch (device pointers) -------------------------------------------------------
// logits/probs [B,Kmax], nK/action [B]; skip: u8[B] or NULL
static int decide_dev(enero_ctx* c, enero_topo* t, int B, const double* loads, const int* sd, const double* bw,
                      const uint8_t* skip, float* logits, float* probs, int* nK, int* action, int T,
                      const EpisodeView* step_ep = nullptr) {
    const TopoDev& d = t->dev;
    const int RPD = d.Kmax * d.E;
    const int64_t rows = (int64_t)B * RPD;
    cudaStream_t st = c->stream;
    TRY(ensure_rows(c, rows));
    TRY(c->util.ensure((size_t)B * d.E * 4));
    TopoView tv = view_of(t);
    IdxTopo idx{rows, d.E, d.off_f, d.off_s, RPD, nK, d.in_off, d.in_first};
    const bool fused0 = c->mp_engine == 1 && T > 1 && !c->prof_unfused;
    if (!fused0) { k_decision_prep<<<(int)(((int64_t)B * d.E + 255) / 256), 256, 0, st>>>(tv, B, loads, sd, skip, nK, c->util.as<float>()); LAUNCH_CHECK(c); }
    if (fused0) {
        // tensor-core engine: iteration 0 rebuilds the feature rows from tables (no k_features, no h_0 / A_0 in HBM)
        const int W = (d.E + 31) / 32;
        TRY(c->bwcap.ensure((size_t)B * d.E * 4)); TRY(c->marks.ensure((size_t)B * d.Kmax * W * 4)); TRY(c->ptab.ensure((size_t)B * d.E * 2 * H * 4));
        const int64_t work = std::max<int64_t>((int64_t)B * d.E, (int64_t)B * d.Kmax);
        k_decision_tables<<<(int)((work + 255) / 256), 256, 0, st>>>(tv, B, W, loads, sd, bw, skip, c->w[ENERO_ACTOR], nK, c->util.as<float>(), c->bwcap.as<float>(),
                                                                     c->marks.as<uint32_t>(), c->ptab.as<float>()); LAUNCH_CHECK(c);
        FeatSrc fs{c->util.as<float>(), c->bwcap.as<float>(), d.capf, c->marks.as<uint32_t>(), c->ptab.as<float>(), W, d.Kmax};
        const int64_t nt = (rows + TC_ROWS - 1) / TC_ROWS;
        const int g = (int)std::min<int64_t>((nt + TC_GROUPS - 1) / TC_GROUPS, (int64_t)c->num_sms * std::max(1, c->occ_tc));
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (c->prof) { CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1)); CUDA_TRY(cudaEventRecord(e0, st)); }
        if (c->math_acc) k_mp_tc<IdxTopo, false, 3, true, true><<<g, TC_TPB, TC_SMEM_BYTES, st>>>(idx, c->w[ENERO_ACTOR], nullptr, nullptr, c->h[1].as<float>(), c->A[1].as<float>(), (int)nt, fs);
        else k_mp_tc<IdxTopo, false, 3, false, true><<<g, TC_TPB, TC_SMEM_BYTES, st>>>(idx, c->w[ENERO_ACTOR], nullptr, nullptr, c->h[1].as<float>(), c->A[1].as<float>(), (int)nt, fs);
        LAUNCH_CHECK(c);
        if (c->prof) { CUDA_TRY(cudaEventRecord(e1, st)); c->prof_ev.emplace_back(e0, e1); }
        for (int it = 1; it < T; ++it)
            TRY(launch_iteration(c, idx, 1, ENERO_ACTOR, it == T - 1, false, c->h[it & 1].as<float>(), c->A[it & 1].as<float>(),
                                 c->h[(it + 1) & 1].as<float>(), c->A[(it + 1) & 1].as<float>()));
    } else {
        k_features<<<(int)((rows + MP_TPB - 1) / MP_TPB), MP_TPB, 0, st>>>(tv, B, RPD, sd, bw, nK, c->util.as<float>(), c->w[ENERO_ACTOR],
                                                                         c->h[0].as<float>(), c->A[0].as<float>()); LAUNCH_CHECK(c);
        TRY(run_iterations(c, idx, ENERO_ACTOR, T, rows, 1, T > 1));
    }
    GraphsTopo gr{B, d.Kmax, d.E, RPD, nK};
    // read-out + soft-max + arg-max (+ Env16.step with the greedy action when the caller runs the episode on the device)
    // One launch for the tail wins while the step is latency bound (-2 % per decision at <= 16 episodes); from a few
    // hundred decisions on, three full-width launches beat CTAs whose warp 0 runs the serial part (+0.8 % at 4 096).
    const bool fused_tail = c->tail_mode < 0 ? B < 512 : c->tail_mode != 0;
    if (!fused_tail) {
        CUDA_TRY(cudaMemsetAsync(logits, 0, (size_t)B * d.Kmax * 4, st));
        k_readout<GraphsTopo><<<(B * d.Kmax + 3) / 4, 128, 0, st>>>(gr, c->w[ENERO_ACTOR], c->h[T & 1].as<float>(), logits); LAUNCH_CHECK(c);
        k_policy<<<(B + 3) / 4, 128, 0, st>>>(B, d.Kmax, nK, logits, probs, action); LAUNCH_CHECK(c);
        if (step_ep) { k_env_step<<<(B + 3) / 4, 128, 0, st>>>(tv, *step_ep, action); LAUNCH_CHECK(c); }
        return ENERO_OK;
    }
    const int ftpb = FIN_TPB;
    if (step_ep) k_decision_finish<true><<<B, ftpb, 0, st>>>(gr, c->w[ENERO_ACTOR], c->h[T & 1].as<float>(), logits, probs, action, tv, *step_ep);
    else k_decision_finish<false><<<B, ftpb, 0, st>>>(gr, c->w[ENERO_ACTOR], c->h[T & 1].as<float>(), logits, probs, action, tv, EpisodeView{});
    LAUNCH_CHECK(c);
    return ENERO_OK;
}

static int value_dev(enero_ctx* c, enero_topo* t, int B, const double* loads, float* values, int T) {
    const TopoDev& d = t->dev;
    const int64_t rows = (int64_t)B * d.E;
    cudaStream_t st = c->stream;
    TRY(ensure_rows(c, rows));
    TRY(c->util.ensure((size_t)B * d.E * 4));
    TRY(c->nK.ensure((size_t)B * 4));
    TRY(c->in_sd.ensure((size_t)B * 8));
    CUDA_TRY(cudaMemsetAsync(c->in_sd.p, 0xFF, (size_t)B * 8, st));       // sd = -1: prep yields nK = 0 ...
    TopoView tv = view_of(t);
    k_decision_prep<<<(int)(((int64_t)B * d.E + 255) / 256), 256, 0, st>>>(tv, B, loads, c->in_sd.as<int>(), nullptr, c->nK.as<int>(), c->util.as<float>()); LAUNCH_CHECK(c);
    // ... and the critic's "one graph per state" is nK = 1 everywhere
    if (c->ones.cap < (size_t)B * 4) {
        TRY(c->ones.ensure((size_t)B * 4));
        std::vector<int> one((size_t)c->ones.cap / 4, 1);
        CUDA_TRY(cudaMemcpyAsync(c->ones.p, one.data(), one.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    k_features_critic<<<(int)((rows + MP_TPB - 1) / MP_TPB), MP_TPB, 0, st>>>(tv, B, c->util.as<float>(), c->w[ENERO_CRITIC],
                                                                            c->h[0].as<float>(), c->A[0].as<float>()); LAUNCH_CHECK(c);
    IdxTopo idx{rows, d.E, d.off_f, d.off_s, d.E, c->ones.as<int>(), d.in_off, d.in_first};
    TRY(run_iterations(c, idx, ENERO_CRITIC, T, rows, 1, T > 1));
    GraphsTopo gr{B, 1, d.E, d.E, c->ones.as<int>()};
    k_readout<GraphsTopo><<<(B + 3) / 4, 128, 0, st>>>(gr, c->w[ENERO_CRITIC], c->h[T & 1].as<float>(), values); LAUNCH_CHECK(c);
    return ENERO_OK;
}

extern "C" int enero_decide_batch(enero_ctx* c, enero_topo* t, int B, const double* loads, const int32_t* sd,
                                  const double* bw, float* logits, float* probs, int32_t* nK, int32_t* action,
                                  int ptr_kind) {
    ENERO_TRACE("enero_decide_batch");
    if (!c || !t || !loads || !sd || !bw || B < 1) return fail(ENERO_ERR_INVALID, "enero_decide_batch: bad argument");
    if (t->dev_id != c->device) return fail(ENERO_ERR_STATE, "enero_decide_batch: topology not uploaded to this device");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int E = t->E, Kmax = t->Kmax;
    TRY(c->logits.ensure((size_t)B * Kmax * 4)); TRY(c->probs.ensure((size_t)B * Kmax * 4));
    TRY(c->nK.ensure((size_t)B * 4)); TRY(c->action.ensure((size_t)B * 4));
    if (ptr_kind == ENERO_HOST) {
        TRY(c->in_loads.ensure((size_t)B * E * 8)); TRY(c->in_sd.ensure((size_t)B * 8)); TRY(c->in_bw.ensure((size_t)B * 8));
        CUDA_TRY(cudaMemcpyAsync(c->in_loads.p, loads, (size_t)B * E * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->in_sd.p, sd, (size_t)B * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->in_bw.p, bw, (size_t)B * 8, cudaMemcpyHostToDevice, st));
        TRY(decide_dev(c, t, B, c->in_loads.as<double>(), c->in_sd.as<int>(), c->in_bw.as<double>(), nullptr,
                       c->logits.as<float>(), c->probs.as<float>(), c->nK.as<int>(), c->action.as<int>(), 5));
        if (logits) CUDA_TRY(cudaMemcpyAsync(logits, c->logits.p, (size_t)B * Kmax * 4, cudaMemcpyDeviceToHost, st));
        if (probs) CUDA_TRY(cudaMemcpyAsync(probs, c->probs.p, (size_t)B * Kmax * 4, cudaMemcpyDeviceToHost, st));
        if (nK) CUDA_TRY(cudaMemcpyAsync(nK, c->nK.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        if (action) CUDA_TRY(cudaMemcpyAsync(action, c->action.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        TRY(decide_dev(c, t, B, loads, sd, bw, nullptr, logits ? logits : c->logits.as<float>(),
                       probs ? probs : c->probs.as<float>(), nK ? nK : c->nK.as<int>(),
                       action ? action : c->action.as<int>(), 5));
    }
    return ENERO_OK;
}

extern "C" int enero_decide_value_batch(enero_ctx* c, enero_topo* t, int B, const double* loads, const int32_t* sd,
                                        const double* bw, float* probs, int32_t* nK, float* values) {
    ENERO_TRACE("enero_decide_value_batch");
    if (!c || !t || !loads || !sd || !bw || !probs || !nK || !values || B < 1) return fail(ENERO_ERR_INVALID, "enero_decide_value_batch: bad argument");
    if (t->dev_id != c->device) return fail(ENERO_ERR_STATE, "enero_decide_value_batch: topology not uploaded to this device");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int E = t->E, Kmax = t->Kmax;
    TRY(c->logits.ensure((size_t)B * Kmax * 4)); TRY(c->probs.ensure((size_t)B * Kmax * 4));
    TRY(c->nK.ensure((size_t)B * 4)); TRY(c->action.ensure((size_t)B * 4)); TRY(c->x_out.ensure((size_t)B * 4));
    TRY(c->in_loads.ensure((size_t)B * E * 8)); TRY(c->in_sd.ensure((size_t)B * 8)); TRY(c->in_bw.ensure((size_t)B * 8));
    TRY(c->x_gid.ensure((size_t)B * 4));      // nK of the actor pass (value_dev reuses c->nK for its own prologue)
    CUDA_TRY(cudaMemcpyAsync(c->in_loads.p, loads, (size_t)B * E * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c->in_sd.p, sd, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c->in_bw.p, bw, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    TRY(decide_dev(c, t, B, c->in_loads.as<double>(), c->in_sd.as<int>(), c->in_bw.as<double>(), nullptr,
                   c->logits.as<float>(), c->probs.as<float>(), c->x_gid.as<int>(), c->action.as<int>(), 5));
    CUDA_TRY(cudaMemcpyAsync(probs, c->probs.p, (size_t)B * Kmax * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nK, c->x_gid.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    TRY(value_dev(c, t, B, c->in_loads.as<double>(), c->x_out.as<float>(), 5));
    CUDA_TRY(cudaMemcpyAsync(values, c->x_out.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ENERO_OK;
}

extern "C" int enero_value_batch(enero_ctx* c, enero_topo* t, int B, const double* loads, float* values, int ptr_kind) {
    ENERO_TRACE("enero_value_batch");
    if (!c || !t || !loads || !values || B < 1) return fail(ENERO_ERR_INVALID, "enero_value_batch: bad argument");
    if (t->dev_id != c->device) return fail(ENERO_ERR_STATE, "enero_value_batch: topology not uploaded to this device");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    if (ptr_kind == ENERO_HOST) {
        TRY(c->in_loads.ensure((size_t)B * t->E * 8)); TRY(c->logits.ensure((size_t)B * 4));
        CUDA_TRY(cudaMemcpyAsync(c->in_loads.p, loads, (size_t)B * t->E * 8, cudaMemcpyHostToDevice, st));
        TRY(value_dev(c, t, B, c->in_loads.as<double>(), c->logits.as<float>(), 5));
        CUDA_TRY(cudaMemcpyAsync(values, c->logits.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        TRY(value_dev(c, t, B, loads, values, 5));
    }
    return ENERO_OK;
}

// ---- device-resident episode batch ---------------------------------------------------------------
// captured greedy decision step (enero_batch_run_greedy): executable graph + what its nodes baked in
struct GreedyGraph {
    cudaGraphExec_t exec = nullptr;
    const void* key[10] = {};
    int kernels = 0;
    void release() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; }
};

struct enero_batch {
    enero_ctx* c = nullptr; enero_topo* t = nullptr;
    GreedyGraph graph;
    int B = 0, Dmax = 0, NN = 0; double pct = 0.15;
    bool reset_done = false, policy_valid = false;
    DevBuf tm, loads, mid_cur, elig, elig_len, it, cur, cur_bw, done, steps, max_uti, max_edge, best, best_step, ospf,
        h_action, h_reward, h_max, last_reward, logits, probs, nK, action, values;
    DevBuf snap;                                    // enero_batch_snapshot image of the mutable state
    DevBuf sap_order, sap_loads, sap_mt, sap_idx, sap_max, sap_act;   // SAP baseline scratch
    DevBuf r_loads, r_sd, r_bw, r_action, r_prob, r_value;            // rollout records
    DevBuf crit_scratch, ls_routing, ls_nrouting;                      // critical-demand selection scratch
    DevBuf ls_loads, ls_mid, ls_elig, ls_elig_len, ls_val, ls_nmoves, ls_moves, ls_move_val, ls_max_edge;
    std::vector<int> h_cur; std::vector<uint8_t> h_done; bool mirror_valid = false;   // host mirror of (cur, done) for action checks
    std::vector<int> h_elig, h_elig_len;           // host mirror of list_eligible_demands
    std::vector<int> h_ls_elig, h_ls_elig_len; int ls_cap = 0;
    int max_len = 0;
};

// A topology has ONE device mirror (the one-process-per-GPU model): if another context moved it to a different
// device since this batch was created, every pointer the batch would hand to a kernel is stale -- refuse loudly.
static int check_batch_device(const enero_batch* b, const char* who) {
    if (b->t->dev_id != b->c->device)
        return fail(ENERO_ERR_STATE, std::string(who) + ": the topology's device mirror now lives on another GPU "
                    "(enero_topo_upload from a context of a different device); use one topology object per device");
    return ENERO_OK;
}

static EpisodeView ep_view(enero_batch* b) {
    EpisodeView v;
    v.B = b->B; v.Dmax = b->Dmax; v.tm = b->tm.as<double>(); v.loads = b->loads.as<double>(); v.mid_cur = b->mid_cur.as<int>();
    v.elig = b->elig.as<int>(); v.elig_len = b->elig_len.as<int>(); v.it = b->it.as<int>(); v.cur = b->cur.as<int>();
    v.cur_bw = b->cur_bw.as<double>(); v.done = b->done.as<uint8_t>(); v.steps = b->steps.as<int>();
    v.max_uti = b->max_uti.as<double>(); v.max_edge = b->max_edge.as<int>(); v.best = b->best.as<double>();
    v.best_step = b->best_step.as<int>(); v.ospf = b->ospf.as<double>(); v.h_action = b->h_action.as<int>();
    v.h_reward = b->h_reward.as<double>(); v.h_max = b->h_max.as<double>(); v.last_reward = b->last_reward.as<double>();
    return v;
}

extern "C" int enero_batch_create(enero_ctx* c, enero_topo* t, int B, double pct, enero_batch** out) {
    if (!c || !t || !out || B < 1 || !(pct > 0)) return fail(ENERO_ERR_INVALID, "enero_batch_create: bad argument");
    if (t->N < 3) return fail(ENERO_ERR_INVALID, "enero_batch_create: need N >= 3 (the finished episode's dummy demand is (1,2))");
    TRY(enero_topo_upload(c, t));
    CUDA_TRY(cudaSetDevice(c->device));
    auto* b = new enero_batch();
    b->c = c; b->t = t; b->B = B; b->pct = pct; b->NN = t->N * t->N;
    b->Dmax = (int)std::ceil((double)(t->N * (t->N - 1)) * pct);       // environment16.py:199
    const size_t BB = (size_t)B, D = (size_t)b->Dmax, NN = (size_t)b->NN, E = (size_t)t->E, K = (size_t)t->Kmax;
    int r = ENERO_OK;
    auto A = [&](DevBuf& buf, size_t bytes) { if (r == ENERO_OK) r = buf.ensure(std::max<size_t>(bytes, 16)); };
    A(b->tm, BB * NN * 8); A(b->loads, BB * E * 8); A(b->mid_cur, BB * NN * 4); A(b->elig, BB * D * 4); A(b->elig_len, BB * 4);
    A(b->it, BB * 4); A(b->cur, BB * 8); A(b->cur_bw, BB * 8); A(b->done, BB); A(b->steps, BB * 4); A(b->max_uti, BB * 8);
    A(b->max_edge, BB * 4); A(b->best, BB * 8); A(b->best_step, BB * 4); A(b->ospf, BB * 8); A(b->h_action, BB * D * 4);
    A(b->h_reward, BB * D * 8); A(b->h_max, BB * D * 8); A(b->last_reward, BB * 8); A(b->logits, BB * K * 4);
    A(b->probs, BB * K * 4); A(b->nK, BB * 4); A(b->action, BB * 4); A(b->values, BB * 4);
    A(b->ls_loads, BB * E * 8); A(b->ls_mid, BB * NN * 4); A(b->ls_val, BB * 8); A(b->ls_nmoves, BB * 4); A(b->ls_max_edge, BB * 4);
    if (r != ENERO_OK) { delete b; return r; }
    *out = b;
    return ENERO_OK;
}

extern "C" int enero_batch_destroy(enero_batch* b) {
    if (!b) return ENERO_OK;
    cudaSetDevice(b->c->device);
    cudaStreamSynchronize(b->c->stream);
    b->graph.release();
    delete b;
    return ENERO_OK;
}

extern "C" int enero_batch_dmax(enero_batch* b) { return b ? b->Dmax : -1; }

extern "C" int enero_batch_reset(enero_batch* b, const double* tms) {
    ENERO_TRACE("enero_batch_reset");
    if (!b || !tms) return fail(ENERO_ERR_INVALID, "enero_batch_reset: null argument");
    TRY(check_batch_device(b, "enero_batch_reset"));
    enero_ctx* c = b->c; enero_topo* t = b->t;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int B = b->B, E = t->E, NN = b->NN, D = b->Dmax;
    CUDA_TRY(cudaMemcpyAsync(b->tm.p, tms, (size_t)B * NN * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(b->mid_cur.p, 0xFF, (size_t)B * NN * 4, st));
    TopoView tv = view_of(t);
    k_route_loads<<<B, 256, (size_t)E * 8, st>>>(tv, b->tm.as<double>(), nullptr, b->loads.as<double>(), b->max_uti.as<double>(), b->max_edge.as<int>());
    LAUNCH_CHECK(c);
    // critical-demand selection on the device (one CTA per episode); the host keeps a mirror of the lists for the
    // getters and the best-routing reconstruction
    TRY(b->crit_scratch.ensure((size_t)B * NN * 4));
    {
        CritView cv{t->dev.cross_off, t->dev.cross_sd, nullptr, nullptr, nullptr, 0, D, D, b->crit_scratch.as<int>()};
        const size_t smem = (size_t)E * 8 + (size_t)((NN + 31) / 32) * 4;
        k_critical_demands<<<B, 256, smem, st>>>(tv, cv, b->loads.as<double>(), b->tm.as<double>(), b->elig.as<int>(), b->elig_len.as<int>());
        LAUNCH_CHECK(c);
    }
    b->h_elig.assign((size_t)B * D, 0); b->h_elig_len.assign(B, 0);
    CUDA_TRY(cudaMemcpyAsync(b->h_elig.data(), b->elig.p, b->h_elig.size() * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(b->h_elig_len.data(), b->elig_len.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    b->max_len = 0;
    for (int i = 0; i < B; ++i) {
        if (b->h_elig_len[i] == 0) return fail(ENERO_ERR_INVALID, "enero_batch_reset: traffic matrix yields no eligible demand");
        b->max_len = std::max(b->max_len, b->h_elig_len[i]);
    }
    k_episode_begin<<<(B + 3) / 4, 128, 0, st>>>(tv, ep_view(b)); LAUNCH_CHECK(c);
    CUDA_TRY(cudaStreamSynchronize(st));
    b->reset_done = true; b->policy_valid = false; b->mirror_valid = false;
    return ENERO_OK;
}

// mutable episode state = (buffer, bytes) list, in a fixed order
static std::vector<std::pair<DevBuf*, size_t>> mutable_state(enero_batch* b) {
    const size_t B = b->B, NN = b->NN, E = b->t->E, D = b->Dmax;
    return {{&b->loads, B * E * 8}, {&b->mid_cur, B * NN * 4}, {&b->it, B * 4}, {&b->cur, B * 8}, {&b->cur_bw, B * 8},
            {&b->done, B}, {&b->steps, B * 4}, {&b->max_uti, B * 8}, {&b->max_edge, B * 4}, {&b->best, B * 8},
            {&b->best_step, B * 4}, {&b->ospf, B * 8}, {&b->h_action, B * D * 4}, {&b->h_reward, B * D * 8},
            {&b->h_max, B * D * 8}};
}
extern "C" int enero_batch_snapshot(enero_batch* b) {
    if (!b || !b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_snapshot: reset first");
    CUDA_TRY(cudaSetDevice(b->c->device));
    size_t tot = 0;
    for (auto& x : mutable_state(b)) tot += (x.second + 255) & ~size_t(255);
    TRY(b->snap.ensure(tot));
    size_t o = 0;
    for (auto& x : mutable_state(b)) {
        CUDA_TRY(cudaMemcpyAsync((char*)b->snap.p + o, x.first->p, x.second, cudaMemcpyDeviceToDevice, b->c->stream));
        o += (x.second + 255) & ~size_t(255);
    }
    return ENERO_OK;
}
extern "C" int enero_batch_restore(enero_batch* b) {
    if (!b || !b->snap.p) return fail(ENERO_ERR_STATE, "enero_batch_restore: no snapshot");
    CUDA_TRY(cudaSetDevice(b->c->device));
    size_t o = 0;
    for (auto& x : mutable_state(b)) {
        CUDA_TRY(cudaMemcpyAsync(x.first->p, (char*)b->snap.p + o, x.second, cudaMemcpyDeviceToDevice, b->c->stream));
        o += (x.second + 255) & ~size_t(255);
    }
    b->policy_valid = false; b->mirror_valid = false;
    return ENERO_OK;
}

extern "C" int enero_batch_policy(enero_batch* b, float* probs_host, int32_t* nK_host) {
    ENERO_TRACE("enero_batch_policy");
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_policy: null batch");
    TRY(check_batch_device(b, "enero_batch_policy"));
    if (!b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_policy: reset first");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(decide_dev(c, b->t, b->B, b->loads.as<double>(), b->cur.as<int>(), b->cur_bw.as<double>(), b->done.as<uint8_t>(),
                   b->logits.as<float>(), b->probs.as<float>(), b->nK.as<int>(), b->action.as<int>(), 5));
    b->policy_valid = true;
    if (probs_host) CUDA_TRY(cudaMemcpyAsync(probs_host, b->probs.p, (size_t)b->B * b->t->Kmax * 4, cudaMemcpyDeviceToHost, c->stream));
    if (nK_host) CUDA_TRY(cudaMemcpyAsync(nK_host, b->nK.p, (size_t)b->B * 4, cudaMemcpyDeviceToHost, c->stream));
    if (probs_host || nK_host) CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

extern "C" int enero_batch_value(enero_batch* b, float* values_host) {
    if (!b || !values_host) return fail(ENERO_ERR_INVALID, "enero_batch_value: null argument");
    TRY(check_batch_device(b, "enero_batch_value"));
    if (!b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_value: reset first");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(value_dev(c, b->t, b->B, b->loads.as<double>(), b->values.as<float>(), 5));
    CUDA_TRY(cudaMemcpyAsync(values_host, b->values.p, (size_t)b->B * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

// host-side check of a host action vector against the candidate counts of the current demands; uses the host
// mirror of (cur, done) when the last call left it valid, else fetches it (one extra stream round trip)
static int check_actions(enero_batch* b, const int32_t* actions) {
    enero_ctx* c = b->c;
    if (!b->mirror_valid) {
        b->h_cur.resize((size_t)b->B * 2); b->h_done.resize(b->B);
        CUDA_TRY(cudaMemcpyAsync(b->h_cur.data(), b->cur.p, b->h_cur.size() * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(b->h_done.data(), b->done.p, b->h_done.size(), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        b->mirror_valid = true;
    }
    for (int i = 0; i < b->B; ++i) {
        if (b->h_done[i]) continue;
        const int sdi = b->h_cur[2 * i] * b->t->N + b->h_cur[2 * i + 1];
        const int k = b->t->mid_off[sdi + 1] - b->t->mid_off[sdi];
        if (actions[i] < 0 || actions[i] >= k) return fail(ENERO_ERR_INVALID, "enero_batch_step: action out of range");
    }
    return ENERO_OK;
}

static int step_dev(enero_batch* b) {
    enero_ctx* c = b->c;
    k_env_step<<<(b->B + 3) / 4, 128, 0, c->stream>>>(view_of(b->t), ep_view(b), b->action.as<int>());
    LAUNCH_CHECK(c);
    b->policy_valid = false;
    return ENERO_OK;
}

extern "C" int enero_batch_step(enero_batch* b, const int32_t* actions, double* reward, uint8_t* done) {
    ENERO_TRACE("enero_batch_step");
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_step: null batch");
    TRY(check_batch_device(b, "enero_batch_step"));
    if (!b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_step: reset first");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    if (actions) {
        TRY(check_actions(b, actions));
        CUDA_TRY(cudaMemcpyAsync(b->action.p, actions, (size_t)b->B * 4, cudaMemcpyHostToDevice, c->stream));
    } else if (!b->policy_valid) {
        return fail(ENERO_ERR_STATE, "enero_batch_step: greedy step needs enero_batch_policy first");
    }
    TRY(step_dev(b));
    b->mirror_valid = false;
    if (reward) CUDA_TRY(cudaMemcpyAsync(reward, b->last_reward.p, (size_t)b->B * 8, cudaMemcpyDeviceToHost, c->stream));
    if (done) CUDA_TRY(cudaMemcpyAsync(done, b->done.p, (size_t)b->B, cudaMemcpyDeviceToHost, c->stream));
    if (reward || done || actions) CUDA_TRY(cudaStreamSynchronize(c->stream));   // otherwise asynchronous
    return ENERO_OK;
}

extern "C" int enero_batch_step_state(enero_batch* b, const int32_t* actions, double* reward, uint8_t* done, double* loads,
                                      int32_t* cur, double* bw, double* max_uti, int32_t* max_edge) {
    ENERO_TRACE("enero_batch_step_state");
    if (!b || !actions) return fail(ENERO_ERR_INVALID, "enero_batch_step_state: null argument");
    TRY(check_batch_device(b, "enero_batch_step_state"));
    if (!b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_step_state: reset first");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    TRY(check_actions(b, actions));
    const size_t B = b->B;
    CUDA_TRY(cudaMemcpyAsync(b->action.p, actions, B * 4, cudaMemcpyHostToDevice, st));
    TRY(step_dev(b));
    b->h_cur.resize(B * 2); b->h_done.resize(B);
    CUDA_TRY(cudaMemcpyAsync(b->h_cur.data(), b->cur.p, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(b->h_done.data(), b->done.p, B, cudaMemcpyDeviceToHost, st));
    if (reward) CUDA_TRY(cudaMemcpyAsync(reward, b->last_reward.p, B * 8, cudaMemcpyDeviceToHost, st));
    if (loads) CUDA_TRY(cudaMemcpyAsync(loads, b->loads.p, B * b->t->E * 8, cudaMemcpyDeviceToHost, st));
    if (bw) CUDA_TRY(cudaMemcpyAsync(bw, b->cur_bw.p, B * 8, cudaMemcpyDeviceToHost, st));
    if (max_uti) CUDA_TRY(cudaMemcpyAsync(max_uti, b->max_uti.p, B * 8, cudaMemcpyDeviceToHost, st));
    if (max_edge) CUDA_TRY(cudaMemcpyAsync(max_edge, b->max_edge.p, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    b->mirror_valid = true;
    if (cur) std::copy(b->h_cur.begin(), b->h_cur.end(), cur);
    if (done) std::copy(b->h_done.begin(), b->h_done.end(), done);
    return ENERO_OK;
}

// One greedy decision step (decide_dev with the env step in its last kernel) captured as a CUDA graph and replayed max_len times: at small
// batch sizes the episode is bound by launch latency (7 launches per decision), and a graph replay hands the whole
// step to the GPU in one submission.  The graph is keyed on everything its nodes baked in (buffer addresses, math
// mode) and re-captured when any of it changes.
static int greedy_step_graph(enero_batch* b) {
    enero_ctx* c = b->c; const TopoDev& d = b->t->dev;
    const int64_t rows = (int64_t)b->B * d.Kmax * d.E;
    TRY(ensure_rows(c, rows));                                   // no allocation may happen while capturing
    TRY(c->util.ensure((size_t)b->B * d.E * 4));
    TRY(c->bwcap.ensure((size_t)b->B * d.E * 4)); TRY(c->marks.ensure((size_t)b->B * d.Kmax * ((d.E + 31) / 32) * 4));
    TRY(c->ptab.ensure((size_t)b->B * d.E * 2 * H * 4));
    const void* key[10] = {c->ptab.p, c->bwcap.p, c->h[0].p, c->h[1].p, c->A[0].p, c->A[1].p, c->util.p, b->loads.p,
                          (const void*)(intptr_t)(c->math_acc + 1 + 4 * c->mp_engine + 16 * (c->tail_mode + 1)), (const void*)c->marks.p};
    GreedyGraph& g = b->graph;
    if (g.exec && !std::memcmp(key, g.key, sizeof(key))) return ENERO_OK;
    g.release();
    c->tile_ctr_next = 16;                                       // the captured step starts by zeroing its own tile counters
    const int64_t l0 = c->launches;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const EpisodeView ep = ep_view(b);
    const int r = decide_dev(c, b->t, b->B, b->loads.as<double>(), b->cur.as<int>(), b->cur_bw.as<double>(), b->done.as<uint8_t>(),
                             b->logits.as<float>(), b->probs.as<float>(), b->nK.as<int>(), b->action.as<int>(), 5, &ep);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    g.kernels = (int)(c->launches - l0);
    c->launches = l0;                                            // nothing ran yet: replays are counted when launched
    c->tile_ctr_next = 16;
    if (r != ENERO_OK) { if (graph) cudaGraphDestroy(graph); return r; }
    if (e != cudaSuccess) { enero_set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); return ENERO_ERR_CUDA; }
    const cudaError_t e2 = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { g.exec = nullptr; enero_set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2)); return ENERO_ERR_CUDA; }
    std::memcpy(g.key, key, sizeof(key));
    return ENERO_OK;
}

extern "C" int enero_batch_run_greedy(enero_batch* b) {
    ENERO_TRACE("enero_batch_run_greedy");
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_run_greedy: null batch");
    TRY(check_batch_device(b, "enero_batch_run_greedy"));
    if (!b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_run_greedy: reset first");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->prof || c->no_graphs) {                               // per-launch event timing needs individual launches
        const EpisodeView ep = ep_view(b);
        for (int s = 0; s < b->max_len; ++s)
            TRY(decide_dev(c, b->t, b->B, b->loads.as<double>(), b->cur.as<int>(), b->cur_bw.as<double>(), b->done.as<uint8_t>(),
                           b->logits.as<float>(), b->probs.as<float>(), b->nK.as<int>(), b->action.as<int>(), 5, &ep));
        b->policy_valid = false;
    } else {
        TRY(greedy_step_graph(b));
        for (int s = 0; s < b->max_len; ++s) CUDA_TRY(cudaGraphLaunch(b->graph.exec, c->stream));
        c->launches += (int64_t)b->graph.kernels * b->max_len;
        b->policy_valid = false;
    }
    b->mirror_valid = false;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

extern "C" int enero_batch_get_state(enero_batch* b, double* loads, int32_t* cur, double* bw, double* max_uti,
                                     int32_t* max_edge, uint8_t* done, int32_t* steps) {
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_get_state: null batch");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t B = b->B;
    if (loads) CUDA_TRY(cudaMemcpyAsync(loads, b->loads.p, B * b->t->E * 8, cudaMemcpyDeviceToHost, st));
    if (cur) CUDA_TRY(cudaMemcpyAsync(cur, b->cur.p, B * 8, cudaMemcpyDeviceToHost, st));
    if (bw) CUDA_TRY(cudaMemcpyAsync(bw, b->cur_bw.p, B * 8, cudaMemcpyDeviceToHost, st));
    if (max_uti) CUDA_TRY(cudaMemcpyAsync(max_uti, b->max_uti.p, B * 8, cudaMemcpyDeviceToHost, st));
    if (max_edge) CUDA_TRY(cudaMemcpyAsync(max_edge, b->max_edge.p, B * 4, cudaMemcpyDeviceToHost, st));
    if (done) CUDA_TRY(cudaMemcpyAsync(done, b->done.p, B, cudaMemcpyDeviceToHost, st));
    if (steps) CUDA_TRY(cudaMemcpyAsync(steps, b->steps.p, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ENERO_OK;
}

extern "C" int enero_batch_get_eligible(enero_batch* b, int32_t* elig, int32_t* elig_len) {
    if (!b || !elig || !elig_len) return fail(ENERO_ERR_INVALID, "enero_batch_get_eligible: null argument");
    const int N = b->t->N;
    for (int i = 0; i < b->B; ++i) {
        elig_len[i] = b->h_elig_len[i];
        for (int k = 0; k < b->Dmax; ++k) {
            const int sdi = k < b->h_elig_len[i] ? b->h_elig[(size_t)i * b->Dmax + k] : 0;
            elig[((size_t)i * b->Dmax + k) * 2] = sdi / N;
            elig[((size_t)i * b->Dmax + k) * 2 + 1] = sdi % N;
        }
    }
    return ENERO_OK;
}

extern "C" int enero_batch_get_history(enero_batch* b, int32_t* action, double* reward, double* max_uti) {
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_get_history: null batch");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)b->B * b->Dmax;
    if (action) CUDA_TRY(cudaMemcpyAsync(action, b->h_action.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (reward) CUDA_TRY(cudaMemcpyAsync(reward, b->h_reward.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (max_uti) CUDA_TRY(cudaMemcpyAsync(max_uti, b->h_max.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

// routing snapshot of the best step, in sp_middlepoints insertion order (script :181-184)
static int best_routing_host(enero_batch* b, std::vector<double>& best, std::vector<double>& ospf,
                             std::vector<int>& routing, std::vector<int>& nr) {
    enero_ctx* c = b->c; enero_topo* t = b->t;
    const int B = b->B, D = b->Dmax, N = t->N;
    std::vector<int> bstep(B), act((size_t)B * D);
    best.resize(B); ospf.resize(B);
    CUDA_TRY(cudaMemcpyAsync(best.data(), b->best.p, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(ospf.data(), b->ospf.p, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(bstep.data(), b->best_step.p, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(act.data(), b->h_action.p, act.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    routing.assign((size_t)B * D * 3, 0); nr.assign(B, 0);
    for (int i = 0; i < B; ++i) {
        int n = 0;
        for (int s = 0; s <= bstep[i]; ++s) {
            const int sdi = b->h_elig[(size_t)i * D + s];
            const int src = sdi / N, dst = sdi % N;
            const int mid = t->mids[t->mid_off[sdi] + act[(size_t)i * D + s]];
            if (mid == dst) continue;
            // the last step() de-allocates the dummy demand (1,2) and drops its entry before the copy
            if (bstep[i] == b->h_elig_len[i] - 1 && src == 1 && dst == 2) continue;
            int* r = &routing[((size_t)i * D + n) * 3];
            r[0] = src; r[1] = dst; r[2] = mid; ++n;
        }
        nr[i] = n;
    }
    return ENERO_OK;
}

extern "C" int enero_batch_get_best(enero_batch* b, double* best, double* ospf, int32_t* routing, int32_t* n_routing) {
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_get_best: null batch");
    CUDA_TRY(cudaSetDevice(b->c->device));
    std::vector<double> vb, vo; std::vector<int> r, n;
    TRY(best_routing_host(b, vb, vo, r, n));
    if (best) std::copy(vb.begin(), vb.end(), best);
    if (ospf) std::copy(vo.begin(), vo.end(), ospf);
    if (routing) std::copy(r.begin(), r.end(), routing);
    if (n_routing) std::copy(n.begin(), n.end(), n_routing);
    return ENERO_OK;
}

// ---- Local Search ----------------------------------------------------------------------------------
extern "C" int enero_batch_local_search(enero_batch* b, int mode, const int32_t* routing, const int32_t* n_routing,
                                        int max_moves, double* start, double* final_val, int32_t* n_moves,
                                        int32_t* moves, double* move_val) {
    ENERO_TRACE("enero_batch_local_search");
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_local_search: null batch");
    TRY(check_batch_device(b, "enero_batch_local_search"));
    if (mode < 0 || mode > 1 || max_moves < 0) return fail(ENERO_ERR_INVALID, "enero_batch_local_search: bad mode / max_moves");
    enero_ctx* c = b->c; enero_topo* t = b->t;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int B = b->B, E = t->E, NN = b->NN, N = t->N, D = b->Dmax;
    std::vector<int> own_r, own_n; std::vector<double> tmpb, tmpo;
    int rstride = D;
    if (mode == 0 && !routing) {
        if (!b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_local_search: no episode to take a routing from");
        TRY(best_routing_host(b, tmpb, tmpo, own_r, own_n));
        routing = own_r.data(); n_routing = own_n.data();
    }
    if (mode == 0 && !n_routing) return fail(ENERO_ERR_INVALID, "enero_batch_local_search: n_routing missing");
    // sp_middlepoints of the routing
    std::vector<int> mid((size_t)B * NN, -1);
    if (mode == 0)
        for (int i = 0; i < B; ++i)
            for (int k = 0; k < n_routing[i]; ++k) {
                const int* r = routing + ((size_t)i * rstride + k) * 3;
                if (r[0] < 0 || r[0] >= N || r[1] < 0 || r[1] >= N || r[2] < 0 || r[2] >= N || r[0] == r[1])
                    return fail(ENERO_ERR_INVALID, "enero_batch_local_search: bad routing entry");
                if (r[2] != r[1]) mid[(size_t)i * NN + r[0] * N + r[1]] = r[2];
            }
    CUDA_TRY(cudaMemcpyAsync(b->ls_mid.p, mid.data(), mid.size() * 4, cudaMemcpyHostToDevice, st));
    TopoView tv = view_of(t);
    k_route_loads<<<B, 256, (size_t)E * 8, st>>>(tv, b->tm.as<double>(), b->ls_mid.as<int>(), b->ls_loads.as<double>(),
                                                 b->ls_val.as<double>(), b->ls_max_edge.as<int>());
    LAUNCH_CHECK(c);
    const int cap = mode == 0 ? D : N * (N - 1);
    b->ls_cap = cap;
    b->h_ls_elig.assign((size_t)B * cap, 0); b->h_ls_elig_len.assign(B, 0);
    TRY(b->ls_elig.ensure((size_t)B * cap * 4)); TRY(b->ls_elig_len.ensure((size_t)B * 4));
    if (mode == 0) {
        // critical demands of the restored routing, selected on the device (same kernel as enero_batch_reset, with
        // the re-routed demands taken out of the direct crossing lists and appended in snapshot order)
        TRY(b->ls_routing.ensure((size_t)B * rstride * 3 * 4)); TRY(b->ls_nrouting.ensure((size_t)B * 4));
        TRY(b->crit_scratch.ensure((size_t)B * NN * 4));
        CUDA_TRY(cudaMemcpyAsync(b->ls_routing.p, routing, (size_t)B * rstride * 3 * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->ls_nrouting.p, n_routing, (size_t)B * 4, cudaMemcpyHostToDevice, st));
        CritView cv{t->dev.cross_off, t->dev.cross_sd, b->ls_mid.as<int>(), b->ls_routing.as<int>(), b->ls_nrouting.as<int>(), rstride,
                    D, cap, b->crit_scratch.as<int>()};
        const size_t csm = (size_t)E * 8 + (size_t)((NN + 31) / 32) * 4;
        k_critical_demands<<<B, 256, csm, st>>>(tv, cv, b->ls_loads.as<double>(), b->tm.as<double>(), b->ls_elig.as<int>(), b->ls_elig_len.as<int>());
        LAUNCH_CHECK(c);
        CUDA_TRY(cudaMemcpyAsync(b->h_ls_elig.data(), b->ls_elig.p, b->h_ls_elig.size() * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(b->h_ls_elig_len.data(), b->ls_elig_len.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        for (int i = 0; i < B; ++i) {
            int n = 0;
            for (int s = 0; s < N; ++s) for (int d = 0; d < N; ++d) if (s != d) b->h_ls_elig[(size_t)i * cap + n++] = s * N + d;
            b->h_ls_elig_len[i] = n;
        }
    }
    TRY(b->ls_moves.ensure((size_t)B * std::max(1, max_moves) * 12)); TRY(b->ls_move_val.ensure((size_t)B * std::max(1, max_moves) * 8));
    if (mode != 0) {
        CUDA_TRY(cudaMemcpyAsync(b->ls_elig.p, b->h_ls_elig.data(), b->h_ls_elig.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->ls_elig_len.p, b->h_ls_elig_len.data(), (size_t)B * 4, cudaMemcpyHostToDevice, st));
    }
    if (start) CUDA_TRY(cudaMemcpyAsync(start, b->ls_val.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    LsView lv{B, cap, max_moves, b->tm.as<double>(), b->ls_loads.as<double>(), b->ls_mid.as<int>(), b->ls_elig.as<int>(),
              b->ls_elig_len.as<int>(), b->ls_val.as<double>(), b->ls_nmoves.as<int>(), b->ls_moves.as<int>(), b->ls_move_val.as<double>()};
    const size_t smem = ls_smem_bytes(E);
    if (max_moves == 0) {
        // reset only (environment15.py:485-544): loads, start value and demand list are in place, no move is made
        CUDA_TRY(cudaMemsetAsync(b->ls_nmoves.p, 0, (size_t)B * 4, st));
    } else {
    if (smem > 200 * 1024) return fail(ENERO_ERR_INVALID, "enero_batch_local_search: topology too large for the shared-memory LS kernel");
    // few episodes: split each episode's demand list over a thread-block cluster so the whole GPU works
    int CL = 1;
    while (CL < LS_MAX_CLUSTER && (int64_t)B * CL * 2 <= c->num_sms) CL *= 2;
    if (CL > 1) {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_local_search<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * CL)); cfg.blockDim = dim3(LS_TPB); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, k_local_search<true>, tv, lv, CL));
        c->launches++;
    } else {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_local_search<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_local_search<false><<<B, LS_TPB, smem, st>>>(tv, lv, 1); LAUNCH_CHECK(c);
    }
    }
    if (final_val) CUDA_TRY(cudaMemcpyAsync(final_val, b->ls_val.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    if (n_moves) CUDA_TRY(cudaMemcpyAsync(n_moves, b->ls_nmoves.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (moves) CUDA_TRY(cudaMemcpyAsync(moves, b->ls_moves.p, (size_t)B * max_moves * 12, cudaMemcpyDeviceToHost, st));
    if (move_val) CUDA_TRY(cudaMemcpyAsync(move_val, b->ls_move_val.p, (size_t)B * max_moves * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ENERO_OK;
}

extern "C" int enero_batch_get_ls_eligible(enero_batch* b, int32_t* elig, int32_t* elig_len, int cap) {
    if (!b || !elig || !elig_len) return fail(ENERO_ERR_INVALID, "enero_batch_get_ls_eligible: null argument");
    if (cap < b->ls_cap) return fail(ENERO_ERR_INVALID, "enero_batch_get_ls_eligible: cap too small");
    const int N = b->t->N;
    for (int i = 0; i < b->B; ++i) {
        elig_len[i] = b->h_ls_elig_len[i];
        for (int k = 0; k < b->h_ls_elig_len[i]; ++k) {
            const int sdi = b->h_ls_elig[(size_t)i * b->ls_cap + k];
            elig[((size_t)i * cap + k) * 2] = sdi / N;
            elig[((size_t)i * cap + k) * 2 + 1] = sdi % N;
        }
    }
    return ENERO_OK;
}

// ---- sampled-action rollouts -------------------------------------------------------------------------------------
extern "C" int enero_batch_rollout(enero_batch* b, uint64_t seed, double* loads, int32_t* sd, double* bw, int32_t* action,
                                   float* prob, float* value, double* reward, int32_t* steps, float* last_value) {
    ENERO_TRACE("enero_batch_rollout");
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_rollout: null batch");
    if (!b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_rollout: reset first");
    TRY(check_batch_device(b, "enero_batch_rollout"));
    enero_ctx* c = b->c; enero_topo* t = b->t;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t B = b->B, D = b->Dmax, E = t->E;
    TRY(b->r_loads.ensure(B * D * E * 8)); TRY(b->r_sd.ensure(B * D * 8)); TRY(b->r_bw.ensure(B * D * 8));
    TRY(b->r_action.ensure(B * D * 4)); TRY(b->r_prob.ensure(B * D * 4)); TRY(b->r_value.ensure(B * D * 4));
    RolloutRec rec{b->r_loads.as<double>(), b->r_sd.as<int>(), b->r_bw.as<double>(), b->r_action.as<int>(), b->r_prob.as<float>(),
                   b->r_value.as<float>()};
    for (int s = 0; s < b->max_len; ++s) {
        // actor over every candidate of the current demands, critic on the same link loads (train script :495-500)
        TRY(decide_dev(c, t, b->B, b->loads.as<double>(), b->cur.as<int>(), b->cur_bw.as<double>(), b->done.as<uint8_t>(),
                       b->logits.as<float>(), b->probs.as<float>(), b->nK.as<int>(), b->action.as<int>(), 5));
        TRY(value_dev(c, t, b->B, b->loads.as<double>(), b->values.as<float>(), 5));
        k_rollout_pick<<<(b->B + 3) / 4, 128, 0, st>>>(view_of(t), ep_view(b), t->Kmax, b->probs.as<float>(), b->nK.as<int>(),
                                                       b->values.as<float>(), seed, rec, b->action.as<int>());
        LAUNCH_CHECK(c);
        TRY(step_dev(b));
    }
    // value of the state each environment is left in (the bootstrap of train script :622-625 uses the last one)
    TRY(value_dev(c, t, b->B, b->loads.as<double>(), b->values.as<float>(), 5));
    b->mirror_valid = false;
    if (loads) CUDA_TRY(cudaMemcpyAsync(loads, b->r_loads.p, B * D * E * 8, cudaMemcpyDeviceToHost, st));
    if (sd) CUDA_TRY(cudaMemcpyAsync(sd, b->r_sd.p, B * D * 8, cudaMemcpyDeviceToHost, st));
    if (bw) CUDA_TRY(cudaMemcpyAsync(bw, b->r_bw.p, B * D * 8, cudaMemcpyDeviceToHost, st));
    if (action) CUDA_TRY(cudaMemcpyAsync(action, b->r_action.p, B * D * 4, cudaMemcpyDeviceToHost, st));
    if (prob) CUDA_TRY(cudaMemcpyAsync(prob, b->r_prob.p, B * D * 4, cudaMemcpyDeviceToHost, st));
    if (value) CUDA_TRY(cudaMemcpyAsync(value, b->r_value.p, B * D * 4, cudaMemcpyDeviceToHost, st));
    if (reward) CUDA_TRY(cudaMemcpyAsync(reward, b->h_reward.p, B * D * 8, cudaMemcpyDeviceToHost, st));
    if (steps) CUDA_TRY(cudaMemcpyAsync(steps, b->steps.p, B * 4, cudaMemcpyDeviceToHost, st));
    if (last_value) CUDA_TRY(cudaMemcpyAsync(last_value, b->values.p, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ENERO_OK;
}

// ---- SAP baseline --------------------------------------------------------------------------------------------
extern "C" int enero_batch_sap(enero_batch* b, const double* tms, int K, uint32_t seed, double* max_uti, int32_t* actions) {
    ENERO_TRACE("enero_batch_sap");
    if (!b || K < 1) return fail(ENERO_ERR_INVALID, "enero_batch_sap: bad argument");
    TRY(check_batch_device(b, "enero_batch_sap"));
    if (!tms && !b->reset_done) return fail(ENERO_ERR_STATE, "enero_batch_sap: no traffic matrices (pass tms or reset the batch first)");
    enero_ctx* c = b->c; enero_topo* t = b->t;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    TRY(enero_topo_build_ksp(t, K));
    if (t->ksp_dev_stale) {
        cudaFree(t->d_ksp_pair_off); cudaFree(t->d_ksp_edge_off); cudaFree(t->d_ksp_edge);
        TRY(upload_vec(&t->d_ksp_pair_off, t->ksp_pair_off)); TRY(upload_vec(&t->d_ksp_edge_off, t->ksp_edge_off));
        TRY(upload_vec(&t->d_ksp_edge, t->ksp_edge));
        t->ksp_dev_stale = false;
    }
    const int B = b->B, N = t->N, NN = b->NN, D = NN - N, E = t->E;
    for (int sdi = 0; sdi < NN; ++sdi)
        if (sdi / N != sdi % N && t->ksp_pair_off[sdi + 1] == t->ksp_pair_off[sdi])
            return fail(ENERO_ERR_INVALID, "enero_batch_sap: a node pair has no path");
    if (tms) CUDA_TRY(cudaMemcpyAsync(b->tm.p, tms, (size_t)B * NN * 8, cudaMemcpyHostToDevice, st));
    TRY(b->sap_order.ensure((size_t)B * D * 4)); TRY(b->sap_loads.ensure((size_t)B * E * 8));
    TRY(b->sap_mt.ensure((size_t)B * 624 * 4)); TRY(b->sap_idx.ensure((size_t)B * 4)); TRY(b->sap_max.ensure((size_t)B * 8));
    if (actions) TRY(b->sap_act.ensure((size_t)B * D * 4));
    k_sap_order<<<dim3((NN + 255) / 256, B), 256, 0, st>>>(N, b->tm.as<double>(), b->sap_order.as<int>()); LAUNCH_CHECK(c);
    KspView ks{t->d_ksp_pair_off, t->d_ksp_edge_off, t->d_ksp_edge};
    k_sap_run<<<(B + 3) / 4, 128, 0, st>>>(view_of(t), ks, B, K, seed, b->tm.as<double>(), b->sap_order.as<int>(), b->sap_loads.as<double>(),
                                           b->sap_mt.as<uint32_t>(), b->sap_idx.as<int>(), b->sap_max.as<double>(),
                                           actions ? b->sap_act.as<int>() : nullptr);
    LAUNCH_CHECK(c);
    if (max_uti) CUDA_TRY(cudaMemcpyAsync(max_uti, b->sap_max.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    if (actions) CUDA_TRY(cudaMemcpyAsync(actions, b->sap_act.p, (size_t)B * D * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (tms) { b->reset_done = false; b->policy_valid = false; }     // the resident TMs no longer belong to the last reset
    return ENERO_OK;
}

#include "train_api.cuh"

// ---- per-operation Local-Search API (callers that drive the hill climbing move by move) -------------------
extern "C" int enero_batch_ls_route_add(enero_batch* b, int episode, int src, int dst, double value, double* loads_host,
                                        double* max_uti, int32_t* max_edge) {
    if (!b || episode < 0 || episode >= b->B) return fail(ENERO_ERR_INVALID, "enero_batch_ls_route_add: bad episode");
    TRY(check_batch_device(b, "enero_batch_ls_route_add"));
    enero_topo* t = b->t; enero_ctx* c = b->c;
    if (src < 0 || dst < 0 || src >= t->N || dst >= t->N || src == dst) return fail(ENERO_ERR_INVALID, "enero_batch_ls_route_add: bad path end points");
    CUDA_TRY(cudaSetDevice(c->device));
    double* loads = b->ls_loads.as<double>() + (size_t)episode * t->E;
    k_ls_route_add<<<1, 32, 0, c->stream>>>(view_of(t), loads, src, dst, value, b->ls_val.as<double>() + episode, b->ls_max_edge.as<int>() + episode);
    LAUNCH_CHECK(c);
    if (loads_host) CUDA_TRY(cudaMemcpyAsync(loads_host, loads, (size_t)t->E * 8, cudaMemcpyDeviceToHost, c->stream));
    if (max_uti) CUDA_TRY(cudaMemcpyAsync(max_uti, b->ls_val.as<double>() + episode, 8, cudaMemcpyDeviceToHost, c->stream));
    if (max_edge) CUDA_TRY(cudaMemcpyAsync(max_edge, b->ls_max_edge.as<int>() + episode, 4, cudaMemcpyDeviceToHost, c->stream));
    if (loads_host || max_uti || max_edge) CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

extern "C" int enero_batch_get_ls_state(enero_batch* b, double* loads, int32_t* max_edge) {
    if (!b) return fail(ENERO_ERR_INVALID, "enero_batch_get_ls_state: null batch");
    enero_ctx* c = b->c;
    CUDA_TRY(cudaSetDevice(c->device));
    if (loads) CUDA_TRY(cudaMemcpyAsync(loads, b->ls_loads.p, (size_t)b->B * b->t->E * 8, cudaMemcpyDeviceToHost, c->stream));
    if (max_edge) CUDA_TRY(cudaMemcpyAsync(max_edge, b->ls_max_edge.p, (size_t)b->B * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}
