// Shared host-side definitions of the CUDA translation units (api.cu, train.cu).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "enero_internal.h"

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            enero_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
            return ENERO_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)
#define TRY(expr) do { int _r = (expr); if (_r != ENERO_OK) return _r; } while (0)

static inline int fail(int code, const std::string& m) { enero_set_error(m); return code; }

// Growth head-room of DevBuf::ensure on this thread.  A re-allocation is a cudaFree (device-wide synchronisation) plus a
// cudaMalloc, ~1.5 ms each: the ~30 scratch buffers of a training lane re-allocating because a minibatch holds a few
// more samples of a topology than any before stalled single PPO updates by 80 ms .. seconds.  The training path
// reserves for the largest share of the minibatch it can expect (ReserveScope) so the buffers settle at the first step.
inline double& devbuf_reserve() { static thread_local double r = 1.0; return r; }
struct ReserveScope {
    double prev;
    explicit ReserveScope(double r) : prev(devbuf_reserve()) { devbuf_reserve() = r; }
    ~ReserveScope() { devbuf_reserve() = prev; }
};

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return ENERO_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        want = std::max(want, (size_t)((double)bytes * devbuf_reserve()) + 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { enero_set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return ENERO_ERR_CUDA; }
        cap = want;
        return ENERO_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One concurrent pipeline of the PPO update (train_api.cuh): its own stream, activation / gradient scratch, tile
// counters and gradient accumulator.  A minibatch holds up to three topologies x {actor, critic}: six independent
// forward+backward passes of a few thousand rows each, latency bound one by one, which run side by side on six lanes
// and meet in a fixed-order sum (deterministic) before clip + Adam.
struct TrainLane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;                       // work enqueued since the last join
    DevBuf t_h[8], t_A[8];                   // h_0..h_T, A_0..A_{T-1} of a training forward
    DevBuf t_agg, t_G, t_dB, t_dA, t_B, t_dagg, t_dhpart, t_dh[2], t_V, t_dlogit, t_partials;
    DevBuf util, t_nK, t_logits, t_loss, t_ent;
    float* grad = nullptr;                   // [ENERO_GRAD_STRIDE + 4]: gradient of this lane's model | its loss sums
    int* tile_ctr = nullptr; int tile_ctr_next = 16;
};
struct TrainGroup {                          // the samples of one topology in the current minibatch + its two lanes
    DevBuf t_loads, t_sd, t_bw, t_action, t_oldp, t_adv, t_ret;
    cudaEvent_t ready = nullptr;
    TrainLane lane[2];                       // [ENERO_ACTOR], [ENERO_CRITIC]
};
constexpr int TRAIN_GROUPS = 3;

// One helper thread per ctx that enqueues the critic lane of a training group while the calling thread enqueues the
// actor lane: an Adam step is ~175 small launches, and the host's 5.5 us per launch (0.96 ms) was level with the GPU's
// time per step.  run() hands a job over, wait() returns its status (and the error text, which is thread local).
struct LaneWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, finished = true, stop = false;
    int rc = 0;
    std::string err;
    explicit LaneWorker(int device) {
        th = std::thread([this, device] {
            cudaSetDevice(device);
            std::unique_lock<std::mutex> lk(m);
            for (;;) {
                cv.wait(lk, [this] { return has_job || stop; });
                if (stop) return;
                std::function<int()> f = std::move(job);
                has_job = false;
                lk.unlock();
                const int r = f();
                const std::string e = r ? enero_last_error() : "";
                lk.lock();
                rc = r; err = e; finished = true;
                cv.notify_all();
            }
        });
    }
    void run(std::function<int()> f) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(f); has_job = true; finished = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return finished; });
        if (rc) enero_set_error(err);
        return rc;
    }
    ~LaneWorker() {
        { std::lock_guard<std::mutex> lk(m); stop = true; cv.notify_all(); }
        if (th.joinable()) th.join();
    }
};

struct enero_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;           // stream every launch / copy of this ctx goes to (enero_ctx_set_stream)
    cudaStream_t own_stream = nullptr;       // the non-blocking stream the ctx created for itself
    float* w[2] = {nullptr, nullptr};
    std::atomic<int64_t> launches{0};         // (the two lanes of a training group are enqueued from two host threads)
    // scratch of the GNN engine
    LaneWorker* lane_worker = nullptr;       // created with the lanes unless ENERO_TRAIN_THREADS=0
    DevBuf bwcap, marks, ptab;                      // tables of the fused first iteration (k_decision_tables)
    int tail_mode = -1;                      // decision tail: 1 = k_decision_finish, 0 = k_readout + k_policy + k_env_step, -1 = by batch
                                             // size (ENERO_FUSED_TAIL=0/1 forces one; measured cross-over in DESIGN.md)
    bool prof_unfused = false;               // ENERO_UNFUSED_FEATURES=1: keep k_features + generic iteration 0 (A/B measurements)
    DevBuf h[2], A[2], util, nK, logits, probs, action, in_loads, in_sd, in_bw, ones;
    DevBuf x_ls, x_gid, x_first, x_second, x_cnt, x_off, x_cur, x_slot, x_src, x_goff, x_err, x_out;
    int occ_iter[2][2][2] = {};              // [math][indexer][last] resident CTAs/SM of k_mp_iter
    int math_acc = ENERO_MATH_DEFAULT;                        // ENERO_MATH_ACCURATE / ENERO_MATH_FAST (enero_ctx_set_math, ENERO_MATH)
    int* tile_ctr = nullptr;                 // [16] dynamic tile counters of k_mp_iter, one per launch of a sequence
    int tile_ctr_next = 16;                  // next unused counter (16 = re-zero first)
    // optional per-launch CUDA-event timing of the dominant kernel (bench.py roofline leg)
    int mp_engine = 0;                       // ENERO_MP_FFMA (k_mp_iter) / ENERO_MP_TC (k_mp_tc); enero_ctx_set_engine, ENERO_MP
    int occ_tc = 1;                          // resident CTAs/SM of k_mp_tc
    bool no_graphs = false;                  // ENERO_NO_GRAPHS=1: enqueue every launch of a greedy episode individually
    bool prof = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    // training state (train_api.cuh): one contiguous gradient buffer
    //   [actor ENERO_GRAD_STRIDE | critic ENERO_GRAD_STRIDE | actor-loss sum, entropy sum, critic-loss sum, pad]
    // so that a sharded PPO step needs exactly one all-reduce; Adam slots of actor [0] and critic [1]
    float* gradbuf = nullptr;
    float* grad[2] = {nullptr, nullptr};      // views into gradbuf
    float* adam_m[2] = {nullptr, nullptr};
    float* adam_v[2] = {nullptr, nullptr};
    TrainGroup* groups = nullptr;             // [TRAIN_GROUPS], created on first use
    int group_cursor = 0;
    float** lane_ptrs = nullptr;              // device array of the lanes' gradient buffers (k_join_lanes)
    DevBuf t_norm, t_hist;
    void* sample_stores = nullptr;            // topology -> device-resident PPO samples (train_api.cuh)
};

// NVTX range over one C-ABI call (header-only NVTX v3: shows up in nsys / ncu timelines, free otherwise)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define ENERO_TRACE(name) NvtxRange _nvtx_range_(name)

#define LAUNCH_CHECK(ctx)                                                     \
    do { (ctx)->launches++; CUDA_TRY(cudaGetLastError()); } while (0)

