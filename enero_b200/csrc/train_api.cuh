// C-ABI glue of the PPO update (included at the end of api.cu so that the k_mp_iter instantiations,
// their shared-memory attributes and occupancy live in one translation unit).
//
// Reference: train_Enero_3top_script.py:264-324 (_critic_step, _actor_step, _train_step_combined),
// :326-365 (ppo_update), :367-377 (get_advantages).
#pragma once
#include <unordered_map>

#include "train_kernels.cuh"

namespace {

struct TrainTopo {           // launch geometry of one model's pass over n samples
    int n, Kmax, E, RPD; int64_t rows; const int* nK;
};

int ensure_ones(enero_ctx* c, int B) {
    if (c->ones.cap >= (size_t)B * 4) return ENERO_OK;
    CUDA_TRY(cudaDeviceSynchronize());                           // lanes may still be reading the old buffer
    TRY(c->ones.ensure((size_t)B * 4));
    std::vector<int> one((size_t)c->ones.cap / 4, 1);
    CUDA_TRY(cudaMemcpyAsync(c->ones.p, one.data(), one.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

int ensure_lanes(enero_ctx* c) {
    if (c->groups) return ENERO_OK;
    c->groups = new TrainGroup[TRAIN_GROUPS];
    for (int g = 0; g < TRAIN_GROUPS; ++g) {
        CUDA_TRY(cudaEventCreateWithFlags(&c->groups[g].ready, cudaEventDisableTiming));
        for (TrainLane& L : c->groups[g].lane) {
            CUDA_TRY(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
            CUDA_TRY(cudaMalloc(&L.grad, sizeof(float) * (ENERO_GRAD_STRIDE + 4)));
            CUDA_TRY(cudaMemset(L.grad, 0, sizeof(float) * (ENERO_GRAD_STRIDE + 4)));
            CUDA_TRY(cudaMalloc(&L.tile_ctr, 16 * sizeof(int)));
        }
    }
    const char* e = std::getenv("ENERO_TRAIN_THREADS");
    if (!(e && std::atoi(e) == 0)) c->lane_worker = new LaneWorker(c->device);
    return ENERO_OK;
}

// forward with saved activations: h_0/A_0 must already be in t_h[0]/t_A[0]
int train_forward(enero_ctx* c, TrainLane& L, const IdxTopo& idx, int which, int T, int64_t rows) {
    if (rows >= (int64_t)1 << 31) return fail(ENERO_ERR_INVALID, "more than 2^31 link-state rows in one launch: split the minibatch");
    for (int it = 0; it < T; ++it) {
        const bool last = (it == T - 1);
        TRY(launch_topo_iteration(c, idx, which, last, it == 0 && !last, L.t_h[it].as<float>(), L.t_A[it].as<float>(),
                                  L.t_h[it + 1].as<float>(), last ? nullptr : L.t_A[it + 1].as<float>(), &L));
    }
    return ENERO_OK;
}

// backward through the readout and the T message-passing iterations; dlogit [n*Kmax] on device
int train_backward(enero_ctx* c, TrainLane& L, enero_topo* t, const TrainTopo& g, int which, int T, const float* dlogit) {
    const TopoDev& d = t->dev;
    cudaStream_t st = L.st;
    const int64_t rows = g.rows;
    const size_t rb = (size_t)rows * H * sizeof(float);
    TRY(L.t_agg.ensure(rb)); TRY(L.t_G.ensure(4 * rb)); TRY(L.t_dB.ensure(rb)); TRY(L.t_dA.ensure(rb));
    TRY(L.t_B.ensure(rb)); TRY(L.t_dagg.ensure(rb)); TRY(L.t_dhpart.ensure(rb));
    TRY(L.t_dh[0].ensure(rb)); TRY(L.t_dh[1].ensure(rb));
    const int n_graphs = g.n * g.Kmax;
    TRY(L.t_V.ensure((size_t)n_graphs * 6 * H * 4));
    const int wg_grid = (int)std::min<int64_t>((rows + WG_ROWS - 1) / WG_ROWS, (int64_t)c->num_sms);
    TRY(L.t_partials.ensure((size_t)wg_grid * WG_PART * 4));
    IdxTopo idx{rows, g.E, d.off_f, d.off_s, g.RPD, g.nK, d.in_off, d.in_first};
    OutTopo out{g.E, d.off_f, d.off_s, g.RPD, g.nK, d.out_off, d.out_second};
    GraphsTopo gr{g.n, g.Kmax, g.E, g.RPD, g.nK};
    float* grad = L.grad;
    const float* w = c->w[which];
    k_readout_bwd<<<(n_graphs + 3) / 4, 128, 0, st>>>(gr, w, L.t_h[T].as<float>(), dlogit, L.t_dh[0].as<float>(), L.t_V.as<float>());
    LAUNCH_CHECK(c);
    k_readout_wgrad<<<(RW_NOUT + RW_OUT_PER_CTA - 1) / RW_OUT_PER_CTA, RW_TPB, 0, st>>>(n_graphs, g.Kmax, g.nK, L.t_V.as<float>(), grad + W_R1_K); LAUNCH_CHECK(c);
    const int bgrid = (int)((rows + BW_TPB - 1) / BW_TPB);
    int cur = 0;
    for (int it = T - 1; it >= 0; --it) {
        const float* ht = L.t_h[it].as<float>();
        const float* At = L.t_A[it].as<float>();
        auto gates = [&](auto kern) {
            kern<<<bgrid, BW_TPB, BWG_SMEM_BYTES, st>>>(idx, w, ht, At, L.t_dh[cur].as<float>(), L.t_agg.as<float>(), L.t_G.as<float>(),
                                                        L.t_B.as<float>(), L.t_dagg.as<float>(), L.t_dhpart.as<float>());
        };
        if (c->math_acc) gates(k_bwd_gates<true>); else gates(k_bwd_gates<false>);
        LAUNCH_CHECK(c);
        auto msgs = [&](auto kern, float* dh_prev) {
            kern<<<bgrid, BW_TPB, 0, st>>>(idx, out, w, At, L.t_B.as<float>(), L.t_dagg.as<float>(), L.t_dhpart.as<float>(),
                                           L.t_dB.as<float>(), L.t_dA.as<float>(), dh_prev);
        };
        if (it > 0) {
            if (c->math_acc) msgs(k_bwd_msgs<true, true>, L.t_dh[cur ^ 1].as<float>());
            else msgs(k_bwd_msgs<true, false>, L.t_dh[cur ^ 1].as<float>());
        } else {
            if (c->math_acc) msgs(k_bwd_msgs<false, true>, nullptr);
            else msgs(k_bwd_msgs<false, false>, nullptr);
        }
        LAUNCH_CHECK(c);
        WgradIter wi{ht, L.t_agg.as<float>(), L.t_G.as<float>(), L.t_dB.as<float>(), L.t_dA.as<float>()};
        k_wgrad<<<wg_grid, WG_TPB, 0, st>>>(idx, wi, (rows + WG_ROWS - 1) / WG_ROWS, L.t_partials.as<float>(), it < T - 1 ? 1 : 0); LAUNCH_CHECK(c);
        cur ^= 1;
    }
    k_reduce_partials<<<(W_MP_FLOATS + 255) / 256, 256, 0, st>>>(wg_grid, L.t_partials.as<float>(), grad); LAUNCH_CHECK(c);
    return ENERO_OK;
}

int ensure_train_rows(TrainLane& L, int T, int64_t rows) {
    const size_t rb = (size_t)rows * H * sizeof(float);
    if (T + 1 > 8) return fail(ENERO_ERR_INVALID, "training supports T <= 7");
    for (int i = 0; i <= T; ++i) TRY(L.t_h[i].ensure(rb));
    for (int i = 0; i < T; ++i) TRY(L.t_A[i].ensure(rb));
    return ENERO_OK;
}

// lane gradients -> the context's gradient buffer, lanes in a fixed order (deterministic), lane buffers zeroed
__global__ void k_join_lanes(float* __restrict__ gradbuf, float* const* __restrict__ lane_grad, int n_lanes) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ENERO_GRAD_STRIDE + 4) return;
    for (int l = 0; l < n_lanes; ++l) {
        float* g = lane_grad[l];
        const float v = g[o];
        g[o] = 0.f;
        const int which = l & 1;                                  // lane order: group-major, actor then critic
        if (o < ENERO_GRAD_STRIDE) gradbuf[which * ENERO_GRAD_STRIDE + o] += v;
        else if (o < ENERO_GRAD_STRIDE + 3) {
            // loss sums: an actor lane holds (loss, entropy, -), a critic lane (critic loss, -, -)
            const int k = o - ENERO_GRAD_STRIDE;
            if (which == 0 && k < 2) gradbuf[2 * ENERO_GRAD_STRIDE + k] += v;
            if (which == 1 && k == 0) gradbuf[2 * ENERO_GRAD_STRIDE + 2] += v;
        }
    }
}

// every lane's work joins the context's stream; their gradients are folded into the gradient buffer
int join_lanes(enero_ctx* c) {
    if (!c->groups) return ENERO_OK;
    bool any = false;
    for (int g = 0; g < TRAIN_GROUPS; ++g)
        for (TrainLane& L : c->groups[g].lane)
            if (L.busy) { CUDA_TRY(cudaStreamWaitEvent(c->stream, L.done, 0)); L.busy = false; any = true; }
    if (!any) return ENERO_OK;
    if (!c->lane_ptrs) {
        std::vector<float*> p;
        for (int g = 0; g < TRAIN_GROUPS; ++g) for (TrainLane& L : c->groups[g].lane) p.push_back(L.grad);
        CUDA_TRY(cudaMalloc(&c->lane_ptrs, p.size() * sizeof(float*)));
        CUDA_TRY(cudaMemcpy(c->lane_ptrs, p.data(), p.size() * sizeof(float*), cudaMemcpyHostToDevice));
    }
    k_join_lanes<<<(ENERO_GRAD_STRIDE + 4 + 255) / 256, 256, 0, c->stream>>>(c->gradbuf, c->lane_ptrs, 2 * TRAIN_GROUPS);
    LAUNCH_CHECK(c);
    c->group_cursor = 0;
    return ENERO_OK;
}

}  // namespace

void train_release_lanes(enero_ctx* c) {
    delete c->lane_worker;
    c->lane_worker = nullptr;
    if (!c->groups) return;
    for (int g = 0; g < TRAIN_GROUPS; ++g) {
        cudaEventDestroy(c->groups[g].ready);
        for (TrainLane& L : c->groups[g].lane) {
            cudaStreamSynchronize(L.st);
            cudaStreamDestroy(L.st); cudaEventDestroy(L.done); cudaFree(L.grad); cudaFree(L.tile_ctr);
        }
    }
    delete[] c->groups;
    c->groups = nullptr;
    cudaFree(c->lane_ptrs);
    c->lane_ptrs = nullptr;
}

extern "C" void* enero_ppo_grad_buffer(enero_ctx* c) { return c ? (void*)c->gradbuf : nullptr; }

extern "C" int enero_ppo_zero_grad(enero_ctx* c) {
    if (!c) return fail(ENERO_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(join_lanes(c));                                          // folds (and clears) whatever the lanes still hold
    CUDA_TRY(cudaMemsetAsync(c->gradbuf, 0, sizeof(float) * ENERO_GRAD_FLOATS, c->stream));
    return ENERO_OK;
}

// the gradient buffer is complete on the context's stream after this (called by everything that reads it)
extern "C" int enero_ppo_grad_ready(enero_ctx* c) {
    if (!c) return fail(ENERO_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaSetDevice(c->device));
    return join_lanes(c);
}

// validate what a kernel could not report (host copies of the demands / actions)
static int validate_samples(const enero_topo* t, int n, const int32_t* sd, const int32_t* action, const char* who) {
    for (int i = 0; i < n; ++i) {
        const int s = sd[2 * i], dd = sd[2 * i + 1];
        if (s < 0 || dd < 0 || s >= t->N || dd >= t->N || s == dd) return fail(ENERO_ERR_INVALID, std::string(who) + ": bad demand");
        const int k = t->mid_off[s * t->N + dd + 1] - t->mid_off[s * t->N + dd];
        if (action[i] < 0 || action[i] >= k) return fail(ENERO_ERR_INVALID, std::string(who) + ": action out of range");
    }
    return ENERO_OK;
}

static int accumulate_dev(enero_ctx* c, TrainGroup& G, enero_topo* t, int n, const double* loads, const int32_t* sd, const double* bw,
                          const int32_t* action, const float* old_prob, const float* advantage, const float* ret,
                          float inv_m, float entropy_beta, float clipping_val);

// next group of lanes for one topology's share of the minibatch; the context's stream first waits for whatever the
// group's lanes were still doing with its sample buffers
static int acquire_group(enero_ctx* c, TrainGroup** out) {
    TRY(ensure_lanes(c));
    TrainGroup& G = c->groups[c->group_cursor % TRAIN_GROUPS];
    c->group_cursor++;
    for (TrainLane& L : G.lane)
        if (L.busy) CUDA_TRY(cudaStreamWaitEvent(c->stream, L.done, 0));
    *out = &G;
    return ENERO_OK;
}
static int ensure_group_inputs(TrainGroup& G, int n, int E) {
    ReserveScope rs(2.0);
    TRY(G.t_loads.ensure((size_t)n * E * 8)); TRY(G.t_sd.ensure((size_t)n * 8)); TRY(G.t_bw.ensure((size_t)n * 8));
    TRY(G.t_action.ensure((size_t)n * 4)); TRY(G.t_oldp.ensure((size_t)n * 4)); TRY(G.t_adv.ensure((size_t)n * 4));
    TRY(G.t_ret.ensure((size_t)n * 4));
    return ENERO_OK;
}

extern "C" int enero_ppo_accumulate(enero_ctx* c, enero_topo* t, int n, const double* loads, const int32_t* sd,
                                    const double* bw, const int32_t* action, const float* old_prob,
                                    const float* advantage, const float* ret, float inv_m, float entropy_beta,
                                    float clipping_val, int ptr_kind) {
    ENERO_TRACE("enero_ppo_accumulate");
    if (!c || !t || !loads || !sd || !bw || !action || !old_prob || !advantage || !ret || n < 1)
        return fail(ENERO_ERR_INVALID, "enero_ppo_accumulate: bad argument");
    TRY(enero_topo_upload(c, t));
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int E = t->dev.E;
    TrainGroup* G = nullptr;
    TRY(acquire_group(c, &G));
    if (ptr_kind == ENERO_HOST) {
        TRY(validate_samples(t, n, sd, action, "enero_ppo_accumulate"));
        TRY(ensure_group_inputs(*G, n, E));
        CUDA_TRY(cudaMemcpyAsync(G->t_loads.p, loads, (size_t)n * E * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G->t_sd.p, sd, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G->t_bw.p, bw, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G->t_action.p, action, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G->t_oldp.p, old_prob, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G->t_adv.p, advantage, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G->t_ret.p, ret, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        loads = G->t_loads.as<double>(); sd = G->t_sd.as<int>(); bw = G->t_bw.as<double>();
        action = G->t_action.as<int>(); old_prob = G->t_oldp.as<float>(); advantage = G->t_adv.as<float>(); ret = G->t_ret.as<float>();
    }
    return accumulate_dev(c, *G, t, n, loads, sd, bw, action, old_prob, advantage, ret, inv_m, entropy_beta, clipping_val);
}

// ---- device-resident sample store ----------------------------------------------------------------------------------
// ppo_update (train script :326-365) walks the same 835-sample buffer 8 times in minibatches of 55: the samples of
// one topology are uploaded ONCE per update (enero_ppo_store) and every minibatch passes only the indices it drew
// (enero_ppo_accumulate_stored: one small H2D copy + a gather kernel) instead of seven host arrays per call.
struct SampleStore {
    int n = 0;
    DevBuf loads, sd, bw, action, oldp, adv, ret, idx;
};
static std::unordered_map<const enero_topo*, SampleStore*>& stores_of(enero_ctx* c) {
    if (!c->sample_stores) c->sample_stores = new std::unordered_map<const enero_topo*, SampleStore*>();
    return *static_cast<std::unordered_map<const enero_topo*, SampleStore*>*>(c->sample_stores);
}
void train_release_stores(enero_ctx* c) {
    if (!c->sample_stores) return;
    auto* m = static_cast<std::unordered_map<const enero_topo*, SampleStore*>*>(c->sample_stores);
    for (auto& kv : *m) delete kv.second;
    delete m;
    c->sample_stores = nullptr;
}

__global__ void k_gather_samples(int n, int E, const int* __restrict__ idx, const double* __restrict__ loads, const int* __restrict__ sd,
                                 const double* __restrict__ bw, const int* __restrict__ action, const float* __restrict__ oldp,
                                 const float* __restrict__ adv, const float* __restrict__ ret, double* __restrict__ o_loads,
                                 int* __restrict__ o_sd, double* __restrict__ o_bw, int* __restrict__ o_action,
                                 float* __restrict__ o_oldp, float* __restrict__ o_adv, float* __restrict__ o_ret) {
    const int i = blockIdx.x;
    const int j = idx[i];
    for (int e = threadIdx.x; e < E; e += blockDim.x) o_loads[(int64_t)i * E + e] = loads[(int64_t)j * E + e];
    if (threadIdx.x == 0) {
        o_sd[2 * i] = sd[2 * j]; o_sd[2 * i + 1] = sd[2 * j + 1]; o_bw[i] = bw[j]; o_action[i] = action[j];
        o_oldp[i] = oldp[j]; o_adv[i] = adv[j]; o_ret[i] = ret[j];
    }
}

extern "C" int enero_ppo_store(enero_ctx* c, enero_topo* t, int n, const double* loads, const int32_t* sd, const double* bw,
                               const int32_t* action, const float* old_prob, const float* advantage, const float* ret) {
    ENERO_TRACE("enero_ppo_store");
    if (!c || !t || n < 0 || (n > 0 && (!loads || !sd || !bw || !action || !old_prob || !advantage || !ret)))
        return fail(ENERO_ERR_INVALID, "enero_ppo_store: bad argument");
    TRY(enero_topo_upload(c, t));
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(validate_samples(t, n, sd, action, "enero_ppo_store"));
    auto& m = stores_of(c);
    if (!m.count(t)) m[t] = new SampleStore();
    SampleStore& s = *m[t];
    const size_t N = std::max(1, n), E = t->dev.E;
    TRY(s.loads.ensure(N * E * 8)); TRY(s.sd.ensure(N * 8)); TRY(s.bw.ensure(N * 8)); TRY(s.action.ensure(N * 4));
    TRY(s.oldp.ensure(N * 4)); TRY(s.adv.ensure(N * 4)); TRY(s.ret.ensure(N * 4)); TRY(s.idx.ensure(N * 4));
    cudaStream_t st = c->stream;
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(s.loads.p, loads, (size_t)n * E * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.sd.p, sd, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.bw.p, bw, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.action.p, action, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.oldp.p, old_prob, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.adv.p, advantage, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.ret.p, ret, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    s.n = n;
    return ENERO_OK;
}

extern "C" int enero_ppo_accumulate_stored(enero_ctx* c, enero_topo* t, int n, const int32_t* idx, float inv_m,
                                           float entropy_beta, float clipping_val) {
    ENERO_TRACE("enero_ppo_accumulate_stored");
    if (!c || !t || !idx || n < 1) return fail(ENERO_ERR_INVALID, "enero_ppo_accumulate_stored: bad argument");
    auto& m = stores_of(c);
    if (!m.count(t)) return fail(ENERO_ERR_STATE, "enero_ppo_accumulate_stored: no samples stored for this topology (enero_ppo_store)");
    SampleStore& s = *m[t];
    for (int i = 0; i < n; ++i)
        if (idx[i] < 0 || idx[i] >= s.n) return fail(ENERO_ERR_INVALID, "enero_ppo_accumulate_stored: sample index out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int E = t->dev.E;
    TrainGroup* G = nullptr;
    TRY(acquire_group(c, &G));
    TRY(ensure_group_inputs(*G, n, E));
    if (s.idx.cap < (size_t)n * 4) TRY(s.idx.ensure((size_t)n * 4));
    CUDA_TRY(cudaMemcpyAsync(s.idx.p, idx, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_gather_samples<<<n, 64, 0, st>>>(n, E, s.idx.as<int>(), s.loads.as<double>(), s.sd.as<int>(), s.bw.as<double>(), s.action.as<int>(),
                                       s.oldp.as<float>(), s.adv.as<float>(), s.ret.as<float>(), G->t_loads.as<double>(), G->t_sd.as<int>(),
                                       G->t_bw.as<double>(), G->t_action.as<int>(), G->t_oldp.as<float>(), G->t_adv.as<float>(),
                                       G->t_ret.as<float>());
    LAUNCH_CHECK(c);
    return accumulate_dev(c, *G, t, n, G->t_loads.as<double>(), G->t_sd.as<int>(), G->t_bw.as<double>(), G->t_action.as<int>(),
                          G->t_oldp.as<float>(), G->t_adv.as<float>(), G->t_ret.as<float>(), inv_m, entropy_beta, clipping_val);
}

static int accumulate_dev(enero_ctx* c, TrainGroup& G, enero_topo* t, int n, const double* loads, const int32_t* sd, const double* bw,
                          const int32_t* action, const float* old_prob, const float* advantage, const float* ret,
                          float inv_m, float entropy_beta, float clipping_val) {
    const TopoDev& d = t->dev;
    const int E = d.E, Kmax = d.Kmax, T = 5;
    TopoView tv = view_of(t);
    TRY(ensure_ones(c, n));
    CUDA_TRY(cudaEventRecord(G.ready, c->stream));               // the group's samples are in place after this point
    // ---- actor lane: K' candidate graphs per sample (enqueued by the calling thread)
    // scratch sized for up to twice this share of the minibatch, at least 24 samples (a rank of a sharded step sees
    // small, uneven shares), never more than the whole minibatch
    const int m_all = inv_m > 0.f ? std::max(n, (int)std::lround(1.0 / inv_m)) : n;
    const double reserve = (double)std::min(m_all, std::max(2 * n, 24)) / n;
    auto actor_lane = [&]() -> int {
        ReserveScope rs(reserve);
        TrainLane& L = G.lane[ENERO_ACTOR];
        cudaStream_t st = L.st;
        CUDA_TRY(cudaStreamWaitEvent(st, G.ready, 0));
        const int RPD = Kmax * E;
        const int64_t rows = (int64_t)n * RPD;
        TRY(ensure_train_rows(L, T, rows));
        TRY(L.util.ensure((size_t)n * E * 4)); TRY(L.t_nK.ensure((size_t)n * 4));
        TRY(L.t_logits.ensure((size_t)n * Kmax * 4)); TRY(L.t_dlogit.ensure((size_t)n * Kmax * 4));
        TRY(L.t_loss.ensure((size_t)n * 4)); TRY(L.t_ent.ensure((size_t)n * 4));
        int* nK = L.t_nK.as<int>();
        float* sums = L.grad + ENERO_GRAD_STRIDE;
        CUDA_TRY(cudaMemsetAsync(L.t_logits.p, 0, (size_t)n * Kmax * 4, st));
        k_decision_prep<<<(int)(((int64_t)n * E + 255) / 256), 256, 0, st>>>(tv, n, loads, sd, nullptr, nK, L.util.as<float>()); LAUNCH_CHECK(c);
        k_features<<<(int)((rows + MP_TPB - 1) / MP_TPB), MP_TPB, 0, st>>>(tv, n, RPD, sd, bw, nK, L.util.as<float>(), c->w[ENERO_ACTOR],
                                                                         L.t_h[0].as<float>(), L.t_A[0].as<float>()); LAUNCH_CHECK(c);
        IdxTopo idx{rows, E, d.off_f, d.off_s, RPD, nK, d.in_off, d.in_first};
        TRY(train_forward(c, L, idx, ENERO_ACTOR, T, rows));
        GraphsTopo gr{n, Kmax, E, RPD, nK};
        k_readout<GraphsTopo><<<(n * Kmax + 3) / 4, 128, 0, st>>>(gr, c->w[ENERO_ACTOR], L.t_h[T].as<float>(), L.t_logits.as<float>()); LAUNCH_CHECK(c);
        k_ppo_actor_head<<<(n + 3) / 4, 128, 0, st>>>(n, Kmax, nK, L.t_logits.as<float>(), action, old_prob, advantage, inv_m,
                                                      entropy_beta, clipping_val, L.t_dlogit.as<float>(), L.t_loss.as<float>(),
                                                      L.t_ent.as<float>()); LAUNCH_CHECK(c);
        k_sum_into<<<1, 64, 0, st>>>(n, L.t_loss.as<float>(), L.t_ent.as<float>(), sums); LAUNCH_CHECK(c);
        TrainTopo g{n, Kmax, E, RPD, rows, nK};
        TRY(train_backward(c, L, t, g, ENERO_ACTOR, T, L.t_dlogit.as<float>()));
        CUDA_TRY(cudaEventRecord(L.done, st));
        L.busy = true;
        return ENERO_OK;
    };
    // ---- critic lane: one graph per sample, concurrently with the actor lane (enqueued by the helper thread)
    auto critic_lane = [&]() -> int {
        ReserveScope rs(reserve);
        TrainLane& L = G.lane[ENERO_CRITIC];
        cudaStream_t st = L.st;
        CUDA_TRY(cudaStreamWaitEvent(st, G.ready, 0));
        const int64_t rows = (int64_t)n * E;
        TRY(ensure_train_rows(L, T, rows));
        TRY(L.util.ensure((size_t)n * E * 4)); TRY(L.t_nK.ensure((size_t)n * 4));
        TRY(L.t_logits.ensure((size_t)n * 4)); TRY(L.t_dlogit.ensure((size_t)n * 4)); TRY(L.t_loss.ensure((size_t)n * 4));
        const int* ones = c->ones.as<int>();
        float* sums = L.grad + ENERO_GRAD_STRIDE;
        k_decision_prep<<<(int)(((int64_t)n * E + 255) / 256), 256, 0, st>>>(tv, n, loads, sd, nullptr, L.t_nK.as<int>(), L.util.as<float>()); LAUNCH_CHECK(c);
        k_features_critic<<<(int)((rows + MP_TPB - 1) / MP_TPB), MP_TPB, 0, st>>>(tv, n, L.util.as<float>(), c->w[ENERO_CRITIC],
                                                                                L.t_h[0].as<float>(), L.t_A[0].as<float>()); LAUNCH_CHECK(c);
        IdxTopo idx{rows, E, d.off_f, d.off_s, E, ones, d.in_off, d.in_first};
        TRY(train_forward(c, L, idx, ENERO_CRITIC, T, rows));
        GraphsTopo gr{n, 1, E, E, ones};
        k_readout<GraphsTopo><<<(n + 3) / 4, 128, 0, st>>>(gr, c->w[ENERO_CRITIC], L.t_h[T].as<float>(), L.t_logits.as<float>()); LAUNCH_CHECK(c);
        k_ppo_critic_head<<<(n + 127) / 128, 128, 0, st>>>(n, L.t_logits.as<float>(), ret, inv_m, L.t_dlogit.as<float>(), L.t_loss.as<float>());
        LAUNCH_CHECK(c);
        k_sum_into<<<1, 32, 0, st>>>(n, L.t_loss.as<float>(), nullptr, sums); LAUNCH_CHECK(c);
        TrainTopo g{n, 1, E, E, rows, ones};
        TRY(train_backward(c, L, t, g, ENERO_CRITIC, T, L.t_dlogit.as<float>()));
        CUDA_TRY(cudaEventRecord(L.done, st));
        L.busy = true;
        return ENERO_OK;
    };
    if (c->lane_worker) {
        c->lane_worker->run(critic_lane);
        const int ra = actor_lane();
        const int rc = c->lane_worker->wait();                   // always collected: the job refers to this frame
        if (ra != ENERO_OK) return ra;
        if (rc != ENERO_OK) return rc;
    } else {
        TRY(actor_lane());
        TRY(critic_lane());
    }
    // no synchronisation: H2D copies from pageable host memory have left the caller's buffers when
    // cudaMemcpyAsync returns, and the device staging buffers are reused in stream order
    return ENERO_OK;
}

extern "C" int enero_ppo_apply(enero_ctx* c, float lr, float beta1, float beta2, float eps, float clip_norm,
                               int64_t step, float* grad_norm_host, int hist_slot) {
    ENERO_TRACE("enero_ppo_apply");
    if (!c || step < 1 || !(clip_norm > 0) || hist_slot >= ENERO_PPO_HISTORY)
        return fail(ENERO_ERR_INVALID, "enero_ppo_apply: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(join_lanes(c));
    // Keras Adam._prepare_local in fp32: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    const float t = (float)step;
    const float alpha = lr * std::sqrt(1.0f - std::pow(beta2, t)) / (1.0f - std::pow(beta1, t));
    TRY(c->t_norm.ensure(16));
    TRY(c->t_hist.ensure((size_t)ENERO_PPO_HISTORY * 16));
    k_clip_adam<<<1, 1024, 0, c->stream>>>(c->w[0], c->w[1], c->grad[0], c->grad[1], c->adam_m[0], c->adam_v[0], c->adam_m[1],
                                           c->adam_v[1], alpha, beta1, beta2, eps, clip_norm, c->t_norm.as<float>(),
                                           c->gradbuf + 2 * ENERO_GRAD_STRIDE, hist_slot >= 0 ? c->t_hist.as<float>() + 4 * hist_slot : nullptr);
    LAUNCH_CHECK(c);
    CUDA_TRY(cudaMemsetAsync(c->gradbuf + 2 * ENERO_GRAD_STRIDE, 0, 16, c->stream));
    if (grad_norm_host) {
        CUDA_TRY(cudaMemcpyAsync(grad_norm_host, c->t_norm.p, 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return ENERO_OK;
}

extern "C" int enero_ppo_history(enero_ctx* c, int first_slot, int n, float* out) {
    if (!c || !out || first_slot < 0 || n < 1 || first_slot + n > ENERO_PPO_HISTORY) return fail(ENERO_ERR_INVALID, "enero_ppo_history: bad argument");
    if (!c->t_hist.p) return fail(ENERO_ERR_STATE, "enero_ppo_history: no step recorded yet");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(out, c->t_hist.as<float>() + 4 * first_slot, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

extern "C" int enero_ppo_grad_download(enero_ctx* c, float* actor, float* critic, float* sums) {
    if (!c) return fail(ENERO_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(join_lanes(c));
    if (actor) CUDA_TRY(cudaMemcpyAsync(actor, c->grad[0], sizeof(float) * ENERO_NPARAMS, cudaMemcpyDeviceToHost, c->stream));
    if (critic) CUDA_TRY(cudaMemcpyAsync(critic, c->grad[1], sizeof(float) * ENERO_NPARAMS, cudaMemcpyDeviceToHost, c->stream));
    if (sums) CUDA_TRY(cudaMemcpyAsync(sums, c->gradbuf + 2 * ENERO_GRAD_STRIDE, sizeof(float) * 3, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

extern "C" int enero_adam_slots_upload(enero_ctx* c, int which, const float* m, const float* v) {
    if (!c || !m || !v || which < 0 || which > 1) return fail(ENERO_ERR_INVALID, "enero_adam_slots_upload: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(c->adam_m[which], m, sizeof(float) * ENERO_NPARAMS, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->adam_v[which], v, sizeof(float) * ENERO_NPARAMS, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}
extern "C" int enero_adam_slots_download(enero_ctx* c, int which, float* m, float* v) {
    if (!c || !m || !v || which < 0 || which > 1) return fail(ENERO_ERR_INVALID, "enero_adam_slots_download: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(m, c->adam_m[which], sizeof(float) * ENERO_NPARAMS, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(v, c->adam_v[which], sizeof(float) * ENERO_NPARAMS, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ENERO_OK;
}

extern "C" int enero_gae(enero_ctx* c, int n, const float* values, const uint8_t* masks, const double* rewards,
                         double gamma, double lmbda, double* returns, double* advantages) {
    ENERO_TRACE("enero_gae");
    if (!c || n < 1 || !values || !masks || !rewards || !returns || !advantages) return fail(ENERO_ERR_INVALID, "enero_gae: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    DevBuf dv, dm, dr, dret, dadv;
    TRY(dv.ensure((size_t)(n + 1) * 4)); TRY(dm.ensure(n)); TRY(dr.ensure((size_t)n * 8)); TRY(dret.ensure((size_t)n * 8)); TRY(dadv.ensure((size_t)n * 8));
    CUDA_TRY(cudaMemcpyAsync(dv.p, values, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dm.p, masks, n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dr.p, rewards, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    k_gae<<<1, 32, 0, st>>>(n, dv.as<float>(), dm.as<unsigned char>(), dr.as<double>(), gamma, lmbda, dret.as<double>(), dadv.as<double>());
    LAUNCH_CHECK(c);
    CUDA_TRY(cudaMemcpyAsync(returns, dret.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(advantages, dadv.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ENERO_OK;
}
