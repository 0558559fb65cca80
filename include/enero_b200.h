/*
 * enero_b200 -- C ABI of the B200-native ENERO hot path.
 *
 * The reference (BNN-UPC/ENERO) has no FFI: its boundary is two Python class
 * contracts (actor/critic `myModel`, gym-graph `Env16`/`Env15`) plus file formats
 * (SURVEY.md section 8b).  This header is the flat C surface the Python shims in
 * enero_b200/ bind with ctypes; each entry point cites the reference code it
 * replaces (paths relative to the reference tree).
 *
 * Conventions: every function returns 0 on success and a negative code on error,
 * with a message available from enero_last_error() (thread-local).  No exceptions
 * cross the boundary.  The caller owns every buffer it passes; the library owns
 * handles until the matching *_destroy.  One ctx per (process, device); a ctx is
 * thread-compatible, not thread-safe.  A topology keeps ONE device mirror (the deployment model is one process
 * per GPU): uploading it from a context of another device moves the mirror, and batches created on the first
 * device then fail with ENERO_ERR_STATE instead of touching freed memory -- use one topology object per device.  Pointers named *_host are host memory,
 * *_dev are device memory on the ctx's device; `ptr_kind` selects per call
 * (ENERO_HOST = 0, ENERO_DEVICE = 1) where both are accepted.
 */
#ifndef ENERO_B200_H_
#define ENERO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ENERO_OK 0
#define ENERO_ERR_INVALID -1
#define ENERO_ERR_CUDA -2
#define ENERO_ERR_STATE -3

#define ENERO_HOST 0
#define ENERO_DEVICE 1

#define ENERO_ACTOR 0
#define ENERO_CRITIC 1

/* transcendental arithmetic of the message-passing kernels (selu / sigmoid / tanh of actorPPOmiddR.py:45-63):
 * FAST = MUFU.EX2 / MUFU.RCP forms (~1e-6 relative per call); ACCURATE = libdevice expf / tanhf and IEEE division,
 * written as TensorFlow's functors write them.  Measured on 17 812 decisions against the fp64 arbiter
 * (profiles/r02_parity_sweep.json): both modes hold every logit element-wise within
 * 1e-5 * |x| + 1e-5 of the arbiter and of the numpy oracle, both pick the arbiter's arg-max on every decision, and
 * neither is closer to the arbiter than the other (fp32 summation rounding dominates, not the transcendentals);
 * ACCURATE costs 46 % of the throughput, so FAST is the default. */
#define ENERO_MATH_FAST 0
#define ENERO_MATH_ACCURATE 1
#define ENERO_MATH_DEFAULT ENERO_MATH_FAST

#define ENERO_MP_FFMA 0
#define ENERO_MP_TC 1
#define ENERO_MP_DEFAULT ENERO_MP_TC

#define ENERO_H 20            /* link_state_dim  (train_Enero_3top_script.py:83) */
#define ENERO_NPARAMS 4201    /* parameters per model, Keras variable order (:238-248) */
#define ENERO_GRAD_STRIDE 4204  /* floats between the actor and critic halves of the gradient buffer */
#define ENERO_GRAD_FLOATS (2 * ENERO_GRAD_STRIDE + 4)
#define ENERO_PPO_HISTORY 4096  /* per-step records kept on the device by enero_ppo_apply */

typedef struct enero_ctx enero_ctx;
typedef struct enero_topo enero_topo;
typedef struct enero_batch enero_batch;

/* ---- library ------------------------------------------------------------ */
const char* enero_last_error(void);
int enero_version(void);
int enero_device_count(void);

/* ---- context -------------------------------------------------------------- */
int enero_ctx_create(int device, enero_ctx** out);
int enero_ctx_destroy(enero_ctx* ctx);
int enero_ctx_sync(enero_ctx* ctx);
/* kernels launched by this ctx since creation (bench.py's gpu_launches) */
int64_t enero_ctx_launch_count(enero_ctx* ctx);
/* raw cudaStream_t the ctx launches on (for CUDA-event timing by the caller) */
void* enero_ctx_stream(enero_ctx* ctx);
/* Stream selection for every later entry point of this ctx, cuBLAS style (one setter instead of a stream argument on
 * each of the ~40 compute calls): `stream` is the caller's cudaStream_t on the ctx's device, NULL = back to the
 * ctx's own non-blocking stream.  Entry points documented as asynchronous only enqueue on it; the others return after
 * it has drained.  The previous stream is synchronised first. */
int enero_ctx_set_stream(enero_ctx* ctx, void* stream);
/* math mode of every later launch on this ctx (ENERO_MATH_*); the environment variable ENERO_MATH=fast|accurate
 * sets the mode a new ctx starts with.  get returns the mode, or -1 for a null ctx. */
int enero_ctx_set_math(enero_ctx* ctx, int mode);
int enero_ctx_get_math(enero_ctx* ctx);
/* engine of the message-passing iterations of the fused paths (decision batch, value batch, training forward):
 * ENERO_MP_FFMA = k_mp_iter (mat-vecs on the FP32 pipe), ENERO_MP_TC = k_mp_tc (tcgen05 tensor cores, 3xTF32
 * fp32-grade products, accumulators in TMEM).  ENERO_MP=ffma|tc sets what a new ctx starts with. */
int enero_ctx_set_engine(enero_ctx* ctx, int engine);
int enero_ctx_get_engine(enero_ctx* ctx);

/* per-launch CUDA-event timing of the dominant kernel (k_mp_iter) between begin and end:
 * total milliseconds and number of launches (bench.py's roofline leg; off by default) */
int enero_ctx_profile_begin(enero_ctx* ctx);
int enero_ctx_profile_end(enero_ctx* ctx, double* mp_iter_ms, int64_t* mp_iter_launches);

/* FP32 FMA peak of the ctx's device measured live with a packed-FFMA2 register loop (no memory traffic): the
 * denominator of bench.py's FP32 roofline fraction (MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only). */
int enero_ctx_fp32_peak(enero_ctx* ctx, double* tflops);

/* ---- weights ---------------------------------------------------------------
 * Flat fp32[4201] in Keras variable order: FirstLayer/kernel[40,20], bias[20],
 * gru/kernel[20,60], recurrent_kernel[20,60], bias[2,60], Readout1/kernel[20,20],
 * bias[20], Readout2/kernel[20,20], bias[20], Readout3/kernel[20,1], bias[1]
 * (actorPPOmiddR.py:12-39; criticPPO.py:12-35).  `which` = ENERO_ACTOR / ENERO_CRITIC. */
int enero_weights_upload(enero_ctx* ctx, int which, const float* flat_host);
int enero_weights_download(enero_ctx* ctx, int which, float* flat_host);

/* ---- topology (host precompute; no GPU needed until enero_topo_upload) -------
 * Replaces Env16.generate_environment: edge ids (environment16.py:557-569), pair
 * lists _first_second (:507-524), compute_SPs (:478-505, closed form SURVEY Q3) and
 * compute_middlepoint_set_remove_rep_actions_no_loop (:395-476).
 * Links are given in edge-id order (grouped by source node in node-insertion
 * order, each node's links in adjacency order).  `K` is NUM_ACTIONS (clamped to N).
 * sp_off/sp_nodes optionally override the BFS shortest paths with the content of
 * a shortest_paths.json (sp_off[N*N+1] offsets into sp_nodes; pass NULL to compute). */
int enero_topo_build(int n_nodes, int n_edges, const int32_t* edge_src, const int32_t* edge_dst,
                     const double* capacity, int K, const int32_t* sp_off, const int32_t* sp_nodes,
                     enero_topo** out);
int enero_topo_destroy(enero_topo* topo);
/* sizes: out[0]=N out[1]=E out[2]=P out[3]=K out[4]=max(first)+1 out[5]=max(second)+1
 *        out[6]=total middlepoints out[7]=Kmax (largest candidate set) out[8]=total sp edges */
int enero_topo_sizes(const enero_topo* topo, int64_t out[9]);
int enero_topo_get_pairs(const enero_topo* topo, int32_t* first, int32_t* second);
/* shortest paths as edge-id lists: off[N*N+1], edges[out[8]] */
int enero_topo_get_sp_edges(const enero_topo* topo, int32_t* off, int32_t* edges);
/* candidate middlepoints src_dst_k_middlepoints: off[N*N+1], mids[out[6]] */
int enero_topo_get_middlepoints(const enero_topo* topo, int32_t* off, int32_t* mids);
int enero_topo_get_capacity_feature(const enero_topo* topo, float* capf);
int enero_topo_upload(enero_ctx* ctx, enero_topo* topo);

/* ---- dataset ingest (host only) ------------------------------------------------------------
 * Defo_results._get_traffic_matrix (defo_process_results.py:216-225): skip two header lines, then
 * `label src dst bw` per line; tm[src][dst] = floor(bw).  tm fp64[N*N] (zero-filled by the callee);
 * the batch form parses n_files files into tms fp64[n_files,N,N] on n_threads host threads (0 = all). */
int enero_parse_demands(const char* path, int n_nodes, double* tm);
int enero_parse_demands_batch(const char* const* paths, int n_files, int n_nodes, double* tms, int n_threads);

/* ---- GNN, signature-compatible path ------------------------------------------
 * actor.myModel.call (actorPPOmiddR.py:42-69) when graph_id != NULL, critic.myModel.call
 * (criticPPO.py:38-64) when graph_id == NULL (single graph, n_graphs must be 1).
 * link_state fp32[n,20]; graph_id int32[n] ascending; first/second int32[p] (arbitrary,
 * honours the max(index)+1 batching quirk Q1); out fp32[n_graphs]. */
int enero_gnn_forward(enero_ctx* ctx, int which, const float* link_state, const int32_t* graph_id,
                      const int32_t* first, const int32_t* second, int64_t n, int64_t p,
                      int n_graphs, int T, float* out, int ptr_kind);

/* ---- fused decision batch ------------------------------------------------------
 * pred_action_node_distrib_sp (script_eval_on_single_topology.py:83-129): for each of B
 * independent decisions, build the K' candidate graphs of demand (s,d,bw) over link loads
 * `loads` (mark_action_sp, environment16.py:711-725; get_graph_features, script :131-160),
 * run the actor over all of them and soft-max the logits.
 * loads fp64[B,E] (current demand already removed), sd int32[B,2], bw fp64[B].
 * Outputs (any may be NULL): logits/probs fp32[B,Kmax] (entries >= nK[b] are 0),
 * nK int32[B], action int32[B] (np.argmax of probs: first maximum). */
int enero_decide_batch(enero_ctx* ctx, enero_topo* topo, int B, const double* loads,
                       const int32_t* sd, const double* bw, float* logits, float* probs,
                       int32_t* nK, int32_t* action, int ptr_kind);
/* actor probabilities AND critic value of the same B states in one stream round trip (the rollout of
 * train_Enero_3top_script.py:495-500 asks for both at every decision).  Host pointers; probs fp32[B,Kmax]. */
int enero_decide_value_batch(enero_ctx* ctx, enero_topo* topo, int B, const double* loads, const int32_t* sd,
                             const double* bw, float* probs, int32_t* nK, float* values);
/* critic value of B link-load states (critic_get_graph_features, train script :204-230) */
int enero_value_batch(enero_ctx* ctx, enero_topo* topo, int B, const double* loads, float* values,
                      int ptr_kind);

/* ---- device-resident episode batch (GraphEnv-v16 semantics for B episodes) -------- */
int enero_batch_create(enero_ctx* ctx, enero_topo* topo, int B, double percentage_demands,
                       enero_batch** out);
int enero_batch_destroy(enero_batch* b);
/* Env16.reset (environment16.py:642-690) with top_K_critical_demands=True for B traffic
 * matrices tms_host fp64[B,N,N] (already floored, defo_process_results.py:216-225). */
int enero_batch_reset(enero_batch* b, const double* tms_host);
/* actor over the current demand of every live episode; probs_host fp32[B,Kmax] / nK_host may be NULL */
int enero_batch_policy(enero_batch* b, float* probs_host, int32_t* nK_host);
/* critic over the current link loads of every episode; values_host fp32[B] */
int enero_batch_value(enero_batch* b, float* values_host);
/* Env16.step (environment16.py:582-640).  actions_host == NULL: greedy np.argmax of the last
 * enero_batch_policy.  Outputs (nullable): reward fp64[B], done u8[B].  With all three pointers NULL the
 * call only enqueues the kernel (asynchronous); otherwise it returns after the stream has drained. */
int enero_batch_step(enero_batch* b, const int32_t* actions_host, double* reward_host, uint8_t* done_host);
/* Env16.step with host actions plus everything the callers read afterwards (reward, done, link loads, next
 * demand, its bandwidth, edgeMaxUti) in one stream round trip.  Outputs nullable except via `actions`. */
int enero_batch_step_state(enero_batch* b, const int32_t* actions_host, double* reward, uint8_t* done, double* loads,
                           int32_t* cur, double* bw, double* max_uti, int32_t* max_edge);
/* whole greedy episodes (play_middRout_games_sp, script :162-188) without host round trips */
int enero_batch_run_greedy(enero_batch* b);
/* save / restore the mutable episode state on device (rollout restarts, benchmarking) */
int enero_batch_snapshot(enero_batch* b);
int enero_batch_restore(enero_batch* b);
/* current state: loads fp64[B,E]; cur int32[B,2]; bw fp64[B]; max_uti fp64[B];
 * max_edge int32[B] (edge id, -1 = the reference's (0,0,0)); done u8[B]; steps int32[B] */
int enero_batch_get_state(enero_batch* b, double* loads, int32_t* cur, double* bw, double* max_uti,
                          int32_t* max_edge, uint8_t* done, int32_t* steps);
/* list_eligible_demands: elig int32[B,Dmax,2], elig_len int32[B]; Dmax via enero_batch_dmax */
int enero_batch_dmax(enero_batch* b);
int enero_batch_get_eligible(enero_batch* b, int32_t* elig, int32_t* elig_len);
/* per-step history: action int32[B,Dmax], reward fp64[B,Dmax], max_uti fp64[B,Dmax] */
int enero_batch_get_history(enero_batch* b, int32_t* action, double* reward, double* max_uti);
/* best (lowest) max utilisation seen, the initial (OSPF) value and the routing snapshot taken at
 * that step (script :181-184): routing int32[B,Dmax,3] = (s,d,mid) in insertion order, n int32[B] */
int enero_batch_get_best(enero_batch* b, double* best, double* ospf, int32_t* routing, int32_t* n_routing);

/* ---- Local Search (GraphEnv-v15 + HILL_CLIMBING) ------------------------------------
 * reset_DRL_hill_sp (environment15.py:485-544) from a routing per episode (routing
 * int32[B,Dmax,3], n int32[B]; NULL = the batch's own best routing), then the hill-climbing
 * loop of play_DRL_GNN_sp_hill_climbing_games (script :473-509) to convergence on device.
 * mode 0 = critical demands of the routing (explore_neighbourhood_DRL_sp), 1 = all N(N-1)
 * demands from plain shortest paths (play_sp_hill_climbing_games, script :437-471).
 * max_moves = 0 performs the reset only.
 * Outputs (nullable): start fp64[B] (max uti at reset), final fp64[B], n_moves int32[B],
 * moves int32[B,max_moves,3] = (action,s,d), move_val fp64[B,max_moves]. */
int enero_batch_local_search(enero_batch* b, int mode, const int32_t* routing, const int32_t* n_routing,
                             int max_moves, double* start, double* final_val, int32_t* n_moves,
                             int32_t* moves, double* move_val);
/* Per-operation view of the Local-Search loads of one episode, for callers that drive the hill climbing
 * move by move through the Env15 methods (script_eval_on_single_topology.py:317-435 does): add `value`
 * (+TM[s][d] = allocate_to_destination_sp, environment15.py:546-563; -TM[s][d] =
 * decrease_links_utilization_sp, :150-165) on every link of the shortest path src->dst, then rescan the
 * maximum utilisation (step_hill_sp, :389-399).  enero_batch_local_search with max_moves = 0 is the reset
 * (reset_DRL_hill_sp / reset_hill_sp) that puts the loads in place.  Outputs nullable: loads fp64[E]. */
int enero_batch_ls_route_add(enero_batch* b, int episode, int src, int dst, double value, double* loads_host,
                             double* max_uti, int32_t* max_edge);
/* Local-Search state: loads fp64[B,E] and the edge id of the maximum utilisation int32[B] (nullable) */
int enero_batch_get_ls_state(enero_batch* b, double* loads, int32_t* max_edge);
/* critical demands chosen by the last enero_batch_local_search reset: elig int32[B,Dcap,2] */
int enero_batch_get_ls_eligible(enero_batch* b, int32_t* elig, int32_t* elig_len, int cap);

/* ---- sampled-action rollouts (train_Enero_3top_script.py:483-625) -------------------------------------------
 * Every episode of the batch from its reset state to `done`, device resident: per decision the actor over all
 * candidates, the critic on the same link loads, action ~ Categorical(probs) from a Philox4x32-10 stream keyed by
 * (seed, episode, step), Env16.step.  Outputs (nullable), [B,Dmax] unless noted, valid for step < steps[b]:
 * loads fp64[B,Dmax,E] (the state the actor saw), sd int32[B,Dmax,2], bw fp64, action int32, prob fp32 (old policy's
 * probability of the action), value fp32 (critic), reward fp64; steps int32[B]; last_value fp32[B] = critic value of
 * the state each episode ended in (the bootstrap value of :622-625). */
int enero_batch_rollout(enero_batch* b, uint64_t seed, double* loads, int32_t* sd, double* bw, int32_t* action,
                        float* prob, float* value, double* reward, int32_t* steps, float* last_value);

/* ---- SAP baseline (GraphEnv-v20 + SAPAgent, results[8] of the evaluation records) -----------------------
 * enero_topo_build_ksp: Env20.compute_SPs (environment20.py:165-196): per node pair the K simple paths with the
 * fewest hops among those of at most 2*diameter links, ties in lexicographic node order (host precompute, idempotent).
 * enero_topo_get_ksp: pair_off int32[N*N+1] -> path index, path_off int32[n_paths+1] -> nodes; sizes out[0]=n_paths,
 * out[1]=total nodes. */
int enero_topo_build_ksp(enero_topo* topo, int K);
int enero_topo_ksp_sizes(const enero_topo* topo, int64_t out[2]);
int enero_topo_get_ksp(const enero_topo* topo, int32_t* pair_off, int32_t* path_off, int32_t* nodes);
/* play_sap_games (script_eval_on_single_topology.py:511-557) for the B traffic matrices tms_host fp64[B,N,N]
 * (NULL = the matrices of the last enero_batch_reset): every demand in descending bandwidth order
 * (environment20.py:124-163) goes to the first of its K shortest paths on which no link would exceed its capacity
 * (SAPAgent.act), else to random.randint(0, len(paths)-1) of Python's Mersenne Twister seeded with `seed`
 * (env.seed(SEED), environment20.py:98-100, 251-254).  Outputs (nullable): max_uti fp64[B] = maxLinkUti[2] of the
 * last step; actions int32[B, N(N-1)] = what SAPAgent.act returned per demand (-1 = no path had room). */
int enero_batch_sap(enero_batch* b, const double* tms_host, int K, uint32_t seed, double* max_uti, int32_t* actions);

/* ---- PPO update (train_Enero_3top_script.py:264-377) ---------------------------------------
 * The gradient buffer is one device array fp32[ENERO_GRAD_FLOATS]:
 *   [0, 4201)                       d total_loss / d actor weights   (Keras variable order)
 *   [STRIDE, STRIDE + 4201)         d total_loss / d critic weights
 *   [2*STRIDE + 0 / 1 / 2]          sum of actor sample losses / entropies / critic sample losses
 * A rank of a sharded PPO step all-reduces (sum) exactly this array, then every rank applies. */
void* enero_ppo_grad_buffer(enero_ctx* ctx);
int enero_ppo_zero_grad(enero_ctx* ctx);
/* enero_ppo_accumulate* run the (up to three topologies) x (actor, critic) passes of a minibatch concurrently on
 * internal streams ("lanes"), each with its own gradient accumulator; enero_ppo_grad_ready joins them into the
 * ctx's stream and folds the lane gradients into the gradient buffer in a fixed order (deterministic).  A caller
 * that reads enero_ppo_grad_buffer itself (the all-reduce of a sharded step) calls it first; enero_ppo_apply,
 * enero_ppo_grad_download and enero_ppo_zero_grad do so themselves. */
int enero_ppo_grad_ready(enero_ctx* ctx);
/* _train_step_combined (train script :293-315), gradient part, for n samples that share one topology:
 * actor forward over the K' candidate graphs of demand sd[i] (bandwidth bw[i]) on link loads loads[i]
 * (the same state pred_action_distrib_sp saw), critic forward on the same loads, the per-sample losses
 * of _actor_step (:273-291) / _critic_step (:264-271), and the backward pass.  Adds the gradient of
 *   inv_m * sum_i [ loss_i - entropy_beta * entropy_i + critic_loss_i ]
 * and the three loss sums into the gradient buffer (inv_m = 1 / minibatch size gives reduce_mean).
 * loads fp64[n,E]; sd int32[n,2]; bw fp64[n]; action int32[n]; old_prob fp32[n] = old policy's
 * probability of the action; advantage, ret fp32[n].  All pointers host (ENERO_HOST) or device. */
int enero_ppo_accumulate(enero_ctx* ctx, enero_topo* topo, int n, const double* loads, const int32_t* sd,
                         const double* bw, const int32_t* action, const float* old_prob,
                         const float* advantage, const float* ret, float inv_m, float entropy_beta,
                         float clipping_val, int ptr_kind);
/* ppo_update's buffer (train script :326-365) kept on the device: the samples of one topology are uploaded once
 * per update (enero_ppo_store, same arrays as enero_ppo_accumulate, host pointers) and each minibatch passes only the
 * indices it drew into that store (idx_host int32[n]); otherwise identical to enero_ppo_accumulate. */
int enero_ppo_store(enero_ctx* ctx, enero_topo* topo, int n, const double* loads, const int32_t* sd, const double* bw,
                    const int32_t* action, const float* old_prob, const float* advantage, const float* ret);
int enero_ppo_accumulate_stored(enero_ctx* ctx, enero_topo* topo, int n, const int32_t* idx_host, float inv_m,
                                float entropy_beta, float clipping_val);
/* tf.clip_by_global_norm(grad, clip_norm) over all 22 tensors + Adam.apply_gradients (train script
 * :317-320; Keras Adam dense update).  `step` = optimizer.iterations + 1.  Zeroes the gradient buffer.
 * grad_norm_host (nullable) receives the global norm before clipping -- and makes the call wait for the stream.
 * hist_slot >= 0 records (norm, actor-loss sum, entropy sum, critic-loss sum) of this step in slot hist_slot of a
 * device history instead, read back for n consecutive slots by enero_ppo_history (out fp32[n,4]): the 128 Adam
 * steps of one ppo_update then run without a single host synchronisation. */
int enero_ppo_apply(enero_ctx* ctx, float lr, float beta1, float beta2, float eps, float clip_norm,
                    int64_t step, float* grad_norm_host, int hist_slot);
int enero_ppo_history(enero_ctx* ctx, int first_slot, int n, float* out);
/* host copies of the gradient buffer halves (any pointer may be NULL): actor/critic fp32[4201], sums fp32[3] */
int enero_ppo_grad_download(enero_ctx* ctx, float* actor, float* critic, float* sums);
/* Adam slot variables m, v of one model, flat fp32[4201] (checkpoint .OPTIMIZER_SLOT entries) */
int enero_adam_slots_upload(enero_ctx* ctx, int which, const float* m, const float* v);
int enero_adam_slots_download(enero_ctx* ctx, int which, float* m, float* v);
/* get_advantages (train script :367-377): GAE over n steps in fp64.  values fp32[n+1], masks u8[n],
 * rewards fp64[n] -> returns fp64[n], advantages fp64[n] ((adv - mean) / (std + 1e-10)).  Host pointers. */
int enero_gae(enero_ctx* ctx, int n, const float* values, const uint8_t* masks, const double* rewards,
              double gamma, double lmbda, double* returns, double* advantages);

#ifdef __cplusplus
}
#endif
#endif /* ENERO_B200_H_ */
