"""Agent and episode drivers of script_eval_on_single_topology.py, on the GPU engine.

    PPOMIDDROUTING_SP.pred_action_node_distrib_sp   script :83-129   -> enero_decide_batch
    play_middRout_games_sp                           script :162-188
    play_DRL_GNN_sp_hill_climbing_games              script :473-509  -> k_local_search
    play_sp_hill_climbing_games                      script :437-471  -> k_local_search (mode 1)
    SAPAgent / play_sap_games                        script :511-557  -> k_sap_order + k_sap_run
"""
from __future__ import annotations

import time as tt

import numpy as np

from . import gym_graph
from .models import ActorModel

SEED = 9
hparamsDRLSP = {'l2': 0.005, 'dropout_rate': 0.1, 'link_state_dim': 20, 'readout_units': 20,
                'learning_rate': 0.0002, 'T': 5}


class PPOMIDDROUTING_SP:
    def __init__(self, env_training, actor=None):
        self.K = env_training.K
        self.actor = actor if actor is not None else ActorModel(hparamsDRLSP)
        if not self.actor.built:
            self.actor.build()
        self.listQValues = None
        self.softMaxQValues = None

    def pred_action_node_distrib_sp(self, env, source, destination):
        """Probabilities over the candidate middlepoints of (source, destination) given the
        environment's current link loads; one fused launch sequence instead of K' feature
        builds + ~70 eager TF ops."""
        t = env.topology
        loads = np.ascontiguousarray(env.edge_state[:, 0][None])
        sd = np.array([[source, destination]], dtype=np.int32)
        bw = np.array([env.TM[source][destination]], dtype=np.float64)
        self.actor.bind()
        logits, probs, nK, action = self.actor.context.decide_batch(t, loads, sd, bw)
        k = int(nK[0])
        self.listQValues = logits[:1, :k]
        self.softMaxQValues = probs[:1, :k]
        return probs[0, :k], {"num_candidates": k, "logits": logits[0, :k]}


def play_middRout_games_sp(tm_id, env_middRout_sp, agent, timesteps):
    """script_eval_on_single_topology.py:162-188, same return tuple
    (best maxUti, seconds, OSPF maxUti, best_routing, list_of_demands_to_change, time_start_DRL).

    With this package's agent and environment the decision loop of :169-186 (actor over every candidate, arg-max,
    env.step, best-routing bookkeeping) runs on the device without a host round trip per decision
    (enero_batch_run_greedy: one CUDA-graph replay per decision); `timesteps` gets one row per strict improvement,
    spaced by the amortised per-decision time.  Any other agent / environment pair goes through the reference's
    per-decision protocol (pred_action_node_distrib_sp -> np.argmax -> env.step)."""
    env = env_middRout_sp
    demand, source, destination = env.reset(tm_id)
    initMaxUti = env.edgeMaxUti[2]
    OSPF_init = initMaxUti
    list_of_demands_to_change = env.list_eligible_demands
    timesteps.append((0, initMaxUti))
    start = tt.time()
    time_start_DRL = start
    if (isinstance(agent, PPOMIDDROUTING_SP) and hasattr(env, "run_greedy_episode")
            and agent.actor.context is env.context):
        agent.actor.bind()
        ep = env.run_greedy_episode()
        end = tt.time()
        dt = (end - start) / max(1, len(ep["max_uti"]))
        for k, v in enumerate(ep["max_uti"]):
            if v < initMaxUti:
                initMaxUti = float(v)
                timesteps.append(((k + 1) * dt, initMaxUti))
        return initMaxUti, end - start, OSPF_init, ep["best_routing"], list_of_demands_to_change, time_start_DRL
    best_routing = env.sp_middlepoints_step.copy()
    done = False
    while not done:
        action_dist, _ = agent.pred_action_node_distrib_sp(env, source, destination)
        action = np.argmax(action_dist)
        _, done, _, demand, source, destination, maxLinkUti, _, _ = env.step(action, demand, source, destination)
        if maxLinkUti[2] < initMaxUti:
            initMaxUti = maxLinkUti[2]
            best_routing = env.sp_middlepoints_step.copy()
            timesteps.append((tt.time() - time_start_DRL, initMaxUti))
    end = tt.time()
    return initMaxUti, end - start, OSPF_init, best_routing, list_of_demands_to_change, time_start_DRL


def _make_env15(dataset_folder, topology_name, percentage_demands, K=100):
    env = gym_graph.make('GraphEnv-v15')
    env.seed(SEED)
    env.generate_environment(dataset_folder, topology_name, 100, K, percentage_demands)
    return env


def play_DRL_GNN_sp_hill_climbing_games(tm_id, best_routing, list_of_demands_to_change, timesteps, time_start_DRL,
                                        dataset_folder, topology_name, percentage_demands=0.15, env=None):
    """script :473-509 -> (final maxUti, seconds)."""
    env = env if env is not None else _make_env15(dataset_folder, topology_name, percentage_demands)
    start = tt.time()
    env.reset_DRL_hill_sp(tm_id, best_routing, list_of_demands_to_change)
    final, moves = env.local_search()
    end = tt.time()
    for (_, _, _, val) in moves:
        timesteps.append((end - time_start_DRL, val))
    return final, end - start


def play_sp_hill_climbing_games(tm_id, dataset_folder, topology_name, percentage_demands=0.15, env=None):
    """script :437-471 -> (final maxUti, seconds)."""
    env = env if env is not None else _make_env15(dataset_folder, topology_name, percentage_demands)
    start = tt.time()
    env.reset_hill_sp(tm_id)
    final, _ = env.local_search()
    return final, tt.time() - start


class SAPAgent:
    """script :511-541.  The per-demand ``act`` lives inside k_sap_run (one device loop per traffic matrix)."""

    def __init__(self, env):
        self.K = env.K


def play_sap_games(tm_id, dataset_folder, topology_name, env=None):
    """script :543-557 -> (final maxLinkUti[2], seconds)."""
    if env is None:
        env = gym_graph.make('GraphEnv-v20')
        env.seed(SEED)
        env.generate_environment(dataset_folder, topology_name, 100, 100)
    env.reset(tm_id)
    SAPAgent(env)
    start = tt.time()
    final, _ = env.run_sap()
    return final, tt.time() - start
