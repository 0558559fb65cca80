"""Drop-in boundary on the GPU: the reference's own usage pattern (gym.make -> generate_environment
-> checkpoint.restore -> play_middRout_games_sp -> hill climbing; script_eval_on_single_topology.py:
599-623) against the reference-recorded goldens."""
import json
import os

import numpy as np
import pytest

from conftest import CASE_NAMES, GOLDEN, load_case

pytestmark = pytest.mark.gpu

hparams = {'l2': 0.005, 'dropout_rate': 0.1, 'link_state_dim': 20, 'readout_units': 20, 'learning_rate': 0.0002, 'T': 5}


def _dataset(case, folder):
    """re-create <folder>/<name>.graph and TM/<name>.<id>.demands from the fixture"""
    from enero_b200 import datasets as ds
    os.makedirs(os.path.join(folder, "TM"), exist_ok=True)
    name = case["name"]
    open(os.path.join(folder, name + ".graph"), "w").write(case["graph_text"])
    for tm_id, tm in case["tms"].items():
        ds.write_demands_file(os.path.join(folder, "TM", "%s.%s.demands" % (name, tm_id)), np.array(tm))
    if case["sp_seeded"]:
        json.dump(case["topo"]["paths"], open(os.path.join(folder, "shortest_paths.json"), "w"))
    return name


@pytest.fixture(scope="module")
def actor():
    import enero_b200.actorPPOmiddR as actor_mod
    from enero_b200.checkpoint import Checkpoint
    from enero_b200.optim import Adam
    a = actor_mod.myModel(hparams, None, None)
    a.build()
    Checkpoint(model=a, optimizer=Adam(learning_rate=2e-4, beta_1=0.9, epsilon=1e-5)).restore(os.path.join(GOLDEN, "ckpt_ACT-1835"))
    return a


@pytest.mark.parametrize("name", CASE_NAMES)
def test_eval_script_flow(name, actor, tmp_path):
    from enero_b200 import agents, gym_graph
    case = load_case(name)
    folder = str(tmp_path)
    topo_name = _dataset(case, folder)
    env = gym_graph.make('GraphEnv-v16')
    env.seed(9)
    env.generate_environment(folder, topo_name, 100, 100, 0.15)
    env.top_K_critical_demands = True
    g = case["topo"]
    assert (env.numNodes, env.numEdges, env.K, env.firstTrueSize) == (g["N"], g["E"], g["K"], len(g["first"]))
    assert env.src_dst_k_middlepoints == g["mids"]
    assert np.asarray(env.first).tolist() == g["first"] and np.asarray(env.second).tolist() == g["second"]
    agent = agents.PPOMIDDROUTING_SP(env, actor)
    for ep in case["episodes"]:
        timesteps = []
        # one worker process per TM in the reference (eval_on_single_topology.py:92-94): a fresh env has
        # an empty sp_middlepoints_step; a re-used one keeps the previous episode's dict (same quirk here)
        env.sp_middlepoints_step = dict()
        best, _, ospf, best_routing, demands, t0 = agents.play_middRout_games_sp(ep["tm_id"], env, agent, timesteps)
        assert (best, ospf) == (ep["best"], ep["ospf"])
        assert [[int(k.split(":")[0]), int(k.split(":")[1]), m] for k, m in best_routing.items()] == ep["best_routing"]
        assert [[s, d, float(b)] for s, d, b in demands] == ep["reset"]["elig"]
        final, _ = agents.play_DRL_GNN_sp_hill_climbing_games(ep["tm_id"], best_routing, demands, timesteps, t0, folder, topo_name)
        assert final == ep["ls"]["final"]
        assert timesteps[-1][1] == ep["ls"]["final"] or not ep["ls"]["moves"]
        if "plain_hill_final" in ep:
            fin_h, _ = agents.play_sp_hill_climbing_games(ep["tm_id"], folder, topo_name)
            assert fin_h == ep["plain_hill_final"]


def test_greedy_driver_device_loop_equals_per_decision_protocol(actor, tmp_path):
    """play_middRout_games_sp runs the decision loop on the device (CUDA-graph replays) for this package's agent;
    a foreign agent object goes through pred_action_node_distrib_sp / np.argmax / env.step one decision at a time.
    Both must return the same tuple and leave the environment in the same state."""
    from enero_b200 import agents, gym_graph
    case = load_case("tiny12")
    folder = str(tmp_path)
    topo_name = _dataset(case, folder)

    class Foreign:                                   # same policy, not an instance of PPOMIDDROUTING_SP
        def __init__(self, inner):
            self.inner = inner

        def pred_action_node_distrib_sp(self, env, s, d):
            return self.inner.pred_action_node_distrib_sp(env, s, d)

    outs = []
    for wrap in (False, True):
        env = gym_graph.make('GraphEnv-v16')
        env.seed(9)
        env.generate_environment(folder, topo_name, 100, 100, 0.15)
        env.top_K_critical_demands = True
        agent = agents.PPOMIDDROUTING_SP(env, actor)
        ts = []
        r = agents.play_middRout_games_sp(case["episodes"][0]["tm_id"], env, Foreign(agent) if wrap else agent, ts)
        outs.append((r[0], r[2], dict(r[3]), list(r[4]), [v for _, v in ts], env.edge_state[:, 0].tolist(),
                     tuple(env.edgeMaxUti), dict(env.sp_middlepoints), env.episode_over))
    assert outs[0] == outs[1]
    assert outs[0][0] == case["episodes"][0]["best"]


def test_two_models_share_a_context_without_clobbering(actor):
    """a second model built on the same context must not silently replace the first one's weights"""
    import enero_b200.actorPPOmiddR as actor_mod
    ls = np.random.default_rng(0).normal(0, 0.3, (12, 20)).astype(np.float32)
    first = np.array([0, 1, 2, 3, 4, 5], np.int32)
    second = np.array([1, 2, 3, 4, 5, 0], np.int32)
    gid = np.zeros(12, np.int32)
    want = actor(ls, gid, first, second, 12)
    other = actor_mod.myModel(hparams, None, None)
    other.build()                                    # random initialisation: different weights, same context
    got_other = other(ls, gid, first, second, 12)
    assert not np.array_equal(got_other, want)
    assert np.array_equal(actor(ls, gid, first, second, 12), want)          # the first model still computes with ITS weights
    assert np.array_equal(other(ls, gid, first, second, 12), got_other)
    for a, b in zip(actor.get_weights(), other.get_weights()):
        if a.size > 1 and np.abs(a).max() > 0:
            assert not np.array_equal(a, b)


def test_reference_style_agent_through_myModel_call(actor, tmp_path):
    """The reference agent's own feature plumbing (mark_action_sp, old_cummax offsets, concat;
    script :83-129) feeding actor(link_state, graph_id, first, second, num_edges): logits within
    1e-5 element-wise (absolute floor 1e-5) of the recorded ones, greedy actions identical, env.step 9-tuples bit-exact."""
    from enero_b200 import gym_graph
    case = load_case("tiny12")
    folder = str(tmp_path)
    topo_name = _dataset(case, folder)
    env = gym_graph.make('GraphEnv-v16')
    env.seed(9)
    env.generate_environment(folder, topo_name, 100, 100, 0.15)
    env.top_K_critical_demands = True
    ep = case["episodes"][0]
    demand, source, destination = env.reset(ep["tm_id"])
    assert (float(demand), source, destination) == (ep["reset"]["demand"], ep["reset"]["s"], ep["reset"]["d"])
    for stp, dc in zip(ep["steps"], ep["decisions"]):
        mids = env.src_dst_k_middlepoints[str(source) + ':' + str(destination)]
        feats = []
        for m in mids:
            env.mark_action_sp(source, m, source, destination)
            if m != destination:
                env.mark_action_sp(m, destination, source, destination)
            ls = np.zeros((env.numEdges, 20), dtype=np.float32)
            ls[:, 0] = np.divide(env.edge_state[:, 0], env.edge_state[:, 1]).astype(np.float32)
            ls[:, 1] = env.link_capacity_feature
            ls[:, 2] = env.edge_state[:, 2].astype(np.float32)
            feats.append(ls)
            env.edge_state[:, 2] = 0
        first = np.asarray(env.first)[:env.firstTrueSize]
        second = np.asarray(env.second)[:env.firstTrueSize]
        offs_f = [k * (int(first.max()) + 1) for k in range(len(feats))]
        offs_s = [k * (int(second.max()) + 1) for k in range(len(feats))]
        r = actor(np.concatenate(feats), np.repeat(np.arange(len(feats)), env.numEdges),
                  np.concatenate([first + o for o in offs_f]), np.concatenate([second + o for o in offs_s]),
                  env.numEdges * len(feats), training=False)
        assert r.shape == (len(mids), 1)
        want = np.array(dc["logits"])
        np.testing.assert_allclose(r.reshape(-1), want, rtol=1e-5, atol=1e-5)
        action = int(np.argmax(r.reshape(-1)))
        assert action == stp["action"]
        reward, done, _, demand, source, destination, mx, _, std = env.step(action, demand, source, destination)
        assert (float(reward), done, [float(demand), source, destination]) == (stp["reward"], stp["done"], stp["next"])
        assert [mx[0], mx[1], float(mx[2])] == stp["edge_max"] and float(std) == stp["std"]
        assert env.edge_state[:, 0].tolist() == stp["loads_after"]
        if done:
            break


def _climb_move_by_move(env, current):
    """The hill-climbing loop a caller writes against the Env15 method surface (the structure of
    script_eval_on_single_topology.py:317-351,395-435,482-505, restated): every candidate move is evaluated by
    de-allocating the demand, allocating it over the candidate middlepoint, scanning edge_state, and undoing it."""
    def demand_routes(s, d):
        key = "%d:%d" % (s, d)
        if key in env.sp_middlepoints:
            m = env.sp_middlepoints[key]
            return [(s, m), (m, d)]
        return [(s, d)]

    def max_utilisation():
        best, pos = -1000000, 0
        for i in env.graph:
            for j in env.graph[i]:
                u = env.edge_state[pos][0] / env.links_bw[i][j]
                if u > best:
                    best = u
                pos += 1
        return best

    moves = []
    while True:
        next_val, next_state = -1000000, None
        for (s, d, _) in env.list_eligible_demands:
            mids = env.src_dst_k_middlepoints["%d:%d" % (s, d)]
            for action, m in enumerate(mids):
                cur = demand_routes(s, d)
                for a, b in cur:
                    env.decrease_links_utilization_sp(a, b, s, d)
                cand = [(s, m)] if m == d else [(s, m), (m, d)]
                for a, b in cand:
                    env.allocate_to_destination_sp(a, b, s, d)
                value = -max_utilisation()
                for a, b in cand:
                    env.decrease_links_utilization_sp(a, b, s, d)
                for a, b in cur:
                    env.allocate_to_destination_sp(a, b, s, d)
                if value > next_val:
                    next_val, next_state = value, (action, s, d)
        if next_val <= current or abs(next_val - current) < 1e-4:
            break
        action, s, d = next_state
        for a, b in demand_routes(s, d):
            env.decrease_links_utilization_sp(a, b, s, d)
        env.sp_middlepoints.pop("%d:%d" % (s, d), None)
        current = env.step_hill_sp(action, s, d)
        moves.append([action, s, d, -current])
    return -current, moves


@pytest.mark.parametrize("name", ["tri9", "tiny12"])
def test_env15_per_operation_api(name, tmp_path):
    """GraphEnv-v15 driven move by move through allocate_to_destination_sp / decrease_links_utilization_sp /
    step_hill_sp / edge_state / graph / links_bw (Env (LS) row) reproduces the move sequence and the
    final value the reference's own loop recorded, and agrees with the one-launch device Local Search"""
    from enero_b200 import gym_graph
    case = load_case(name)
    folder = str(tmp_path)
    topo_name = _dataset(case, folder)
    env = gym_graph.make('GraphEnv-v15')
    env.seed(9)
    env.generate_environment(folder, topo_name, 100, 100, 0.15)
    for ep in case["episodes"]:
        routing = {"%d:%d" % (s, d): m for s, d, m in ep["best_routing"]}
        cur = env.reset_DRL_hill_sp(ep["tm_id"], dict(routing), None)
        assert cur == ep["ls"]["reset_val"]
        assert [[s, d, bw] for s, d, bw in env.list_eligible_demands] == ep["ls"]["elig"]
        assert env.edge_state[:, 0].tolist() == ep["ls"]["loads0"]
        final, moves = _climb_move_by_move(env, cur)
        assert final == ep["ls"]["final"]
        assert moves == ep["ls"]["moves"]
        # the fused device Local Search from the same reset gives the same answer
        env.reset_DRL_hill_sp(ep["tm_id"], dict(routing), None)
        f2, m2 = env.local_search()
        assert f2 == final and [[a, s, d, v] for a, s, d, v in m2] == moves
