"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.json by running the
reference's OWN environment / Local-Search code (unmodified, via
oracle/ref_loader.py) on seeded synthetic datasets.  Run in the build container:

    python -m oracle.gen_golden

The GNN cannot be run through the reference (TensorFlow is not installable), so
the greedy agent that drives the reference environments here is
oracle.enero_oracle.actor_forward fed with features built by the reference's own
``mark_action_sp`` / ``edge_state`` (script_eval_on_single_topology.py:83-160
restated in numpy), with the shipped ckpt_ACT-1835 weights.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from enero_b200 import checkpoint as ckpt          # noqa: E402  (file-format reader)
from enero_b200 import datasets as ds              # noqa: E402  (synthetic generators)
from oracle import enero_oracle as orc             # noqa: E402
from oracle import ref_loader as rl                # noqa: E402

CKPT = os.path.join(rl.REFERENCE_ROOT, "modelsSP_3top_15_B_NEW", "ckpt_ACT-1835")

# name, N, L, topo seed, #TMs, pre-seed shortest_paths.json?, plain hill climbing?
CASES = [
    ("tiny12", 12, 16, 1000, 2, False, True),
    ("eli20", 20, 30, 1001, 2, True, False),     # cfg1: EliBackbone-sized
    ("tri9", 9, 13, 1007, 2, False, True),
    ("zoo30", 30, 50, 1000, 1, True, False),     # cfg2: 100-link
    ("hur24", 24, 37, 1002, 1, True, False),     # cfg3: HurricaneElectric-sized (the topology of workloads.run_cfg3)
]


def _f(x):
    return float(x)


class RefAgent:
    """pred_action_node_distrib_sp (script :83-129) against the reference env,
    numpy instead of TF for the tensor plumbing, oracle GNN for the actor."""

    def __init__(self, weights, trace):
        self.w = weights
        self.trace = trace

    def pred_action_node_distrib_sp(self, env, source, destination):
        mids = env.src_dst_k_middlepoints[str(source) + ':' + str(destination)]
        feats, marks = [], []
        for m in mids:
            env.mark_action_sp(source, m, source, destination)
            if m != destination:
                env.mark_action_sp(m, destination, source, destination)
            ls = np.zeros((env.numEdges, 20), dtype=np.float32)
            ls[:, 0] = np.divide(env.edge_state[:, 0], env.edge_state[:, 1]).astype(np.float32)
            ls[:, 1] = np.asarray(env.link_capacity_feature, dtype=np.float32)
            ls[:, 2] = env.edge_state[:, 2].astype(np.float32)
            marks.append([_f(v) for v in env.edge_state[:, 2]])
            feats.append(ls)
            env.edge_state[:, 2] = 0
        first = np.asarray(env.first, dtype=np.int32)[:env.firstTrueSize]
        second = np.asarray(env.second, dtype=np.int32)[:env.firstTrueSize]
        b = orc.batch_candidates(feats, first, second)
        logits = orc.actor_forward(self.w, b["link_state"], b["graph_id"], b["first"],
                                   b["second"], b["num_edges"]).reshape(-1)
        probs = orc.softmax(logits)
        self.trace.append({"s": int(source), "d": int(destination),
                           "loads_before": [_f(v) for v in env.edge_state[:, 0]],
                           "marks": marks if len(self.trace) < 3 else None,
                           "logits": [_f(v) for v in logits],
                           "probs": [_f(v) for v in probs]})
        return probs, b


def run_case(name, N, L, seed, n_tms, seed_sp, plain_hill, out_dir):
    folder = tempfile.mkdtemp(prefix="enero_golden_")
    try:
        ds.make_synthetic_dataset(folder, name, N, L, seed, n_tms, write_sp_cache=seed_sp)
        graph_text = open(os.path.join(folder, name + ".graph")).read()
        env16 = rl.make_env16(folder, name)
        fns = rl.load_eval_functions(folder, name)
        weights = dict(zip(orc.WEIGHT_NAMES, ckpt.Bundle.load(CKPT).model_weights()))
        paths = json.load(open(os.path.join(folder, "shortest_paths.json")))
        edges = [None] * env16.numEdges
        for key, pos in env16.edgesDict.items():
            i, j = key.split(':')
            edges[pos] = [int(i), int(j)]
        case = {
            "name": name, "graph_text": graph_text, "tms": {}, "sp_seeded": seed_sp,
            "topo": {
                "N": env16.numNodes, "E": env16.numEdges, "K": env16.K,
                "edges": edges,
                "cap": [_f(c) for c in env16.edge_state[:, 1]],
                "link_capacity_feature": [_f(c) for c in np.asarray(env16.link_capacity_feature)],
                "first": [int(v) for v in np.asarray(env16.first)],
                "second": [int(v) for v in np.asarray(env16.second)],
                "paths": paths,
                "mids": {k: [int(m) for m in v] for k, v in env16.src_dst_k_middlepoints.items()},
            },
            "episodes": [],
        }
        for tm_id in range(n_tms):
            tm = ds.parse_demands_file(os.path.join(folder, "TM", "%s.%d.demands" % (name, tm_id)), N)
            case["tms"][str(tm_id)] = [[_f(v) for v in row] for row in tm]
            trace = []
            agent = RefAgent(weights, trace)
            ep = {"tm_id": tm_id}
            # --- one greedy DRL episode, instrumented copy of the loop at script :162-188
            demand, s, d = env16.reset(tm_id)
            ep["reset"] = {"demand": _f(demand), "s": int(s), "d": int(d),
                           "elig": [[int(a), int(b), _f(c)] for a, b, c in env16.list_eligible_demands],
                           "loads": [_f(v) for v in env16.edge_state[:, 0]],
                           "edge_max": [int(env16.edgeMaxUti[0]), int(env16.edgeMaxUti[1]), _f(env16.edgeMaxUti[2])]}
            steps = []
            while True:
                probs, _ = agent.pred_action_node_distrib_sp(env16, s, d)
                a = int(np.argmax(probs))
                r, done, _, demand, s, d, mx, _, std = env16.step(a, demand, s, d)
                steps.append({"action": a, "reward": _f(r), "done": bool(done),
                              "next": [_f(demand), int(s), int(d)],
                              "edge_max": [int(mx[0]), int(mx[1]), _f(mx[2])], "std": _f(std),
                              "loads_after": [_f(v) for v in env16.edge_state[:, 0]]})
                if done:
                    break
            ep["steps"] = steps
            ep["decisions"] = trace
            # --- the reference's own driver on a fresh env (pins best-routing bookkeeping)
            timesteps = []
            env16.sp_middlepoints_step = dict()     # as in a fresh worker process (script :599-603)
            best, _, ospf, best_routing, dem_list, t0 = fns["play_middRout_games_sp"](
                tm_id, env16, RefAgent(weights, []), timesteps)
            ep["best"] = _f(best)
            ep["ospf"] = _f(ospf)
            ep["best_routing"] = [[int(k.split(':')[0]), int(k.split(':')[1]), int(m)]
                                  for k, m in best_routing.items()]
            # --- Local Search from the DRL routing (script :473-509), instrumented
            env15 = rl.make_env15(folder, name)
            cur = env15.reset_DRL_hill_sp(tm_id, dict(best_routing), dem_list)
            ls = {"reset_val": _f(cur),
                  "elig": [[int(a), int(b), _f(c)] for a, b, c in env15.list_eligible_demands],
                  "loads0": [_f(v) for v in env15.edge_state[:, 0]], "moves": []}
            hc = fns["HILL_CLIMBING"](env15)
            while True:
                nxt, state = hc.explore_neighbourhood_DRL_sp(env15)
                if nxt <= cur or abs((-1) * nxt - (-1) * cur) < 1e-4:
                    ls["last_next"] = _f(nxt)
                    break
                a, s2, d2 = state
                key = str(s2) + ':' + str(d2)
                if key in env15.sp_middlepoints:
                    m = env15.sp_middlepoints[key]
                    env15.decrease_links_utilization_sp(s2, m, s2, d2)
                    env15.decrease_links_utilization_sp(m, d2, s2, d2)
                    del env15.sp_middlepoints[key]
                else:
                    env15.decrease_links_utilization_sp(s2, d2, s2, d2)
                cur = env15.step_hill_sp(a, s2, d2)
                ls["moves"].append([int(a), int(s2), int(d2), _f(-cur)])
            ls["final"] = _f(-cur)
            # cross-check with the un-instrumented reference driver
            fin, _ = fns["play_DRL_GNN_sp_hill_climbing_games"](tm_id, dict(best_routing), dem_list, [], t0)
            assert fin == -cur, (fin, -cur)
            ep["ls"] = ls
            if plain_hill:
                fin_h, _ = fns["play_sp_hill_climbing_games"](tm_id)
                ep["plain_hill_final"] = _f(fin_h)
            case["episodes"].append(ep)
        with open(os.path.join(out_dir, "case_%s.json" % name), "w") as fp:
            json.dump(case, fp)
        print("wrote", name, "E=%d" % env16.numEdges, "steps=%s" % [len(e["steps"]) for e in case["episodes"]],
              "ls_moves=%s" % [len(e["ls"]["moves"]) for e in case["episodes"]])
    finally:
        shutil.rmtree(folder, ignore_errors=True)


def main():
    out_dir = os.path.join(_ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for c in CASES:
        if only and c[0] not in only:
            continue
        run_case(*c, out_dir=out_dir)
    # the shipped checkpoint's weights, so GPU-box tests need no /root/reference
    b = ckpt.Bundle.load(CKPT)
    np.savez(os.path.join(out_dir, "ckpt_ACT-1835_weights.npz"),
             **dict(zip(orc.WEIGHT_NAMES, b.model_weights())))
    for ext in (".index", ".data-00000-of-00001"):
        shutil.copyfile(CKPT + ext, os.path.join(out_dir, "ckpt_ACT-1835" + ext))


if __name__ == "__main__":
    main()
