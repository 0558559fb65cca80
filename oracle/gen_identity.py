"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/identity_configs.json: oracle greedy action sequences,
best / OSPF utilisations and Local-Search results on synthetic TMs of every BASELINE.json configuration size
(cfg3 and the three cfg4 topologies of enero_b200.workloads, plus one full greedy episode at 180 links), so that the
GPU identity tests at those sizes need neither /root/reference nor minutes of oracle time on the GPU box.

    OMP_NUM_THREADS=1 python -m oracle.gen_identity

The topologies and TMs are re-created in the tests from (N, L, seed) and the TM seeds (datasets.zoo_like /
uniform_tm are deterministic); the file stores only the oracle's answers.  tests/test_oracle_golden.py re-runs one
episode per entry on the CPU to keep this file honest.
"""
import json
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

# name, (N, L, topology seed), number of TMs, run the oracle's Local Search?
SPECS = [
    ("cfg3_hur24", (24, 37, 1002), 8, True),
    ("cfg4_top1", (20, 31, 1003), 8, True),
    ("cfg4_top2", (23, 25, 1004), 8, True),
    ("cfg4_top3", (17, 31, 1005), 8, True),
    ("links180", (60, 90, 1021), 1, False),
]
FIRST_TM_SEED = 2000


def build(spec):
    from enero_b200 import datasets as ds
    n, l, seed = spec
    links, caps = ds.zoo_like(n, l, seed)
    p = os.path.join(tempfile.mkdtemp(prefix="enero_ident_"), "t.graph")
    ds.write_graph_file(p, n, links, caps)
    gf = ds.parse_graph_file(p)
    return gf, ds.sigma_for_mean_utilisation(gf)


def _episode(args):
    name, spec, j, do_ls = args
    from enero_b200 import datasets as ds
    from oracle import enero_oracle as orc
    g = _episode.__dict__
    if g.get("name") != name:
        gf, sigma = build(spec)
        z = np.load(os.path.join(_ROOT, "tests", "golden", "ckpt_ACT-1835_weights.npz"))
        g.update(name=name, gf=gf, sigma=sigma, w={k: z[k] for k in z.files},
                 ot=orc.Topology(gf, K=100, paths=ds.bfs_lexmin_paths(gf)))
    tm = ds.uniform_tm(g["gf"].num_nodes, g["sigma"], FIRST_TM_SEED + j)
    best, ospf, routing, actions = orc.play_middrout_greedy(g["w"], orc.Env16(g["ot"]), tm)
    ls_final, n_moves = None, None
    if do_ls:
        e15 = orc.Env15(g["ot"])
        ls_final, moves = orc.local_search(e15, e15.reset_DRL_hill_sp(tm, routing))
        ls_final, n_moves = float(ls_final), len(moves)
    return name, j, {"tm_seed": FIRST_TM_SEED + j, "actions": actions, "best": float(best), "ospf": float(ospf),
                     "routing": [[int(s), int(d), int(m)] for (s, d), m in routing.items()],
                     "ls_final": ls_final, "ls_moves": n_moves}


def main():
    jobs = [(name, spec, j, ls) for name, spec, n, ls in SPECS for j in range(n)]
    with mp.get_context("spawn").Pool(min(os.cpu_count(), len(jobs))) as pool:
        res = pool.map(_episode, jobs, chunksize=1)
    out = {}
    for name, spec, n, ls in SPECS:
        gf, sigma = build(spec)
        eps = [r[2] for r in sorted((r for r in res if r[0] == name), key=lambda r: r[1])]
        out[name] = {"spec": list(spec), "sigma": float(sigma), "episodes": eps}
        print(name, "E=%d" % (2 * spec[1]), "decisions", [len(e["actions"]) for e in eps])
    with open(os.path.join(_ROOT, "tests", "golden", "identity_configs.json"), "w") as fp:
        json.dump(out, fp)


if __name__ == "__main__":
    main()
