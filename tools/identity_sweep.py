"""One-off sweep (not a test): greedy action sequences, best DRL utilisation and Local-Search results of the GPU
path against the oracle on many fresh synthetic TMs of the golden topologies.
    python tools/identity_sweep.py tiny12 400 eli20 60"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import case_graph_file, golden_weights, load_case      # noqa: E402
from enero_b200 import datasets as ds                                 # noqa: E402
from enero_b200.engine import EpisodeBatch                            # noqa: E402
from enero_b200.runtime import Context                                # noqa: E402
from enero_b200.topology import Topology                              # noqa: E402
from oracle import enero_oracle as orc                                # noqa: E402


def main(argv):
    w = golden_weights()
    ctx = Context(0)
    ctx.upload_weights(0, [w[k] for k in orc.WEIGHT_NAMES])
    out = {}
    ls_limit = int(os.environ.get("SWEEP_LS_LIMIT", "1000000"))
    for name, n in zip(argv[0::2], argv[1::2]):
        n = int(n)
        case = load_case(name)
        gf = case_graph_file(case, tempfile.mkdtemp())
        topo, ot = Topology(gf, K=100), orc.Topology(gf, K=100)
        sigma = ds.sigma_for_mean_utilisation(gf)
        tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 90000 + j) for j in range(n)])
        eb = EpisodeBatch(ctx, topo, n)
        eb.reset(tms)
        eb.run_greedy()
        best, ospf, _ = eb.best()
        act, _, _ = eb.history()
        steps = eb.state()["steps"]
        ls = eb.local_search(None, mode=0)
        bad_actions = bad_best = bad_ls = decisions = 0
        for i in range(n):
            obest, oospf, orouting, oactions = orc.play_middrout_greedy(w, orc.Env16(ot), tms[i])
            decisions += len(oactions)
            bad_actions += act[i, :steps[i]].tolist() != oactions
            bad_best += (best[i], ospf[i]) != (obest, oospf)
            if i < ls_limit:                     # the oracle's Local Search is slow on large graphs
                e15 = orc.Env15(ot)
                ofinal, _ = orc.local_search(e15, e15.reset_DRL_hill_sp(tms[i], orouting))
                bad_ls += ls["final"][i] != ofinal
        out[name] = {"tms": n, "decisions": decisions, "action_sequences_differ": int(bad_actions),
                     "best_utilisation_differs": int(bad_best), "local_search_final_differs": int(bad_ls)}
        eb.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1:])
