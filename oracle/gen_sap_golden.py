"""TEST INFRASTRUCTURE ONLY -- tests/golden/sap_<case>.json from the reference's OWN Env20 / SAPAgent /
play_sap_games (environment20.py, script_eval_on_single_topology.py:511-557), run unmodified through
oracle/ref_loader.py on the tri9 and tiny12 datasets of the golden cases (small enough for the reference's
exhaustive all_simple_paths), on the cases' own TMs and on scaled-up copies that overflow the links so that the
random-fallback branch (environment20.py:251-254) is exercised.

    python -m oracle.gen_sap_golden
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tests"))

from conftest import load_case                      # noqa: E402
from enero_b200 import datasets as ds               # noqa: E402
from oracle import ref_loader as rl                 # noqa: E402

SCALES = (1.0, 2.5, 6.0)


def run_case(name):
    case = load_case(name)
    folder = tempfile.mkdtemp(prefix="enero_sap_")
    try:
        open(os.path.join(folder, name + ".graph"), "w").write(case["graph_text"])
        os.makedirs(os.path.join(folder, "TM"))
        N = case["topo"]["N"]
        tms = []
        for tm_id in sorted(case["tms"], key=int):
            for sc in SCALES:
                tms.append(np.floor(np.array(case["tms"][tm_id]) * sc))
        for i, tm in enumerate(tms):
            ds.write_demands_file(os.path.join(folder, "TM", "%s.%d.demands" % (name, i)), tm)
        fns = rl.load_eval_functions(folder, name)
        gym = rl.load_gym()
        out = {"name": name, "tms": [tm.tolist() for tm in tms], "episodes": []}
        for i in range(len(tms)):
            # instrumented copy of play_sap_games (script :543-557)
            env = gym.make("GraphEnv-v20")
            env.seed(9)
            env.generate_environment(folder, name, 100, 100)
            demand, s, d = env.reset(i)
            agent = fns["SAPAgent"](env)
            acts = []
            while True:
                a = agent.act(env, demand, s, d)
                acts.append(int(a))
                done, _, demand, s, d, mx, _, _ = env.step(a, demand, s, d)
                if done:
                    break
            fin, _ = fns["play_sap_games"](i)                  # the un-instrumented reference driver
            assert fin == mx[2], (fin, mx[2])
            out["episodes"].append({"actions": acts, "max_uti": float(mx[2])})
            if i == 0:
                out["all_paths"] = {k: [[int(v) for v in p] for p in ps] for k, ps in env.allPaths.items()}
        with open(os.path.join(_ROOT, "tests", "golden", "sap_%s.json" % name), "w") as fp:
            json.dump(out, fp)
        print(name, "episodes", len(tms), "fallbacks", [sum(1 for a in e["actions"] if a < 0) for e in out["episodes"]],
              "max_uti", [round(e["max_uti"], 3) for e in out["episodes"]])
    finally:
        shutil.rmtree(folder, ignore_errors=True)


if __name__ == "__main__":
    for n in ("tri9", "tiny12"):
        run_case(n)
