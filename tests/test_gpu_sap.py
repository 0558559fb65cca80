"""SAP baseline on the GPU (k_sap_order + k_sap_run) against the episodes the reference's own Env20 / SAPAgent
recorded (tests/golden/sap_*.json: per-demand actions and the final maximum utilisation, random-fallback branch
included -- Python's Mersenne Twister is reproduced on the device) and against the oracle on larger graphs.
Integer / fp64 work: everything bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, case_graph_file, load_case
from oracle import enero_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from enero_b200.runtime import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["tri9", "tiny12"])
def test_sap_matches_reference_recordings(ctx, name, tmp_path):
    from enero_b200.engine import EpisodeBatch
    from enero_b200.topology import Topology
    gold = json.load(open(os.path.join(GOLDEN, "sap_%s.json" % name)))
    topo = Topology(case_graph_file(load_case(name), tmp_path), K=100)
    tms = np.array(gold["tms"])
    eb = EpisodeBatch(ctx, topo, len(tms))
    mx, acts = eb.sap(tms, want_actions=True)
    for i, ep in enumerate(gold["episodes"]):
        assert acts[i].tolist() == ep["actions"], "TM %d" % i
        assert mx[i] == ep["max_uti"]
    # the matrices of a reset batch, one at a time, give the same answers (batch independence)
    eb1 = EpisodeBatch(ctx, topo, 1)
    for i in (1, 2):
        eb1.reset(tms[i:i + 1])
        assert eb1.sap(None)[0] == gold["episodes"][i]["max_uti"]
    eb.close()
    eb1.close()


@pytest.mark.parametrize("name,n_tms,scale", [("eli20", 6, 3.0), ("zoo30", 4, 2.0)])
def test_sap_matches_oracle_on_larger_graphs(ctx, name, n_tms, scale, tmp_path):
    from enero_b200 import datasets as ds
    from enero_b200.engine import EpisodeBatch
    from enero_b200.topology import Topology
    gf = case_graph_file(load_case(name), tmp_path)
    topo, ot = Topology(gf, K=100), orc.Topology(gf, K=100)
    paths = topo.k_shortest_paths(5)
    assert paths == orc.k_shortest_simple_paths(ot, 5, exhaustive=False)
    sigma = ds.sigma_for_mean_utilisation(gf)
    tms = np.stack([np.floor(ds.uniform_tm(gf.num_nodes, sigma, 3000 + j) * (scale if j % 2 else 1.0)) for j in range(n_tms)])
    eb = EpisodeBatch(ctx, topo, n_tms)
    mx, acts = eb.sap(tms, want_actions=True)
    fallbacks = 0
    for i in range(n_tms):
        omx, oacts = orc.play_sap(ot, paths, tms[i])
        assert acts[i].tolist() == oacts and mx[i] == omx
        fallbacks += sum(1 for a in oacts if a < 0)
    assert fallbacks > 0                       # the random branch was exercised
    eb.close()


def test_env20_shim_play_sap_games(ctx, tmp_path):
    """the reference's call sequence (script :543-557) through the GraphEnv-v20 shim"""
    from enero_b200 import datasets as ds
    from enero_b200.agents import play_sap_games
    case = load_case("tiny12")
    gold = json.load(open(os.path.join(GOLDEN, "sap_tiny12.json")))
    folder = str(tmp_path)
    open(os.path.join(folder, "tiny12.graph"), "w").write(case["graph_text"])
    os.makedirs(os.path.join(folder, "TM"))
    for i, tm in enumerate(gold["tms"]):
        ds.write_demands_file(os.path.join(folder, "TM", "tiny12.%d.demands" % i), np.array(tm))
    from enero_b200 import gym_graph
    env = gym_graph.make('GraphEnv-v20', context=ctx)
    env.seed(9)
    env.generate_environment(folder, "tiny12", 100, 100)
    assert env.K == 5 and env.allPaths == gold["all_paths"]
    for i in (0, 2, 5):
        final, secs = play_sap_games(i, folder, "tiny12", env=env)
        assert final == gold["episodes"][i]["max_uti"]
