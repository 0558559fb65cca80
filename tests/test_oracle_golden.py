"""The oracle (oracle/enero_oracle.py) against outputs recorded from the
reference's own environment / Local-Search code (tests/golden, made by
oracle/gen_golden.py).  Integer / fp64 quantities: bit-exact."""
import numpy as np
import pytest

from conftest import CASE_NAMES, case_graph_file, load_case
from oracle import enero_oracle as orc


@pytest.fixture(scope="module", params=CASE_NAMES)
def case_topo(request, tmp_path_factory):
    case = load_case(request.param)
    gf = case_graph_file(case, tmp_path_factory.mktemp(request.param))
    # cases generated without a pre-seeded JSON pin the BFS closed form against
    # the reference's all_simple_paths enumeration (environment16.py:491-495)
    topo = orc.Topology(gf, K=100)
    return case, topo


def test_topology_tables(case_topo):
    case, t = case_topo
    g = case["topo"]
    assert (t.N, t.E, t.K) == (g["N"], g["E"], g["K"])
    assert [list(e) for e in t.edges] == g["edges"]
    assert t.cap.tolist() == g["cap"]
    assert t.link_capacity_feature.tolist() == g["link_capacity_feature"]
    assert t.first.tolist() == g["first"]
    assert t.second.tolist() == g["second"]
    for key, p in g["paths"].items():
        s, d = map(int, key.split(":"))
        assert t.paths[(s, d)] == p
    for key, m in g["mids"].items():
        s, d = map(int, key.split(":"))
        assert t.mids[(s, d)] == m, key


def test_env16_episode_and_ls(case_topo, weights):
    case, t = case_topo
    for ep in case["episodes"]:
        tm = np.array(case["tms"][str(ep["tm_id"])])
        env = orc.Env16(t)
        demand, s, d = env.reset(tm)
        r = ep["reset"]
        assert (float(demand), s, d) == (r["demand"], r["s"], r["d"])
        assert [[a, b, float(c)] for a, b, c in env.elig] == r["elig"]
        assert env.loads.tolist() == r["loads"]
        assert [env.edge_max[0], env.edge_max[1], float(env.edge_max[2])] == r["edge_max"]
        for st, dec in zip(ep["steps"], ep["decisions"]):
            assert (s, d) == (dec["s"], dec["d"])
            assert env.loads.tolist() == dec["loads_before"]
            if dec["marks"] is not None:
                feats = env.candidate_features(s, d)
                for f, mk in zip(feats, dec["marks"]):
                    assert f[:, 2].tolist() == np.asarray(mk).astype(np.float32).tolist()
            reward, done, _, nb, s, d, mx, _, std = env.step(st["action"])
            assert float(reward) == st["reward"] and done == st["done"]
            assert [float(nb), s, d] == st["next"]
            assert [mx[0], mx[1], float(mx[2])] == st["edge_max"]
            assert float(std) == st["std"]
            assert env.loads.tolist() == st["loads_after"]
        # full greedy driver incl. GNN: actions, best value and routing snapshot
        rec = []
        best, ospf, routing, actions = orc.play_middrout_greedy(weights, orc.Env16(t), tm, rec)
        assert actions == [st["action"] for st in ep["steps"]]
        assert (float(best), float(ospf)) == (ep["best"], ep["ospf"])
        assert [[k[0], k[1], m] for k, m in routing.items()] == ep["best_routing"]
        np.testing.assert_allclose(rec[0]["logits"], ep["decisions"][0]["logits"], rtol=1e-6, atol=1e-7)
        # Local Search
        e15 = orc.Env15(t)
        cur = e15.reset_DRL_hill_sp(tm, routing)
        ls = ep["ls"]
        assert float(cur) == ls["reset_val"]
        assert [[a, b, float(c)] for a, b, c in e15.elig] == ls["elig"]
        assert e15.loads.tolist() == ls["loads0"]
        final, moves = orc.local_search(e15, cur)
        assert [[a, s_, d_, float(v)] for a, s_, d_, v in moves] == ls["moves"]
        assert float(final) == ls["final"]


@pytest.mark.parametrize("name", ["tri9", "tiny12"])
def test_plain_hill_climbing(name, tmp_path):
    case = load_case(name)
    t = orc.Topology(case_graph_file(case, tmp_path), K=100)
    for ep in case["episodes"]:
        tm = np.array(case["tms"][str(ep["tm_id"])])
        e15 = orc.Env15(t)
        cur = e15.reset_hill_sp(tm)
        final, _ = orc.local_search(e15, cur)
        assert float(final) == ep["plain_hill_final"]
