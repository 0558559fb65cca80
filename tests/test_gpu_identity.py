"""Greedy-action / best-utilisation / Local-Search identity at every BASELINE.json configuration size, against the
oracle's answers recorded in tests/golden/identity_configs.json (oracle/gen_identity.py):

    cfg3   (24 nodes, 74 links)                  8 TMs   greedy actions, best, OSPF, Local-Search final
    cfg4   (20,31) (23,25) (17,31) topologies    8 TMs each, same checks
    180 links (60 nodes)                         one full greedy episode of 531 decisions

north_star: "the greedy action sequence and final max link utilisation must be identical on the shipped checkpoint".
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

FIXTURE = json.load(open(os.path.join(GOLDEN, "identity_configs.json")))


@pytest.fixture(scope="module")
def ctx(weights):
    from enero_b200.runtime import Context
    from oracle import enero_oracle as orc
    c = Context(0)
    c.upload_weights(0, [weights[k] for k in orc.WEIGHT_NAMES])
    yield c
    c.close()


@pytest.mark.parametrize("name", sorted(FIXTURE))
@pytest.mark.parametrize("mode", ["fast", "accurate"])
def test_identity_at_config_size(ctx, name, mode, tmp_path):
    from enero_b200 import datasets as ds
    from enero_b200.engine import EpisodeBatch
    from enero_b200.topology import Topology
    fx = FIXTURE[name]
    n, l, seed = fx["spec"]
    links, caps = ds.zoo_like(n, l, seed)
    p = os.path.join(str(tmp_path), "t.graph")
    ds.write_graph_file(p, n, links, caps)
    gf = ds.parse_graph_file(p)
    assert ds.sigma_for_mean_utilisation(gf) == fx["sigma"]
    topo = Topology(gf, K=100)
    eps = fx["episodes"]
    tms = np.stack([ds.uniform_tm(n, fx["sigma"], e["tm_seed"]) for e in eps])
    ctx.math = mode
    try:
        eb = EpisodeBatch(ctx, topo, len(eps))
        eb.reset(tms)
        eb.run_greedy()
        best, ospf, routing = eb.best()
        act, _, _ = eb.history()
        steps = eb.state()["steps"]
        for i, e in enumerate(eps):
            assert act[i, :steps[i]].tolist() == e["actions"], "%s TM %d: greedy action sequences differ" % (name, i)
            assert (best[i], ospf[i]) == (e["best"], e["ospf"])
            assert [list(r) for r in routing[i]] == e["routing"]
        if eps[0]["ls_final"] is not None:
            ls = eb.local_search(None, mode=0)
            assert ls["final"].tolist() == [e["ls_final"] for e in eps]
            assert ls["n_moves"].tolist() == [e["ls_moves"] for e in eps]
        eb.close()
    finally:
        ctx.math = "fast"
