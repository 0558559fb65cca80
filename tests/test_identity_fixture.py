"""tests/golden/identity_configs.json is what the GPU identity tests trust at the BASELINE configuration sizes; this
CPU test re-runs the oracle on one episode per entry (a prefix of the 531-decision episode at 180 links) so the
fixture cannot drift from oracle/enero_oracle.py."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import enero_oracle as orc
from oracle import gen_identity as gi

FIXTURE = json.load(open(os.path.join(GOLDEN, "identity_configs.json")))


@pytest.mark.parametrize("name", sorted(FIXTURE))
def test_fixture_matches_oracle(name, weights):
    from enero_b200 import datasets as ds
    fx = FIXTURE[name]
    assert [tuple(s[1]) for s in gi.SPECS if s[0] == name] == [tuple(fx["spec"])]
    gf, sigma = gi.build(tuple(fx["spec"]))
    assert sigma == fx["sigma"]
    ot = orc.Topology(gf, K=100, paths=ds.bfs_lexmin_paths(gf))
    e = fx["episodes"][0]
    tm = ds.uniform_tm(gf.num_nodes, sigma, e["tm_seed"])
    if len(e["actions"]) > 100:                       # 180 links: the first 12 decisions only (seconds, not a minute)
        env = orc.Env16(ot)
        _, s, d = env.reset(tm)
        assert env.edge_max[2] == e["ospf"]
        for want in e["actions"][:12]:
            probs, _, _ = orc.pred_action_distrib(weights, env, s, d)
            assert int(np.argmax(probs)) == want
            _, _, _, _, s, d, _, _, _ = env.step(want)
        return
    best, ospf, routing, actions = orc.play_middrout_greedy(weights, orc.Env16(ot), tm)
    assert actions == e["actions"] and (best, ospf) == (e["best"], e["ospf"])
    assert [[s, d, m] for (s, d), m in routing.items()] == e["routing"]
