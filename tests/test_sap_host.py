"""SAP baseline, host side (no GPU): the K-shortest-simple-path table of csrc/ksp.cpp and the oracle's restatement
against the table the reference's own Env20.compute_SPs built (tests/golden/sap_*.json, oracle/gen_sap_golden.py),
and the oracle's play_sap against the reference-recorded episodes (random-fallback branch included)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, case_graph_file, load_case
from oracle import enero_oracle as orc


def _sap(name):
    return json.load(open(os.path.join(GOLDEN, "sap_%s.json" % name)))


@pytest.mark.parametrize("name", ["tri9", "tiny12"])
def test_ksp_table_matches_reference(name, tmp_path):
    from enero_b200.topology import Topology
    gold = _sap(name)
    gf = case_graph_file(load_case(name), tmp_path)
    assert Topology(gf, K=100).k_shortest_paths(5) == gold["all_paths"]
    ot = orc.Topology(gf, K=100)
    assert orc.k_shortest_simple_paths(ot, 5, exhaustive=True) == gold["all_paths"]
    assert orc.k_shortest_simple_paths(ot, 5, exhaustive=False) == gold["all_paths"]


def test_ksp_table_medium_graph(tmp_path):
    """60 links (eli20): csrc/ksp.cpp against the oracle's Yen-based route (the exhaustive one is infeasible here)"""
    from enero_b200.topology import Topology
    gf = case_graph_file(load_case("eli20"), tmp_path)
    assert Topology(gf, K=100).k_shortest_paths(5) == orc.k_shortest_simple_paths(orc.Topology(gf, K=100), 5, exhaustive=False)


@pytest.mark.parametrize("name", ["tri9", "tiny12"])
def test_oracle_sap_matches_reference_episodes(name, tmp_path):
    gold = _sap(name)
    ot = orc.Topology(case_graph_file(load_case(name), tmp_path), K=100)
    for tm, ep in zip(gold["tms"], gold["episodes"]):
        mx, acts = orc.play_sap(ot, gold["all_paths"], np.array(tm))
        assert acts == ep["actions"] and mx == ep["max_uti"]
