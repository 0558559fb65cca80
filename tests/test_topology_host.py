"""Host-side (C++) topology precompute against the reference-recorded goldens: bit-exact."""
import numpy as np
import pytest

from conftest import CASE_NAMES, case_graph_file, load_case
from enero_b200.topology import Topology


@pytest.mark.parametrize("name", CASE_NAMES)
def test_tables_match_reference(name, tmp_path):
    case = load_case(name)
    g = case["topo"]
    t = Topology(case_graph_file(case, tmp_path), K=100)
    assert (t.N, t.E, t.K) == (g["N"], g["E"], g["K"])
    assert [list(e) for e in t.edges] == g["edges"]
    assert t.first.tolist() == g["first"]
    assert t.second.tolist() == g["second"]
    assert t.P == len(g["first"])
    assert t.off_first == max(g["first"]) + 1 and t.off_second == max(g["second"]) + 1
    assert t.link_capacity_feature.tolist() == g["link_capacity_feature"]
    assert t.paths_dict() == g["paths"]
    assert t.middlepoint_dict() == g["mids"]
    assert t.Kmax == max(len(v) for v in g["mids"].values())


def test_paths_override_equals_bfs(tmp_path):
    case = load_case("tiny12")
    gf = case_graph_file(case, tmp_path)
    a = Topology(gf, K=100)
    b = Topology(gf, K=100, paths=case["topo"]["paths"])
    assert a.middlepoint_dict() == b.middlepoint_dict()
    assert np.array_equal(a.sp_edge, b.sp_edge)


def test_bad_input_fails_loudly(tmp_path):
    from enero_b200 import _lib
    case = load_case("tri9")
    gf = case_graph_file(case, tmp_path)
    gf.links_bw[gf.edge_list()[0]] = 0.0
    with pytest.raises(_lib.EneroError):
        Topology(gf)
