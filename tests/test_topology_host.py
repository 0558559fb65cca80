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


def test_native_demands_parser_matches_python(tmp_path):
    """enero_parse_demands_batch == the line-by-line Python parser == defo_process_results.py:216-225"""
    import os
    import time
    from enero_b200 import datasets as ds
    rng = np.random.default_rng(0)
    n = 23
    paths = []
    for j in range(40):
        tm = rng.uniform(0, 5000, (n, n))
        np.fill_diagonal(tm, 0.0)
        p = os.path.join(str(tmp_path), "t.%d.demands" % j)
        with open(p, "w") as fp:
            fp.write("DEMANDS %d\nlabel src dest bw\n" % (n * (n - 1)))
            k = 0
            for s in range(n):
                for d in range(n):
                    if s != d:
                        # fractional, exponent and integer spellings: float() / strtod must agree
                        v = tm[s, d]
                        txt = ("%.3f" % v) if k % 3 == 0 else (("%.6e" % v) if k % 3 == 1 else str(int(v)))
                        fp.write("demand_%d %d %d %s\n" % (k, s, d, txt))
                        k += 1
        paths.append(p)
    t0 = time.perf_counter()
    want = np.stack([ds.parse_demands_file(p, n) for p in paths])
    t1 = time.perf_counter()
    got = ds.parse_demands_files(paths, n)
    t2 = time.perf_counter()
    assert np.array_equal(got, want)
    print('python %.1f ms, native %.1f ms' % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))      # informative, not asserted
    from enero_b200 import _lib
    with pytest.raises(_lib.EneroError):
        ds.parse_demands_files(paths + [os.path.join(str(tmp_path), "missing.demands")], n)
    with pytest.raises(_lib.EneroError):
        ds.parse_demands_files(paths[:1], 5)           # node ids out of range for a 5-node topology


def test_link_failure_variants():
    """generate_link_failure_topologies.py:123-177: k links removed, connected, distinct, order kept"""
    from enero_b200 import datasets as ds
    links, caps = ds.zoo_like(20, 31, 1001)
    vs = ds.link_failure_variants(links, caps, 20, 2, 6, seed=3)
    assert len(vs) == 6 and len({tuple(v[2]) for v in vs}) == 6
    for vl, vc, removed in vs:
        assert len(vl) == len(links) - 2 and len(vc) == len(vl) and len(removed) == 2
        assert vl == [l for l in links if l not in removed]                      # order preserved
        t = Topology(_gf_from(vl, vc, 20), K=100)                                # strongly connected or build fails
        assert t.E == 2 * len(vl)
    with pytest.raises(ValueError):
        ds.link_failure_variants(links, caps, 20, len(links), 1, seed=0)


def _gf_from(links, caps, n):
    import os
    import tempfile
    from enero_b200 import datasets as ds
    d = tempfile.mkdtemp()
    p = os.path.join(d, "v.graph")
    ds.write_graph_file(p, n, links, caps)
    return ds.parse_graph_file(p)
