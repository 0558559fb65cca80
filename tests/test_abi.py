"""The C-ABI contract on CPU: libenero_b200.so loads, exports every function include/enero_b200.h
declares, the ctypes table binds exactly that set, the product path never imports the oracle, and a
compute call without a CUDA device fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "enero_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(enero_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from enero_b200 import _lib
    lib = ctypes.CDLL(_lib.library_path()) if os.path.isfile(_lib.library_path()) else _lib.load()
    names = _declared()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from enero_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_header_cites_the_reference():
    src = open(HEADER).read()
    for cite in ("actorPPOmiddR.py:42-69", "criticPPO.py:38-64", "environment16.py:582-640", "environment15.py:485-544",
                 "train_Enero_3top_script.py:264", "script_eval_on_single_topology.py:83-129"):
        assert cite in src, cite


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "enero_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt and f.endswith(".py") and "import" in txt.split("oracle/")[0][-40:]:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_no_cpu_fallback_without_a_gpu():
    """in a process that sees no CUDA device every compute entry point must raise"""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from enero_b200 import _lib\n"
        "from enero_b200.runtime import Context\n"
        "try:\n"
        "    Context(0)\n"
        "except _lib.EneroError as e:\n"
        "    print('RAISED', e)\n"
        "else:\n"
        "    print('NO ERROR')\n" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert "RAISED" in out.stdout and "no CPU fallback" in out.stdout, out.stdout + out.stderr


def test_topology_precompute_needs_no_gpu(tmp_path):
    """host-side part of the ABI (enero_topo_build and getters) runs without a device"""
    from conftest import case_graph_file, load_case
    from enero_b200.topology import Topology
    t = Topology(case_graph_file(load_case("tri9"), tmp_path), K=100)
    assert t.E == 26 and t.P > 0 and t.Kmax >= 1
