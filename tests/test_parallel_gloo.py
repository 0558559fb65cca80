"""N > 1 host logic on CPU: two processes, gloo backend (episodes shard with no
collective; the PPO minibatch shards with one gradient all-reduce)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

from enero_b200 import parallel as par

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 64, 835):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                b, e = par.shard_range(n, r, world)
                assert 0 <= b <= e <= n and (e - b) in (n // world, n // world + 1)
                cover += list(range(b, e))
            assert cover == list(range(n))


def test_shard_indices_partitions():
    inds = np.random.RandomState(1).permutation(55)
    for world in (1, 2, 4, 8):
        got = np.concatenate([par.shard_indices(inds, r, world) for r in range(world)])
        assert sorted(got.tolist()) == list(range(55))
    # the last minibatch of an epoch has 10 samples (835 = 15 * 55 + 10): with 8 ranks some get 1, some 2
    sizes = [len(par.shard_indices(inds[:10], r, 8)) for r in range(8)]
    assert sum(sizes) == 10 and max(sizes) - min(sizes) <= 1


def test_shard_by_cost_balances():
    costs = np.random.RandomState(2).randint(50, 3353, size=500).astype(float)
    parts = par.shard_by_cost(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(500))
    loads = np.array([costs[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01


def test_world2_gloo(tmp_path):
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "helpers", "gloo_worker.py"), str(tmp_path)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    res = [json.load(open(os.path.join(str(tmp_path), "rank%d.json" % r))) for r in range(2)]
    # rank 0 holds all 11 records in unit order; rank 1 holds none
    assert [r["tm_id"] for r in res[0]["gathered"]] == list(range(11))
    assert res[1]["gathered"] is None
    # the two shards partition the minibatch
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(55))
    assert abs(len(res[0]["mine"]) - len(res[1]["mine"])) <= 1
    # all-reduced gradient is identical on both ranks and equals the single-process sum
    assert res[0]["grad_head"] == res[1]["grad_head"]
    want = np.zeros(8412, dtype=np.float64)
    for i in range(55):
        want += (1.0 / 55) * np.sin(np.arange(8412) * 0.001 * (i + 1))
    np.testing.assert_allclose(res[0]["grad_sum"], want.sum(), rtol=1e-4)
    np.testing.assert_allclose(res[0]["grad_head"], want[:4], rtol=1e-4, atol=1e-6)
    assert res[0]["max_ms"] == res[1]["max_ms"] == 11.0
