"""Tensor-core engine (k_mp_tc) against the FP32-pipe engine (k_mp_iter) on the GPU box: logits of golden decisions (vs each
other and vs the reference-recorded oracle logits), throughput and per-launch time at 4 096 episodes of cfg2, and greedy
episode identity over the whole batch.    python tools/engine_compare.py"""
import sys, os, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import *
from oracle import enero_oracle as orc
from enero_b200.runtime import Context
from enero_b200.topology import Topology
from enero_b200.engine import EpisodeBatch
from enero_b200 import datasets as ds
w = golden_weights()
c = Context(0)
c.upload_weights(0, [w[k] for k in orc.WEIGHT_NAMES])
for name in ("tri9", "zoo30"):
    case = load_case(name); gf = case_graph_file(case, tempfile.mkdtemp())
    topo = Topology(gf, K=100)
    ep = case["episodes"][0]
    tm = np.array(case["tms"][str(ep["tm_id"])])
    decs = ep["decisions"][:16]
    loads = np.array([dc["loads_before"] for dc in decs])
    sd = np.array([[dc["s"], dc["d"]] for dc in decs], dtype=np.int32)
    bw = np.array([tm[dc["s"], dc["d"]] for dc in decs])
    out = {}
    for eng in ("ffma", "tc"):
        c.engine = eng
        out[eng] = c.decide_batch(topo, loads, sd, bw)
        print(name, eng, out[eng][0][0, :4], flush=True)
    lg_f, lg_t = out["ffma"][0], out["tc"][0]
    want = np.zeros_like(lg_f)
    for b, dc in enumerate(decs):
        want[b, :len(dc["logits"])] = dc["logits"]
    print(name, "max |tc - ffma|", np.abs(lg_t - lg_f).max(), "max |tc - oracle|", np.abs(lg_t - want).max(), "max |ffma - oracle|", np.abs(lg_f - want).max(),
          "actions equal", np.array_equal(out["tc"][3], out["ffma"][3]), flush=True)
gf = case_graph_file(load_case("zoo30"), tempfile.mkdtemp())
topo = Topology(gf, K=100)
sigma = ds.sigma_for_mean_utilisation(gf)
B = 4096
tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 2000 + j) for j in range(B)])
eb = EpisodeBatch(c, topo, B)
eb.reset(tms)
eb.snapshot()
for eng in ("ffma", "tc", "ffma", "tc"):
    c.engine = eng
    eb.restore()
    for _ in range(3):
        eb.policy(want_probs=False); eb.step(None, want_outputs=False)
    c.sync()
    c.profile_begin()
    t0 = time.perf_counter()
    for _ in range(20):
        eb.policy(want_probs=False); eb.step(None, want_outputs=False)
    ms, n = c.profile_end()
    dt = time.perf_counter() - t0
    print(eng, "decisions/s %.0f" % (B * 20 / dt), "mp launch avg %.3f ms over %d" % (ms / n, n), flush=True)
eb.restore(); c.engine = "tc"; eb.run_greedy(); a_tc = eb.history()[0].copy(); b_tc = eb.best()[0].copy()
eb.restore(); c.engine = "ffma"; eb.run_greedy(); a_ff = eb.history()[0].copy(); b_ff = eb.best()[0].copy()
print("greedy episodes: action sequences differ in", int((a_tc != a_ff).any(axis=1).sum()), "of", B, "; best differs in", int((b_tc != b_ff).sum()))
