"""Where a k_mp_tc tile spends its time: build with  ENERO_NVCC_EXTRA=-DENERO_TC_TIMING python -m enero_b200.build --force  and
run this on the GPU box; thread 0 of every group accumulates clock64() per segment of the tile loop (operand store, MMA waits,
gather, gates, read-back), printed as shares.  Without the macro only the throughput lines are printed."""
import sys, os, tempfile, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import *
from oracle import enero_oracle as orc
from enero_b200.runtime import Context
from enero_b200.topology import Topology
from enero_b200.engine import EpisodeBatch
from enero_b200 import datasets as ds, _lib
w = golden_weights()
c = Context(0)
c.upload_weights(0, [w[k] for k in orc.WEIGHT_NAMES])
gf = case_graph_file(load_case("zoo30"), tempfile.mkdtemp())
topo = Topology(gf, K=100)
sigma = ds.sigma_for_mean_utilisation(gf)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 2000 + j) for j in range(B)])
eb = EpisodeBatch(c, topo, B)
eb.reset(tms)
eb.snapshot()
lib = _lib.load()
for eng in ("ffma", "tc", "tc"):
    c.engine = eng
    eb.restore()
    for _ in range(3):
        eb.policy(want_probs=False); eb.step(None, want_outputs=False)
    c.sync()
    if hasattr(lib, "enero_debug_tc_counters"):
        out = (ctypes.c_ulonglong * 8)()
        lib.enero_debug_tc_counters(c.handle, out)
    c.profile_begin()
    t0 = time.perf_counter()
    for _ in range(10):
        eb.policy(want_probs=False); eb.step(None, want_outputs=False)
    ms, n = c.profile_end()
    dt = time.perf_counter() - t0
    print(eng, "decisions/s %.0f" % (B * 10 / dt), "mp launch avg %.3f ms over %d" % (ms / n, n), flush=True)
    if hasattr(lib, "enero_debug_tc_counters") and eng == "tc":
        lib.enero_debug_tc_counters(c.handle, out)
        v = np.array(list(out), dtype=np.float64)
        names = ["load h+store operand+sync", "MMA1 wait", "ldB+gather+store agg+sync", "MMA2 wait", "gates+store h'+operand+sync", "MMA3 wait", "ld A'+store", "-"]
        tot = v.sum()
        for nme, x in zip(names, v):
            print("   %-32s %6.1f %%" % (nme, 100 * x / tot))
        print("   occ_tc", int(v[7]), "cycles per CTA per launch: %.0f" % ((tot - v[7]) / n / (148 * max(1, int(v[7])))))
