import sys, os, tempfile, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
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
gf = case_graph_file(load_case("zoo30"), tempfile.mkdtemp())
topo = Topology(gf, K=100)
sigma = ds.sigma_for_mean_utilisation(gf)
B = 4096
tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 2000 + j) for j in range(B)])
eb = EpisodeBatch(c, topo, B)
eb.reset(tms)
for _ in range(3):
    eb.policy(want_probs=False); eb.step(None, want_outputs=False)
c.sync()
import torch
t0 = time.perf_counter()
for _ in range(20):
    eb.policy(want_probs=False); eb.step(None, want_outputs=False)
c.sync()
print(os.environ.get("ENERO_UNFUSED_FEATURES", "0"), "decisions/s %.0f  step ms %.3f" % (B * 20 / (time.perf_counter() - t0), (time.perf_counter() - t0) / 20 * 1e3))
