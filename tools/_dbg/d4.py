import sys, os, tempfile
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
c.engine = sys.argv[2] if len(sys.argv) > 2 else "tc"
gf = case_graph_file(load_case("zoo30"), tempfile.mkdtemp())
topo = Topology(gf, K=100)
sigma = ds.sigma_for_mean_utilisation(gf)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 2000 + j) for j in range(B)])
eb = EpisodeBatch(c, topo, B)
eb.reset(tms)
for _ in range(3):
    eb.policy(want_probs=False); eb.step(None, want_outputs=False)
c.sync()
print("ok")
