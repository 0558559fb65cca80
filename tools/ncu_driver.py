"""Short driver for ncu captures of the message-passing kernel: three decision steps of B episodes of cfg2.
    ncu --set full --clock-control none --import-source on -k regex:k_mp_tc -s 6 -c 2 -o gpurun_out/rep python tools/ncu_driver.py 1024 tc"""
import sys, os, tempfile
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
