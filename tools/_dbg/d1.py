import sys, os, tempfile
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from conftest import *
from oracle import enero_oracle as orc
from enero_b200.runtime import Context
from enero_b200.topology import Topology
w = golden_weights()
c = Context(0)
c.upload_weights(0, [w[k] for k in orc.WEIGHT_NAMES])
print("ctx ok", flush=True)
case = load_case("tri9"); gf = case_graph_file(case, tempfile.mkdtemp())
topo, ot = Topology(gf, K=100), orc.Topology(gf, K=100)
env = orc.Env16(ot)
ep = case["episodes"][0]
_, s, d = env.reset(np.array(case["tms"][str(ep["tm_id"])]))
feats = env.candidate_features(s, d)
b = orc.batch_candidates(feats, ot.first, ot.second)
print("calling gnn_forward", flush=True)
for mode in ("fast", "accurate"):
    c.math = mode
    got = c.gnn_forward(0, b["link_state"], b["graph_id"], b["first"], b["second"], len(feats))
    print(mode, got[:4], flush=True)
want = orc.actor_forward(w, b["link_state"], b["graph_id"], b["first"], b["second"], b["num_edges"]).reshape(-1)
print(want[:4], flush=True)
lg = c.decide_batch(topo, env.loads[None], np.array([[s, d]], np.int32), np.array([env.tm[s, d]]))
print(lg[0][0, :4], flush=True)
