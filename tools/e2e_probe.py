import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from enero_b200.engine import EpisodeBatch
from enero_b200.runtime import Context
from enero_b200.topology import Topology
wd, wl = bench.load_weights()
gf, sigma = bench.make_topology()
topo = Topology(gf, K=100)
ctx = Context(0); ctx.upload_weights(0, wl)
for B in (256, 1024, 2048):
    tms = bench.make_tms(gf, sigma, B, 2000)
    eb = EpisodeBatch(ctx, topo, B)
    for rep in range(3):
        t0 = time.perf_counter(); eb.reset(tms); t1 = time.perf_counter(); eb.run_greedy(); t2 = time.perf_counter(); r = eb.best(); ctx.sync(); t3 = time.perf_counter()
    dec = int(eb.state()["steps"].sum())
    print(B, "reset %.1f ms  greedy %.1f ms  best %.1f ms  -> %.0f dec/s" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, dec/(t3-t0)))
    t0 = time.perf_counter(); ls = eb.local_search(None, mode=0, want_moves=False); t1 = time.perf_counter()
    print("   LS %.1f ms for %d episodes, moves %d, %.0f TM/s" % ((t1-t0)*1e3, B, int(ls["n_moves"].sum()), B/(t1-t0)))
    eb.close()
