"""The benchmark configurations of BASELINE.json / SURVEY.md section 8(d) as runnable workloads on
synthetic Zoo-shaped data (no dataset ships with the reference, README.md:42):

    python -m enero_b200.workloads cfg3 [--tms 10000]     full Enero (DRL + LS) over many TMs, one topology
    python -m enero_b200.workloads cfg4 [--episodes 1]    PPO training step, 3 topologies
    python -m enero_b200.workloads cfg5 [--topologies 8 --tms 64]   link-failure-style sweep

Under torchrun every rank takes a contiguous / cost-balanced shard of the units and rank 0 prints one
JSON line; there is no data-path collective for cfg3 / cfg5 and one gradient all-reduce per Adam step
for cfg4.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

from . import datasets as ds
from . import parallel as par
from .evaluate import evaluate_tms
from .runtime import Context
from .topology import Topology

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WEIGHT_NAMES = ("msg_kernel", "msg_bias", "gru_kernel", "gru_recurrent_kernel", "gru_bias",
                "r1_kernel", "r1_bias", "r2_kernel", "r2_bias", "r3_kernel", "r3_bias")


def shipped_actor_weights():
    """ckpt_ACT-1835 of modelsSP_3top_15_B_NEW (a byte copy of the shipped checkpoint lives in tests/golden)"""
    from . import checkpoint as ckpt
    return ckpt.load_actor_weights(os.path.join(GOLDEN, "ckpt_ACT-1835"))


def synthetic_topology(n, l, seed):
    links, caps = ds.zoo_like(n, l, seed)
    tmp = tempfile.mkdtemp(prefix="enero_wl_")
    p = os.path.join(tmp, "t.graph")
    ds.write_graph_file(p, n, links, caps)
    gf = ds.parse_graph_file(p)
    return gf, ds.sigma_for_mean_utilisation(gf)


def run_cfg3(ctx, n_tms=10000, rank=0, world=1, batch=2048, topo_spec=(24, 37, 1002)):
    """Full Enero inference over n_tms synthetic TMs on a HurricaneElectric-sized graph (N=24, 37 links)."""
    gf, sigma = synthetic_topology(*topo_spec)
    topo = Topology(gf, K=100)
    b, e = par.shard_range(n_tms, rank, world)
    tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 2000 + j) for j in range(b, e)])
    t0 = time.perf_counter()
    res = evaluate_tms(ctx, topo, tms, batch=batch)
    ctx.sync()
    dt = time.perf_counter() - t0
    return {"tms": int(e - b), "seconds": dt, "decisions": int(res["decisions"].sum()), "ls_moves": int(res["ls_moves"].sum()),
            "mean_ospf": float(res["ospf"].mean()), "mean_drl": float(res["drl"].mean()), "mean_enero": float(res["enero"].mean()),
            "links": topo.E, "nodes": topo.N}


def run_cfg5(ctx, n_topologies=8, n_tms=64, rank=0, world=1, first_seed=1000, variants_per_base=4):
    """Link-failure sweep on synthetic data: base Zoo-shaped topologies of 50-250 undirected links
    (100-500 directed), each with `variants_per_base` link-failure variants (1-3 links removed, connected,
    distinct: generate_link_failure_topologies.py:123-177) that share the base topology's traffic matrices,
    n_tms TMs per topology, sharded over the ranks by estimated cost with no collective."""
    rng = np.random.default_rng(5)
    specs = []                                # (n, links, caps, sigma-source seed, label)
    n_bases = max(1, (n_topologies + variants_per_base - 1) // variants_per_base)
    for i in range(n_bases):
        l = int(rng.integers(50, 251))
        n = max(4, int(round(l / 1.65)))
        links, caps = ds.zoo_like(n, l, first_seed + i)
        fails = 1 + i % 3
        for j, (vl, vc, _) in enumerate(ds.link_failure_variants(links, caps, n, fails, variants_per_base, 7000 + i)):
            if len(specs) < n_topologies:
                specs.append((n, vl, vc, first_seed + i, "syn%d_%d_%d" % (i, fails, j)))
    costs = [ds.max_decisions(n) * min(99, n - 1) * (4 * len(vl)) * 4.0 for n, vl, _, _, _ in specs]      # ~ D * K' * (P + E)
    mine = par.shard_by_cost(costs, world)[rank]
    out = {"topologies": len(mine), "tms": 0, "decisions": 0, "ls_moves": 0, "seconds": 0.0, "build_seconds": 0.0, "links": [],
           "drl_seconds": 0.0, "ls_seconds": 0.0}
    tmp = tempfile.mkdtemp(prefix="enero_cfg5_")
    sigma_cache = {}
    for i in mine:
        n, vl, vc, base_seed, label = specs[i]
        t0 = time.perf_counter()
        p = os.path.join(tmp, label + ".graph")
        ds.write_graph_file(p, n, vl, vc)
        gf = ds.parse_graph_file(p)
        topo = Topology(gf, K=100)
        t1 = time.perf_counter()
        if base_seed not in sigma_cache:      # the variants of one base share its traffic matrices
            sigma_cache[base_seed] = ds.sigma_for_mean_utilisation(gf)
        tms = np.stack([ds.uniform_tm(n, sigma_cache[base_seed], 2000 + j) for j in range(n_tms)])
        res = evaluate_tms(ctx, topo, tms, batch=n_tms)
        ctx.sync()
        out["build_seconds"] += t1 - t0
        out["seconds"] += time.perf_counter() - t1
        out["tms"] += n_tms
        out["decisions"] += int(res["decisions"].sum())
        out["ls_moves"] += int(res["ls_moves"].sum())
        out["drl_seconds"] += float(res["cost_drl"].sum())
        out["ls_seconds"] += float(res["cost_ls"].sum())
        out["links"].append(topo.E)
        topo.close()
    return out


def run_cfg4(ctx, ppo_episodes=1, process_group=None, specs=((20, 31, 1003), (23, 25, 1004), (17, 31, 1005))):
    """One PPO episode as in train_Enero_3top_script.py:471-631 on three synthetic topologies: 835 sampled
    decisions (285 / 304 / 246 per topology), GAE, 8 epochs x 16 minibatches of clip + Adam."""
    from .train import TrainingRun
    run = TrainingRun.synthetic(ctx, specs, process_group=process_group)
    try:
        run.ppo_episode()                   # warm-up: first-touch allocations of the training buffers
        return run.run(1, ppo_episodes)
    finally:
        run.close()


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--tms", type=int, default=None)
    ap.add_argument("--topologies", type=int, default=8)
    ap.add_argument("--episodes", type=int, default=1)
    ap.add_argument("--batch", type=int, default=2048)
    args = ap.parse_args()
    rank, world, local = par.init_process_group()
    ctx = Context(local)
    w = shipped_actor_weights()
    ctx.upload_weights(0, w)
    if args.config == "cfg3":
        r = run_cfg3(ctx, args.tms or 10000, rank, world, args.batch)
        secs = par.max_over_ranks(r["seconds"], device="cuda:%d" % local if world > 1 else None)
        recs = par.gather_records([r])
        if rank == 0:
            tms = sum(x["tms"] for x in recs)
            print(json.dumps({"workload": "cfg3", "n_gpus": world, "tms": tms, "seconds": secs, "tms_per_s": tms / secs,
                              "decisions_per_s": sum(x["decisions"] for x in recs) / secs,
                              "ls_moves": sum(x["ls_moves"] for x in recs), "mean_ospf": recs[0]["mean_ospf"],
                              "mean_drl": recs[0]["mean_drl"], "mean_enero": recs[0]["mean_enero"],
                              "links": recs[0]["links"], "nodes": recs[0]["nodes"]}))
    elif args.config == "cfg5":
        r = run_cfg5(ctx, args.topologies, args.tms or 64, rank, world)
        secs = par.max_over_ranks(r["seconds"] + r["build_seconds"], device="cuda:%d" % local if world > 1 else None)
        recs = par.gather_records([r])
        if rank == 0:
            print(json.dumps({"workload": "cfg5", "n_gpus": world, "topologies": sum(x["topologies"] for x in recs),
                              "tms": sum(x["tms"] for x in recs), "seconds": secs,
                              "decisions_per_s": sum(x["decisions"] for x in recs) / secs,
                              "build_seconds": sum(x["build_seconds"] for x in recs),
                              "drl_seconds": sum(x["drl_seconds"] for x in recs), "ls_seconds": sum(x["ls_seconds"] for x in recs),
                              "ls_moves": sum(x["ls_moves"] for x in recs),
                              "links": sorted(l for x in recs for l in x["links"])}))
    else:
        import torch.distributed as dist
        pg = dist.group.WORLD if world > 1 else None
        stats = run_cfg4(ctx, args.episodes, pg)
        if rank == 0:
            print(json.dumps({"workload": "cfg4", "n_gpus": world, "episodes": stats}))
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    _main()
