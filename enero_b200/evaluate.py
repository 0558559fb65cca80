"""Batched replacement of the eval workers (script_eval_on_single_topology.py:574-652 and its
zoo / link-failure twins): the reference starts one Python process per traffic matrix
(eval_on_single_topology.py:92-94); here all the traffic matrices of a topology run as ONE
device-resident batch and the per-TM result records are written in the reference's own formats, so
the figure scripts (figures_5_and_6.py:93-109, figure_7.py:71-89, figure_8.py:77-100,
figure_9.py:110-125) read them unchanged:

    <out>/<diff>/<topo>/<topo>.<tm_id>.pckl        pickle of results = np.zeros(17):
        [3] Enero (DRL + LS) max utilisation   [6] number of links   [7] plain hill climbing
        [9] DRL only   [11] OSPF (initial)     [14] DRL cost (s)     [15] plain HC cost (s)
        [16] DRL + LS cost (s)                 [8] SAP baseline, [13] its cost (with sap=True, as
        script_eval_on_zoo_topologies.py:622,633 and the link-failure script store them; 1 otherwise, the constant
        script_eval_on_single_topology.py:619 stores)     [4] = [12] = 1 (simulated annealing is disabled upstream)
    <out>/<diff>/<topo>/<topo>.<tm_id>.timesteps   JSON list of (t, maxUti, time_start_DRL, DRL maxUti)

Costs are wall-clock seconds of the batch divided by its size (one number for all TMs of a batch):
the reference's per-process timings have no equivalent when thousands of episodes share a launch.
"""
from __future__ import annotations

import json
import os
import pickle
import time as tt

import numpy as np

from . import datasets as ds
from .engine import EpisodeBatch
from .runtime import Context
from .topology import Topology


LS_MAX_MOVES = 1024          # recorded moves per TM (the reference's LS on <=500-link graphs makes tens)


def evaluate_tms(ctx: Context, topo: Topology, tms, percentage_demands: float = 0.15, plain_hill_climbing: bool = False,
                 batch: int = 1024, sap: bool = False):
    """Full Enero inference for traffic matrices `tms` fp64[n,N,N] -> dict of per-TM arrays:
    ospf, drl, enero, (plain_hc), decisions, ls_moves, drl_history (list of improvement steps),
    and amortised per-TM costs cost_drl, cost_ls, cost_hc in seconds."""
    tms = np.ascontiguousarray(tms, dtype=np.float64)
    n = tms.shape[0]
    out = {k: np.zeros(n) for k in ("ospf", "drl", "enero", "plain_hc", "cost_drl", "cost_ls", "cost_hc")}
    out["sap"], out["cost_sap"] = np.ones(n), np.ones(n)
    out["decisions"] = np.zeros(n, dtype=np.int64)
    out["ls_moves"] = np.zeros(n, dtype=np.int64)
    out["drl_steps"] = [None] * n
    out["ls_steps"] = [None] * n
    batches = {}                                  # one device-resident batch per chunk size, reused by every chunk
    try:
        for b0 in range(0, n, batch):
            chunk = tms[b0:b0 + batch]
            B = chunk.shape[0]
            if B not in batches:
                batches[B] = EpisodeBatch(ctx, topo, B, percentage_demands)
            eb = batches[B]
            t0 = tt.perf_counter()
            eb.reset(chunk)
            eb.run_greedy()
            best, ospf, routing, n_routing = eb.best_arrays()
            t1 = tt.perf_counter()
            ls = eb.local_search(None, mode=0, want_moves=True, max_moves=LS_MAX_MOVES)
            if int(ls["n_moves"].max()) >= LS_MAX_MOVES:
                raise RuntimeError("Local Search hit the %d-move record limit" % LS_MAX_MOVES)
            t2 = tt.perf_counter()
            st = eb.state()
            _, _, hmax = eb.history()
            sl = slice(b0, b0 + B)
            out["ospf"][sl], out["drl"][sl], out["enero"][sl] = ospf, best, ls["final"]
            out["decisions"][sl], out["ls_moves"][sl] = st["steps"], ls["n_moves"]
            out["cost_drl"][sl] = (t1 - t0) / B
            out["cost_ls"][sl] = (t2 - t1) / B
            for i in range(B):
                out["drl_steps"][b0 + i] = hmax[i, :st["steps"][i]].copy()
                out["ls_steps"][b0 + i] = ls["move_val"][i, :ls["n_moves"][i]].copy()
            if plain_hill_climbing:
                t3 = tt.perf_counter()
                hc = eb.local_search(None, mode=1, want_moves=False)
                out["plain_hc"][sl] = hc["final"]
                out["cost_hc"][sl] = (tt.perf_counter() - t3) / B
            if sap:
                t4 = tt.perf_counter()
                out["sap"][sl] = eb.sap(None)
                out["cost_sap"][sl] = (tt.perf_counter() - t4) / B
    finally:
        for eb in batches.values():
            eb.close()
    return out


def result_record(res, i, num_links):
    """results[17] of one TM (script_eval_on_single_topology.py:597,631-642)."""
    r = np.zeros(17)
    r[3] = res["enero"][i]
    r[4] = 1
    r[6] = num_links
    r[7] = res["plain_hc"][i]
    r[8] = res["sap"][i]
    r[9] = res["drl"][i]
    r[11] = res["ospf"][i]
    r[12] = 1
    r[13] = res["cost_sap"][i]
    r[14] = res["cost_drl"][i]
    r[15] = res["cost_hc"][i]
    r[16] = res["cost_drl"][i] + res["cost_ls"][i]
    return r


def timesteps_record(res, i, time_start_drl=0.0):
    """(t, maxUti, time_start_DRL, DRL maxUti) rows (script :163-188,473-509,625-627): one row for the
    initial state, one per strict improvement of the DRL episode, one per Local-Search move; t advances
    by the amortised per-step cost."""
    rows = [(0, float(res["ospf"][i]))]
    steps = res["drl_steps"][i]
    dt = res["cost_drl"][i] / max(1, len(steps))
    cur = float(res["ospf"][i])
    for k, v in enumerate(steps):
        if v < cur:
            cur = float(v)
            rows.append(((k + 1) * dt, cur))
    moves = res["ls_steps"][i]
    dl = res["cost_ls"][i] / max(1, len(moves))
    for k, v in enumerate(moves):
        rows.append((res["cost_drl"][i] + (k + 1) * dl, float(v)))
    drl = float(res["drl"][i])
    return [(t, v, time_start_drl, drl) for t, v in rows]


def write_records(res, tm_ids, out_folder, differentiation_str, topology_name, num_links):
    path = os.path.join(out_folder + differentiation_str, topology_name)
    os.makedirs(path, exist_ok=True)
    for i, tm_id in enumerate(tm_ids):
        base = os.path.join(path, "%s.%s" % (topology_name, tm_id))
        with open(base + ".pckl", "wb") as f:
            pickle.dump(result_record(res, i, num_links), f, pickle.HIGHEST_PROTOCOL)
        with open(base + ".timesteps", "w") as fp:
            json.dump(timesteps_record(res, i), fp)
    return path


def evaluate_topology(ctx: Context, dataset_folder, topology_name, tm_ids, out_folder=None, differentiation_str="",
                      percentage_demands=0.15, K=100, plain_hill_climbing=True, rank=0, world=1, batch=1024, sap=False):
    """One topology of eval_on_single_topology.py / eval_on_zoo_topologies.py: this rank's share of
    `tm_ids` (contiguous shard, no collective) through DRL + LS (+ plain hill climbing), records written
    under out_folder if given.  -> (tm ids of this rank, result dict)."""
    from .parallel import shard_range
    topo = Topology.from_dataset(dataset_folder, topology_name, K=K)
    b, e = shard_range(len(tm_ids), rank, world)
    mine = list(tm_ids[b:e])
    tms = ds.parse_demands_files([os.path.join(dataset_folder, "TM", "%s.%s.demands" % (topology_name, t)) for t in mine], topo.N)
    res = evaluate_tms(ctx, topo, tms, percentage_demands, plain_hill_climbing, batch, sap) if mine else None
    if out_folder is not None and mine:
        write_records(res, mine, out_folder, differentiation_str, topology_name, topo.E)
    return mine, res
