#!/usr/bin/env python
"""ENERO hot-path benchmark: routing decisions/sec (GNN actor over all candidate
middlepoints + env step) on BASELINE.json config 2 -- a synthetic 100-link Zoo-shaped
topology (zoo_like(N=30, L=50, seed=1000)), uniform traffic matrices, shipped
ckpt_ACT-1835 weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one routing decision for each of the B device-resident episodes of this rank:
candidate features -> actor forward over every candidate middlepoint -> soft-max -> greedy
action -> Env16.step.  N > 1 (torchrun): every rank owns B independent episodes (weak
scaling, no data-path collective); time = max over ranks of the device time.

--impl reference times the reference's own CPU path: the reference environment semantics
and the restated actor (oracle/, TensorFlow is not installable) on all host cores, one
process per traffic matrix like eval_on_single_topology.py:92-94.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "routing decisions/sec (GNN actor + env step)"
UNIT = "decisions/s"
WORKLOAD = "cfg2: zoo_like(N=30,L=50,seed=1000) E=100 links, uniform TMs seed 2000+j, ckpt_ACT-1835, greedy DRL decisions"
TOPO = (30, 50, 1000)
GOLD = os.path.join(ROOT, "tests", "golden")


def load_weights():
    z = np.load(os.path.join(GOLD, "ckpt_ACT-1835_weights.npz"))
    names = ("msg_kernel", "msg_bias", "gru_kernel", "gru_recurrent_kernel", "gru_bias",
             "r1_kernel", "r1_bias", "r2_kernel", "r2_bias", "r3_kernel", "r3_bias")
    return {k: z[k] for k in names}, [z[k] for k in names]


def make_topology():
    from enero_b200 import datasets as ds
    n, l, seed = TOPO
    links, caps = ds.zoo_like(n, l, seed)
    tmp = tempfile.mkdtemp(prefix="enero_bench_")
    p = os.path.join(tmp, "bench.graph")
    ds.write_graph_file(p, n, links, caps)
    gf = ds.parse_graph_file(p)
    return gf, ds.sigma_for_mean_utilisation(gf)


def make_tms(gf, sigma, count, first_seed):
    from enero_b200 import datasets as ds
    return np.stack([ds.uniform_tm(gf.num_nodes, sigma, first_seed + j) for j in range(count)])


# ------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread every 5 ms
    (nvidia_ml_py), nvidia-smi every 100 ms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML indexes physical ones
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _run_nvml(self):
        n = self._nvml
        bits = ((getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap"))
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM))
        except Exception:
            mx = 0.0
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                mask = int(get_reasons(self._h))
                self.rows.append([sm, mx] + ["Active" if mask & b else "Not Active" for b, _ in bits])
            except Exception:
                pass
            self._stop.wait(0.005)

    def _run_smi(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for nme, v in zip(self.NAMES, r[2:6]):
                if str(v).lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self._nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------------------
# CPU path (oracle): one decision = restated actor over all candidates + env step
# ------------------------------------------------------------------------------------------
_WORKER = {}


def _cpu_worker(args):
    tms, n_dec, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    from enero_b200 import datasets as ds            # noqa: F401
    from oracle import enero_oracle as orc
    if "topo" not in _WORKER:                        # per-process one-off, like the reference's generate_environment
        _WORKER["wd"] = load_weights()[0]
        _WORKER["topo"] = orc.Topology(make_topology()[0], K=100)
    wd, topo = _WORKER["wd"], _WORKER["topo"]
    tms = np.asarray(tms)
    if tms.ndim == 2:
        tms = tms[None]
    n, spent = 0, 0.0
    for tm in tms:                                   # greedy episodes, one TM after the other, until the budget is used
        env = orc.Env16(topo)
        _, s, d = env.reset(tm)
        t0 = time.perf_counter()
        while n < n_dec:
            probs, _, _ = orc.pred_action_distrib(wd, env, s, d)
            _, done, _, _, s, d, _, _, _ = env.step(int(np.argmax(probs)))
            n += 1
            if done:
                break
        spent += time.perf_counter() - t0
        if n >= n_dec:
            break
    return n, spent


def cpu_decisions_per_sec(gf, sigma, procs, n_dec, first_seed=2000, pool=None):
    """`procs` worker processes (the reference's Pool fan-out), each `n_dec` decisions of its own TM."""
    tms = make_tms(gf, sigma, procs, first_seed)
    jobs = [(tms[i], n_dec, i) for i in range(procs)]
    if pool is None:
        n_tms = max(1, -(-n_dec // 100))             # ~131 decisions per TM on this topology
        res = [_cpu_worker((make_tms(gf, sigma, n_tms, first_seed), n_dec, 0))]
        wall = res[0][1]
    else:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    return total / wall, total, wall


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    gf, sigma = make_topology()
    cores = os.cpu_count() or 1
    # bounded sample: each step = `per` decisions on each of `cores` processes (~0.03-0.05 s per decision)
    import multiprocessing as mp
    # one single-threaded worker per core (the reference's Pool fan-out): keep BLAS from oversubscribing
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    per = 6
    vals = []
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [(make_tms(gf, sigma, 1, 1999)[0], 1, 0)] * cores, chunksize=1)   # start-up, untimed
        for i in range(args.warmup + args.steps):
            v, total, wall = cpu_decisions_per_sec(gf, sigma, cores, per, 2000 + 64 * i, pool)
            if i >= args.warmup:
                vals.append((v, total, wall))
    tot = sum(x[1] for x in vals)
    wall = sum(x[2] for x in vals)
    value = tot / wall
    sample = "%d decisions per step (%d processes x %d decisions, restated numpy actor + reference env semantics)" % (cores * per, cores, per)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, len(vals)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "TensorFlow 2.4.1 not installable: restated CPU path (oracle port)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from enero_b200.engine import EpisodeBatch
    from enero_b200.runtime import Context
    from enero_b200.topology import Topology

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the enero_b200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wd, wl = load_weights()
    gf, sigma = make_topology()
    topo = Topology(gf, K=100)
    ctx = Context(local_rank)
    ctx.upload_weights(0, wl)
    B = args.batch
    tms = make_tms(gf, sigma, B, 2000 + rank * B)
    eb = EpisodeBatch(ctx, topo, B)
    eb.reset(tms)
    _, n_elig = eb.eligible()
    horizon = int(n_elig.min())
    eb.snapshot()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def one_step(i):
        if i % horizon == 0 and i:
            eb.restore()                      # episode horizon reached: rewind the device state (D2D copy)
        eb.policy(want_probs=False)
        eb.step(None, want_outputs=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        with torch.cuda.stream(stream):
            e0.record()
        for i in range(args.steps):
            one_step(args.warmup + i)
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms / 1000.0)

    # ---- roofline leg: per-launch CUDA-event time of the dominant kernel (k_mp_iter), same steps
    eb.restore()
    for i in range(2):
        one_step(i)
    ctx.sync()
    ctx.profile_begin()
    for i in range(2, 2 + min(args.steps, 8)):
        one_step(i)
    mp_ms, mp_n = ctx.profile_end()
    st = eb.state()
    # live rows of the measured decisions ~ mean candidates x E (from the nK of the last policy)
    _, nK = eb.policy(want_probs=True)
    mean_k = float(nK[nK > 0].mean()) if (nK > 0).any() else 0.0
    E, P = topo.E, topo.P
    # algorithmic bytes of one message-passing iteration in this (multi-launch) formulation, per decision:
    # per live row read h (80 B) + write h' (80 B) + write A' (80 B); per message pair gather A (80 B)
    bytes_per_dec_iter = mean_k * (E * 240.0 + P * 80.0)
    flops_per_dec_iter = mean_k * (E * 2.0 * (400 + 400 + 2400) + P * 20.0 * 2)
    launch_ms = mp_ms / max(1, mp_n)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = B * bytes_per_dec_iter / (launch_ms / 1000.0) / 1e9 if launch_ms else 0.0
    fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    fp32_ach = B * flops_per_dec_iter / (launch_ms / 1000.0) / 1e12 if launch_ms else 0.0
    # measured DRAM traffic of one k_mp_iter launch, from the committed ncu --set full capture (taken at
    # 1024 episodes per GPU on this workload; scaled linearly with the batch)
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "r01_v7_k_mp_iter_ncu_full.csv")
    if os.path.isfile(prof):
        import csv
        rows = {r[0]: r for r in csv.reader(open(prof)) if r}
        try:
            mb = float(rows["dram__bytes_read.sum"][2]) + float(rows["dram__bytes_write.sum"][2])
            traffic = mb * 1e6 * B / 1024.0
            traffic_src = "profiles/r01_v7_k_mp_iter_ncu_full.csv (dram__bytes_read.sum + dram__bytes_write.sum, launch 1, x B/1024)"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": B * bytes_per_dec_iter, "kernel": "k_mp_iter<IdxTopo>",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6.65 TB/s",
                "launch_ms": launch_ms, "launches_timed": mp_n,
                "note": "the kernel is FP32-FMA bound, not HBM bound (SURVEY 8d): see fp32",
                "fp32": {"achieved_tflops": fp32_ach, "peak_tflops": fp32_peak, "frac": fp32_ach / fp32_peak,
                         "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz"}}

    # ---- end to end through the public API with host buffers: whole greedy episodes
    e2e = None
    cpu = None
    extras = None
    if rank == 0 or world > 1:
        Be = min(B, args.e2e_batch)
        tms_e = np.ascontiguousarray(tms[:Be])
        eb2 = EpisodeBatch(ctx, topo, Be)
        t_best = None
        for rep in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eb2.reset(tms_e)                          # H2D of the traffic matrices, loads on device, demand lists
            eb2.run_greedy()                          # all decisions, device resident
            best, ospf, routing, n_routing = eb2.best_arrays()   # D2H of the results
            ctx.sync()
            dt = time.perf_counter() - t0
            if rep:                                    # first repetition warms allocations
                t_best = dt if t_best is None else min(t_best, dt)
        dec = int(eb2.state()["steps"].sum())
        if world > 1:
            t = torch.tensor([t_best], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_best = float(t.item())
        steps_e = dec / Be
        e2e = {"value": world * dec / t_best, "unit": UNIT,
               "h2d_bytes_per_step": int(tms_e.nbytes / steps_e + Be * eb2.Dmax * 4 / steps_e),
               "d2h_bytes_per_step": int((Be * E * 8 + Be * eb2.Dmax * 4 + Be * 28) / steps_e),
               "path": "EpisodeBatch.reset(host TMs) + run_greedy() + best(): %d episodes x %.0f decisions, wall clock incl. copies" % (Be, steps_e)}
        # second stage on the same batch (Local Search reported separately): hill climbing to
        # convergence from the DRL routing, on device; informative extras, not part of `value`
        t0 = time.perf_counter()
        ls = eb2.local_search(None, mode=0, want_moves=False)
        ctx.sync()
        t_ls = time.perf_counter() - t0
        extras = {"local_search": {"tms_per_s": Be / t_ls, "ms": 1000.0 * t_ls, "episodes": Be, "moves_applied": int(ls["n_moves"].sum()),
                                   "mean_max_uti_drl": float(best.mean()), "mean_max_uti_enero": float(ls["final"].mean()),
                                   "mean_max_uti_ospf": float(ospf.mean())},
                  "full_enero_tms_per_s": Be / (t_best + t_ls)}
        eb2.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = 1
        v, total, wall = cpu_decisions_per_sec(gf, sigma, 1, args.cpu_decisions, 2000, None)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d greedy decisions over consecutive TMs (restated numpy actor + env step), %.1f s" % (total, wall)}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "episodes_per_gpu": B, "decisions_per_step": world * B,
                       "mean_candidates": mean_k, "links": E, "pairs": P,
                       "l2": "inputs larger than L2 (%.0f MB of link-state rows per step)" % (B * topo.Kmax * E * 320 / 1e6),
                       "parallelism": "episodes sharded x%d, no collective" % world},
            "clocks": clk.summary(), "gpu_launches": int(launches), "roofline": roofline,
            "e2e": e2e, "cpu_baseline": cpu, "extras": extras,
        }
        print(json.dumps(line), flush=True)
    eb.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096, help="episodes per GPU")
    ap.add_argument("--e2e-batch", type=int, default=4096)
    ap.add_argument("--cpu-decisions", type=int, default=1000, help="CPU-baseline sample (about 15 s on one core)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
