#!/usr/bin/env python
"""ENERO hot-path benchmark: routing decisions/sec (GNN actor over all candidate
middlepoints + env step) on BASELINE.json config 2 -- a synthetic 100-link Zoo-shaped
topology (zoo_like(N=30, L=50, seed=1000)), uniform traffic matrices, shipped
ckpt_ACT-1835 weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
                    [--config cfg2|cfg3|cfg4|cfg5] [--no-sweep] [--no-cpu]

One "step" = one routing decision for each of the B device-resident episodes of this rank:
candidate features -> actor forward over every candidate middlepoint -> soft-max -> greedy
action -> Env16.step.  N > 1 (torchrun): every rank owns B independent episodes (weak
scaling, no data-path collective); time = max over ranks of the device time.

--impl reference times the reference's own CPU path (TensorFlow is not installable, so: the
reference environment semantics + the actor restated op for op in torch-CPU, index_add_ for the
segment sums -- oracle/) the way the reference deploys it: one single-threaded worker process per
traffic matrix on every host core (eval_on_single_topology.py:92-94), each running WHOLE greedy
episodes so that reset is amortised exactly as in the reference.

--config cfg3|cfg4|cfg5 run the other BASELINE.json configurations (full Enero over many TMs, one PPO
training episode, a link-failure sweep slice) and print one JSON line each in the same format.
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
# measured DRAM traffic of the dominant kernel, per engine (ncu --set full captures summarised by tools/ncu_summary.py)
NCU_FULL = {"tc": os.path.join("profiles", "r02_k_mp_tc_ncu_full.csv"), "ffma": os.path.join("profiles", "r01_v7_k_mp_iter_ncu_full.csv")}


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
# CPU path (oracle): whole greedy episodes, reference env semantics + torch-CPU actor
# ------------------------------------------------------------------------------------------
_WORKER = {}


def _cpu_episodes(args):
    """(tms [n,N,N], torch threads) -> (decisions, seconds): n whole greedy episodes, reset included, one after the
    other -- what one worker process of eval_on_single_topology.py does for its traffic matrix."""
    tms, threads = args[0], args[1]
    max_dec = args[2] if len(args) > 2 else 0          # 0 = whole episodes; tests bound the work
    import torch
    torch.set_num_threads(int(threads))
    from oracle import enero_oracle as orc
    from oracle import ppo_torch as pt
    if "topo" not in _WORKER:                        # per-process one-off, like the reference's generate_environment
        wd = load_weights()[0]
        _WORKER["w"] = {k: torch.tensor(wd[k]) for k in orc.WEIGHT_NAMES}
        _WORKER["topo"] = orc.Topology(make_topology()[0], K=100)
    w, topo = _WORKER["w"], _WORKER["topo"]
    n = 0
    t0 = time.perf_counter()
    for tm in np.asarray(tms):
        env = orc.Env16(topo)
        _, s, d = env.reset(tm)
        done = False
        while not done:
            b = orc.batch_candidates(env.candidate_features(s, d), topo.first, topo.second)
            with torch.no_grad():
                logits = pt.actor_logits(w, b)
                a = int(torch.argmax(torch.softmax(logits, dim=0)))
            _, done, _, _, s, d, _, _, _ = env.step(a)
            n += 1
            if max_dec and n >= max_dec:
                return n, time.perf_counter() - t0
    return n, time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    gf, sigma = make_topology()
    cores = os.cpu_count() or 1
    import multiprocessing as mp
    # one single-threaded worker per core (the reference's Pool fan-out): keep BLAS / OpenMP from oversubscribing
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    vals = []
    with mp.get_context("spawn").Pool(cores) as pool:
        # start-up (imports, topology precompute), untimed: a 2-decision prefix is not possible with whole episodes,
        # so the first step doubles as the warm-up of every worker
        for i in range(args.warmup + args.steps):
            tms = make_tms(gf, sigma, cores, 2000 + 64 * i)
            t0 = time.perf_counter()
            res = pool.map(_cpu_episodes, [(tms[j:j + 1], 1, args.ref_decisions) for j in range(cores)], chunksize=1)
            wall = time.perf_counter() - t0
            if i >= args.warmup:
                vals.append((sum(r[0] for r in res), wall))
    tot = sum(x[0] for x in vals)
    wall = sum(x[1] for x in vals)
    value = tot / wall
    # BASELINE.md section 2 also asks for the single-process numbers: one thread and all threads
    os.environ["OMP_NUM_THREADS"] = str(cores)
    one_tm = make_tms(gf, sigma, 1, 1999)
    _cpu_episodes((one_tm, 1, args.ref_decisions))
    n1, t1 = _cpu_episodes((one_tm, 1, args.ref_decisions))
    na, ta = _cpu_episodes((one_tm, cores, args.ref_decisions))
    sample = ("%d whole greedy episodes per step (one per process, %d single-threaded processes, ~%d decisions each): torch-CPU "
              "actor with index_add_ segment sums + reference env semantics" % (cores, cores, tot // max(1, len(vals) * cores)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, len(vals)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "TensorFlow 2.4.1 not installable: restated CPU path (oracle port)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_process_1_thread": n1 / t1, "single_process_all_threads": na / ta},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def _ncu_traffic(B, engine):
    """measured DRAM bytes of one message-passing launch from the committed ncu --set full capture of the engine in
    use (taken at 1024 episodes per GPU on this workload; scaled linearly with the batch)"""
    import csv
    for rel in (NCU_FULL[engine],):
        prof = os.path.join(ROOT, rel)
        if not os.path.isfile(prof):
            continue
        rows = {r[0]: r for r in csv.reader(open(prof)) if r}
        try:
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = [float(x) * unit[rows["dram__bytes_read.sum"][1]] for x in rows["dram__bytes_read.sum"][2:] if x]
            wr = [float(x) * unit[rows["dram__bytes_write.sum"][1]] for x in rows["dram__bytes_write.sum"][2:] if x]
            mb = (sum(rd) + sum(wr)) / len(rd) / 1e6
            return mb * 1e6 * B / 1024.0, rel + " (dram__bytes_read.sum + dram__bytes_write.sum, mean of the %d launches captured, x B/1024)" % len(rd)
        except Exception:
            continue
    return None, None


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
    step_s = ms / 1000.0 / args.steps

    # ---- roofline leg: per-launch CUDA-event time of the dominant kernel (k_mp_iter), same steps
    eb.restore()
    for i in range(2):
        one_step(i)
    ctx.sync()
    ctx.profile_begin()
    for i in range(2, 2 + min(args.steps, 8)):
        one_step(i)
    mp_ms, mp_n = ctx.profile_end()
    # live rows of the measured decisions ~ mean candidates x E (from the nK of the last policy)
    _, nK = eb.policy(want_probs=True)
    mean_k = float(nK[nK > 0].mean()) if (nK > 0).any() else 0.0
    E, P = topo.E, topo.P
    launch_s = mp_ms / max(1, mp_n) / 1000.0
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = ctx.fp32_peak_tflops()                       # measured live on this GPU (FFMA2 register loop)
    # algorithmic FLOPs of one message-passing iteration in the reference's dense per-pair formulation,
    # K' * (1600 P + 4800 E) per decision; the kernel's factored formulation executes K' * (6400 E + 40 P)
    flops_alg = B * mean_k * (1600.0 * P + 4800.0 * E)
    flops_exec = B * mean_k * (6400.0 * E + 40.0 * P)
    achieved = flops_alg / launch_s / 1e12 if launch_s else 0.0
    # API-boundary bytes of one decision (inputs materialised as the reference's call signature has them)
    bytes_api = mean_k * E * 80.0 + mean_k * E * 4.0 + 2.0 * mean_k * P * 4.0 + mean_k * 4.0
    traffic, traffic_src = _ncu_traffic(B, ctx.engine)
    tc = ctx.engine == "tc"
    # tensor-core engine: executed tensor FLOPs = 3 TF32 products (3xTF32) x 2 x 24 (K padded) x (80 + 64 + 32) columns per
    # row and iteration, over every row of the launch (padding rows of a tile are multiplied too); TF32 dense peak
    # = half of the measured bf16 peak (MEASURED_PEAKS.json bf16_tflops; nominal 1.1 vs 2.25 PFLOP/s)
    rows_launched = float(B) * topo.Kmax * E
    tensor = None
    if tc:
        tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
        t_exec = rows_launched * 3 * 2 * 24 * (80 + 64 + 32)
        tensor = {"executed_tflops": t_exec / launch_s / 1e12 if launch_s else None, "peak_tflops": tf32_peak,
                  "frac": t_exec / launch_s / 1e12 / tf32_peak if launch_s else None,
                  "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 dense = half of bf16)" if peaks else "fallback 1.59 PFLOP/s bf16 / 2",
                  "note": "the tensor pipe is not the limit (ncu: 16 % active); gather / gate latency and instruction issue are"}
    roofline = {
        "bound": "fp32",
        "kernel": ("k_mp_tc<IdxTopo> (tcgen05 3xTF32; 5 launches per decision step, ~97 % of its kernel time)" if tc
                   else "k_mp_iter<IdxTopo> (5 launches per decision step, ~94 % of it)"),
        "engine": ctx.engine, "tensor": tensor,
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
        "peak_source": "enero_ctx_fp32_peak: FFMA2 register loop timed on this GPU just now (nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4)",
        "flops_per_launch_algorithmic": flops_alg, "flops_per_launch_executed": flops_exec,
        "fp32_frac_executed": flops_exec / launch_s / 1e12 / fp32_peak if launch_s and fp32_peak else None,
        "launch_ms": launch_s * 1e3, "launches_timed": mp_n,
        "traffic": traffic, "traffic_source": traffic_src,
        "dram_frac": (traffic / launch_s / 1e9 / hbm_peak) if (traffic and launch_s) else None,
        "algorithmic_bytes_per_decision": bytes_api,
        "algorithmic_frac": B * bytes_api / step_s / 1e9 / hbm_peak,
        "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6.65 TB/s",
        "note": "achieved = SURVEY 8(d) algorithmic FLOPs (dense per-pair formulation) / launch time, against the FP32-pipe peak: the "
                "rate at which the reference's fp32 work gets done (with the tensor-core engine the mat-muls run as 3xTF32 on tcgen05, so "
                "this can exceed what the FP32 pipe alone could reach; fp32_frac_executed counts the factored formulation's FLOPs). "
                "Arithmetic intensity ~290 FLOP/B against the API-boundary bytes: not HBM bound (SURVEY 8d). dram_frac = measured DRAM bytes "
                "per launch / launch time / HBM peak; algorithmic_frac = 8(d)(i) bytes per step / step time / HBM peak",
    }

    # ---- end to end through the public API with host buffers: whole greedy episodes
    e2e = None
    cpu = None
    extras = {}
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
        eb2.run_greedy()                          # all decisions, device resident (one CUDA-graph replay per decision)
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
    if rank == 0:
        # second stage on the same batch (Local Search reported separately): hill climbing to
        # convergence from the DRL routing, on device; informative extras, not part of `value`
        # (each stage twice, the second timed: the first call pays allocations, the K-shortest-path tables and
        # module loading; it is reported as first_call_ms)
        t_ls = t_sap = first_ls = first_sap = None
        for rep in range(2):
            eb2.reset(tms_e); eb2.run_greedy(); ctx.sync()
            t0 = time.perf_counter()
            ls = eb2.local_search(None, mode=0, want_moves=False)
            ctx.sync()
            t_ls = time.perf_counter() - t0
            t0 = time.perf_counter()
            sap = eb2.sap(None)
            t_sap = time.perf_counter() - t0
            if rep == 0:
                first_ls, first_sap = t_ls, t_sap
        extras = {"local_search": {"tms_per_s": Be / t_ls, "ms": 1000.0 * t_ls, "first_call_ms": 1000.0 * first_ls, "episodes": Be, "moves_applied": int(ls["n_moves"].sum()),
                                   "mean_max_uti_drl": float(best.mean()), "mean_max_uti_enero": float(ls["final"].mean()),
                                   "mean_max_uti_ospf": float(ospf.mean())},
                  "full_enero_tms_per_s": Be / (t_best + t_ls),
                  "sap_baseline": {"tms_per_s": Be / t_sap, "ms": 1000.0 * t_sap, "first_call_ms": 1000.0 * first_sap, "mean_max_uti": float(sap.mean())}}
    eb2.close()
    eb.close()

    # ---- batch sweep: whole greedy episodes, device resident, by episodes in flight
    if rank == 0 and world == 1 and not args.no_sweep:
        sweep = []
        for Bs in (1, 16, 256, 1024, 4096, 16384):
            tms_s = tms[:Bs] if Bs <= B else np.concatenate([tms] * (Bs // B))[:Bs]
            ebs = EpisodeBatch(ctx, topo, Bs)
            best_t = None
            for rep in range(3):
                ebs.reset(tms_s)
                ctx.sync()
                l0 = ctx.launch_count
                t0 = time.perf_counter()
                ebs.run_greedy()
                dt = time.perf_counter() - t0
                if rep:
                    best_t = dt if best_t is None else min(best_t, dt)
            st = ebs.state()["steps"]
            sweep.append({"episodes": Bs, "decisions_per_s": float(st.sum()) / best_t,
                          "decision_latency_us": 1e6 * best_t / int(st.max()),
                          "launches_per_decision": (ctx.launch_count - l0) / int(st.max())})
            ebs.close()
        extras["batch_sweep"] = sweep
    if rank == 0 and world == 1 and not args.no_cpu:
        n_tms = max(1, args.cpu_decisions // 131)
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        n, wall = _cpu_episodes((make_tms(gf, sigma, n_tms, 2000), 1))
        cpu = {"value": n / wall, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d whole greedy episodes (%d decisions, reset included) on one thread: torch-CPU actor with index_add_ "
                         "segment sums + reference env semantics, %.1f s" % (n_tms, n, wall)}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "episodes_per_gpu": B, "decisions_per_step": world * B,
                       "mean_candidates": mean_k, "links": E, "pairs": P, "math": ctx.math, "engine": ctx.engine,
                       "l2": "inputs larger than L2 (%.0f MB of link-state rows per step)" % (B * topo.Kmax * E * 320 / 1e6),
                       "parallelism": "episodes sharded x%d, no collective" % world},
            "clocks": clk.summary(), "gpu_launches": int(launches), "roofline": roofline,
            "e2e": e2e, "cpu_baseline": cpu, "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, one JSON line each
# ------------------------------------------------------------------------------------------
def run_config(args, rank, world, local_rank):
    from enero_b200 import parallel as par
    from enero_b200 import workloads as wl
    from enero_b200.runtime import Context
    dist = None
    if world > 1:
        # torch only where there is a process group: once torch's CUDA runtime hooks are loaded every kernel launch of
        # the library costs ~2 us more on the host (cfg4: 128 Adam steps 0.088 s -> 0.13 s)
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the enero_b200 path has no CPU fallback)")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Context(local_rank)                      # raises without a CUDA device: there is no CPU fallback
    ctx.upload_weights(0, wl.shipped_actor_weights())
    dev = "cuda:%d" % local_rank if world > 1 else None
    base = {"n_gpus": world, "steps": 1, "warmup": 1, "higher_is_better": True, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    if args.config == "cfg3":
        n_tms = args.tms or 10000
        wl.run_cfg3(ctx, 256, rank, world)                                  # warm-up
        with ClockSampler(local_rank) as clk:                               # three passes, the fastest reported (host-side jitter)
            runs = [wl.run_cfg3(ctx, n_tms, rank, world, args.batch if args.batch <= 4096 else 2048) for _ in range(3)]
        r = min(runs, key=lambda x: x["seconds"])
        secs = par.max_over_ranks(r["seconds"], device=dev)
        recs = par.gather_records([r])
        if rank == 0:
            tms = sum(x["tms"] for x in recs)
            print(json.dumps(dict(base, metric="full Enero inference (greedy DRL + Local Search), traffic matrices/sec", unit="TMs/s",
                                  value=tms / secs, ms_per_step=1e3 * secs, scaling="strong", clocks=clk.summary(),
                                  config={"workload": "cfg3: zoo_like(24,37,1002) E=74, %d uniform TMs, host TMs in -> records out" % tms,
                                          "parallelism": "TMs sharded x%d, no collective" % world},
                                  extras={"decisions_per_s": sum(x["decisions"] for x in recs) / secs,
                                          "ls_moves": sum(x["ls_moves"] for x in recs), "mean_ospf": recs[0]["mean_ospf"],
                                          "mean_drl": recs[0]["mean_drl"], "mean_enero": recs[0]["mean_enero"],
                                          "seconds_all_passes": [x["seconds"] for x in runs],
                                          "reference": "4.5 s per TM on topologies up to 100 links (README.md:17)"})), flush=True)
    elif args.config == "cfg5":
        n_top, n_tms = args.topologies, args.tms or 64
        with ClockSampler(local_rank) as clk:
            r = wl.run_cfg5(ctx, n_top, n_tms, rank, world)
        secs = par.max_over_ranks(r["seconds"] + r["build_seconds"], device=dev)
        recs = par.gather_records([r])
        if rank == 0:
            units = sum(x["tms"] for x in recs)
            print(json.dumps(dict(base, metric="link-failure sweep, (topology, TM) evaluations/sec (DRL + LS)", unit="evaluations/s",
                                  value=units / secs, ms_per_step=1e3 * secs, scaling="strong", clocks=clk.summary(), warmup=0,
                                  config={"workload": "cfg5 slice: %d link-failure variants x %d TMs, 100-500 directed links" % (n_top, n_tms),
                                          "parallelism": "topologies sharded x%d by estimated cost, no collective" % world},
                                  extras={"links": sorted(l for x in recs for l in x["links"]),
                                          "decisions_per_s": sum(x["decisions"] for x in recs) / secs,
                                          "build_seconds": sum(x["build_seconds"] for x in recs),
                                          "drl_seconds": sum(x["drl_seconds"] for x in recs),
                                          "ls_seconds": sum(x["ls_seconds"] for x in recs)})), flush=True)
    else:
        pg = dist.group.WORLD if world > 1 else None
        with ClockSampler(local_rank) as clk:
            stats = wl.run_cfg4(ctx, args.episodes, pg)
        if rank == 0:
            n = sum(s["samples"] for s in stats)
            secs = sum(s["collect_s"] + s["update_s"] for s in stats)
            print(json.dumps(dict(base, metric="PPO training samples/sec (rollouts + 8 epochs x 16 minibatches of clip+Adam)", unit="samples/s",
                                  value=n / secs, ms_per_step=1e3 * secs / len(stats), steps=len(stats), scaling="strong",
                                  clocks=clk.summary(),
                                  config={"workload": "cfg4: zoo_like (20,31,1003) (23,25,1004) (17,31,1005), 835-sample buffer, minibatch 55",
                                          "parallelism": ("minibatch sharded x%d, one NCCL all-reduce of 8412 floats per Adam step" % world)
                                          if world > 1 else "single GPU"},
                                  extras={"collect_s": [s["collect_s"] for s in stats], "update_s": [s["update_s"] for s in stats],
                                          "eval_s": [s["eval_s"] for s in stats], "adam_steps": stats[0]["adam_steps"],
                                          "adam_step_ms": 1e3 * stats[-1]["update_s"] / stats[-1]["adam_steps"]})), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096, help="episodes per GPU")
    ap.add_argument("--e2e-batch", type=int, default=4096)
    ap.add_argument("--cpu-decisions", type=int, default=1000, help="CPU-baseline sample (about 15 s on one core)")
    ap.add_argument("--ref-decisions", type=int, default=0,
                    help="--impl reference: cap every worker at this many decisions per step (0 = whole episodes; tests use a small cap)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 1..16384-episode batch sweep of the N=1 run")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--tms", type=int, default=None, help="cfg3 / cfg5: traffic matrices")
    ap.add_argument("--topologies", type=int, default=8, help="cfg5: link-failure variants")
    ap.add_argument("--episodes", type=int, default=2, help="cfg4: PPO episodes")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config != "cfg2":
        run_config(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
