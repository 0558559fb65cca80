"""Parity sweep of the GNN arithmetic (not a test): GPU-fast, GPU-accurate and the numpy fp32 oracle against the
fp64 arbiter (oracle.enero_oracle.actor_forward64) on every decision of 692 greedy episodes, plus greedy-sequence
/ best-utilisation / Local-Search identity of both GPU math modes and what each mode costs.

    python tools/parity_sweep.py gen  [out.npz]      # CPU, build container: oracle greedy episodes, all cores
    python tools/parity_sweep.py run  [in.npz]       # GPU box: both math modes against the recorded states

`gen` records, for every decision on the ORACLE's greedy path: the state the actor saw (link loads, demand), the
oracle's fp32 logits and the fp64 arbiter's logits on that same state, and per episode the action sequence, best
utilisation, OSPF utilisation and (small graphs) the Local-Search result.  `run` feeds exactly those states to
enero_decide_batch in each math mode, so every implementation is compared on bit-identical inputs, and then
replays the whole episodes on the device.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import case_graph_file, golden_weights, load_case      # noqa: E402

PLAN = (("tri9", 200), ("tiny12", 400), ("eli20", 60), ("zoo30", 32))
LS_ORACLE = {"tri9": 200, "tiny12": 400, "eli20": 60, "zoo30": 4}      # oracle Local Search is slow on 100 links
FIRST_SEED = 90000
DEFAULT = os.path.join(ROOT, "tools", "_sweep", "states.npz")
FLOOR = 1e-3          # |x| below this is compared absolutely (relative error of a value near zero is meaningless)


def _episode(args):
    name, j, do_ls = args
    from enero_b200 import datasets as ds
    from oracle import enero_oracle as orc
    g = _episode.__dict__
    if g.get("name") != name:
        gf = case_graph_file(load_case(name), tempfile.mkdtemp())
        g.update(name=name, gf=gf, ot=orc.Topology(gf, K=100), w=golden_weights(), sigma=ds.sigma_for_mean_utilisation(gf))
    ot, w = g["ot"], g["w"]
    tm = ds.uniform_tm(g["gf"].num_nodes, g["sigma"], FIRST_SEED + j)
    env = orc.Env16(ot)
    _, s, d = env.reset(tm)
    best = ospf = env.edge_max[2]
    routing = dict(env.sp_middlepoints)
    rec = []
    while True:
        loads = env.loads.copy()
        feats = env.candidate_features(s, d)
        b = orc.batch_candidates(feats, ot.first, ot.second)
        l32 = orc.actor_forward(w, b["link_state"], b["graph_id"], b["first"], b["second"], b["num_edges"]).reshape(-1)
        l64 = orc.actor_forward64(w, b["link_state"], b["graph_id"], b["first"], b["second"], b["num_edges"]).reshape(-1)
        a = int(np.argmax(orc.softmax(l32)))
        rec.append((loads, s, d, tm[s, d], l32, l64, a))
        _, done, _, _, s, d, mx, _, _ = env.step(a)
        if mx[2] < best:
            best = mx[2]
            routing = dict(env.sp_middlepoints)
        if done:
            break
    ls_final = np.nan
    if do_ls:
        e15 = orc.Env15(ot)
        ls_final, _ = orc.local_search(e15, e15.reset_DRL_hill_sp(tm, routing))
    return name, j, tm, rec, best, ospf, ls_final


def gen(path):
    import multiprocessing as mp
    jobs = [(name, j, j < LS_ORACLE[name]) for name, n in PLAN for j in range(n)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(_episode, jobs, chunksize=4)
    out = {}
    for name, n in PLAN:
        eps = sorted((r for r in res if r[0] == name), key=lambda r: r[1])
        K = max(len(x[4]) for r in eps for x in r[3])
        dec = [x for r in eps for x in r[3]]
        nd = len(dec)
        l32 = np.zeros((nd, K), np.float32)
        l64 = np.zeros((nd, K))
        nk = np.zeros(nd, np.int32)
        for i, x in enumerate(dec):
            nk[i] = len(x[4])
            l32[i, :nk[i]] = x[4]
            l64[i, :nk[i]] = x[5]
        out.update({
            name + "/tms": np.stack([r[2] for r in eps]),
            name + "/ep_len": np.array([len(r[3]) for r in eps], np.int32),
            name + "/loads": np.stack([x[0] for x in dec]),
            name + "/sd": np.array([(x[1], x[2]) for x in dec], np.int32),
            name + "/bw": np.array([x[3] for x in dec]),
            name + "/l32": l32, name + "/l64": l64, name + "/nk": nk,
            name + "/action": np.array([x[6] for x in dec], np.int32),
            name + "/best": np.array([r[4] for r in eps]), name + "/ospf": np.array([r[5] for r in eps]),
            name + "/ls_final": np.array([r[6] for r in eps]),
        })
        print(name, "episodes", n, "decisions", nd, flush=True)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print("wrote %s in %.0f s" % (path, time.time() - t0))


def rel_err(got, want, nk):
    """max over the live entries of |got - want| / max(|want|, FLOOR)"""
    m = np.arange(got.shape[1])[None, :] < nk[:, None]
    e = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), FLOOR)
    return float(e[m].max()), float(np.sqrt((e[m] ** 2).mean()))


def needed_atol(got, want, nk, rtol=1e-5):
    """smallest absolute floor under which every live element satisfies |got - want| <= rtol * |want| + atol"""
    m = np.arange(got.shape[1])[None, :] < nk[:, None]
    d = np.abs(got.astype(np.float64) - want) - rtol * np.abs(want)
    return float(max(0.0, d[m].max()))


def softmax_rows(x, nk):
    m = np.arange(x.shape[1])[None, :] < nk[:, None]
    z = np.where(m, x.astype(np.float64), -np.inf)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def run(path):
    from enero_b200.engine import EpisodeBatch
    from enero_b200.runtime import Context
    from enero_b200.topology import Topology
    from oracle import enero_oracle as orc
    z = np.load(path)
    w = golden_weights()
    ctx = Context(0)
    ctx.upload_weights(0, [w[k] for k in orc.WEIGHT_NAMES])
    report = {"floor": FLOOR, "topologies": {}, "modes": {}}
    agg = {m: {"logit_vs_fp64": 0.0, "prob_vs_fp64": 0.0, "logit_vs_oracle": 0.0, "atol_needed_vs_fp64": 0.0,
               "atol_needed_vs_oracle": 0.0, "argmax_vs_oracle": 0, "argmax_vs_fp64": 0,
               "sequences_differ": 0, "best_differs": 0, "ls_differs": 0, "ls_checked": 0} for m in ("fast", "accurate")}
    agg["oracle_fp32"] = {"logit_vs_fp64": 0.0, "prob_vs_fp64": 0.0, "atol_needed_vs_fp64": 0.0, "argmax_vs_fp64": 0}
    total_dec = total_eps = 0
    for name, n in PLAN:
        gf = case_graph_file(load_case(name), tempfile.mkdtemp())
        topo = Topology(gf, K=100)
        g = lambda k: z[name + "/" + k]
        loads, sd, bw, l32, l64, nk, act = g("loads"), g("sd"), g("bw"), g("l32"), g("l64"), g("nk"), g("action")
        nd, K = l32.shape
        total_dec += nd
        total_eps += n
        p64 = softmax_rows(l64, nk)
        a64 = p64.argmax(axis=1)
        o = agg["oracle_fp32"]
        e, _ = rel_err(l32, l64, nk)
        o["logit_vs_fp64"] = max(o["logit_vs_fp64"], e)
        o["atol_needed_vs_fp64"] = max(o["atol_needed_vs_fp64"], needed_atol(l32, l64, nk))
        p32 = softmax_rows(l32, nk)     # the oracle's own fp32 softmax adds < 1 ulp; the logits carry the error
        o["prob_vs_fp64"] = max(o["prob_vs_fp64"], rel_err(p32, p64, nk)[0])
        o["argmax_vs_fp64"] += int((act != a64).sum())
        per = {"episodes": n, "decisions": int(nd), "oracle_argmax_vs_fp64": int((act != a64).sum())}
        for mode in ("fast", "accurate"):
            ctx.math = mode
            a = agg[mode]
            lg = np.zeros((nd, topo.Kmax), np.float32)
            pr = np.zeros((nd, topo.Kmax), np.float32)
            ac = np.zeros(nd, np.int32)
            for i0 in range(0, nd, 2048):
                sl = slice(i0, min(nd, i0 + 2048))
                lg[sl], pr[sl], k2, ac[sl] = ctx.decide_batch(topo, loads[sl], sd[sl], bw[sl])
                assert np.array_equal(k2, nk[sl])
            lg, pr = lg[:, :K], pr[:, :K]
            e64, rms64 = rel_err(lg, l64, nk)
            a["logit_vs_fp64"] = max(a["logit_vs_fp64"], e64)
            a["prob_vs_fp64"] = max(a["prob_vs_fp64"], rel_err(pr, p64, nk)[0])
            a["logit_vs_oracle"] = max(a["logit_vs_oracle"], rel_err(lg, l32.astype(np.float64), nk)[0])
            a["atol_needed_vs_fp64"] = max(a["atol_needed_vs_fp64"], needed_atol(lg, l64, nk))
            a["atol_needed_vs_oracle"] = max(a["atol_needed_vs_oracle"], needed_atol(lg, l32.astype(np.float64), nk))
            a["argmax_vs_oracle"] += int((ac != act).sum())
            a["argmax_vs_fp64"] += int((ac != a64).sum())
            # whole episodes on the device
            eb = EpisodeBatch(ctx, topo, n)
            eb.reset(g("tms"))
            eb.run_greedy()
            best, ospf, _ = eb.best()
            hist, _, _ = eb.history()
            steps = eb.state()["steps"]
            ls = eb.local_search(None, mode=0)
            off = np.concatenate([[0], np.cumsum(g("ep_len"))])
            differ = []
            for i in range(n):
                want = act[off[i]:off[i + 1]].tolist()
                if hist[i, :steps[i]].tolist() != want:
                    t = next((q for q in range(min(len(want), steps[i])) if hist[i, q] != want[q]), None)
                    differ.append({"episode": i, "first_step": t,
                                   "oracle_logits_top2": sorted(l32[off[i] + t, :nk[off[i] + t]].tolist())[-2:] if t is not None else None,
                                   "fp64_action": int(a64[off[i] + t]) if t is not None else None,
                                   "oracle_action": int(want[t]) if t is not None else None,
                                   "gpu_action": int(hist[i, t]) if t is not None else None})
            a["sequences_differ"] += len(differ)
            a["best_differs"] += int(((best != g("best")) | (ospf != g("ospf"))).sum())
            lsf = g("ls_final")
            chk = ~np.isnan(lsf)
            a["ls_checked"] += int(chk.sum())
            a["ls_differs"] += int((ls["final"][chk] != lsf[chk]).sum())
            per[mode] = {"logit_max_rel_vs_fp64": e64, "logit_rms_rel_vs_fp64": rms64, "sequences_differ": differ}
            eb.close()
        report["topologies"][name] = per
    # throughput of both modes on the benchmark topology (cfg2), 4096 device-resident episodes
    from enero_b200 import datasets as ds
    gf = case_graph_file(load_case("zoo30"), tempfile.mkdtemp())
    topo = Topology(gf, K=100)
    sigma = ds.sigma_for_mean_utilisation(gf)
    B = 4096
    tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 2000 + j) for j in range(B)])
    eb = EpisodeBatch(ctx, topo, B)
    eb.reset(tms)
    eb.snapshot()
    for mode in ("fast", "accurate", "fast", "accurate"):
        ctx.math = mode
        eb.restore()
        for _ in range(3):
            eb.policy(want_probs=False)
            eb.step(None, want_outputs=False)
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(20):
            eb.policy(want_probs=False)
            eb.step(None, want_outputs=False)
        ctx.sync()
        agg[mode]["decisions_per_s"] = B * 20 / (time.perf_counter() - t0)
    report["modes"] = agg
    report["episodes"], report["decisions"] = total_eps, total_dec
    print(json.dumps(report, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "parity_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    path = sys.argv[2] if len(sys.argv) > 2 else DEFAULT
    gen(path) if mode == "gen" else run(path)
