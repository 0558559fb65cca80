"""GPU tests of the batched eval worker (result-record writers) and of the
training loop (train_Enero_3top_script.py:471-722) on synthetic data."""
import json
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN, CASE_NAMES, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(weights):
    from enero_b200.runtime import Context
    from oracle import enero_oracle as orc
    c = Context(0)
    c.upload_weights(0, [weights[k] for k in orc.WEIGHT_NAMES])
    yield c
    c.close()


def _dataset(case, tmp):
    folder = os.path.join(str(tmp), case["name"])
    os.makedirs(os.path.join(folder, "TM"), exist_ok=True)
    with open(os.path.join(folder, case["name"] + ".graph"), "w") as fp:
        fp.write(case["graph_text"])
    from enero_b200 import datasets as ds
    for tm_id, tm in case["tms"].items():
        ds.write_demands_file(os.path.join(folder, "TM", "%s.%s.demands" % (case["name"], tm_id)), np.array(tm))
    return folder


@pytest.mark.parametrize("name", CASE_NAMES)
def test_eval_worker_records(name, ctx, tmp_path):
    """the .pckl / .timesteps records of a batched run carry the values the reference's own env / LS code
    produced for the same TMs (tests/golden, recorded by oracle/gen_golden.py)"""
    from enero_b200.evaluate import evaluate_topology
    case = load_case(name)
    folder = _dataset(case, tmp_path)
    tm_ids = [ep["tm_id"] for ep in case["episodes"]]
    want_hc = all(ep.get("plain_hill_final") is not None for ep in case["episodes"])
    sap_gold = os.path.join(GOLDEN, "sap_%s.json" % name)
    want_sap = os.path.isfile(sap_gold)        # the reference's own Env20 ran on these cases (oracle/gen_sap_golden.py)
    mine, res = evaluate_topology(ctx, folder, name, tm_ids, out_folder=str(tmp_path) + "/out", differentiation_str="X",
                                  plain_hill_climbing=want_hc, sap=want_sap)
    assert mine == tm_ids
    for i, ep in enumerate(case["episodes"]):
        base = os.path.join(str(tmp_path) + "/outX", name, "%s.%s" % (name, ep["tm_id"]))
        r = pickle.load(open(base + ".pckl", "rb"))
        assert r.shape == (17,)
        assert r[11] == ep["ospf"] and r[9] == ep["best"]
        assert r[3] == ep["ls"]["final"]
        assert r[6] == len(case["topo"]["edges"])
        if want_hc:
            assert r[7] == ep["plain_hill_final"]
        assert r[4] == r[12] == 1
        if want_sap:                               # results[8] (script_eval_on_zoo_topologies.py:622,633); TM i <-> gold episode 3*i (scale 1.0)
            assert r[8] == json.load(open(sap_gold))["episodes"][3 * i]["max_uti"] and r[13] > 0
        else:
            assert r[8] == r[13] == 1
        ts = json.load(open(base + ".timesteps"))
        assert ts[0][1] == ep["ospf"] and all(row[3] == ep["best"] for row in ts)
        vals = [row[1] for row in ts]
        assert all(b <= a for a, b in zip(vals, vals[1:]))          # improvements only
        assert vals[-1] == ep["ls"]["final"]


def test_sharded_eval_partitions(ctx, tmp_path):
    """two ranks' shards of the TM list cover it once; per-TM results do not depend on the sharding"""
    from enero_b200.evaluate import evaluate_topology
    case = load_case("eli20")
    folder = _dataset(case, tmp_path)
    tm_ids = [ep["tm_id"] for ep in case["episodes"]]
    full = evaluate_topology(ctx, folder, "eli20", tm_ids, plain_hill_climbing=False)
    parts = [evaluate_topology(ctx, folder, "eli20", tm_ids, plain_hill_climbing=False, rank=r, world=2) for r in range(2)]
    assert parts[0][0] + parts[1][0] == tm_ids
    got = np.concatenate([p[1]["enero"] for p in parts if p[1] is not None])
    assert np.array_equal(got, full[1]["enero"])


@pytest.mark.parametrize("mode", ["device", "host"])
def test_training_loop_runs_and_checkpoints(ctx, tmp_path, mode):
    """one PPO episode on three small synthetic topologies: buffer sizes follow the quota rule, the log
    protocol and the checkpoint files match the reference's layout, a restored run resumes the optimiser"""
    from enero_b200 import checkpoint as ckpt
    from enero_b200.train import TrainingRun
    specs = ((8, 11, 1003), (9, 12, 1004), (7, 10, 1005))
    log = os.path.join(str(tmp_path), "expLogs.txt")
    cdir = os.path.join(str(tmp_path), "models")
    run = TrainingRun.synthetic(ctx, specs, n_train_tms=6, n_eval_tms=3, folder=os.path.join(str(tmp_path), "data"),
                                log_path=log, checkpoint_dir=cdir, eval_episodes=3, rollout_mode=mode)
    assert run.quotas == [int(np.ceil(0.15 * n * (n - 1))) * m for (n, _, _), m in zip(specs, (5, 4, 6))]
    w0 = np.concatenate([w.reshape(-1) for w in ctx.download_weights(0)])
    st = run.ppo_episode()
    assert st["samples"] == sum(run.quotas) and st["adam_steps"] == 8 * ((st["samples"] + 54) // 55)
    w1 = np.concatenate([w.reshape(-1) for w in ctx.download_weights(0)])
    assert np.abs(w1 - w0).max() > 0 and np.isfinite(w1).all()
    assert run.agent.optimizer.iterations == st["adam_steps"]
    tags = [l.split(",")[0] for l in open(log).read().splitlines()]
    assert tags[:9] == ["a", "c", ";", "+", "<", ">", "ENTR", "REW", "lr"] and tags[9].startswith("MAX REWD")
    for kind in ("ACT", "CRT"):
        assert os.path.isfile(os.path.join(cdir, "ckpt_%s-1.index" % kind))
        assert os.path.isfile(os.path.join(cdir, "ckpt_%s-1.data-00000-of-00001" % kind))
    b = ckpt.Bundle.load(os.path.join(cdir, "ckpt_ACT-1"))
    assert b.optimizer_scalars()["iter"] == st["adam_steps"] and b.optimizer_scalars()["save_counter"] == 1
    np.testing.assert_array_equal(np.concatenate([w.reshape(-1) for w in b.model_weights()]), w1)
    # resume: weights, Adam slots and the step counter come back from the files
    run2 = TrainingRun.synthetic(ctx, specs, n_train_tms=6, n_eval_tms=3, folder=os.path.join(str(tmp_path), "data"),
                                 checkpoint_dir=cdir, counter=2, eval_episodes=3)
    run2.restore(1)
    assert run2.agent.optimizer.iterations == st["adam_steps"]
    m, _ = run2.agent.download_slots(0)
    np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in m]),
                                  np.concatenate([x.reshape(-1) for x in b.adam_slots("m")]))
    st2 = run2.ppo_episode()
    assert os.path.isfile(os.path.join(cdir, "ckpt_ACT-2.index")) and np.isfinite(st2["actor_loss"])


def test_evaluate_tms_independent_of_chunking(ctx, tmp_path):
    """evaluate_tms walks the matrices in chunks on ONE device batch per chunk size (re-used, reset per chunk): the
    per-TM results do not depend on the chunk size, including a ragged last chunk"""
    from enero_b200 import datasets as ds
    from enero_b200.evaluate import evaluate_tms
    from enero_b200.topology import Topology
    case = load_case("tiny12")
    folder = _dataset(case, tmp_path)
    gf = ds.parse_graph_file(os.path.join(folder, "tiny12.graph"))
    topo = Topology(gf, K=100)
    sigma = ds.sigma_for_mean_utilisation(gf)
    tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 700 + j) for j in range(7)])
    whole = evaluate_tms(ctx, topo, tms, batch=1024)
    pieces = evaluate_tms(ctx, topo, tms, batch=3)               # chunks of 3, 3, 1
    for k in ("ospf", "drl", "enero", "decisions", "ls_moves"):
        assert np.array_equal(whole[k], pieces[k]), k
    for a, b in zip(whole["drl_steps"] + whole["ls_steps"], pieces["drl_steps"] + pieces["ls_steps"]):
        assert np.array_equal(a, b)
    topo.close()


def test_rollouts_side_by_side_equal_one_after_the_other(ctx, tmp_path, monkeypatch):
    """TrainingRun rolls the topologies out on helper contexts from one host thread each; the buffer (Philox streams
    keyed by seed / episode / topology / slot) and the greedy evaluation are those of the sequential run"""
    from enero_b200.train import TrainingRun
    specs = ((8, 11, 1003), (9, 12, 1004), (7, 10, 1005))
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("ENERO_ROLLOUT_THREADS", mode)
        run = TrainingRun.synthetic(ctx, specs, n_train_tms=6, n_eval_tms=3, folder=os.path.join(str(tmp_path), "data" + mode),
                                    eval_episodes=3)
        try:
            buf = run.collect()
            ev = run.evaluate()
        finally:
            run.close()
        out[mode] = (buf, ev)
    a, b = out["1"][0], out["0"][0]
    assert len(a["tensors"]) == len(b["tensors"]) > 0
    for x, y in zip(a["tensors"], b["tensors"]):
        assert x["sd"] == y["sd"] and x["bw"] == y["bw"] and np.array_equal(x["loads"], y["loads"]) and x["old_prob"] == y["old_prob"]
    assert a["actions"] == b["actions"] and a["rewards"] == b["rewards"] and a["masks"] == b["masks"]
    assert np.array_equal(np.asarray(a["values"]), np.asarray(b["values"]))
    for x, y in zip(out["1"][1], out["0"][1]):
        assert np.array_equal(x, y)
