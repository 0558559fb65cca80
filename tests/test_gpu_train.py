"""GPU parity of the PPO update (kernel family 5: backward kernels, PPO loss head, global-norm
clip + Adam, GAE) against the torch-CPU restatement in oracle/ppo_torch.py.

Tolerances (fp32, written here as north_star asks).  A gradient entry is a sum over thousands of link-state
rows of products of O(1) activations with cancellation, so a per-element relative error is meaningless for the
entries that cancel to ~0; errors are measured against the largest magnitude of the flat gradient:
        |got - want| <= tol * max|want|.
The fp64 arbiter (the same torch graph in float64, oracle/ppo_torch.to_params(dtype=float64)) shows where fp32
itself stands: torch's own fp32 gradients are up to 1.3e-5 * max|g| away from the fp64 ones
(tools/train_parity.py, profiles/r02_train_parity.json), the CUDA path 1.4e-6 ... 1.45e-5 in either math mode.
So the test holds the CUDA gradients to the arbiter at  max(1e-5, 2 x torch-fp32's own distance to it)
and to the fp32 restatement at 2e-5 (two fp32 roundings apart); losses and the clip scale at 1e-5.
"""
import numpy as np
import pytest

from conftest import case_graph_file, load_case
from oracle import enero_oracle as orc
from oracle import ppo_torch as ppo_o

pytestmark = pytest.mark.gpu

GRAD_TOL_ARBITER = 1e-5     # vs fp64, unless torch fp32 itself is further away (then 2x its distance)
GRAD_TOL_FP32 = 2e-5        # vs the fp32 restatement
RTOL = 1e-5
BETA = 0.01


def _critic_weights(weights):
    rng = np.random.default_rng(7)
    return {k: (weights[k] + rng.normal(0, 0.05, weights[k].shape)).astype(np.float32) for k in orc.WEIGHT_NAMES}


def _rollout(name, weights, n_samples, seed, tmp):
    """n_samples decisions with sampled actions on the oracle env; returns (gpu topology, oracle samples,
    gpu sample records)."""
    from enero_b200.topology import Topology
    case = load_case(name)
    gf = case_graph_file(case, tmp)
    topo, ot = Topology(gf, K=100), orc.Topology(gf, K=100)
    rng = np.random.default_rng(seed)
    tm_ids = sorted(case["tms"].keys())
    o_samples, g_samples = [], []
    ti = 0
    while len(o_samples) < n_samples:
        tm = np.array(case["tms"][tm_ids[ti % len(tm_ids)]])
        ti += 1
        env = orc.Env16(ot)
        _, s, d = env.reset(tm)
        while len(o_samples) < n_samples:
            probs, batch, _ = orc.pred_action_distrib(weights, env, s, d)
            a = int(rng.choice(len(probs), p=probs.astype(np.float64) / probs.astype(np.float64).sum()))
            onehot = np.zeros(len(probs), np.float32)
            onehot[a] = 1.0
            # a perturbed "old" policy so that ratios land on both sides of the clip interval
            oldp = probs * np.exp(rng.normal(0, 0.15, len(probs))).astype(np.float32)
            oldp = (oldp / oldp.sum()).astype(np.float32)
            adv = np.float32(rng.normal())
            ret = np.float32(rng.normal())
            smp = dict(batch)
            smp.update(link_state_critic=env.critic_features(), first_critic=ot.first, second_critic=ot.second,
                       num_edges_critic=ot.E, old_act=onehot, old_policy_probs=oldp, advantage=adv)
            smp["return"] = ret
            o_samples.append(smp)
            g = {"topology": topo, "loads": env.loads.copy(), "sd": (s, d), "bw": float(env.tm[s, d]),
                 "action": a, "old_prob": np.float32(oldp[a]), "advantage": adv, "return": ret}
            g_samples.append(g)
            _, done, _, _, s, d, _, _, _ = env.step(a)
            if done:
                break
    return topo, o_samples, g_samples


@pytest.fixture(scope="module")
def agent(weights):
    from enero_b200.ppo import PPOActorCritic
    from enero_b200.runtime import Context
    ctx = Context(0)
    ag = PPOActorCritic(context=ctx, entropy_beta=BETA)
    ag.actor.set_weights([weights[k] for k in orc.WEIGHT_NAMES])
    cw = _critic_weights(weights)
    ag.critic.set_weights([cw[k] for k in orc.WEIGHT_NAMES])
    yield ag, cw
    ctx.close()


def _dist(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


def _assert_grad(got, fp32, fp64, what):
    scale = float(np.abs(fp64).max())
    tol = max(GRAD_TOL_ARBITER, 2.0 * _dist(fp32, fp64) / scale)
    assert _dist(got, fp64) <= tol * scale, "%s vs fp64 arbiter: %.3e > %.2e * %.3e" % (what, _dist(got, fp64), tol, scale)
    assert _dist(got, fp32) <= GRAD_TOL_FP32 * scale, "%s vs torch fp32: %.3e > %.1e * %.3e" % (what, _dist(got, fp32), GRAD_TOL_FP32, scale)


@pytest.mark.parametrize("names,n", [(("tri9",), 6), (("tiny12",), 9), (("eli20", "zoo30"), 7)])
def test_minibatch_gradients(agent, weights, names, n, tmp_path_factory):
    """_train_step_combined's tape.gradient over one minibatch (train script :293-317), including a
    minibatch that mixes topologies (the 835-sample buffer holds three)."""
    ag, cw = agent
    o_all, g_all = [], []
    for i, name in enumerate(names):
        _, o, g = _rollout(name, weights, n, 11 + i, tmp_path_factory.mktemp("r" + name))
        o_all += o
        g_all += g
    import torch
    ag.memory = g_all
    wa, wc = ppo_o.to_params(weights), ppo_o.to_params(cw)
    oa, oc, al, cl, ent = ppo_o.gradients(wa, wc, o_all, BETA)
    ta, tc, *_ = ppo_o.gradients(ppo_o.to_params(weights, torch.float64), ppo_o.to_params(cw, torch.float64), o_all, BETA)
    m = len(g_all)
    before = ag.context.math
    try:
        for mode in ("fast", "accurate"):
            ag.context.math = mode
            ga, gc, sums = ag.gradients(np.arange(len(g_all)))
            np.testing.assert_allclose(sums[1] / m, ent, rtol=RTOL)
            np.testing.assert_allclose(sums[0] / m - BETA * sums[1] / m, al, rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(sums[2] / m, cl, rtol=RTOL)
            _assert_grad(ga, oa, ta, mode + " actor gradient")
            _assert_grad(gc, oc, tc, mode + " critic gradient")
            # per-tensor check so that a small tensor cannot hide behind a large one
            o = 0
            for k in orc.WEIGHT_NAMES:
                sz = weights[k].size
                if np.abs(ta[o:o + sz]).max() > 1e-2 * np.abs(ta).max():
                    _assert_grad(ga[o:o + sz], oa[o:o + sz], ta[o:o + sz], mode + " actor " + k)
                if np.abs(tc[o:o + sz]).max() > 1e-2 * np.abs(tc).max():
                    _assert_grad(gc[o:o + sz], oc[o:o + sz], tc[o:o + sz], mode + " critic " + k)
                o += sz
    finally:
        ag.context.math = before
    ag.memory = []


def _oracle_steps(weights, cw, o, batches, lr, tdtype, ndtype):
    """three _train_step_combined calls restated: torch gradients in `tdtype`, clip + Keras Adam in `ndtype`"""
    import torch
    wa, wc = ppo_o.to_params(weights, tdtype), ppo_o.to_params(cw, tdtype)
    fa = ppo_o.flat([weights[k] for k in orc.WEIGHT_NAMES], ndtype)
    fc = ppo_o.flat([cw[k] for k in orc.WEIGHT_NAMES], ndtype)
    slots = [np.zeros_like(fa) for _ in range(4)]       # ma, va, mc, vc
    outs = []
    for step, inds in enumerate(batches, start=1):
        oa, oc, al, cl, ent = ppo_o.gradients(wa, wc, [o[i] for i in inds], BETA)
        (ca, cc), norm = ppo_o.clip_by_global_norm([oa, oc], 0.5, ndtype)
        outs.append((al, cl, ent, float(norm)))
        fa, slots[0], slots[1] = ppo_o.adam_step(fa, slots[0], slots[1], ca, lr, step, dtype=ndtype)
        fc, slots[2], slots[3] = ppo_o.adam_step(fc, slots[2], slots[3], cc, lr, step, dtype=ndtype)
        with torch.no_grad():
            for w, f in ((wa, fa), (wc, fc)):
                off = 0
                for k in orc.WEIGHT_NAMES:
                    sz = w[k].numel()
                    w[k].copy_(torch.tensor(f[off:off + sz], dtype=tdtype).reshape(w[k].shape))
                    off += sz
    return fa, fc, slots, outs


def test_train_steps_match_oracle(agent, weights, tmp_path_factory):
    """Three consecutive _train_step_combined calls (gradient, clip_by_global_norm(0.5), Keras Adam) against the
    same three steps restated in fp32 and, as the arbiter, in fp64.  An Adam update is lr_t * m / (sqrt(v) + eps):
    for entries whose gradient is far above eps it is ~lr * sign(g) and insensitive to rounding; for entries with
    |g| ~ eps = 1e-5 it is g / eps, i.e. it amplifies the 1e-5-of-max|g| gradient rounding by max|g| / eps ~ 1e3.
    Tolerances therefore: first moments (linear in g) 2e-5 of their largest entry against the fp32 restatement; the
    weight UPDATES within max(1e-5, 2 x the fp32 restatement's own distance to the fp64 arbiter) of the arbiter,
    measured against the largest update."""
    import torch
    ag, cw = agent
    _, o, g = _rollout("eli20", weights, 12, 5, tmp_path_factory.mktemp("steps"))
    ag.memory = g
    lr = float(ag.optimizer.learning_rate)
    batches = ([0, 1, 2, 3], [4, 5, 6, 7, 8], [9, 10, 11, 0])
    fa, fc, slots, outs = _oracle_steps(weights, cw, o, batches, lr, torch.float32, np.float32)
    ta, tc, tslots, _ = _oracle_steps(weights, cw, o, batches, lr, torch.float64, np.float64)
    for inds, (al, cl, ent, norm) in zip(batches, outs):
        got = ag._train_step_combined(np.array(inds))
        np.testing.assert_allclose(ag.last_grad_norm, norm, rtol=2e-5)
        np.testing.assert_allclose(got, (al, cl, ent), rtol=2e-5, atol=1e-6)
    ga = np.concatenate([w.reshape(-1) for w in ag.context.download_weights(0)])
    gc = np.concatenate([w.reshape(-1) for w in ag.context.download_weights(1)])
    fa0 = ppo_o.flat([weights[k] for k in orc.WEIGHT_NAMES], np.float64)
    fc0 = ppo_o.flat([cw[k] for k in orc.WEIGHT_NAMES], np.float64)
    for got_w, f32, f64, w0, what in ((ga, fa, ta, fa0, "actor"), (gc, fc, tc, fc0, "critic")):
        upd = float(np.abs(f64 - w0).max())
        assert upd > 0.5 * lr
        tol = max(1e-5, 2.0 * _dist(f32, f64) / upd)
        # the weights themselves carry an fp32 representation error of 6e-8 * |w| on top of the update error
        floor = float(np.abs(w0).max()) * 1.2e-7
        assert _dist(got_w, f64) <= tol * upd + floor, "%s updates: %.3e > %.2e * %.3e" % (what, _dist(got_w, f64), tol, upd)
    m, v = ag.download_slots(0)
    gm = np.concatenate([x.reshape(-1) for x in m])
    gv = np.concatenate([x.reshape(-1) for x in v])
    assert _dist(gm, slots[0]) <= 3e-5 * np.abs(slots[0]).max()
    assert _dist(gv, slots[1]) <= 1e-4 * np.abs(slots[1]).max()
    # restore the module-scoped agent for other tests
    ag.actor.set_weights([weights[k] for k in orc.WEIGHT_NAMES])
    ag.critic.set_weights([cw[k] for k in orc.WEIGHT_NAMES])
    ag.memory = []


def test_gae(agent):
    """get_advantages (train script :367-377) in fp64."""
    from enero_b200.ppo import get_advantages
    ag, _ = agent
    rng = np.random.default_rng(3)
    n = 835
    values = [np.float32(v) for v in rng.normal(0, 1, n + 1)]
    masks = [bool(b) for b in (rng.random(n) > 0.02)]
    rewards = [np.float64(np.around(r, 2)) for r in rng.normal(0, 0.5, n)]
    want_ret, want_adv = orc.get_advantages(values, masks, rewards)
    ret, adv = get_advantages(values, masks, rewards, context=ag.context)
    np.testing.assert_allclose(ret, np.array(want_ret, dtype=np.float64), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(adv, want_adv, rtol=1e-10, atol=1e-12)


def test_accumulate_rejects_bad_action(agent, weights, tmp_path_factory):
    from enero_b200._lib import EneroError
    ag, _ = agent
    _, _, g = _rollout("tri9", weights, 2, 1, tmp_path_factory.mktemp("bad"))
    g[0]["action"] = 999
    ag.memory = g
    with pytest.raises(EneroError):
        ag.gradients(np.arange(2))
    ag.memory = []


def _run_two_ranks(tmp_path, backend):
    import os
    import socket
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "helpers"))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank) if backend == "nccl" else "0",
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(here, "helpers", "ppo_worker.py"), str(tmp_path), backend],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "rank1.npz"))
    assert np.array_equal(r0["wa"], r1["wa"]) and np.array_equal(r0["wc"], r1["wc"])
    import ppo_worker
    ag = ppo_worker.build_agent_and_samples(str(tmp_path))
    wa0 = np.concatenate([w.reshape(-1) for w in ag.actor.get_weights()])
    wa, wc, losses = ppo_worker.run_steps(ag)
    ag.context.close()
    upd = np.abs(wa - wa0).max()
    assert upd > 0
    assert np.abs(r0["wa"] - wa).max() <= 1e-3 * upd
    np.testing.assert_allclose(r0["losses"], losses, rtol=1e-4, atol=1e-6)


def test_sharded_step_matches_single(tmp_path):
    """the minibatch split over 2 ranks + one all-reduce of the gradient buffer gives the
    weights of the single-process step (sum order differs: 1e-6 of the update scale) and identical
    weights on both ranks.  gloo, both ranks on one GPU: runs on any box."""
    _run_two_ranks(tmp_path, "gloo")


def test_sharded_step_nccl(tmp_path):
    """the same step with one GPU per rank and the all-reduce carried by NCCL on the library's stream (what a
    torchrun training job runs); needs two GPUs"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run_two_ranks(tmp_path, "nccl")
