"""Gradient parity report (not a test): minibatch gradients of the CUDA PPO path in both math modes and of the
torch fp32 restatement, each against the torch fp64 arbiter (oracle/ppo_torch.py with dtype=float64).
    python tools/train_parity.py            # GPU box; writes gpurun_out/train_parity.json
Errors are max |g - g64| / max|g64| over a tensor ("norm-wise") and the element-wise maximum of
|g - g64| / max(|g64|, 1e-3 * max|g64|)."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                                          # noqa: E402
from conftest import golden_weights                                   # noqa: E402
from oracle import enero_oracle as orc                                # noqa: E402
from oracle import ppo_torch as ppo_o                                 # noqa: E402
from test_gpu_train import BETA, _critic_weights, _rollout            # noqa: E402


def errs(g, g64):
    g = g.astype(np.float64)
    scale = np.abs(g64).max()
    return {"normwise": float(np.abs(g - g64).max() / scale),
            "elementwise": float((np.abs(g - g64) / np.maximum(np.abs(g64), 1e-3 * scale)).max())}


def main():
    from enero_b200.ppo import PPOActorCritic
    from enero_b200.runtime import Context
    w = golden_weights()
    cw = _critic_weights(w)
    ctx = Context(0)
    ag = PPOActorCritic(context=ctx, entropy_beta=BETA)
    ag.actor.set_weights([w[k] for k in orc.WEIGHT_NAMES])
    ag.critic.set_weights([cw[k] for k in orc.WEIGHT_NAMES])
    out = {}
    for names, n in ((("tri9",), 6), (("tiny12",), 9), (("eli20", "zoo30"), 7), (("eli20",), 55)):
        o_all, g_all = [], []
        for i, name in enumerate(names):
            _, o, g = _rollout(name, w, n, 11 + i, tempfile.mkdtemp())
            o_all += o
            g_all += g
        a64, c64, *_ = ppo_o.gradients(ppo_o.to_params(w, torch.float64), ppo_o.to_params(cw, torch.float64), o_all, BETA)
        a32, c32, *_ = ppo_o.gradients(ppo_o.to_params(w), ppo_o.to_params(cw), o_all, BETA)
        rec = {"torch_fp32": {"actor": errs(a32, a64), "critic": errs(c32, c64)}}
        ag.memory = g_all
        for mode in ("fast", "accurate"):
            ctx.math = mode
            ga, gc, _ = ag.gradients(np.arange(len(g_all)))
            rec["gpu_" + mode] = {"actor": errs(ga, a64), "critic": errs(gc, c64),
                                  "actor_vs_torch32": errs(ga, a32.astype(np.float64)), "critic_vs_torch32": errs(gc, c32.astype(np.float64))}
        out["+".join(names) + "x%d" % n] = rec
        print(json.dumps({"+".join(names): rec}), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
