"""One rank of the sharded-PPO-step tests (tests/test_gpu_train.py::test_sharded_step_matches_single /
::test_sharded_step_nccl).  argv[2] = backend: "gloo" -- both ranks on cuda:0 (a one-GPU box), the all-reduce of the
CUDA gradient buffer staged by gloo; "nccl" -- one GPU per rank (LOCAL_RANK), the all-reduce enqueued on the
library's stream exactly as a torchrun training job does."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_agent_and_samples(tmp, process_group=None, device=0):
    from conftest import golden_weights
    from enero_b200.ppo import PPOActorCritic
    from enero_b200.runtime import Context
    from oracle import enero_oracle as orc
    from test_gpu_train import BETA, _critic_weights, _rollout
    weights = golden_weights()
    ctx = Context(device)
    ag = PPOActorCritic(context=ctx, entropy_beta=BETA, process_group=process_group)
    ag.actor.set_weights([weights[k] for k in orc.WEIGHT_NAMES])
    cw = _critic_weights(weights)
    ag.critic.set_weights([cw[k] for k in orc.WEIGHT_NAMES])
    g = []
    for i, name in enumerate(("eli20", "tiny12")):
        d = os.path.join(tmp, "w%d_%s" % (os.getpid(), name))
        os.makedirs(d, exist_ok=True)
        g += _rollout(name, weights, 8, 21 + i, d)[2]
    ag.memory = g
    return ag


def run_steps(ag):
    rng = np.random.RandomState(3)
    inds = np.arange(len(ag.memory))
    losses = []
    for _ in range(2):
        rng.shuffle(inds)
        for start in range(0, len(inds), 6):
            losses.append(ag._train_step_combined(inds[start:start + 6].copy()))
    wa = np.concatenate([w.reshape(-1) for w in ag.context.download_weights(0)])
    wc = np.concatenate([w.reshape(-1) for w in ag.context.download_weights(1)])
    return wa, wc, np.array(losses)


def main(out_dir, backend="gloo"):
    import torch.distributed as dist
    from enero_b200 import parallel as par
    rank, world, local = par.init_process_group(backend)
    ag = build_agent_and_samples(out_dir, process_group=dist.group.WORLD, device=local if backend == "nccl" else 0)
    wa, wc, losses = run_steps(ag)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), wa=wa, wc=wc, losses=losses)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "gloo")
