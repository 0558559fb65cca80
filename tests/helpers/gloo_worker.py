"""world_size-2 worker for tests/test_parallel_gloo.py (CPU, gloo): exercises the host-side
sharding logic of enero_b200.parallel exactly as a 2-GPU job would use it."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(out_dir):
    import torch
    import torch.distributed as dist
    from enero_b200 import parallel as par

    rank, world, _ = par.init_process_group("gloo")
    assert world == 2 and dist.get_backend() == "gloo"
    res = {"rank": rank}

    # ---- episodes of an eval sweep: contiguous shard, records gathered on rank 0 in unit order
    n_units = 11
    b, e = par.shard_range(n_units, rank, world)
    records = [{"tm_id": i, "value": float(i) * 0.5} for i in range(b, e)]
    allrec = par.gather_records(records)
    res["gathered"] = allrec

    # ---- one PPO minibatch: strided shard of a shuffled index vector; the per-rank partial
    # "gradients" (here: a deterministic function of the sample index, scaled by the GLOBAL 1/m)
    # summed by one all-reduce must equal the single-process total
    rng = np.random.RandomState(9)
    inds = np.arange(55)
    rng.shuffle(inds)
    mine = par.shard_indices(inds, rank, world)
    res["mine"] = [int(i) for i in mine]
    m = len(inds)
    g = np.zeros(8412, dtype=np.float32)
    for i in mine:
        g += np.float32(1.0 / m) * np.sin(np.arange(8412, dtype=np.float32) * np.float32(0.001) * np.float32(i + 1))
    t = torch.from_numpy(g)
    par.allreduce_sum_(t)
    res["grad_sum"] = float(t.double().sum())
    res["grad_head"] = [float(v) for v in t[:4]]

    # ---- timings are reported as the slowest rank's
    res["max_ms"] = par.max_over_ranks(10.0 + rank)
    dist.barrier()
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as fp:
        json.dump(res, fp)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
