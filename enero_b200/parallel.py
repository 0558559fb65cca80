"""One process per GPU: process-group setup and the (small) host logic of sharding ENERO's hot path.

* Inference / Local Search / link-failure sweeps: independent (topology, TM) episodes are split over
  the ranks with no data-path collective; rank 0 gathers the result records
  (eval_on_single_topology.py:92-94 fans the same units out as OS processes).
* PPO update: the samples of a minibatch are split over the ranks and the flat gradient buffer is
  summed by one all-reduce (enero_b200/ppo.py).
"""
from __future__ import annotations

import os

import numpy as np


def env_rank():
    """(rank, world, local_rank) from the torchrun environment (1 process = defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend=None):
    """NCCL when a CUDA device is visible, gloo otherwise; rendezvous from MASTER_ADDR/MASTER_PORT.
    Returns (rank, world, local_rank); a single process needs no group."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n, rank, world):
    """balanced contiguous split of n units: ranks < n % world get one extra -> (begin, end)"""
    base, extra = divmod(int(n), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_indices(inds, rank, world):
    """strided split of one (shuffled) minibatch; the union over ranks is the minibatch, once each"""
    return inds[rank::world]


def shard_by_cost(costs, world):
    """size-balanced assignment of episodes with unequal cost (cost ~ D*K'*(P+E)): longest
    processing time first onto the least loaded rank.  -> list of index lists, one per rank."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    out = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(int(i))
        load[r] += float(costs[i])
    return [sorted(x) for x in out]


def allreduce_sum_(tensor, group=None):
    """in-place sum over ranks (the PPO gradient buffer); no-op without a process group"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def max_over_ranks(value, device=None, group=None):
    """max of a python float over ranks (device timings are reported as the slowest rank's)"""
    import sys
    if "torch" not in sys.modules:                  # no torch, no process group: a single rank
        return float(value)
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_records(records, group=None):
    """rank 0 receives the concatenation of every rank's result records in rank order; others get None"""
    import sys
    if "torch" not in sys.modules:
        return list(records)
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(records)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = [None] * world if rank == 0 else None
    dist.gather_object(list(records), out, dst=0, group=group)
    if rank != 0:
        return None
    return [r for part in out for r in part]
