"""Multi-GPU plumbing: environments are independent, so the job shards by global env id (one
process per GPU, contiguous slices) and the only collective on the path is the all-reduce (sum)
of the 12-element episode-statistics vector."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

INT_KEYS = ("episodes", "completed", "truncated", "steps", "successes", "elements", "sum_n", "sum_n_success",
            "ring_items", "sum_n_ring")
FLOAT_KEYS = ("sum_return", "sum_length")


def shard_range(global_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [start, start+count) of global env ids owned by `rank` (remainder spread
    over the first ranks).  Philox subsequences are keyed by the global id, so results do not
    depend on `world`."""
    base, rem = divmod(int(global_envs), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def allreduce_stats(stats: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum the episode statistics of every rank (NCCL over NVLink on GPUs, gloo on CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    ti = torch.tensor([int(stats[k]) for k in INT_KEYS], dtype=torch.int64, device=dev)
    tf = torch.tensor([float(stats[k]) for k in FLOAT_KEYS], dtype=torch.float64, device=dev)
    dist.all_reduce(ti, op=dist.ReduceOp.SUM)
    dist.all_reduce(tf, op=dist.ReduceOp.SUM)
    out = {k: int(v) for k, v in zip(INT_KEYS, ti.tolist())}
    out.update({k: float(v) for k, v in zip(FLOAT_KEYS, tf.tolist())})
    return out


def allreduce_stats_device(stats_dev: torch.Tensor) -> torch.Tensor:
    """In-place all-reduce (sum) of a device-resident ``mg_episode_stats`` (the 12 x int64 tensor written by
    ``BatchedBoudaryEnv.stats_async``: 10 int64 counters followed by the bit patterns of 2 float64 sums), enqueued
    behind the step kernels without any host synchronisation -- the one collective of the data-parallel job
    (every 64 steps)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats_dev
    n = len(INT_KEYS)
    dist.all_reduce(stats_dev[:n], op=dist.ReduceOp.SUM)
    dist.all_reduce(stats_dev[n:].view(torch.float64), op=dist.ReduceOp.SUM)
    return stats_dev


def stats_from_tensor(stats_dev: torch.Tensor) -> Dict[str, float]:
    """Host dict of a (possibly all-reduced) device statistics tensor; synchronises."""
    n = len(INT_KEYS)
    host = stats_dev.cpu()
    out = {k: int(v) for k, v in zip(INT_KEYS, host[:n].tolist())}
    out.update({k: float(v) for k, v in zip(FLOAT_KEYS, host[n:].view(torch.float64).tolist())})
    return out
