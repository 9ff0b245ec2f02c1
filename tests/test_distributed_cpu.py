"""CPU, world_size 2, gloo: env sharding and the episode-statistics all-reduce (the N>1 host logic)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reinforcementlearning4meshgeneration_b200.distributed import FLOAT_KEYS, INT_KEYS, allreduce_stats, shard_range


def test_shard_range_partitions_all_envs():
    for total, world in [(1048576, 8), (10, 4), (7, 8), (65536, 1)]:
        seen = []
        for r in range(world):
            s, c = shard_range(total, r, world)
            seen.extend(range(s, s + c))
        assert seen == list(range(total))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = {k: (rank + 1) * (i + 1) for i, k in enumerate(INT_KEYS)}
    stats.update({k: 0.5 * (rank + 1) for k in FLOAT_KEYS})
    out = allreduce_stats(stats, torch.device("cpu"))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_stats_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        for i, k in enumerate(INT_KEYS):
            assert res[r][k] == 3 * (i + 1)
        for k in FLOAT_KEYS:
            assert res[r][k] == 1.5
