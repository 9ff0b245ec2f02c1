#!/usr/bin/env python
"""BASELINE config 5: SAC-style rollout loop with a torch actor consuming device-resident
observations -- obs never leave the GPU.  Actor = the reference's SAC MlpPolicy shape
(net_arch [128,128,128], ReLU; v2/src/mesh_rl/algorithms/sb3_algos.py:56-62): 18 -> 128 -> 128 ->
128 -> (3 mean, 3 log-std), tanh squash, affine map to the action box (E:78-80).  Random weights
(no learner here: the replay buffer / SAC update is SURVEY.md section 8f rank 2).

    python examples/sac_rollout.py --envs 65536 --steps 500
    torchrun --nproc-per-node 8 examples/sac_rollout.py --envs 131072   # env-sharded, stats all-reduced
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning4meshgeneration_b200 import ACTION_HIGH, ACTION_LOW, BatchedBoudaryEnv  # noqa: E402
from reinforcementlearning4meshgeneration_b200.distributed import allreduce_stats  # noqa: E402


class Actor(torch.nn.Module):
    def __init__(self, obs_dim=18, act_dim=3, hidden=(128, 128, 128)):
        super().__init__()
        layers, d = [], obs_dim
        for h in hidden:
            layers += [torch.nn.Linear(d, h), torch.nn.ReLU()]
            d = h
        self.body = torch.nn.Sequential(*layers)
        self.mu = torch.nn.Linear(d, act_dim)
        self.log_std = torch.nn.Linear(d, act_dim)
        self.register_buffer("low", torch.from_numpy(ACTION_LOW.copy()))
        self.register_buffer("high", torch.from_numpy(ACTION_HIGH.copy()))

    @torch.no_grad()
    def forward(self, obs):
        h = self.body(obs)
        mu, log_std = self.mu(h), self.log_std(h).clamp(-20, 2)
        a = torch.tanh(mu + log_std.exp() * torch.randn_like(mu))
        return (self.low + (a + 1) * 0.5 * (self.high - self.low)).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--stats-every", type=int, default=64)
    ap.add_argument("--graph", action="store_true", help="capture policy + env step in one CUDA graph (mg_step is capturable)")
    ap.add_argument("--tf32", action="store_true", help="TF32 matmuls for the policy (the env stays FP64)")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(0)                                  # identical policy replicas on every rank
    actor = Actor().to(dev)
    env = BatchedBoudaryEnv(None, num_envs=args.envs, device=dev, random_polygons=dict(min_verts=64, max_verts=512),
                            seed=2026, env_id_offset=rank * args.envs)
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
    obs = env.reset()
    graph = None
    if args.graph:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):                               # warm-up outside the capture
                env.step(actor(env.obs))
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            env.step(actor(env.obs))                         # env.obs is the env's persistent output buffer
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            r = env.step(actor(obs))
            obs = r.obs
        if (t + 1) % args.stats_every == 0:
            s = allreduce_stats(env.stats(reset=True), dev)
            if rank == 0 and s["episodes"]:
                print(f"step {t+1}: episodes {s['episodes']} completed {s['completed']} mean return "
                      f"{s['sum_return']/s['episodes']:.3f} mean elements {s['elements']/s['episodes']:.1f}")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"{world} GPU(s) x {args.envs} envs: {world*args.envs*args.steps/dt:.3e} env-steps/s with the MLP policy in the loop")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
