#!/usr/bin/env python
"""BASELINE config 5: SAC-style rollout loop with a torch actor consuming device-resident
observations -- obs never leave the GPU.  Actor = the reference's SAC MlpPolicy shape
(net_arch [128,128,128], ReLU; v2/src/mesh_rl/algorithms/sb3_algos.py:56-62): 18 -> 128 -> 128 ->
128 -> (3 mean, 3 log-std), tanh squash, affine map to the action box (E:78-80).  Random weights
(no learner here: the replay buffer / SAC update is SURVEY.md section 8f rank 2).

    python examples/sac_rollout.py --envs 65536 --steps 500 --graph
    torchrun --nproc-per-node 8 examples/sac_rollout.py --envs 65536 --graph   # env-sharded, stats all-reduced on the device
Prints one JSON line (env-steps/s with the policy in the loop, clocks) on rank 0.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning4meshgeneration_b200 import ACTION_HIGH, ACTION_LOW, BatchedBoudaryEnv  # noqa: E402


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
        self.register_buffer("tick", torch.zeros(1))          # forward-call counter of the deterministic exploration
        self.register_buffer("phase", torch.tensor([0.0, 2.1, 4.2]))

    @torch.no_grad()
    def forward(self, obs, stochastic=True):
        h = self.body(obs)
        mu, log_std = self.mu(h), self.log_std(h).clamp(-20, 2)
        if stochastic:
            eps = torch.randn_like(mu)
        else:      # reproducible exploration (tests): a fixed function of the observation and of the call counter
            eps = 1.5 * torch.sin(997.0 * obs.sum(dim=1, keepdim=True) + 0.7 * self.tick + self.phase)
            self.tick += 1
        a = torch.tanh(mu + log_std.exp() * eps)
        return (self.low + (a + 1) * 0.5 * (self.high - self.low)).contiguous()


def capture_step(env, actor, stochastic=True):
    """One CUDA graph of policy forward + env step (mg_step keeps no host state: capturable).  The env state is
    the same before and after the call."""
    snap = env.snapshot()                                    # the warm-up below must not advance the rollout
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            env.step(actor(env.obs, stochastic))
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    env.restore(snap)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        act = actor(env.obs, stochastic)                     # env.obs is the env's persistent output buffer
        env.step(act)
    env.restore(snap)
    torch.cuda.synchronize()
    return g, act


def rollout(env, actor, steps, graph=False, stochastic=True, stats_every=64, record=None):
    """``steps`` transitions of ``env`` under ``actor``; ``graph``: True (capture here) or a (graph, action buffer) pair
    from ``capture_step`` -- the policy forward and the env step are then replayed from one CUDA graph.  ``record``
    (optional list) receives a clone of (action, obs, reward, terminated | truncated) per step.  Returns the all-reduced
    device statistics tensor."""
    from reinforcementlearning4meshgeneration_b200.distributed import allreduce_stats_device
    obs = env.obs
    g = None
    if graph:
        g, act = capture_step(env, actor, stochastic) if graph is True else graph
    stats = None
    for t in range(steps):
        if g is not None:
            g.replay()
        else:
            act = actor(obs, stochastic)
            obs = env.step(act).obs
        if record is not None:
            record.append((act.clone(), env.obs.clone(), env.reward.clone(), (env.terminated | env.truncated).clone()))
        if (t + 1) % stats_every == 0:
            stats = allreduce_stats_device(env.stats_async())    # the job's one collective: enqueued, never waited on
    return stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=300, help="untimed steps that de-synchronise the episodes")
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
    env.reset()
    import bench                                          # clock sampler of the repo's benchmark
    from reinforcementlearning4meshgeneration_b200.distributed import stats_from_tensor
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    rollout(env, actor, args.warmup, graph=False, stats_every=args.stats_every)
    graph = capture_step(env, actor) if args.graph else False      # set-up, not a step: outside the timed region
    if args.graph:
        rollout(env, actor, 8, graph=graph, stats_every=args.stats_every)
    env.stats(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = rollout(env, actor, args.steps, graph=graph, stats_every=args.stats_every)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        s = stats_from_tensor(stats) if stats is not None else {}
        line = {"metric": "env_steps_per_sec", "value": world * args.envs * args.steps / (float(ms[0]) * 1e-3), "unit": "env-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms[0]) / args.steps,
                "higher_is_better": True, "scaling": "weak", "dtype": "f64 env / " + ("tf32" if args.tf32 else "f32") + " policy",
                "data": "synthetic",
                "config": {"workload": "c5: SAC-style rollout, actor MLP 18-128-128-128-(3+3) (random weights) consuming device-resident "
                                       "observations; random star polygons 64..512 vertices, in-kernel auto-reset",
                           "envs_per_gpu": args.envs, "global_envs": args.envs * world, "cuda_graph": bool(args.graph),
                           "stats_allreduce_every": args.stats_every, "parallelism": f"env-sharded x{world}, identical policy replicas"},
                "clocks": clocks,
                "episodes": s.get("episodes"), "mean_return": (s["sum_return"] / s["episodes"]) if s.get("episodes") else None,
                "mean_elements": (s["elements"] / s["episodes"]) if s.get("episodes") else None}
        import json
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
