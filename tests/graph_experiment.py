#!/usr/bin/env python
"""Hand-run experiment (gpurun): cost of one env step when it is replayed from a CUDA graph instead of being launched
eagerly, under the step's scheduling options (programmatic dependent launches, resets on the side stream).

    python tests/graph_experiment.py [--envs 65536] [--steps 300] [--actor]
Prints one line per variant: mode, options, ms per step."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=304)
    ap.add_argument("--burn", type=int, default=1200)
    ap.add_argument("--actor", action="store_true")
    ap.add_argument("--dot", default=None, help="write the captured graph's topology here (first graph variant)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    env = BatchedBoudaryEnv(None, num_envs=args.envs, device=dev, random_polygons=dict(min_verts=64, max_verts=512), seed=2026)
    env.reset()
    actor = None
    if args.actor:
        from sac_rollout import Actor
        torch.manual_seed(0)
        actor = Actor().to(dev)
    act_buf = torch.empty(args.envs, 3, dtype=torch.float32, device=dev)
    from reinforcementlearning4meshgeneration_b200 import ACTION_HIGH, ACTION_LOW
    low = torch.from_numpy(ACTION_LOW.copy()).to(dev)
    span = torch.from_numpy((ACTION_HIGH - ACTION_LOW).copy()).to(dev)

    def one_step():
        if actor is not None:
            env.step(actor(env.obs, True))
        else:
            # uniform actions from torch's graph-safe generator (a replayed mg_sample_actions would repeat its step index)
            env.step(torch.addcmul(low, torch.rand_like(act_buf), span))

    for _ in range(args.burn):
        one_step()
    torch.cuda.synchronize()
    base = env.snapshot()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    first = True
    for opts in ({}, {"pdl": 0}, {"reset_side": 0}, {"pdl": 0, "reset_side": 0}):
        for k in ("pdl", "reset_side"):
            env.set_option(k, opts.get(k, 1))
        for per_graph in (0, 1, 8):
            env.restore(base)
            torch.cuda.synchronize()
            if per_graph == 0:
                ms = timed(one_step, args.steps) / args.steps
            else:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        one_step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                if first and args.dot:
                    g.enable_debug_mode()
                with torch.cuda.graph(g):
                    for _ in range(per_graph):
                        one_step()
                if first and args.dot:
                    g.debug_dump(args.dot)
                    first = False
                env.restore(base)
                for _ in range(8):
                    g.replay()
                ms = timed(g.replay, args.steps // per_graph) / (args.steps // per_graph * per_graph)
                del g
            print(f"{'actor' if actor is not None else 'synthetic'} steps_per_graph={per_graph} options={opts} ms_per_step={ms:.4f} "
                  f"env_steps_per_s={args.envs / ms * 1e3:.4g}", flush=True)


if __name__ == "__main__":
    main()
