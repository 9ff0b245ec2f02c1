#!/usr/bin/env python
"""Hand-run experiment (gpurun): host-buffer steps of one env of N envs (mg_step_host) against two envs of N / 2 envs
whose steps are pipelined with mg_step_host_begin / _end -- the host reads one half's results and enqueues its next
step while the other half runs on the GPU.

    python tests/e2e_pipeline.py [--envs 65536] [--steps 200]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv  # noqa: E402

GEN = dict(min_verts=64, max_verts=512)


def pinned_out(n):
    return dict(obs=torch.zeros((n, 18), dtype=torch.float32).pin_memory(), reward=torch.zeros(n, dtype=torch.float64).pin_memory(),
                terminated=torch.zeros(n, dtype=torch.uint8).pin_memory(), truncated=torch.zeros(n, dtype=torch.uint8).pin_memory(),
                terminal_obs=torch.zeros((n, 18), dtype=torch.float32).pin_memory(), n_elements=torch.zeros(n, dtype=torch.int32).pin_memory())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--burn", type=int, default=1500)
    ap.add_argument("--parts", type=int, default=2)
    args = ap.parse_args()
    N, P = args.envs, args.parts
    rng = np.random.default_rng(0)
    lo, hi = np.array([-1, -1.5, 0], np.float32), np.array([1, 1.5, 1.5], np.float32)

    def make(n, offset):
        env = BatchedBoudaryEnv(None, num_envs=n, random_polygons=GEN, seed=2026, env_id_offset=offset)
        env.reset()
        for t in range(args.burn):
            env.step(env.sample_actions(2026, t))
        torch.cuda.synchronize()
        return env

    acts = [torch.from_numpy(rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)).pin_memory() for _ in range(4)]
    full, out = make(N, 0), pinned_out(N)
    view = out["reward"].numpy()
    for k in range(12):
        full.step_host(acts[k % 4], out)
    t0 = time.perf_counter()
    for k in range(args.steps):
        full.step_host(acts[k % 4], out)
        _ = float(view[0])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"one env of {N}: {N * args.steps / dt:.4g} env-steps/s, {dt / args.steps * 1e6:.1f} us per step", flush=True)
    full.close()

    n = N // P
    parts = [make(n, p * n) for p in range(P)]
    outs = [pinned_out(n) for _ in range(P)]
    views = [o["reward"].numpy() for o in outs]
    pacts = [[a[p * n:(p + 1) * n].clone().pin_memory() for a in acts] for p in range(P)]
    for k in range(12):
        for p in range(P):
            parts[p].step_host(pacts[p][k % 4], outs[p])
    t0 = time.perf_counter()
    for p in range(P):
        parts[p].step_host_begin(pacts[p][0], outs[p])
    for k in range(args.steps):
        for p in range(P):
            parts[p].step_host_end()
            _ = float(views[p][0])                           # the host reads this part's results (and would run its policy)
            parts[p].step_host_begin(pacts[p][(k + 1) % 4], outs[p])
    for p in range(P):
        parts[p].step_host_end()
    dt = time.perf_counter() - t0
    steps = args.steps + 1
    print(f"{P} envs of {n}, pipelined: {N * steps / dt:.4g} env-steps/s, {dt / steps * 1e6:.1f} us per step of all parts", flush=True)


if __name__ == "__main__":
    main()
