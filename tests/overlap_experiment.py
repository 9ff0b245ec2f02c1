"""Experiment (run by hand on a GPU box): does phase A of one half-batch overlap phase B/C of the other when the
batch is split over two handles on two streams?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
GEN = dict(min_verts=64, max_verts=512)
def run(parts, total=65536, steps=500, burn=1500):
    n = total // parts
    envs = [BatchedBoudaryEnv(None, num_envs=n, random_polygons=GEN, seed=2026, env_id_offset=i * n) for i in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    for e, s in zip(envs, streams):
        with torch.cuda.stream(s):
            e.reset()
    def step(t):
        for e, s in zip(envs, streams):
            with torch.cuda.stream(s):
                e.step(e.sample_actions(2026, t))
    for t in range(burn): step(t)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(burn, burn + steps): step(t)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{parts} handle(s) x {n} envs: {total*steps/dt:.4g} env-steps/s ({dt/steps*1e6:.1f} us per step of {total})", flush=True)
    for e in envs: e.close()
for p in (1, 2, 4):
    run(p)
