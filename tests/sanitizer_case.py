"""Small end-to-end case for compute-sanitizer (memcheck / racecheck) runs on the GPU box:
domain mode and random-polygon mode, a few hundred steps, auto-reset on."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv

z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "domains.npz"))
env = BatchedBoudaryEnv([z["half_wheel"], z["star"], z["boundary16"]], num_envs=48)
env.reset()
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    env.step(env.sample_actions(1, t))
print("domain mode", env.stats())
env2 = BatchedBoudaryEnv(None, num_envs=32, random_polygons=dict(min_verts=16, max_verts=64, min_coarse=4, max_coarse=8), seed=3)
env2.reset()
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    env2.step(env2.sample_actions(2, t))
torch.cuda.synchronize()
print("random mode", env2.stats())
