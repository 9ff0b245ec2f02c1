"""Profiling aid (run by hand on a GPU box): per-kernel device times of the host-facing step (mg_step_host with pinned
buffers: the screen kernel reads the actions and writes most results over PCIe) next to the device-resident step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
N = 65536
env = BatchedBoudaryEnv(None, num_envs=N, random_polygons=bench.GEN, seed=bench.SEED)
env.reset()
for t in range(1500):
    env.step(env.sample_actions(1, t))
torch.cuda.synchronize()
rng = np.random.default_rng(0)
lo, hi = np.array([-1, -1.5, 0], np.float32), np.array([1, 1.5, 1.5], np.float32)
acts = [torch.from_numpy(rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)).pin_memory() for _ in range(4)]
out = dict(obs=torch.empty((N, 18)).pin_memory(), reward=torch.empty(N, dtype=torch.float64).pin_memory(),
           terminated=torch.empty(N, dtype=torch.uint8).pin_memory(), truncated=torch.empty(N, dtype=torch.uint8).pin_memory(),
           terminal_obs=torch.empty((N, 18)).pin_memory(), n_elements=torch.empty(N, dtype=torch.int32).pin_memory())
for k in range(5):
    env.step_host(acts[k % 4], out)
env.set_kernel_timing(True)
for k in range(100):
    env.step_host(acts[k % 4], out)
print("step_host, pinned buffers :", {k: round(v * 1e3, 1) if k != "steps" else v for k, v in env.kernel_times().items()})
dact = [a.to(env.device) for a in acts]
for k in range(100):
    env.step(dact[k % 4])
print("device-resident step      :", {k: round(v * 1e3, 1) if k != "steps" else v for k, v in env.kernel_times().items()})
env.set_kernel_timing(False)
