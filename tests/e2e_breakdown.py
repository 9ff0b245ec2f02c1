"""Profiling aid (run by hand on a GPU box): where the host-facing step (mg_step_host) spends its time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
N = 65536
env = BatchedBoudaryEnv(None, num_envs=N, random_polygons=dict(min_verts=64, max_verts=512), seed=2026)
env.reset()
for t in range(1500):
    env.step(env.sample_actions(1, t))
torch.cuda.synchronize()
rng = np.random.default_rng(0)
lo, hi = np.array([-1, -1.5, 0], np.float32), np.array([1, 1.5, 1.5], np.float32)
acts = [torch.from_numpy(rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)) for _ in range(4)]
pin = torch.empty((N, 3), dtype=torch.float32).pin_memory()
out = dict(obs=torch.empty((N, 18)).pin_memory(), reward=torch.empty(N, dtype=torch.float64).pin_memory(),
           terminated=torch.empty(N, dtype=torch.uint8).pin_memory(), truncated=torch.empty(N, dtype=torch.uint8).pin_memory(),
           terminal_obs=torch.empty((N, 18)).pin_memory(), n_elements=torch.empty(N, dtype=torch.int32).pin_memory())
env.set_host_delta(True)
K = 200
def run(name, f):
    for k in range(5): f(k)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(K): f(k)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"{name:48s} {dt*1e6:8.1f} us/step  {N/dt:.3e} env-steps/s", flush=True)
run("pinned act copy only", lambda k: pin.copy_(acts[k % 4]))
run("step_host (delta)", lambda k: env.step_host(pin, out))
run("copy + step_host + read", lambda k: (pin.copy_(acts[k % 4]), env.step_host(pin, out), float(out["reward"][0])))
dact = torch.empty((N, 3), device=env.device)
run("device step + sync", lambda k: (env.step(dact), torch.cuda.synchronize()))
run("H2D act + device step + sync", lambda k: (dact.copy_(pin, non_blocking=True), env.step(dact), torch.cuda.synchronize()))
rew_h = torch.empty(N, dtype=torch.float64).pin_memory()
run("H2D + step + D2H reward + sync", lambda k: (dact.copy_(pin, non_blocking=True), env.step(dact), rew_h.copy_(env.reward, non_blocking=True), torch.cuda.synchronize()))
env.set_host_delta(False)
run("step_host (full copies)", lambda k: env.step_host(pin, out))
