#!/usr/bin/env python
"""Benchmark of the batched BoudaryEnv hot path (env-steps/s, whole job).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--impl reference]

One "step" = one pass of the hot path over the whole batch: the synthetic policy kernel
(uniform actions in the action box) + the step kernel (every env advances
one transition, auto-reset included).  Workloads (BASELINE.json configs):
  c3 (default): random star polygons, 64..512 vertices, 65536 envs per GPU, in-kernel auto-reset
  c2          : paper domains d1/d2/d3 (120/196/272 vertices), 4096 envs per GPU
  c1          : BoudaryEnv(boundary()), 30 vertices, 4096 envs per GPU
Multi-GPU: launched by torchrun, one rank per GPU, envs sharded by global env id (weak scaling),
the only collective is the all-reduce of the 10-element episode-statistics vector.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
ENVS_PER_GPU = {"c3": 65536, "c2": 4096, "c1": 4096}
GEN = dict(min_verts=64, max_verts=512)
SEED = 2026


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def golden_domains():
    z = np.load(os.path.join(ROOT, "tests", "golden", "domains.npz"))
    return {k: z[k] for k in z.files if not k.startswith("area__")}


def workload_domains(workload):
    d = golden_domains()
    if workload == "c2":
        return [d["boundary16"], d["boundary15"], d["test1"]]
    if workload == "c1":
        return [d["boundary0"]]
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement of the reference env) on the host cores
# --------------------------------------------------------------------------------------------
def cpu_port_throughput(polys, seconds_budget=12.0, threads=None):
    """Steps/s of oracle/liboracle.so (kind "port") with one env per host thread (ctypes releases
    the GIL), uniform random actions, auto-reset, on the given sample of polygons."""
    from oracle.c_oracle import OracleEnv
    threads = threads or os.cpu_count() or 1
    envs = [OracleEnv(polys[i % len(polys)]) for i in range(threads)]
    # calibrate on one thread
    t0 = time.perf_counter()
    envs[0].run_random(1, 2000)
    per_step = (time.perf_counter() - t0) / 2000
    steps = max(2000, int(seconds_budget / per_step))
    done = [0] * threads

    def work(i):
        done[i] = envs[i].run_random(100 + i, steps)["steps"]

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, threads, steps, dt


def sample_polygons(workload, n_sample=32):
    """Polygons for the CPU arm: the fixed domains, or a sample of generated random polygons copied
    from the device generator (c3) -- falls back to the committed domains when no GPU is present."""
    doms = workload_domains(workload)
    if doms is not None:
        return doms, f"{workload}: {len(doms)} domain(s)"
    try:
        import torch
        if torch.cuda.is_available():
            from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
            env = BatchedBoudaryEnv(None, num_envs=n_sample, random_polygons=GEN, seed=SEED)
            env.reset()
            polys = [env.get_state(e)["xy"] for e in range(n_sample)]
            env.close()
            return polys, f"c3: {n_sample} random polygons copied from the device generator (seed {SEED})"
    except Exception:
        pass
    d = golden_domains()
    return [d["boundary16"], d["boundary15"], d["test1"], d["test3"]], "c3 stand-in: d1/d2/d3/test3 (no GPU for the generator)"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure
    Python and is not present on the GPU box, so this arm times its C restatement (oracle port)
    on all host cores; each step of this arm is one bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    polys, sample = sample_polygons(args.workload)
    vals = []
    total_steps = 0
    t_all = time.perf_counter()
    budget = max(1.0, min(8.0, 100.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        v, threads, steps, dt = cpu_port_throughput(polys, seconds_budget=budget)
        if i >= args.warmup:
            vals.append(v)
            total_steps += steps * threads
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(1, args.steps + args.warmup),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "note": "CPU arm: C port of the reference env (oracle/), all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
    from reinforcementlearning4meshgeneration_b200.distributed import allreduce_stats

    N = args.envs or ENVS_PER_GPU[args.workload]
    doms = workload_domains(args.workload)
    if doms is None:
        env = BatchedBoudaryEnv(None, num_envs=N, device=dev, random_polygons=GEN, seed=SEED, env_id_offset=rank * N)
    else:
        env = BatchedBoudaryEnv(doms, num_envs=N, device=dev)
    env.reset()
    state_bytes = N * env.max_verts * 32
    need_flush = state_bytes < 256 * 1024 * 1024           # L2 is ~126 MB
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev) if need_flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # sampled under the same load from burn-in to the end of the timed region
    step_idx = 0
    # burn-in (setup, untimed): all envs start their first episode together; run until resets are
    # spread over the steps so that the timed region sees the steady-state mix of episode phases
    for _ in range(args.burn_in):
        env.step(env.sample_actions(SEED, step_idx))
        step_idx += 1
    for _ in range(args.warmup):
        env.step(env.sample_actions(SEED, step_idx))
        step_idx += 1
    env.stats(reset=True)
    launches0 = env.launch_count

    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    if need_flush:
        for k in range(K):
            flush_buf.fill_(k & 0xFF)                       # evict the env state from L2 (not timed)
            ev[k][0].record()
            a = env.sample_actions(SEED, step_idx)
            ev[k][1].record()
            env.step(a)
            ev[k][2].record()
            step_idx += 1
    else:
        for k in range(K):
            ev[k][0].record()
            a = env.sample_actions(SEED, step_idx)
            ev[k][1].record()
            env.step(a)
            ev[k][2].record()
            step_idx += 1
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if need_flush:
        total_ms = sum(e[0].elapsed_time(e[2]) for e in ev)
    else:
        total_ms = ev[0][0].elapsed_time(ev[-1][2])
    kern_ms = sum(e[1].elapsed_time(e[2]) for e in ev)      # step kernel only
    clocks = sampler.stop() if rank == 0 else None
    launches = env.launch_count - launches0
    stats = env.stats(reset=True)

    # ---- optional: device time of phase A alone (state frozen at steady state), B+C by difference ----
    phase_times = None
    if args.phase_times:
        from reinforcementlearning4meshgeneration_b200._lib import check
        check(env._L.mg_set_phase_mask(env._h, 1), env._h, "mg_set_phase_mask")
        a = env.sample_actions(SEED, step_idx)
        for _ in range(5):
            env.step(a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            env.step(a)
        e1.record()
        torch.cuda.synchronize(dev)
        check(env._L.mg_set_phase_mask(env._h, 3), env._h, "mg_set_phase_mask")
        ta = e0.elapsed_time(e1) / 50
        phase_times = {"decide_ms": ta, "apply_reset_ms": kern_ms / K - ta}
        env.stats(reset=True)

    # ---- e2e: host buffers through the C ABI (mg_step_host), H2D + D2H inside the timed region ----
    Ke = max(3, min(K, 100))
    pinned = dict(
        act=torch.empty((N, 3), dtype=torch.float32).pin_memory(), obs=torch.empty((N, 18), dtype=torch.float32).pin_memory(),
        reward=torch.empty(N, dtype=torch.float64).pin_memory(), terminated=torch.empty(N, dtype=torch.uint8).pin_memory(),
        truncated=torch.empty(N, dtype=torch.uint8).pin_memory(), terminal_obs=torch.empty((N, 18), dtype=torch.float32).pin_memory(),
        n_elements=torch.empty(N, dtype=torch.int32).pin_memory())
    rng = np.random.default_rng(rank)
    lo, hi = np.array([-1, -1.5, 0], np.float32), np.array([1, 1.5, 1.5], np.float32)
    host_actions = [torch.from_numpy(rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)) for _ in range(4)]
    out = {k: v for k, v in pinned.items() if k != "act"}
    reward_view = out["reward"].numpy()                       # host view the caller reads its results through
    env.set_host_delta(True)      # persistent pinned result buffers: only the rows that changed cross PCIe
    for k in range(2):
        pinned["act"].copy_(host_actions[k % 4])
        env.step_host(pinned["act"], out)
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        pinned["act"].copy_(host_actions[k % 4])             # the policy's output lands in pinned host memory
        env.step_host(pinned["act"], out)
        _ = float(reward_view[0])                             # the caller reads the result on the host
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    h2d_meas, d2h_meas = env.last_host_bytes()
    h2d = N * 3 * 4
    # obs + reward + flags + element counts for every env; terminal observations only for finished envs
    # (mg_step_host ships them compacted: ~0.3 % of the envs per step in steady state)
    d2h = N * (18 * 4 + 8 + 1 + 1 + 4)
    env.stats(reset=True)

    # ---- reductions over ranks --------------------------------------------------------------
    t = torch.tensor([total_ms, kern_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, kern_ms_max, e2e_ms_max = [float(x) for x in t.tolist()]
    gstats = allreduce_stats(stats, dev) if world > 1 else stats
    total_env_steps = N * K * world
    value = total_env_steps / (total_ms_max * 1e-3)
    e2e_value = N * Ke * world / (e2e_ms_max * 1e-3)

    if rank == 0:
        peak, peak_src = load_peaks()
        # algorithmic bytes: B = 28 n + 14 n s + 226 per env-step, summed from the device counters
        alg_bytes = 28.0 * gstats["sum_n"] + 14.0 * gstats["sum_n_success"] + 226.0 * gstats["steps"]
        alg_bytes_per_launch_per_gpu = alg_bytes / max(1, K) / world
        kern_s_per_launch = kern_ms_max * 1e-3 / K
        achieved = alg_bytes_per_launch_per_gpu / kern_s_per_launch / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": {"c3": "c3: random star polygons 64..512 vertices, in-kernel auto-reset",
                                    "c2": "c2: paper domains d1/d2/d3 (boundary16/boundary15/test1)",
                                    "c1": "c1: BoudaryEnv(boundary()) 30 vertices"}[args.workload],
                       "envs_per_gpu": N, "global_envs": N * world, "policy": "uniform actions in the action box (Philox)",
                       "parallelism": f"env-sharded x{world}", "burn_in_steps": args.burn_in,
                       "l2": "L2 flushed between timed steps" if need_flush else f"state {state_bytes >> 20} MiB > L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_meas, "d2h_bytes_per_step": d2h_meas, "steps": Ke,
                    "note": "mg_step_host, delta rows (observations of changed envs, terminal observations of finished envs)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "mg_step_kernel",
                         "kernel_ms_per_launch": kern_ms_max / K,
                         "alg_bytes_per_launch": alg_bytes_per_launch_per_gpu},
            "episode_stats": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in gstats.items()},
            "mean_boundary_n": gstats["sum_n"] / max(1, gstats["steps"]),
            "success_rate": gstats["successes"] / max(1, gstats["steps"]),
            "wall_s_timed_region": t_wall,
        }
        if phase_times:
            line["phase_times"] = phase_times
        if world == 1 and not args.no_cpu_baseline:
            polys, sample = sample_polygons(args.workload)
            v, threads, steps, dt = cpu_port_throughput(polys, seconds_budget=10.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{sample}; {steps} steps x {threads} threads, {dt:.1f} s"}
        else:
            line["cpu_baseline"] = None
        emit(line)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of this run, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    # libraries chat on file descriptor 1 (NCCL prints its version banner there): route fd 1 to stderr for the
    # duration of the run and keep the real stdout for the one JSON line
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", choices=["c1", "c2", "c3"], default="c3")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--phase-times", action="store_true", help="also time phase A alone (profiling aid)")
    ap.add_argument("--burn-in", type=int, default=1500, help="untimed setup steps that de-synchronise the episodes")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
