#!/usr/bin/env python
"""Benchmark of the batched BoudaryEnv hot path (env-steps/s, whole job).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c2|c1] [--impl reference]

One "step" = one pass of the hot path over the whole batch: the synthetic policy kernel
(uniform actions in the action box) + mg_step (screen, decide, update and observe kernels:
every env advances one transition, auto-reset included) + every 64 steps the all-reduce of the episode statistics.
Workloads (BASELINE.json configs):
  c3 (default): random star polygons, 64..512 vertices, 65536 envs per GPU, in-kernel auto-reset
  c4          : the same with 131072 envs per GPU (1 M envs on 8 GPUs)
  c2          : paper domains d1/d2/d3 (120/196/272 vertices), 4096 envs per GPU
  c1          : BoudaryEnv(boundary()), 30 vertices, 4096 envs per GPU
Launches: the step's kernels are replayed from CUDA graphs (c3 / c4: four steps per graph launch; c2 / c1, which flush
L2 between timed steps: the whole timed loop as one graph with external event-record nodes around every step);
`--no-graph` launches every kernel eagerly.  Timed on the device with CUDA events either way (`launch_mode` in the line).
`e2e` = the same step through mg_step_host with pinned host buffers, one call and one synchronisation per step.
Multi-GPU: launched by torchrun, one rank per GPU, envs sharded by global env id (weak scaling); the only
collective is the all-reduce of the 12-element episode-statistics vector, enqueued every 64 steps inside the
timed loop (NCCL, no host synchronisation).
`--impl reference` times the reference's CPU implementation of the same path on the host cores: the C port of
the reference env (oracle/, kind "port") and, when the reference tree itself is importable (this build
container; it is not on the GPU box), the reference's own Python env on os.cpu_count() processes.  That arm never
imports the product package.
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
ENVS_PER_GPU = {"c3": 65536, "c4": 131072, "c2": 4096, "c1": 4096}
GEN = dict(min_verts=64, max_verts=512)
SEED = 2026
STATS_INTERVAL = 64
GRAPH_STEPS = 4          # steps per CUDA-graph launch in the timed loop (divides STATS_INTERVAL)
WORKLOAD_TEXT = {
    "c3": "c3: random star polygons 64..512 vertices, in-kernel auto-reset",
    "c4": "c4: random star polygons 64..512 vertices, in-kernel auto-reset, 131072 envs per GPU",
    "c2": "c2: paper domains d1/d2/d3 (boundary16/boundary15/test1)",
    "c1": "c1: BoudaryEnv(boundary()) 30 vertices",
}


def make_config(args, world):
    """The workload description both arms print (so that the driver can check they ran the same thing)."""
    N = args.envs or ENVS_PER_GPU[args.workload]
    state_bytes = N * (GEN["max_verts"] if args.workload in ("c3", "c4") else 272) * 32
    return {"workload": WORKLOAD_TEXT[args.workload], "envs_per_gpu": N, "global_envs": N * world,
            "policy": "uniform actions in the action box (Philox)", "parallelism": f"env-sharded x{world}",
            "burn_in_steps": args.burn_in, "stats_allreduce_every": STATS_INTERVAL,
            "l2": "L2 flushed between timed steps" if state_bytes < 256 * 1024 * 1024 else f"state {state_bytes >> 20} MiB > L2"}


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
# CPU arm: the reference env on the host cores (C port always; the Python reference when importable)
# --------------------------------------------------------------------------------------------
def sample_polygons(workload):
    """Polygons for the CPU arm: the fixed domains, or (c3 / c4) the committed sample of the device generator's
    polygons, tests/golden/c3_polys.npz (recorded once by oracle/record_c3_polys.py: seed 2026, global env ids
    0..63, first episode).  Never touches the product package."""
    doms = workload_domains(workload)
    if doms is not None:
        return doms, f"{workload}: {len(doms)} domain(s)"
    p = os.path.join(ROOT, "tests", "golden", "c3_polys.npz")
    if os.path.exists(p):
        z = np.load(p)
        polys = [z[k] for k in sorted(z.files, key=lambda s: int(s[1:]))]
        return polys, f"{workload}: {len(polys)} polygons of the device generator (seed {SEED}, tests/golden/c3_polys.npz)"
    d = golden_domains()
    return [d["boundary16"], d["boundary15"], d["test1"], d["test3"]], f"{workload} stand-in: d1/d2/d3/test3 (fixture missing)"


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


def reference_root():
    """Where the reference's own Python tree is importable from, or None (it is not on the GPU box)."""
    for r in (os.environ.get("MESHGEN_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if r and os.path.isfile(os.path.join(r, "v2", "src", "mesh_rl", "envs", "boundary_env.py")):
            return r
    return None


def _python_ref_worker(args):
    """One worker process: the unmodified reference BoudaryEnv (v2) on one polygon, uniform random float32
    actions, reset on done, for `seconds` of wall time after a short warm-up."""
    root, xy, seed, seconds = args
    os.environ["MESHGEN_REFERENCE_ROOT"] = root
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    env = ref_loader.make_env(np.asarray(xy))
    env.reset()
    rng = np.random.default_rng(seed)
    lo, hi = ref_loader.LOW, ref_loader.HIGH

    def run(budget):
        n = 0
        t_end = time.perf_counter() + budget
        while time.perf_counter() < t_end:
            for _ in range(20):
                obs, _, te, tr, _ = env.step(rng.uniform(lo, hi).astype(np.float32))
                n += 1
                if te or tr or obs is None:
                    env.reset()
        return n

    run(min(1.0, seconds / 5))
    t0 = time.perf_counter()
    n = run(seconds)
    return n, time.perf_counter() - t0


def python_reference_throughput(polys, seconds=10.0, procs=None):
    """Steps/s of the reference's own Python env on `procs` worker processes."""
    import multiprocessing as mp
    root = reference_root()
    if root is None:
        return None
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_python_ref_worker, [(root, np.asarray(polys[i % len(polys)]), 1000 + i, seconds) for i in range(procs)])
    rate = sum(n / dt for n, dt in res)
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": f"unmodified reference v2 BoudaryEnv under the stub loader (oracle/ref_loader.py), {procs} processes x "
                      f"{seconds:.0f} s, {sum(n for n, _ in res)} steps", "per_core": rate / procs}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores; each step of this arm is
    one bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    polys, sample = sample_polygons(args.workload)
    vals = []
    t_all = time.perf_counter()
    budget = max(1.0, min(8.0, 100.0 / max(1, args.steps + args.warmup)))
    threads = steps = 0
    for i in range(args.warmup + args.steps):
        v, threads, steps, dt = cpu_port_throughput(polys, seconds_budget=budget)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(1, args.steps + args.warmup),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"C port of the reference env (oracle/liboracle.so); {sample}; {steps} steps x {threads} threads per timed step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    py = python_reference_throughput(polys, seconds=args.python_ref_seconds) if not args.no_python_reference else None
    line["cpu_baseline_python"] = py if py is not None else {
        "unavailable": "the reference's Python tree is not importable on this box (it exists only in the build container)"}
    emit(line)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node (pinned buffers are then allocated and the
    PCIe traffic of mg_step_host is served locally).  Best effort: returns the cpulist or None."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not out:
            return None
        dom_bus = out[-12:] if len(out) >= 12 else out              # 00000000:1B:00.0 -> 0000:1b:00.0
        path = f"/sys/bus/pci/devices/{dom_bus}/local_cpulist"
        if not os.path.exists(path):
            return None
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{min(cpus)}-{max(cpus)} ({len(cpus)} cpus)"
    except Exception:
        return None
    return None



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
    numa = bind_to_gpu_numa_node(local)      # the e2e leg is host-side work: keep it next to the GPU
    from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
    from reinforcementlearning4meshgeneration_b200.distributed import allreduce_stats, allreduce_stats_device, stats_from_tensor

    N = args.envs or ENVS_PER_GPU[args.workload]
    doms = workload_domains(args.workload)
    if doms is None:
        env = BatchedBoudaryEnv(None, num_envs=N, device=dev, random_polygons=GEN, seed=SEED, env_id_offset=rank * N)
    else:
        env = BatchedBoudaryEnv(doms, num_envs=N, device=dev)
    env.reset()
    config = make_config(args, world)
    need_flush = config["l2"].startswith("L2 flushed")
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev) if need_flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # sampled under the same load from burn-in to the end of the timed region
    step_idx = 0
    stats_dev = torch.zeros(12, dtype=torch.int64, device=dev)
    # burn-in (setup, untimed): all envs start their first episode together; run until resets are
    # spread over the steps so that the timed region sees the steady-state mix of episode phases
    for _ in range(args.burn_in):
        env.step(env.sample_actions(SEED, step_idx))
        step_idx += 1
    # The launches of a step (policy kernel + the kernels of mg_step) are replayed from CUDA graphs -- GRAPH_STEPS steps
    # per launch, single steps for the remainder -- when the state is larger than L2 (no flush between steps): mg_step
    # keeps no step state on the host, and the policy's step index lives in device memory (mg_sample_actions_seq).
    use_graph = not need_flush and not args.no_graph
    K = args.steps
    graphs, launches_per_step = {}, 0
    graph_note = None
    if use_graph:
        step_ctr = torch.tensor([step_idx, 0], dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                        # (torch's capture recipe: first use of the path on a side stream)
            env.step(env.sample_actions(SEED, step_ctr))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        try:
            for n in (1, GRAPH_STEPS):
                g = torch.cuda.CUDAGraph()
                l0 = env.launch_count
                with torch.cuda.graph(g):
                    for _ in range(n):
                        env.step(env.sample_actions(SEED, step_ctr))
                launches_per_step = (env.launch_count - l0) // n
                graphs[n] = g
        except Exception as ex:                              # no capture on this box: the same steps, launched one by one
            graph_note = f"graph capture failed ({type(ex).__name__}: {str(ex)[:120]}); eager launches"
            graphs.clear()
            use_graph = False
            torch.cuda.synchronize(dev)
            step_idx = int(step_ctr[0].item())

    # The configurations whose state fits in L2 flush it between timed steps.  There the whole timed loop -- flush, event,
    # policy kernel, mg_step, event, K times -- is one graph whose events are external event-record nodes, so that a step
    # is timed on the device without the flush and without the gaps of separately launched kernels.
    loop_graph, loop_events, loop_reduces = None, None, 0
    if need_flush and not args.no_graph and world == 1:
        try:
            step_ctr = torch.tensor([step_idx, 0], dtype=torch.int64, device=dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                env.step(env.sample_actions(SEED, step_ctr))
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            loop_events = [(torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
                           for _ in range(K)]
            loop_graph = torch.cuda.CUDAGraph()
            l0 = env.launch_count
            with torch.cuda.graph(loop_graph):
                for k in range(K):
                    flush_buf.fill_(k & 0xFF)               # evict the env state from L2 (not timed)
                    loop_events[k][0].record()
                    env.step(env.sample_actions(SEED, step_ctr))
                    if (k + 1) % STATS_INTERVAL == 0:
                        env.stats_async(stats_dev)          # (one rank: the all-reduce is the device-side sum itself)
                        loop_reduces += 1
                    loop_events[k][1].record()
            launches_per_step = (env.launch_count - l0 - loop_reduces) // K
        except Exception as ex:
            graph_note = f"graph capture failed ({type(ex).__name__}: {str(ex)[:120]}); eager launches"
            loop_graph = None
            torch.cuda.synchronize(dev)
            step_idx = int(step_ctr[0].item())

    def run_steps(k0, k1, record=None):
        """Steps k0..k1-1 of a loop; returns the statistics all-reduces it enqueued (every STATS_INTERVAL steps)."""
        nonlocal step_idx
        reduces, k = 0, k0
        while k < k1:
            if use_graph:
                n = GRAPH_STEPS if (k1 - k >= GRAPH_STEPS and k % STATS_INTERVAL + GRAPH_STEPS <= STATS_INTERVAL) else 1
                graphs[n].replay()
            else:
                n = 1
                if need_flush:
                    flush_buf.fill_(k & 0xFF)               # evict the env state from L2 (not timed)
                record[k][0].record()
                a = env.sample_actions(SEED, step_idx)
                record[k][1].record()
                env.step(a)
            k += n
            step_idx += n
            if k % STATS_INTERVAL == 0:                     # the job's one collective, inside the timed loop
                allreduce_stats_device(env.stats_async(stats_dev))
                reduces += 1
            if not use_graph:
                record[k - 1][2].record()
        return reduces

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(max(K, args.warmup))]
    run_steps(0, args.warmup, ev)
    if world > 1:                                            # warm the collective up (communicator set-up is not a step)
        allreduce_stats_device(env.stats_async(stats_dev))
    env.stats(reset=True)
    launches0 = env.launch_count

    if loop_graph is not None:
        step_ctr[0] = step_idx                               # (the warm-up steps above were launched with host-side indices)
        loop_graph.replay()                                  # one untimed pass of the graph itself
        env.stats(reset=True)
    barrier()
    t_wall0 = time.perf_counter()
    if use_graph:
        ev[0][0].record()
    if loop_graph is not None:
        loop_graph.replay()
        n_reduces = loop_reduces
    else:
        n_reduces = run_steps(0, K, ev)
    if use_graph:
        ev[0][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if loop_graph is not None:
        total_ms = sum(e[0].elapsed_time(e[1]) for e in loop_events)
        kern_ms = total_ms
        launches = K * launches_per_step + n_reduces
        step_idx = int(step_ctr[0].item())
        loop_graph = True                                    # (the graph itself is released below)
        loop_events = None
    elif use_graph:
        total_ms = ev[0][0].elapsed_time(ev[0][2])
        kern_ms = total_ms                                   # (includes the policy kernel: no events inside a graph)
        launches = K * launches_per_step + n_reduces           # + mg_stats_kernel per reduce
        step_idx = int(step_ctr[0].item())
    else:
        if need_flush:
            total_ms = sum(e[0].elapsed_time(e[2]) for e in ev[:K])
        else:
            total_ms = ev[0][0].elapsed_time(ev[K - 1][2])
        kern_ms = sum(e[1].elapsed_time(e[2]) for e in ev[:K])   # mg_step (+ the statistics reduce on its steps) only
        launches = env.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    stats = env.stats(reset=True)
    graphs.clear()

    # ---- per-kernel device times (CUDA events inside mg_step, a separate short pass on the same steady state) ----
    env.set_kernel_timing(True)
    Kt = max(8, min(K, 200))
    for _ in range(Kt):
        if need_flush:
            flush_buf.fill_(1)
        env.step(env.sample_actions(SEED, step_idx))
        step_idx += 1
    ktimes = env.kernel_times()
    env.set_kernel_timing(False)
    kstats = env.stats(reset=True)

    # ---- e2e: host buffers through the C ABI (mg_step_host), H2D + D2H inside the timed region ----
    Ke = max(3, min(K, 100))
    pinned = dict(
        act=torch.empty((N, 3), dtype=torch.float32).pin_memory(), obs=torch.empty((N, 18), dtype=torch.float32).pin_memory(),
        reward=torch.empty(N, dtype=torch.float64).pin_memory(), terminated=torch.empty(N, dtype=torch.uint8).pin_memory(),
        truncated=torch.empty(N, dtype=torch.uint8).pin_memory(), terminal_obs=torch.empty((N, 18), dtype=torch.float32).pin_memory(),
        n_elements=torch.empty(N, dtype=torch.int32).pin_memory())
    rng = np.random.default_rng(rank)
    lo, hi = np.array([-1, -1.5, 0], np.float32), np.array([1, 1.5, 1.5], np.float32)
    # The policy's output of every step sits in pinned host memory: a pool of pre-drawn action buffers, one per timed step
    # (at most 64), so that no env sees the same action twice within the window -- a failed action repeated on an unchanged
    # state fails again, and with a short cycle most envs would run into the 100-failures truncation inside the window.
    n_act = max(4, min(Ke, 64))
    host_actions = [torch.from_numpy(rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)).pin_memory() for _ in range(n_act)]
    out = {k: v for k, v in pinned.items() if k != "act"}
    reward_view = out["reward"].numpy()                       # host view the caller reads its results through
    for k in range(2 * n_act + 4):                            # (mg_step_host builds a graph per action buffer on its second use)
        env.step_host(host_actions[k % n_act], out)
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        env.step_host(host_actions[k % n_act], out)           # H2D of this step's actions is inside mg_step_host
        _ = float(reward_view[0])                             # the caller reads the result on the host
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    h2d_meas, d2h_meas = env.last_host_bytes()
    env.stats(reset=True)

    # ---- reductions over ranks --------------------------------------------------------------
    t = torch.tensor([total_ms, kern_ms, e2e_s * 1e3, ktimes["screen_ms"], ktimes["decide_ms"], ktimes["update_ms"], ktimes["observe_ms"]],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, kern_ms_max, e2e_ms_max, screen_ms, decide_ms, update_ms, observe_ms = [float(x) for x in t.tolist()]
    gstats = allreduce_stats(stats, dev) if world > 1 else stats
    gk = allreduce_stats(kstats, dev) if world > 1 else kstats
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
        # bytes the memoised two-kernel design has to move: every env's 128-byte record + action + outputs + the 32
        # record bytes a failed step changes (screen kernel); ring, keys and stamps only for the steps that reach the
        # ring kernel
        per_env_screen = 128 + 12 + 14 + 32
        ksteps = max(1, gk["steps"])
        per = 1.0 / Kt / world
        screen_bytes = per_env_screen * ksteps * per
        # decide: ring (16 n) + records per item; update: ring + keys + stamps + ids (32 n) read, shifted tail (16 n on
        # average) written per accepted element; observe: ring + keys + stamps (28 n) per changed env
        decide_items = gk["ring_items"] - (gk["successes"] - 0)          # upper bound on the decide list: every ring item
        decide_bytes = (16.0 * gk["sum_n_ring"] + 226.0 * gk["ring_items"]) * per
        update_bytes = (48.0 * gk["sum_n_success"] + 226.0 * gk["successes"]) * per
        observe_bytes = (28.0 * gk["sum_n_success"] + 226.0 * gk["successes"]) * per

        def entry(name, ms, b, units, **extra):
            d = {"name": name, "ms": ms, "alg_bytes": b, "frac": b / (ms * 1e-3) / 1e9 / peak if ms > 0 else None, "units": units}
            d.update(extra)
            return d
        per_kernel = [
            entry("mg_step_screen_kernel", screen_ms, screen_bytes, "every env: 128 B record + 12 B action read, 14 B results + 32 B record written"),
            entry("mg_step_decide_kernel", decide_ms, decide_bytes, "per item that needs the boundary: 16 n (ring) + 226 (upper bound: counts every ring item)",
                  items_per_launch=gk["ring_items"] * per),
            entry("mg_step_update_kernel", update_ms, update_bytes, "per accepted element: 32 n read (ring, keys, stamps, ids) + ~16 n written (shifted tail) + 226",
                  items_per_launch=gk["successes"] * per),
            entry("mg_step_observe_kernel", observe_ms, observe_bytes, "per changed env: 28 n (ring, keys, stamps) + 226; resets on top",
                  items_per_launch=gk["successes"] * per),
        ]
        ring_bytes = decide_bytes + update_bytes + observe_bytes
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get(args.workload)
                traffic_src = tj.get("source")
            except Exception:
                traffic = None
        memo_bytes = screen_bytes + ring_bytes
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms_max / K, "timed_ms": total_ms_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_meas, "d2h_bytes_per_step": d2h_meas, "steps": Ke,
                    "pcie_gbs_per_rank": (h2d_meas + d2h_meas) / (e2e_ms_max * 1e-3 / Ke) / 1e9, "numa_binding": numa,
                    "note": "mg_step_host with pinned host buffers: actions read over PCIe by the screen kernel (no staging copy); "
                            "rewards, flags, element counts, changed observation rows and terminal rows written by the step "
                            "kernels into the caller's arrays; one sync"},
            "gpu_launches": int(launches),
            "launch_mode": ({"cuda_graph": True, "steps_per_graph": GRAPH_STEPS, "kernels_per_step": launches_per_step,
                             "note": "policy kernel + the kernels of mg_step replayed from CUDA graphs (mg_step keeps no host state; "
                                     "the policy's step index lives in device memory); --no-graph launches them one by one"}
                            if use_graph else
                            {"cuda_graph": True, "steps_per_graph": K, "kernels_per_step": launches_per_step,
                             "note": "the whole timed loop is one graph: L2 flush, external event, policy kernel + the kernels of "
                                     "mg_step, external event, per step; a step is timed between its two events"}
                            if loop_graph else {"cuda_graph": False, "note": graph_note}),
            "collective": {"op": "all_reduce(sum) of mg_episode_stats (10 x int64 + 2 x float64)", "backend": "nccl" if world > 1 else "none (1 rank: device-side sum only)",
                           "every_steps": STATS_INTERVAL, "inside_timed_loop": n_reduces},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": max(per_kernel, key=lambda k: k["ms"])["name"],
                         "kernel_ms_per_launch": kern_ms_max / K,
                         "kernel_ms_note": ("CUDA events around the whole graph-replayed loop / steps: one mg_step (its four kernels, the "
                                            "reset kernel beside two of them) + the 2-us policy kernel -- no events inside a graph"
                                            if use_graph else "external events around policy kernel + mg_step of every step inside the loop graph"
                                            if loop_graph else "CUDA events around every mg_step call of the timed loop"),
                         "alg_bytes_per_launch": alg_bytes_per_launch_per_gpu,
                         "model": "SURVEY 8d: sum(28 n + 14 n s + 226) per env-step, i.e. every env's boundary crossing HBM once per "
                                  "step.  The kernels memoise the state-only predicates, so ~93 % of the steps never read their "
                                  "boundary: `frac` can exceed 1; `memo_frac` uses the bytes this design must move (per_kernel)",
                         "memo_alg_bytes_per_launch": memo_bytes,
                         "memo_frac": memo_bytes / kern_s_per_launch / 1e9 / peak,
                         "per_kernel": per_kernel,
                         "per_kernel_note": f"CUDA events around each kernel inside mg_step, {ktimes['steps']} steps after the timed region "
                                            "(plain launches; decide = 0 when fused into the update launch; mg_step_reset_kernel "
                                            "runs on a side stream next to the update / observe kernels and is not in this list)"},
            "episode_stats": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in gstats.items()},
            "mean_boundary_n": gstats["sum_n"] / max(1, gstats["steps"]),
            "success_rate": gstats["successes"] / max(1, gstats["steps"]),
            "ring_fraction": gstats["ring_items"] / max(1, gstats["steps"]),
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            polys, sample = sample_polygons(args.workload)
            v, threads, steps, dt = cpu_port_throughput(polys, seconds_budget=10.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"C port of the reference env (oracle/liboracle.so); {sample}; {steps} steps x {threads} threads, {dt:.1f} s"}
            py = python_reference_throughput(polys, seconds=args.python_ref_seconds) if not args.no_python_reference else None
            line["cpu_baseline_python"] = py if py is not None else {
                "unavailable": "the reference's Python tree is not importable on this box (it exists only in the build container); "
                               "measured there: profiles/r2_bench_reference_container.json"}
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
    ap.add_argument("--workload", choices=["c1", "c2", "c3", "c4"], default="c3")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-python-reference", action="store_true", help="skip the Python reference even when its tree is importable")
    ap.add_argument("--python-ref-seconds", type=float, default=10.0)
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
