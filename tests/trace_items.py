"""Item timeline of the warp-per-item step kernels (profiling aid, run by hand under gpurun with the MG_TRACE variant):

    MESHGEN_LIB=$PWD/reinforcementlearning4meshgeneration_b200/lib/variants/trace.so python tests/trace_items.py c3 gpurun_out/trace_c3

Records {kernel, kind, n, SM, start, end} of every item of a few steady-state steps and prints, per kernel launch: the
span, the item-duration distribution by kind, the slot occupancy over time (how full the 20 warp slots of every SM are
in each tenth of the launch) and the length of the tail (time after the last item STARTED)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KIND = {0: "decide_new", 1: "decide_rule", 2: "apply", 3: "observe", 4: "reset", 10: "decide_new(fail)", 11: "decide_rule(fail)"}


def main():
    import torch
    import bench
    from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv

    workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/trace"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    N = bench.ENVS_PER_GPU[workload]
    doms = bench.workload_domains(workload)
    if doms is None:
        env = BatchedBoudaryEnv(None, num_envs=N, device="cuda:0", random_polygons=bench.GEN, seed=bench.SEED)
    else:
        env = BatchedBoudaryEnv(doms, num_envs=N, device="cuda:0")
    env.reset()
    for k in range(600):
        env.step(env.sample_actions(bench.SEED, k))
    torch.cuda.synchronize()
    os.environ["MESHGEN_TRACE_FILE"] = out + ".bin"
    env.set_option("trace", 1 << 18)
    for k in range(steps):
        env.step(env.sample_actions(bench.SEED, 600 + k))
    torch.cuda.synchronize()
    env.set_option("trace_dump", 0)
    report(out + ".bin", out + ".txt", steps)


def report(path, txt, steps=4):
    dt = np.dtype([("kernel", "i4"), ("kind", "i4"), ("n", "i4"), ("sm", "i4"), ("t0", "u8"), ("t1", "u8")])
    r = np.fromfile(path, dtype=dt)
    lines = [f"{len(r)} item records"]
    if len(r) == 0:
        open(txt, "w").write("\n".join(lines) + "\n")
        return
    order = np.argsort(r["t0"])
    r = r[order]
    blocks = r[r["kernel"] == 0]           # block-start records (n = block index)
    r = r[r["kernel"] != 0]
    # launches of each kernel: the (steps - 1) largest gaps between consecutive item starts separate them
    t_base = int(r["t0"].min())
    names = {1: "update", 2: "observe", 3: "reset"}
    groups = []
    for k in (1, 2, 3):
        rk = r[r["kernel"] == k]
        if len(rk) == 0:
            continue
        gaps = np.diff(rk["t0"].astype(np.int64))
        cuts = np.sort(np.argsort(gaps)[-(steps - 1):] + 1) if steps > 1 and len(gaps) >= steps - 1 else np.array([], dtype=np.int64)
        for part in np.split(rk, cuts):
            if len(part):
                groups.append((int(part["t0"].min()), k, part))
    groups.sort(key=lambda g: g[0])
    for L, (_, k, q) in enumerate(groups):
        t0, t1 = int(q["t0"].min()), int(q["t1"].max())
        span = (t1 - t0) / 1e3
        dur = (q["t1"] - q["t0"]).astype(np.float64) / 1e3
        last_start = (int(q["t0"].max()) - t0) / 1e3
        lines.append(f"== launch {L}: kernel {names[k]}  items {len(q)}  start +{(t0 - t_base) / 1e3:.1f} us  end +{(t1 - t_base) / 1e3:.1f} us  span {span:.1f} us  "
                     f"last item started at {last_start:.1f} us (tail {span - last_start:.1f} us)  sum of item time {dur.sum():.0f} us "
                     f"= {dur.sum() / span / (148 * 20) * 100:.0f} % of 2960 slots")
        for kind in np.unique(q["kind"]):
            d = dur[q["kind"] == kind]
            nn = q["n"][q["kind"] == kind]
            lines.append(f"   {KIND.get(int(kind), kind):18s} count {len(d):6d}  mean {d.mean():6.1f}  p50 {np.median(d):6.1f}  p90 {np.percentile(d, 90):6.1f}  "
                         f"max {d.max():6.1f} us   mean n {nn.mean():6.1f}")
            # duration vs n (4 size classes)
            for lo, hi in ((0, 128), (128, 256), (256, 384), (384, 10000)):
                m = (nn > lo) & (nn <= hi)
                if m.any():
                    lines.append(f"        n in ({lo:4d},{hi:5d}]  count {int(m.sum()):6d}  mean {d[m].mean():6.1f} us")
        # occupancy over time: fraction of the 2960 slots busy in each tenth of the span
        edges = np.linspace(t0, t1, 11)
        occ = []
        for a, b in zip(edges[:-1], edges[1:]):
            overlap = np.clip(np.minimum(q["t1"], b) - np.maximum(q["t0"], a), 0, None).astype(np.float64)
            occ.append(overlap.sum() / (b - a) / (148 * 20))
        lines.append("   slot occupancy per tenth of the launch: " + " ".join(f"{o:.2f}" for o in occ))
        per_sm = np.bincount(q["sm"], minlength=148)
        lines.append(f"   items per SM: min {per_sm.min()} max {per_sm.max()}")
        b = blocks[(blocks["t0"] >= t0 - 20000) & (blocks["t0"] <= t1)]
        if len(b):
            bs = (b["t0"].astype(np.int64) - t0) / 1e3
            hist = np.histogram(bs, bins=[-20, 0, 2, 5, 10, 20, 30, 40, 50, 60, 80, 1000])[0]
            lines.append(f"   block starts: {len(b)} blocks; starts per interval [-20,0,2,5,10,20,30,40,50,60,80,..) us: {hist.tolist()}; blocks per SM "
                         f"min {np.bincount(b['sm'], minlength=148).min()} max {np.bincount(b['sm'], minlength=148).max()}")
        st = (q["t0"].astype(np.int64) - t0) / 1e3
        lines.append("   item starts per 5 us: " + " ".join(str(int(x)) for x in np.histogram(st, bins=np.arange(0, span + 5, 5))[0]))
    open(txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "report":
        report(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 4)
    else:
        main()
