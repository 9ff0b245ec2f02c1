"""TEST INFRASTRUCTURE: validate the C restatement of smooth_pave (M:816-1140, M:1284-1316, reached from
BoudaryEnv.move() when every reference candidate is on the not-valid list, E:548-583) against the live reference, and
record golden traces for the GPU tests.

    python oracle/sweep_smooth_vs_reference.py [moves per domain] [seed] [--record]

Same action stream as sweep_move_vs_reference.py; here the reference's smooth_pave RUNS and the episode goes on.  After
every move: observation, done, is_complete, element count, boundary ids, and the coordinates of EVERY vertex
(self.boundary.vertices: front and interior) are compared -- exactly (the oracle calls the same libm)."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader as rl  # noqa: E402
from oracle.c_oracle import OracleEnv  # noqa: E402
from oracle.sweep_move_vs_reference import actions  # noqa: E402


def run_domain(name, xy, T, seed, record=False):
    t = rl.TracedEnv(xy)
    env = t.env
    calls = {"n": 0}
    orig = env.smooth_pave

    def counted(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    env.smooth_pave = counted
    o = OracleEnv(xy, original_area=float(env.original_area))
    o.set_smoothing(True)
    pol, typ = actions(seed, T)
    V = 4 * len(xy) + 64
    rec = dict(xy0=np.asarray(xy, np.float64), original_area=np.float64(env.original_area), polar=pol, type=typ,
               obs=np.zeros((T, 18), np.float32), obs_none=np.zeros(T, np.uint8), done=np.zeros(T, np.uint8),
               complete=np.zeros(T, np.uint8), smooth=np.zeros(T, np.uint8), n_elements=np.zeros(T, np.int32),
               n_boundary=np.zeros(T, np.int32), ref_index=np.full(T, -1, np.int32), reset_after=np.zeros(T, np.uint8),
               n_vertices=np.zeros(T, np.int32), vertex_xy=np.zeros((T, V, 2), np.float64), boundary_ids=np.full((T, V), -1, np.int32))
    mism = episodes = smooths = 0
    max_dev = 0.0
    for i in range(T):
        p = [float(pol[i, 0]), float(pol[i, 1])]
        ty = float(typ[i])
        before = calls["n"]
        with contextlib.redirect_stdout(io.StringIO()):
            obs, rew, done, info = env.move(p, ty)
        smooth = calls["n"] > before
        oo, _, od, oinfo, osm = o.move(p, ty)
        ok = osm == smooth and od == bool(done) and oinfo["is_complete"] == bool(info["is_complete"])
        ok = ok and ((obs is None) == (oo is None)) and (obs is None or np.array_equal(np.asarray(obs, np.float32), oo))
        ids = t.boundary_ids()
        oid, oxy = o.boundary()
        ok = ok and ids == oid.tolist() and len(env.generated_meshes) == o.n_elements
        allxy = np.array([[float(v.x), float(v.y)] for v in env.boundary.vertices])
        ovx = o.vertex_xy()
        same = allxy.shape == ovx.shape and np.array_equal(allxy, ovx)
        if allxy.shape == ovx.shape and not same:
            max_dev = max(max_dev, float(np.max(np.abs(allxy - ovx))))
        ok = ok and same
        if not ok:
            mism += 1
            print(f"{name}: MISMATCH at move {i} (smooth={smooth}/{osm}): ref done={done} {info} n={len(ids)} el={len(env.generated_meshes)}; "
                  f"oracle done={od} {oinfo} n={len(oid)} el={o.n_elements}; vertex coords equal={same} max dev {max_dev:.3e}; obs equal="
                  f"{(obs is None) == (oo is None) and (obs is None or np.array_equal(np.asarray(obs, np.float32), oo))}")
            break
        nv = len(allxy)
        if nv > V:
            raise RuntimeError("vertex capacity of the recorder")
        rec["obs"][i] = 0 if oo is None else oo
        rec["obs_none"][i] = oo is None
        rec["done"][i], rec["complete"][i], rec["smooth"][i] = od, oinfo["is_complete"], osm
        rec["n_elements"][i], rec["n_boundary"][i], rec["ref_index"][i] = o.n_elements, o.n, o.ref_index
        rec["n_vertices"][i] = nv
        rec["vertex_xy"][i, :nv] = allxy
        rec["boundary_ids"][i, :len(ids)] = ids
        smooths += smooth
        if done:
            episodes += 1
            rec["reset_after"][i] = 1
            env.reset()
            t._rebuild_ids()
            env.not_valid_points = []
            env.last_not_valid_points = []
            o.reset()
    if record:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"smooth_{name}.npz"), **rec)
    return mism, episodes, smooths


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 5
    record = "--record" in sys.argv
    names = ["boundary0", "star", "half_wheel", "tool", "dolphine3", "boundary16", "easy1_1", "basic2", "bird", "fat"]
    doms = {"boundary0": rl.BOUNDARY0_XY}
    for d in names[1:]:
        doms[d] = rl.load_domain_xy(d)
    bad = 0
    for k, name in enumerate(names):
        m, ep, sm = run_domain(name, doms[name], T, seed + k, record=record and name in ("boundary0", "tool", "bird"))
        bad += m
        print(f"{name}: {T} moves, {ep} episodes, {sm} smooth_pave calls, mismatches {m}")
    print("TOTAL mismatches", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
