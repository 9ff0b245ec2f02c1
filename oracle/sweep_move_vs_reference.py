"""TEST INFRASTRUCTURE: validate the C restatement of BoudaryEnv.move() (oracle_move) against the live reference
(v2/src/mesh_rl/envs/boundary_env.py:459-594) on the shipped domains, and record golden traces for the GPU tests.

    python oracle/sweep_move_vs_reference.py [steps per domain] [seed] [--record]

Actions: new_point = (r, phi) with r ~ U(0.05, 0.5), phi ~ U(0.2, 2.9) (Python floats), type from {0.1, 0.5, 0.9} with
probabilities 0.15 / 0.7 / 0.15 -- the value ranges general/EBRD.py feeds it.  An episode is compared step by step
(observation, done, is_complete, element count, boundary ids and coordinates, number of not-valid points) until it
is done or until the reference reaches smooth_pave (every candidate excluded), which is outside the restated path;
then the env is reset.  --record writes tests/golden/move_<domain>.npz for three domains."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader as rl  # noqa: E402
from oracle.c_oracle import OracleEnv  # noqa: E402


class Smoothing(Exception):
    pass


def actions(seed, T):
    rng = np.random.default_rng(seed)
    pol = np.stack([rng.uniform(0.05, 0.5, T), rng.uniform(0.2, 2.9, T)], axis=1)
    typ = rng.choice([0.1, 0.5, 0.9], size=T, p=[0.15, 0.7, 0.15])
    return pol, typ


def run_domain(name, xy, T, seed, record=False):
    t = rl.TracedEnv(xy)
    env = t.env

    def no_smooth(*a, **k):
        raise Smoothing()
    env.smooth_pave = no_smooth
    o = OracleEnv(xy, original_area=float(env.original_area))
    pol, typ = actions(seed, T)
    rec = dict(xy0=np.asarray(xy, np.float64), original_area=np.float64(env.original_area), polar=pol, type=typ,
               obs=np.zeros((T, 18), np.float32), obs_none=np.zeros(T, np.uint8), done=np.zeros(T, np.uint8),
               complete=np.zeros(T, np.uint8), smooth=np.zeros(T, np.uint8), n_elements=np.zeros(T, np.int32),
               n_boundary=np.zeros(T, np.int32), ref_index=np.full(T, -1, np.int32), reset_after=np.zeros(T, np.uint8))
    mism = elems = episodes = smooths = 0
    for i in range(T):
        p = [float(pol[i, 0]), float(pol[i, 1])]
        ty = float(typ[i])
        smooth = False
        try:
            obs, rew, done, info = env.move(p, ty)
        except Smoothing:
            smooth, obs, done, info = True, None, True, {"is_complete": False}
        oo, _, od, oinfo, osm = o.move(p, ty)
        ok = osm == smooth
        if not smooth:
            ok = ok and od == bool(done) and oinfo["is_complete"] == bool(info["is_complete"])
            ok = ok and ((obs is None) == (oo is None)) and (obs is None or np.array_equal(np.asarray(obs, np.float32), oo))
            ids = t.boundary_ids()
            bxy = np.array([[float(v.x), float(v.y)] for v in env.updated_boundary.vertices])
            oid, oxy = o.boundary()
            ok = ok and ids == oid.tolist() and np.array_equal(bxy, oxy)
            ok = ok and len(env.generated_meshes) == o.n_elements and len(env.not_valid_points) == o.n_excluded
        if not ok:
            mism += 1
            print(f"{name}: MISMATCH at step {i}: ref done={done} info={info} smooth={smooth}; oracle done={od} {oinfo} smooth={osm}")
            break
        rec["obs"][i] = 0 if oo is None else oo
        rec["obs_none"][i] = oo is None
        rec["done"][i], rec["complete"][i], rec["smooth"][i] = od, oinfo["is_complete"], osm
        rec["n_elements"][i], rec["n_boundary"][i], rec["ref_index"][i] = o.n_elements, o.n, o.ref_index
        elems = max(elems, o.n_elements)
        if done or smooth:
            episodes += 1
            smooths += smooth
            rec["reset_after"][i] = 1
            env.reset()
            t._rebuild_ids()
            env.not_valid_points = []
            env.last_not_valid_points = []
            o.reset()
    if record:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"move_{name}.npz"), **rec)
    return mism, episodes, smooths, elems


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
        m, ep, sm, el = run_domain(name, doms[name], T, seed + k, record=record and name in ("boundary0", "dolphine3", "easy1_1"))
        bad += m
        print(f"{name}: {T} moves, {ep} episodes ({sm} ended at smooth_pave), up to {el} elements, mismatches {m}")
    print("TOTAL mismatches", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
