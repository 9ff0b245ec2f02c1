"""TEST INFRASTRUCTURE: record golden traces from the LIVE reference env into tests/golden/.

Run in the build container only (needs /root/reference):
    python oracle/record_golden.py

Writes
  tests/golden/domains.npz         polygons of a selection of ui/domains/*.json (+ boundary())
                                   exactly as read_polygon builds them (px / 100.0), plus the
                                   reference's poly_area() per domain
  tests/golden/trace_<name>.npz    per-step outputs of the reference BoudaryEnv for a seeded
                                   action stream with auto-reset on done

Trace keys (T steps, n0 = original polygon size):
  xy0[n0,2] original_area area_range[2] reset_obs[18] reset_ref_index reset_base_length
  actions[T,3] f32      np.random.default_rng(seed).uniform(low, high).astype(f32), one draw per step
  obs[T,18] f32         observation returned to the agent (after auto-reset when done)
  terminal_obs[T,18]    last observation of the episode where done, else 0
  reward[T] f64  terminated[T] truncated[T] success[T] u8
  n_elements[T] i32     len(generated_meshes) after the step (before reset)
  n_boundary[T] i32     len(updated_boundary.vertices) after the step (before reset)
  ref_index[T] i32      index of the reference point after the step (before reset); -1 if None
  ids[T,n0] i16         boundary vertex ids after the step (before reset), padded with -1
  new_xy[T,2] f64       coordinates of the vertex inserted by this step (NaN if none)
  base_length[T] f64, current_area[T] f64, cand_head_key[T] f64 (key of the list head)
  obs_none[T] u8        the reference returned None as next state (empty candidate list)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import ref_loader as rl

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

DOMAINS = ["boundary16", "boundary15", "test1", "random1_1", "test3", "dolphine3", "easy1_1", "basic2",
           "boundary_hole_r3", "half_wheel", "star", "tool", "bird", "fat", "boundary11", "boundary6"]
TRACES = [("boundary0", 7, 4096), ("boundary16", 1234, 1024), ("boundary15", 1234, 1024), ("test1", 1234, 1024),
          ("dolphine3", 123, 768), ("easy1_1", 5, 768), ("half_wheel", 9, 768), ("star", 2, 512)]


def record(name, xy, seed, T):
    t = rl.TracedEnv(xy)
    n0 = len(xy)
    st0 = t.state()
    out = dict(
        xy0=np.asarray(xy, np.float64), original_area=np.float64(t.env.original_area),
        area_range=np.array(t.env.estimated_area_range, np.float64), reset_obs=t.obs.copy(),
        reset_ref_index=np.int32(st0["ref_index"]), reset_base_length=np.float64(st0["base_length"]),
        seed=np.int64(seed), actions=rl.action_stream(seed, T),
        obs=np.zeros((T, 18), np.float32), terminal_obs=np.zeros((T, 18), np.float32), reward=np.zeros(T, np.float64),
        terminated=np.zeros(T, np.uint8), truncated=np.zeros(T, np.uint8), success=np.zeros(T, np.uint8),
        n_elements=np.zeros(T, np.int32), n_boundary=np.zeros(T, np.int32), ref_index=np.zeros(T, np.int32),
        ids=np.full((T, n0), -1, np.int16), new_xy=np.full((T, 2), np.nan, np.float64),
        base_length=np.zeros(T, np.float64), current_area=np.zeros(T, np.float64), cand_head_key=np.zeros(T, np.float64), obs_none=np.zeros(T, np.uint8))
    seen = n0
    for i in range(T):
        r = t.step(out["actions"][i])
        st = r["pre_reset_state"]
        out["obs"][i] = r["obs"]
        if r["terminal_obs"] is not None:
            out["terminal_obs"][i] = r["terminal_obs"]
        out["reward"][i] = r["reward"]
        out["terminated"][i] = r["terminated"]
        out["truncated"][i] = r["truncated"]
        out["success"][i] = r["success"]
        out["n_elements"][i] = r["n_elements"]
        out["obs_none"][i] = r["obs_none"]
        out["n_boundary"][i] = st["n"]
        out["ref_index"][i] = st["ref_index"]
        out["ids"][i, :st["n"]] = st["ids"]
        out["base_length"][i] = st["base_length"]
        out["current_area"][i] = st["current_area"]
        out["cand_head_key"][i] = st["candidates"][0][1] if st["candidates"] else np.inf
        mx = max(st["ids"])
        if mx >= seen:
            out["new_xy"][i] = st["xy"][st["ids"].index(mx)]
            seen = mx + 1
        if r["terminated"] or r["truncated"]:
            seen = n0
    np.savez_compressed(os.path.join(GOLDEN, f"trace_{name}.npz"), **out)
    print(f"{name}: T={T} elements={int(out['success'].sum())} episodes={int(out['terminated'].sum() + out['truncated'].sum())}")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    doms = {"boundary0": rl.BOUNDARY0_XY}
    for d in DOMAINS:
        doms[d] = rl.load_domain_xy(d)
    areas = {}
    for k, xy in doms.items():
        areas[k] = float(rl.make_env(xy).original_area)
    np.savez_compressed(os.path.join(GOLDEN, "domains.npz"), **{k: v for k, v in doms.items()},
                        **{"area__" + k: np.float64(a) for k, a in areas.items()})
    for name, seed, T in TRACES:
        record(name, doms[name], seed, T)


if __name__ == "__main__":
    main()
