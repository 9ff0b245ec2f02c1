"""GPU parity, round-2 hardening (pytest -m gpu): per-step internal state against the golden trace of the live
reference, sentinel paths, the workload generator against the host densifier, random polygons across episode
boundaries, the memoised verdicts, and the new ABI entry points (observation delta, snapshot header, stream
ordering, device-side statistics, kernel timing)."""
import numpy as np
import pytest

from helpers import action_stream, assert_rollout_matches, load_domains, load_trace

pytestmark = pytest.mark.gpu

REWARD_TOL = 1e-9
LOW_A = np.array([-1.0, -1.5, 0.0], np.float32)
HIGH_A = np.array([1.0, 1.5, 1.5], np.float32)


def _mk(domains, n, **kw):
    from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
    return BatchedBoudaryEnv(domains, num_envs=n, **kw)


def test_boundary0_internal_state_at_every_step_of_the_golden_trace():
    """BoudaryEnv(boundary()), seed 7, 4096 steps -- at EVERY step the boundary vertex ids, the
    coordinates of the inserted vertex, the reference index, the boundary size, base length and remaining area of
    the CUDA env equal what the live reference recorded (state before the reset on done steps)."""
    import torch
    tr = load_trace("boundary0")
    T = len(tr["reward"])
    env = _mk([tr["xy0"]], 1, auto_reset=False)
    env.reset()
    acts = torch.from_numpy(tr["actions"]).to(env.device)
    for t in range(T):
        r = env.step(acts[t:t + 1])
        s = env.get_state(0)
        n = int(tr["n_boundary"][t])
        assert s["n"] == n, f"boundary size differs at step {t}"
        assert np.array_equal(s["ids"], tr["ids"][t][:n].astype(np.int32)), f"boundary vertex ids differ at step {t}"
        if not tr["obs_none"][t]:
            assert s["ref_index"] == int(tr["ref_index"][t]), f"reference index differs at step {t}"
            assert s["base_length"] == float(tr["base_length"][t]), f"base length differs at step {t}"
        assert s["current_area"] == float(tr["current_area"][t]) or \
            abs(s["current_area"] - float(tr["current_area"][t])) <= 1e-12 * abs(float(tr["current_area"][t])), f"area differs at step {t}"
        assert s["n_elements"] == int(tr["n_elements"][t])
        if not np.isnan(tr["new_xy"][t, 0]):
            j = int(np.argmax(s["ids"]))
            assert np.array_equal(s["xy"][j], tr["new_xy"][t]), f"inserted vertex differs at step {t}"
        head = min((c[1] for c in s["candidates"]), default=np.inf)
        assert head == float(tr["cand_head_key"][t]), f"head of the candidate list differs at step {t}"
        te, tru = bool(r.terminated[0]), bool(r.truncated[0])
        assert te == bool(tr["terminated"][t]) and tru == bool(tr["truncated"][t])
        if te or tru:
            env.reset()


def test_memoised_verdicts_agree_with_a_fresh_evaluation():
    """EnvHot::flags (would a rule -1 / +1 step be accepted?) is memoised when the state changes.  Replaying every
    env through the oracle with actions forced to rule -1 / +1 shows the verdict the reference would reach: the
    step creates an element exactly when the flag is set."""
    import torch
    from oracle.c_oracle import OracleEnv
    doms, areas = load_domains()
    names = ["boundary16", "dolphine3", "easy1_1", "star", "basic2", "half_wheel"]
    per, T = 8, 150
    N = per * len(names)
    env_domain = np.repeat(np.arange(len(names)), per)
    env = _mk([doms[k] for k in names], N, env_domain=env_domain, auto_reset=False)
    env.reset()
    streams = [action_stream(4000 + e, T) for e in range(N)]
    oracles = [OracleEnv(doms[names[env_domain[e]]], original_area=areas[names[env_domain[e]]]) for e in range(N)]
    checked = accepted = 0
    for t in range(T):
        a = np.stack([s[t] for s in streams])
        env.step(torch.from_numpy(a).to(env.device))
        for e in range(N):
            oracles[e].step(streams[e][t])
        if t % 10 == 9:
            for e in range(0, N, 3):
                o = oracles[e]
                if o.n <= 5 or o.ref_index < 0:
                    continue
                flags = env.get_state(e)["memo_flags"]
                for bit, rule in ((1 | 4, -1.0), (2 | 8, 1.0)):      # accepted, or valid with the intersection test pending
                    # probe on a copy of the oracle's state: rebuild it by replaying the stream so far
                    p = OracleEnv(doms[names[env_domain[e]]], original_area=areas[names[env_domain[e]]])
                    for u in range(t + 1):
                        p.step(streams[e][u])
                    p.step(np.array([rule, 0.0, 0.5], np.float32))
                    ok = bool(p.last_info()["success"])
                    if flags & bit & 3:
                        assert ok, f"memoised 'accepted' is wrong at step {t} env {e} rule {rule}"
                    if not (flags & bit):
                        assert not ok, f"memoised 'rejected' is wrong at step {t} env {e} rule {rule}"
                    checked += 1
                    accepted += ok
    assert checked > 100 and accepted > 3


def test_sentinel_paths():
    """(a) stepping an env that already finished (n <= 5 on entry, E:428-430): reward 10, done, every step -- as the
    reference does; (b) a polygon without any reference candidate (all interior angles >= 0.972 pi): the reference
    returns None as observation and raises on the next step; here the env reports a zero observation and is
    truncated by its first step; (c) a polygon with a zero-length edge (division by zero in the reference,
    C:664): IEEE semantics, finite outputs, no hang."""
    import torch
    from oracle.c_oracle import OracleEnv
    # (a) auto-reset off on both sides: finished envs keep being stepped, like a careless caller of the reference would
    doms, areas = load_domains()
    xy, area = doms["tool"], areas["tool"]
    N, T = 32, 300
    env = _mk([xy], N, auto_reset=False)
    env.reset()
    oracles = [OracleEnv(xy, original_area=area) for _ in range(N)]
    streams = [action_stream(600 + e, T) for e in range(N)]
    n_after = 0
    for t in range(T):
        a = np.stack([s[t] for s in streams])
        r = env.step(torch.from_numpy(a).to(env.device))
        rew, te, tru = r.reward.cpu().numpy(), r.terminated.cpu().numpy(), r.truncated.cpu().numpy()
        for e in range(N):
            if oracles[e].crashed:
                continue
            finished = oracles[e].n <= 5
            o, er, ete, etr, _ = oracles[e].step(streams[e][t])
            if oracles[e].crashed:
                continue
            assert bool(te[e]) == ete and bool(tru[e]) == etr, f"flags differ t={t} env={e}"
            assert abs(rew[e] - er) <= REWARD_TOL * max(1.0, abs(er)), f"reward differs t={t} env={e}"
            if finished:
                assert er == 10.0 and (ete or etr)
                n_after += 1
    assert n_after > 100, "no env finished with n <= 5 in this run"
    # (b)
    k = 240
    ang = -2 * np.pi * np.arange(k) / k          # clockwise regular polygon: interior angle pi - 2 pi / k > 0.972 pi
    env = _mk([np.stack([5 + 3 * np.cos(ang), 5 + 3 * np.sin(ang)], axis=1)], 2)
    obs = env.reset().cpu().numpy()
    assert env.get_state(0)["ref_index"] == -1 and not obs.any()
    r = env.step(torch.zeros((2, 3), device=env.device))
    assert r.truncated.cpu().numpy().all() and not r.terminated.cpu().numpy().any() and (r.reward.cpu().numpy() == 0).all()
    # (c)
    sq = [(0, 0), (0, 1), (0, 2), (0, 3), (1, 3), (2, 3), (3, 3), (3, 2), (3, 1), (3, 0), (2, 0), (2, 0), (1, 0)]
    env = _mk([np.array(sq, np.float64)], 8)
    env.reset()
    for t in range(60):
        r = env.step(env.sample_actions(5, t))
    assert torch.isfinite(r.obs).all() and torch.isfinite(r.reward).all()


def test_device_generator_against_the_host_densifier():
    """the in-kernel workload generator densifies its coarse star polygon like ui/tk-ui.py:252-276.
    The coarse polygon (pixels, clockwise) and the spacing are read back and pushed through the host densifier
    (domains.densify, a restatement of calculate_density checked against the reference's own output in
    tests/test_domain_tools_cpu.py): same vertex count, coordinates within 1e-12 (the kernel emits
    prev + A (j+1) (cur - prev) / L, the reference prev + A (j+1) (cos, sin)(atan2(...)): ~1e-16 relative apart)."""
    from reinforcementlearning4meshgeneration_b200.domains import densify
    N = 64
    env = _mk(None, N, random_polygons=dict(min_verts=64, max_verts=512), seed=31)
    env.reset()
    worst = 0.0
    for e in range(N):
        for ep in (0, 3):
            d = env.debug_polygon(e, ep)
            pts = [(float(x), float(y)) for x, y in d["coarse_px"]]
            assert len(pts) >= 8 and len({p for p in pts}) >= 8
            dense = np.array(densify(pts, [1.0] * len(pts), d["spacing_px"]), np.float64) / 100.0
            assert len(dense) == d["n"], f"env {e} episode {ep}: {len(dense)} vertices on the host, {d['n']} on the device"
            worst = max(worst, float(np.abs(dense - d["xy"]).max()))
        assert np.array_equal(env.debug_polygon(e, 0)["xy"], env.get_state(e)["xy"])
    assert worst <= 1e-12, worst


def test_random_polygons_with_auto_reset_across_episode_boundaries():
    """Config-3 soak in the test suite: random polygons, in-kernel auto-reset, device Philox actions; every picked
    env is replayed through the CPU oracle across ALL its episodes (each episode's polygon and area are read back
    from the generator's counter)."""
    import torch
    from oracle.c_oracle import OracleEnv
    N, T = 512, 420
    env = _mk(None, N, random_polygons=dict(min_verts=64, max_verts=512), seed=77)
    env.reset()
    dev = env.device
    pick = list(range(0, N, 8))
    idx = torch.tensor(pick, device=dev, dtype=torch.long)
    K = len(pick)
    rec = dict(act=torch.zeros((T, K, 3), device=dev), obs=torch.zeros((T, K, 18), device=dev),
               tobs=torch.zeros((T, K, 18), device=dev), rew=torch.zeros((T, K), dtype=torch.float64, device=dev),
               te=torch.zeros((T, K), dtype=torch.uint8, device=dev), tr=torch.zeros((T, K), dtype=torch.uint8, device=dev),
               ne=torch.zeros((T, K), dtype=torch.int32, device=dev))
    for t in range(T):
        a = env.sample_actions(4242, t)
        rec["act"][t] = a[idx]
        r = env.step(a)
        rec["obs"][t] = r.obs[idx]; rec["tobs"][t] = r.terminal_obs[idx]; rec["rew"][t] = r.reward[idx]
        rec["te"][t] = r.terminated[idx]; rec["tr"][t] = r.truncated[idx]; rec["ne"][t] = r.n_elements[idx]
    rec = {k: v.cpu().numpy() for k, v in rec.items()}
    episodes = 0
    for k, e in enumerate(pick):
        ep = 0
        poly = env.debug_polygon(e, ep)
        o = OracleEnv(poly["xy"], original_area=poly["original_area"])
        for t in range(T):
            obs, r, te, tru, _ = o.step(rec["act"][t, k])
            assert bool(rec["te"][t, k]) == te and bool(rec["tr"][t, k]) == tru, f"flags differ env {e} step {t}"
            assert abs(rec["rew"][t, k] - r) <= REWARD_TOL * max(1.0, abs(r)), f"reward differs env {e} step {t}"
            assert int(rec["ne"][t, k]) == o.n_elements
            if te or tru:
                exp_t = np.zeros(18, np.float32) if obs is None else obs
                assert np.array_equal(rec["tobs"][t, k], exp_t), f"terminal obs differs env {e} step {t}"
                ep += 1
                episodes += 1
                poly = env.debug_polygon(e, ep)
                o = OracleEnv(poly["xy"], original_area=poly["original_area"])
                obs = o.obs()
            assert np.array_equal(rec["obs"][t, k], obs), f"obs differs env {e} step {t} (episode {ep})"
    assert episodes >= K // 2


def test_random_mode_element_log_returns_the_episode_polygon():
    """mg_get_elements in random-polygon mode: original vertex coordinates are the episode's generated polygon
    (round 1 returned zeros), inserted vertices follow, quads index into them."""
    import torch
    N = 16
    env = _mk(None, N, random_polygons=dict(min_verts=64, max_verts=256), seed=5, auto_reset=False)
    env.reset()
    polys = [env.get_state(e)["xy"].copy() for e in range(N)]
    for t in range(120):
        env.step(env.sample_actions(8, t))
    some = 0
    for e in range(N):
        quads, vxy, ne = env.get_elements(e)
        n0 = len(polys[e])
        assert np.array_equal(vxy[:n0], polys[e])
        assert len(quads) == ne and (quads < len(vxy)).all() and (quads >= 0).all()
        s = env.get_state(e)
        for j, vid in enumerate(s["ids"]):
            assert np.array_equal(vxy[vid], s["xy"][j])
        some += ne
    assert some > 10


def test_observation_delta_equals_full_writes():
    """mg_set_obs_delta: only the rows that changed are written, into the same persistent buffer -- identical to
    the full-write mode at every step, across a masked reset and a snapshot restore."""
    import torch
    doms, _ = load_domains()
    N, T = 192, 200
    a = _mk([doms["star"], doms["boundary16"]], N, obs_delta=True)
    b = _mk([doms["star"], doms["boundary16"]], N, obs_delta=False)
    assert torch.equal(a.reset(), b.reset())
    snap = None
    for t in range(T):
        act = a.sample_actions(3, t)
        ra, rb = a.step(act), b.step(act.clone())
        assert torch.equal(ra.obs, rb.obs), f"obs differ at step {t}"
        assert torch.equal(ra.reward, rb.reward) and torch.equal(ra.terminated, rb.terminated)
        d = (rb.terminated | rb.truncated).bool()
        assert torch.equal(ra.terminal_obs[d], rb.terminal_obs[d])
        if t == 70:
            m = torch.zeros(N, dtype=torch.uint8)
            m[::5] = 1
            assert torch.equal(a.reset(m), b.reset(m))
        if t == 100:
            snap = a.snapshot()
        if t == 140:
            oa = a.restore(snap)
            ob = b.restore(snap)
            assert torch.equal(oa, ob)


def test_snapshot_header_rejects_mismatched_handles():
    import torch
    from reinforcementlearning4meshgeneration_b200 import MeshgenError
    doms, _ = load_domains()
    a = _mk([doms["star"]], 8)
    a.reset()
    blob = a.snapshot()
    b = _mk([doms["star"]], 8, log_capacity=77)
    b.reset()
    with pytest.raises(MeshgenError):
        b.restore(blob)
    with pytest.raises(MeshgenError):
        a.restore(blob[: blob.numel() // 2])
    c = _mk([doms["star"], doms["boundary16"]], 8)
    c.reset()
    with pytest.raises(MeshgenError):
        c.restore(blob)
    r1 = _mk(None, 8, random_polygons=dict(min_verts=64, max_verts=128), seed=1)
    r2 = _mk(None, 8, random_polygons=dict(min_verts=64, max_verts=128), seed=2)
    r1.reset(); r2.reset()
    with pytest.raises(MeshgenError):
        r2.restore(r1.snapshot())
    a.restore(blob)          # the matching handle still loads it


def test_reset_then_host_step_is_ordered():
    """mg_reset / mg_step run on the caller's stream, mg_step_host on the library's private stream: a reset of a large
    random-polygon batch (about a millisecond of kernel time) directly followed by step_host must see the reset
    state (round-1 advisory: there was no ordering between the two streams)."""
    N = 16384
    kw = dict(random_polygons=dict(min_verts=64, max_verts=512), seed=9)
    rng = np.random.default_rng(0)
    act = rng.uniform(LOW_A, HIGH_A, size=(N, 3)).astype(np.float32)
    import torch
    ref = _mk(None, N, **kw)
    ref.reset()
    torch.cuda.synchronize()
    exp = ref.step_host(act)
    for _ in range(3):
        env = _mk(None, N, **kw)
        env.reset()                       # no synchronisation here
        got = env.step_host(act)
        for k in ("obs", "reward", "terminated", "truncated", "n_elements"):
            assert np.array_equal(got[k], exp[k]), k
        env.close()


def test_device_side_statistics_and_kernel_timing():
    import torch
    from reinforcementlearning4meshgeneration_b200.distributed import stats_from_tensor
    doms, _ = load_domains()
    env = _mk([doms["boundary16"]], 4096)
    env.reset()
    env.set_kernel_timing(True)
    for t in range(64):
        env.step(env.sample_actions(1, t))
    kt = env.kernel_times()
    env.set_kernel_timing(False)
    assert kt["steps"] == 64 and all(0 < kt[k] < 50 for k in ("screen_ms", "decide_ms", "update_ms", "observe_ms"))
    dev_stats = stats_from_tensor(env.stats_async())
    host_stats = env.stats()
    assert dev_stats == host_stats and host_stats["steps"] == 64 * 4096
    assert 0 < host_stats["ring_items"] < host_stats["steps"] and host_stats["successes"] <= host_stats["ring_items"]
    assert host_stats["sum_n_success"] <= host_stats["sum_n_ring"] <= host_stats["sum_n"]


def test_current_device_is_left_alone():
    """Every ABI call works on the handle's device and restores the caller's current device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    doms, _ = load_domains()
    torch.cuda.set_device(0)
    env = _mk([doms["star"]], 4, device="cuda:1")
    env.reset()
    env.step(env.sample_actions(0, 0))
    env.stats()
    assert torch.cuda.current_device() == 0
    x = torch.zeros(4, device="cuda")
    assert x.device.index == 0
    # host steps on the other device: plain launches, then the cached graph (captured and replayed on cuda:1)
    ref = _mk([doms["star"]], 4, device="cuda:0")
    ref.reset()
    ref.step(ref.sample_actions(0, 0))
    a = torch.zeros((4, 3), dtype=torch.float32).pin_memory()
    rng = np.random.default_rng(2)
    outs = [dict(obs=torch.zeros((4, 18)).pin_memory(), reward=torch.zeros(4, dtype=torch.float64).pin_memory(),
                 terminated=torch.zeros(4, dtype=torch.uint8).pin_memory(), truncated=torch.zeros(4, dtype=torch.uint8).pin_memory(),
                 terminal_obs=torch.zeros((4, 18)).pin_memory(), n_elements=torch.zeros(4, dtype=torch.int32).pin_memory()) for _ in range(2)]
    for t in range(8):
        a.copy_(torch.from_numpy(rng.uniform(LOW_A, HIGH_A, size=(4, 3)).astype(np.float32)))
        env.step_host(a, outs[0])
        ref.step_host(a, outs[1])
        assert torch.cuda.current_device() == 0
        for k in ("obs", "reward", "terminated", "truncated", "n_elements"):
            assert torch.equal(outs[0][k], outs[1][k]), (t, k)


def test_graph_captured_policy_and_env_rollout_equals_eager():
    """BASELINE config 5: the actor MLP (18-128-128-128-(3+3), sb3_algos.py:56-62) consuming the env's device-resident
    observation buffer and mg_step, captured together in ONE CUDA graph and replayed.  The env side must be bit for
    bit what an eager env produces from the same actions, and the actions must be the policy's output for the
    previous observation (cuBLAS may pick another GEMM algorithm under capture, so the policy is compared at float32
    round-off, the env exactly).  Exploration is the actor's reproducible variant (a function of the observation and
    of a device-side call counter that the graph increments itself)."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import sac_rollout
    torch.manual_seed(0)
    actor = sac_rollout.Actor().cuda()
    kw = dict(random_polygons=dict(min_verts=64, max_verts=256), seed=123)
    N, T = 2048, 150
    env = _mk(None, N, **kw)
    obs0 = env.reset().clone()
    rec = []
    sac_rollout.rollout(env, actor, T, graph=True, stochastic=False, record=rec)
    torch.cuda.synchronize()
    assert env.stats()["successes"] > 100
    tick0 = float(actor.tick) - T             # the three warm-up calls ran before the first replay, the capture ran nothing
    eager = _mk(None, N, **kw)
    prev = eager.reset().clone()
    assert torch.equal(prev, obs0)
    for t in range(T):
        act, obs, rew, done = rec[t]
        actor.tick.fill_(tick0 + t)
        assert torch.allclose(act, actor(prev, False), rtol=1e-3, atol=1e-4), f"graph policy output differs at step {t}"
        r = eager.step(act)
        assert torch.equal(r.obs, obs), f"graph replay differs from the eager env at step {t}"
        assert torch.equal(r.reward, rew) and torch.equal((r.terminated | r.truncated), done)
        prev = r.obs.clone()


@pytest.mark.parametrize("name", ["boundary0", "dolphine3", "easy1_1"])
def test_move_against_golden_trace_and_oracle(name):
    """BoudaryEnv.move() (E:459-594) on the CUDA env -- env 0 replays the golden trace recorded from the live
    reference, the other envs fresh action streams checked against the C oracle: observation (static point environment)
    bit-exact, done / is_complete / element count / boundary ids and coordinates / reference index exact, up to the step
    where the reference would call smooth_pave (option smooth_pave = 0: reported as `exhausted`, the way these traces were
    recorded), after which the env is reset.  The smoothing itself: test_move_with_smooth_pave_*."""
    import os
    import torch
    from helpers import GOLDEN
    from oracle.c_oracle import OracleEnv
    z = np.load(os.path.join(GOLDEN, f"move_{name}.npz"))
    tr = {k: z[k] for k in z.files}
    T = min(len(tr["type"]), 600)
    N = 4
    env = _mk([tr["xy0"]], N, auto_reset=False)
    env.set_option("smooth_pave", 0)
    env.reset()
    rng = np.random.default_rng(17)
    pol = np.stack([np.stack([rng.uniform(0.05, 0.5, T), rng.uniform(0.2, 2.9, T)], axis=1) for _ in range(N)], axis=1)   # [T, N, 2]
    typ = np.stack([rng.choice([0.1, 0.5, 0.9], size=T, p=[0.15, 0.7, 0.15]) for _ in range(N)], axis=1)
    pol[:, 0], typ[:, 0] = tr["polar"][:T], tr["type"][:T]
    oracles = [OracleEnv(tr["xy0"], original_area=float(tr["original_area"])) for _ in range(N)]
    n_acc = 0
    for t in range(T):
        r = env.move(pol[t], typ[t])
        obs, done, comp, exh, nel = (r[k].cpu().numpy() for k in ("obs", "done", "is_complete", "exhausted", "n_elements"))
        reset_mask = np.zeros(N, np.uint8)
        for e in range(N):
            oo, _, od, oinfo, osm = oracles[e].move(pol[t, e], typ[t, e])
            assert bool(exh[e]) == osm and bool(done[e]) == od and bool(comp[e]) == oinfo["is_complete"], f"{name}: flags differ t={t} env={e}"
            assert np.array_equal(obs[e], np.zeros(18, np.float32) if oo is None else oo), f"{name}: obs differs t={t} env={e}"
            assert int(nel[e]) == oracles[e].n_elements
            if e == 0:
                assert bool(tr["smooth"][t]) == osm and bool(tr["done"][t]) == od and int(tr["n_elements"][t]) == int(nel[0])
                assert np.array_equal(obs[0], tr["obs"][t])
            if t % 7 == 0 or od:
                s = env.get_state(e)
                ids, xy = oracles[e].boundary()
                assert s["n"] == oracles[e].n and np.array_equal(s["ids"], ids) and np.array_equal(s["xy"], xy), f"{name}: boundary differs t={t} env={e}"
                if oo is not None:
                    assert s["ref_index"] == oracles[e].ref_index
            if od:
                reset_mask[e] = 1
                oracles[e].reset()
        n_acc = max(n_acc, int(nel.max()))
        if reset_mask.any():
            env.reset(torch.from_numpy(reset_mask))
    assert n_acc > 20


def test_move_facade_like_the_data_generation_scripts():
    """general/EBRD.py:544: ``state, reward, done, info = env.move(action, round(type_values, 2))`` on the single-env facade."""
    from reinforcementlearning4meshgeneration_b200.boundary_env import BoudaryEnv
    from oracle.c_oracle import OracleEnv
    tr = load_trace("boundary0")
    env = BoudaryEnv(tr["xy0"])
    env.reset()
    o = OracleEnv(tr["xy0"], original_area=float(tr["original_area"]))
    o.set_smoothing(True)
    rng = np.random.default_rng(4)
    for t in range(200):
        p, ty = [float(rng.uniform(0.05, 0.5)), float(rng.uniform(0.2, 2.9))], float(rng.choice([0.1, 0.5, 0.9], p=[0.15, 0.7, 0.15]))
        state, reward, done, info = env.move(p, round(ty, 2))
        oo, _, od, oinfo, osm = o.move(p, round(ty, 2))
        assert reward == 0 and done == od and info["is_complete"] == oinfo["is_complete"] and not info.get("needs_smoothing", False)
        assert (state is None) == (oo is None) and (state is None or np.array_equal(state, oo))
        assert len(env.generated_meshes) == o.n_elements
        if done:
            env.reset()
            o.reset()
    env.close()


def test_results_do_not_depend_on_the_tuning_options():
    """mg_set_option only changes how a step is scheduled (side-stream resets, fused decision, grid sizes): random
    polygons with auto-reset across episode boundaries must come out bit-identical under every setting."""
    import torch
    N, T = 2048, 260
    kw = dict(random_polygons=dict(min_verts=32, max_verts=96), max_verts=96, seed=31, obs_delta=False)
    settings = [dict(), dict(reset_side=0), dict(fuse_decide=0), dict(update_blocks=3, observe_blocks=2, reset_blocks=1),
                dict(fuse_decide=0, reset_side=0), dict(pdl=0)]
    ref = None
    for opts in settings:
        env = _mk(None, N, **kw)
        for k, v in opts.items():
            env.set_option(k, v)
        env.reset()
        got = []
        for t in range(T):
            r = env.step(env.sample_actions(5, t))
            got.append([x.clone() for x in (r.obs, r.reward, r.terminated, r.truncated, r.terminal_obs, r.n_elements)])
        torch.cuda.synchronize()
        st = env.stats()
        assert st["episodes"] > N // 2, "the run must cross episode boundaries"
        if ref is None:
            ref, ref_stats = got, st
        else:
            for t in range(T):
                done = (ref[t][2] | ref[t][3]).bool()
                for i, (a, b) in enumerate(zip(ref[t], got[t])):
                    if i == 4:                       # terminal rows are defined where the episode ended
                        a, b = a[done], b[done]
                    assert torch.equal(a, b), (opts, t, i)
            for k, v in st.items():              # the two float sums are accumulated with atomics (order-dependent)
                if isinstance(v, float):
                    assert abs(v - ref_stats[k]) <= 1e-9 * max(1.0, abs(v)), (opts, k)
                else:
                    assert v == ref_stats[k], (opts, k)
        env.close()


def test_host_step_reads_pinned_actions_in_place():
    """mg_step_host with a pinned action buffer (read by the screen kernel over PCIe, no staging copy) against the
    same call with a pageable one (cudaMemcpyAsync into the staging buffer)."""
    import torch
    doms, _ = load_domains()
    N, T = 1024, 40
    envs = [_mk([doms["boundary16"], doms["test1"]], N) for _ in range(2)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(3)
    pinned = torch.empty((N, 3), dtype=torch.float32).pin_memory()
    for t in range(T):
        a = rng.uniform(LOW_A, HIGH_A, size=(N, 3)).astype(np.float32)
        pinned.copy_(torch.from_numpy(a))
        r0 = envs[0].step_host(a)
        r1 = envs[1].step_host(pinned)
        for k in ("obs", "reward", "terminated", "truncated", "n_elements"):
            assert np.array_equal(r0[k], r1[k]), (t, k)
    h2d, d2h = envs[1].last_host_bytes()
    assert h2d == N * 12 and d2h >= N * 10
    for e in envs:
        e.close()


def test_sac_learner_on_the_device_replay_buffer():
    """examples/sac_train.py: transitions of the batched env go through mg_replay_add into the device
    replay ring and the SAC learner fits its critics on them -- a fixed buffer, 300 gradient steps: the TD loss must fall
    and everything must stay finite; the policy then drives the env through the same tensors (no host copy)."""
    import importlib.util
    import os
    import torch
    spec = importlib.util.spec_from_file_location(
        "sac_train", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "sac_train.py"))
    sac = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sac)
    from reinforcementlearning4meshgeneration_b200.replay import DeviceReplayBuffer
    doms, _ = load_domains()
    env = _mk([doms["boundary16"]], 1024)
    torch.manual_seed(0)
    learner = sac.SAC(env.device, gamma=0.5)
    buf = DeviceReplayBuffer(env, 32)
    obs = env.reset().clone()
    for t in range(32):
        act = env.sample_actions(11, t)
        r = env.step(act)
        buf.add(obs, act, r)
        obs.copy_(r.obs)
    assert len(buf) == 32 * 1024
    losses = []
    for _ in range(300):
        lq, la = learner.update(buf.sample(512))
        losses.append(float(lq))
        assert torch.isfinite(lq) and torch.isfinite(la)
    first, last = np.mean(losses[:20]), np.mean(losses[-20:])
    assert last < 0.7 * first, (first, last)
    low = torch.from_numpy(LOW_A).to(env.device)
    high = torch.from_numpy(HIGH_A).to(env.device)
    with torch.no_grad():
        for _ in range(8):
            a = learner.actor(obs)[0]
            r = env.step((low + (a + 1) * 0.5 * (high - low)).contiguous())
            obs.copy_(r.obs)
    assert torch.isfinite(r.reward).all() and torch.isfinite(r.obs).all()
    env.close()


SMOOTH_TOL = 1e-9      # smoothed coordinates: device libm vs glibc (tan / cos / sqrt / atan2 of non-quantised arguments)


def _check_move_state(env, e, oracle, where):
    """front ids exact, front + interior coordinates within SMOOTH_TOL, element log exact"""
    s = env.get_state(e)
    ids, xy = oracle.boundary()
    assert s["n"] == oracle.n and np.array_equal(s["ids"], ids), f"{where}: front ids differ"
    assert np.max(np.abs(s["xy"] - xy)) <= SMOOTH_TOL, f"{where}: front coordinates differ by {np.max(np.abs(s['xy'] - xy)):.3e}"
    quads, vxy, _ = env.get_elements(e)
    ovx = oracle.vertex_xy()
    assert len(quads) == oracle.n_elements and vxy.shape == ovx.shape, f"{where}: mesh sizes differ"
    assert np.max(np.abs(vxy - ovx)) <= SMOOTH_TOL, f"{where}: vertex coordinates differ by {np.max(np.abs(vxy - ovx)):.3e}"


@pytest.mark.parametrize("name", ["boundary0", "tool", "bird"])
def test_move_with_smooth_pave_against_golden_trace_and_oracle(name):
    """SURVEY 8f-4, second half: where every reference candidate is on the not-valid list the reference smooths the whole
    mesh (smooth_pave, M:816-1140 / M:1284-1316) and goes on.  Env 0 replays the golden trace recorded from the live
    reference WITH its smoothing (oracle/sweep_smooth_vs_reference.py --record), the other envs fresh move streams checked
    against the C oracle (pinned on the same reference runs, bit-equal coordinates).  Contract: flags, element counts,
    front ids, reference index and the 4-decimal observations exact; coordinates of every vertex within 1e-9."""
    import os
    import torch
    from helpers import GOLDEN
    from oracle.c_oracle import OracleEnv
    z = np.load(os.path.join(GOLDEN, f"smooth_{name}.npz"))
    tr = {k: z[k] for k in z.files}
    T = len(tr["type"])
    N = 6
    env = _mk([tr["xy0"]], N, auto_reset=False, log_capacity=1024)
    env.reset()
    rng = np.random.default_rng(23)
    pol = np.stack([np.stack([rng.uniform(0.05, 0.5, T), rng.uniform(0.2, 2.9, T)], axis=1) for _ in range(N)], axis=1)   # [T, N, 2]
    typ = np.stack([rng.choice([0.1, 0.5, 0.9], size=T, p=[0.15, 0.7, 0.15]) for _ in range(N)], axis=1)
    pol[:, 0], typ[:, 0] = tr["polar"][:T], tr["type"][:T]
    oracles = [OracleEnv(tr["xy0"], original_area=float(tr["original_area"])) for _ in range(N)]
    for o in oracles:
        o.set_smoothing(True)
    n_smooth = 0
    for t in range(T):
        r = env.move(pol[t], typ[t])
        obs, done, comp, exh, nel = (r[k].cpu().numpy() for k in ("obs", "done", "is_complete", "exhausted", "n_elements"))
        reset_mask = np.zeros(N, np.uint8)
        for e in range(N):
            oo, _, od, oinfo, osm = oracles[e].move(pol[t, e], typ[t, e])
            where = f"{name} t={t} env={e} (smoothing {osm})"
            assert not exh[e], f"{where}: the kernel could not smooth"
            assert bool(done[e]) == od and bool(comp[e]) == oinfo["is_complete"], f"{where}: flags differ"
            assert np.array_equal(obs[e], np.zeros(18, np.float32) if oo is None else oo), f"{where}: observation differs"
            assert int(nel[e]) == oracles[e].n_elements
            n_smooth += osm
            if e == 0:
                assert bool(tr["smooth"][t]) == osm and bool(tr["done"][t]) == od and int(tr["n_elements"][t]) == int(nel[0])
                assert np.array_equal(obs[0], tr["obs"][t])
                nv = int(tr["n_vertices"][t])
                _, vxy, _ = env.get_elements(0)
                assert vxy.shape[0] == nv and np.max(np.abs(vxy - tr["vertex_xy"][t, :nv])) <= SMOOTH_TOL, f"{where}: golden vertex coordinates"
                s0 = env.get_state(0)
                assert s0["ids"].tolist() == tr["boundary_ids"][t, :s0["n"]].tolist() and s0["n"] == int(tr["n_boundary"][t])
            if osm or od or t % 5 == 0:
                _check_move_state(env, e, oracles[e], where)
                if oo is not None:
                    assert env.get_state(e)["ref_index"] == oracles[e].ref_index
            if od:
                reset_mask[e] = 1
                oracles[e].reset()
        if reset_mask.any():
            env.reset(torch.from_numpy(reset_mask))
    assert n_smooth >= 5
    env.close()


def test_move_with_smooth_pave_on_more_domains_and_coordinate_agreement():
    """Differential run of move() + smooth_pave over six more reference domains (8 envs x 250 moves each, fresh streams)
    against the oracle; also reports how far the smoothed coordinates are from the reference's (device libm vs glibc)."""
    import torch
    from oracle.c_oracle import OracleEnv
    doms, areas = load_domains()
    rng = np.random.default_rng(41)
    worst, n_smooth = 0.0, 0
    for name in ["star", "half_wheel", "dolphine3", "easy1_1", "boundary16", "test1"]:
        xy = doms[name]
        N, T = 8, 250
        env = _mk([xy], N, auto_reset=False, log_capacity=2048)
        env.reset()
        oracles = [OracleEnv(xy, original_area=float(areas[name])) for _ in range(N)]
        for o in oracles:
            o.set_smoothing(True)
        for t in range(T):
            pol = np.stack([rng.uniform(0.05, 0.5, N), rng.uniform(0.2, 2.9, N)], axis=1)
            typ = rng.choice([0.1, 0.5, 0.9], size=N, p=[0.15, 0.7, 0.15])
            r = env.move(pol, typ)
            obs, done, comp, exh, nel = (r[k].cpu().numpy() for k in ("obs", "done", "is_complete", "exhausted", "n_elements"))
            reset_mask = np.zeros(N, np.uint8)
            for e in range(N):
                oo, _, od, oinfo, osm = oracles[e].move(pol[e], typ[e])
                where = f"{name} t={t} env={e} (smoothing {osm})"
                assert not exh[e] and bool(done[e]) == od and bool(comp[e]) == oinfo["is_complete"], f"{where}: flags differ"
                assert np.array_equal(obs[e], np.zeros(18, np.float32) if oo is None else oo), f"{where}: observation differs"
                assert int(nel[e]) == oracles[e].n_elements
                if osm:
                    n_smooth += 1
                    _check_move_state(env, e, oracles[e], where)
                    _, vxy, _ = env.get_elements(e)
                    worst = max(worst, float(np.max(np.abs(vxy - oracles[e].vertex_xy()))))
                if od:
                    reset_mask[e] = 1
                    oracles[e].reset()
            if reset_mask.any():
                env.reset(torch.from_numpy(reset_mask))
        env.close()
    print(f"smooth_pave: {n_smooth} smoothings, largest coordinate deviation from the oracle {worst:.3e}")
    assert n_smooth >= 20 and worst <= SMOOTH_TOL


def test_move_with_smooth_pave_on_random_polygons():
    """move() + smooth_pave in random-polygon mode: the original polygon of an episode is not stored there, the smoothing
    kernel regenerates it from the generator's counter.  Polygons read back from the device, replayed through the oracle."""
    import torch
    from oracle.c_oracle import OracleEnv
    N, T = 8, 220
    env = _mk(None, N, random_polygons=dict(min_verts=32, max_verts=96), max_verts=96, seed=77, auto_reset=False, log_capacity=1024)
    env.reset()
    polys = [env.debug_polygon(e) for e in range(N)]
    oracles = [OracleEnv(p["xy"], original_area=p["original_area"]) for p in polys]
    for o in oracles:
        o.set_smoothing(True)
    rng = np.random.default_rng(5)
    n_smooth = 0
    for t in range(T):
        pol = np.stack([rng.uniform(0.05, 0.5, N), rng.uniform(0.2, 2.9, N)], axis=1)
        typ = rng.choice([0.1, 0.5, 0.9], size=N, p=[0.15, 0.7, 0.15])
        r = env.move(pol, typ)
        obs, done, comp, exh, nel = (r[k].cpu().numpy() for k in ("obs", "done", "is_complete", "exhausted", "n_elements"))
        reset_mask = np.zeros(N, np.uint8)
        for e in range(N):
            oo, _, od, oinfo, osm = oracles[e].move(pol[e], typ[e])
            where = f"random polygon t={t} env={e} (smoothing {osm})"
            assert not exh[e] and bool(done[e]) == od and bool(comp[e]) == oinfo["is_complete"], f"{where}: flags differ"
            assert np.array_equal(obs[e], np.zeros(18, np.float32) if oo is None else oo), f"{where}: observation differs"
            assert int(nel[e]) == oracles[e].n_elements
            if osm:
                n_smooth += 1
                _check_move_state(env, e, oracles[e], where)
            if od:
                reset_mask[e] = 1
                oracles[e].reset()
        if reset_mask.any():
            env.reset(torch.from_numpy(reset_mask))
    assert n_smooth >= 3
    env.close()


def test_synthetic_policy_with_a_device_side_step_index_replays_from_a_graph():
    """mg_sample_actions_seq (the step index lives in device memory and is advanced by the launch itself): a loop of
    policy + env step captured once and replayed -- what bench.py times -- draws the same actions and produces the
    same rollout (observations, rewards, statistics) as eager launches with the step index passed from the host."""
    import torch
    N, T, G = 2048, 60, 4
    gen = dict(min_coarse=5, max_coarse=10, min_verts=16, max_verts=48)
    eager = _mk(None, N, random_polygons=gen, seed=5)
    eager.reset()
    exp_a, exp_obs, exp_rew = [], [], []
    for t in range(T):
        a = eager.sample_actions(77, 100 + t)
        exp_a.append(a.clone())
        r = eager.step(a)
        exp_obs.append(r.obs.clone()); exp_rew.append(r.reward.clone())
    exp_stats = eager.stats()
    env = _mk(None, N, random_polygons=gen, seed=5)
    env.reset()
    ctr = torch.tensor([100, 0], dtype=torch.int64, device=env.device)
    a_log = torch.zeros((G, N, 3), device=env.device)
    obs_log = torch.zeros((G, N, 18), device=env.device)
    rew_log = torch.zeros((G, N), dtype=torch.float64, device=env.device)
    snap = env.snapshot()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        env.step(env.sample_actions(77, ctr))                    # first use outside the capture, then rewind
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    assert ctr.tolist() == [101, 0]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(G):
            a = env.sample_actions(77, ctr)
            a_log[k].copy_(a)
            r = env.step(a)
            obs_log[k].copy_(r.obs); rew_log[k].copy_(r.reward)
    env.restore(snap)
    env.stats(reset=True)
    ctr.copy_(torch.tensor([100, 0]))
    for rep in range(T // G):
        g.replay()
        for k in range(G):
            t = rep * G + k
            assert torch.equal(a_log[k], exp_a[t]), f"actions differ at step {t}"
            assert torch.equal(obs_log[k], exp_obs[t]), f"obs differ at step {t}"
            assert torch.equal(rew_log[k], exp_rew[t]), f"reward differs at step {t}"
    assert ctr.tolist() == [100 + T, 0]
    got = env.stats()
    for k, v in exp_stats.items():                # (the two float sums are accumulated with atomics: order-dependent rounding)
        assert got[k] == v if isinstance(v, int) else abs(got[k] - v) <= 1e-9 * max(1.0, abs(v)), k
    assert got["successes"] > 0


def test_host_steps_replayed_from_cached_graphs_equal_plain_launches():
    """mg_step_host replays the launches of a step from a CUDA graph once it has seen the same buffers twice
    (mg_set_option "host_graph"): rotating pinned action buffers, persistent pinned result buffers in delta mode,
    random polygons with auto-reset, a partial reset and an option change in the middle -- every result array equals
    the one of an env that launches its kernels one by one."""
    import torch
    N, T = 4096, 90
    gen = dict(min_coarse=5, max_coarse=10, min_verts=16, max_verts=48)
    envs = [_mk(None, N, random_polygons=gen, seed=11) for _ in range(2)]
    envs[1].set_option("host_graph", 0)
    rng = np.random.default_rng(4)
    acts = [torch.empty((N, 3), dtype=torch.float32).pin_memory() for _ in range(3)]
    outs = []
    for e in envs:
        e.reset()
        outs.append(dict(obs=torch.zeros((N, 18), dtype=torch.float32).pin_memory(), reward=torch.zeros(N, dtype=torch.float64).pin_memory(),
                         terminated=torch.zeros(N, dtype=torch.uint8).pin_memory(), truncated=torch.zeros(N, dtype=torch.uint8).pin_memory(),
                         terminal_obs=torch.zeros((N, 18), dtype=torch.float32).pin_memory(),
                         n_elements=torch.zeros(N, dtype=torch.int32).pin_memory()))
    launches0 = [e.launch_count for e in envs]
    for t in range(T):
        a = acts[t % 3]
        a.copy_(torch.from_numpy(rng.uniform(LOW_A, HIGH_A, size=(N, 3)).astype(np.float32)))
        if t == 40:                                  # state changes behind the graphs' back: partial reset on the caller's stream
            mask = torch.zeros(N, dtype=torch.bool, device=envs[0].device)
            mask[::7] = True
            for e in envs:
                e.reset(mask)
        if t == 60:                                  # a different launch configuration: new graphs
            for e in envs:
                e.set_option("pdl", 0)
        for e, o in zip(envs, outs):
            e.step_host(a, o)
        for k in ("obs", "reward", "terminated", "truncated", "n_elements"):
            assert torch.equal(outs[0][k], outs[1][k]), (t, k)
    assert envs[0].launch_count - launches0[0] == envs[1].launch_count - launches0[1]
    s0, s1 = envs[0].stats(), envs[1].stats()
    assert all(s0[k] == s1[k] for k in s0 if isinstance(s0[k], int)) and s0["steps"] == N * T and s0["episodes"] > 0
    for e in envs:
        e.close()


def test_host_step_begin_end_halves_equal_one_env():
    """mg_step_host_begin / _end: two handles that hold the two halves of the envs (env_id_offset keeps the generator's
    counters), their host steps interleaved -- begin A, begin B, end A, ... -- give the rows a single handle gives; a
    second begin without end, and a device step while a host step is in flight, are refused."""
    import torch
    from reinforcementlearning4meshgeneration_b200._lib import MeshgenError
    N, T = 2048, 50
    gen = dict(min_coarse=5, max_coarse=10, min_verts=16, max_verts=48)
    full = _mk(None, N, random_polygons=gen, seed=9)
    parts = [_mk(None, N // 2, random_polygons=gen, seed=9, env_id_offset=p * (N // 2)) for p in range(2)]

    def outs(n):
        return dict(obs=torch.zeros((n, 18), dtype=torch.float32).pin_memory(), reward=torch.zeros(n, dtype=torch.float64).pin_memory(),
                    terminated=torch.zeros(n, dtype=torch.uint8).pin_memory(), truncated=torch.zeros(n, dtype=torch.uint8).pin_memory(),
                    terminal_obs=torch.zeros((n, 18), dtype=torch.float32).pin_memory(), n_elements=torch.zeros(n, dtype=torch.int32).pin_memory())
    for e in [full] + parts:
        e.reset()
    o_full, o_parts = outs(N), [outs(N // 2) for _ in range(2)]
    rng = np.random.default_rng(12)
    a_full = torch.empty((N, 3), dtype=torch.float32).pin_memory()
    a_parts = [torch.empty((N // 2, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    for t in range(T):
        a = torch.from_numpy(rng.uniform(LOW_A, HIGH_A, size=(N, 3)).astype(np.float32))
        a_full.copy_(a)
        full.step_host(a_full, o_full)
        for p in range(2):
            a_parts[p].copy_(a[p * (N // 2):(p + 1) * (N // 2)])
            parts[p].step_host_begin(a_parts[p], o_parts[p])
        if t == 3:
            with pytest.raises(MeshgenError):
                parts[0].step_host_begin(a_parts[0], o_parts[0])
            with pytest.raises(MeshgenError):
                parts[0].step(torch.zeros((N // 2, 3), device=parts[0].device))
        for p in range(2):
            parts[p].step_host_end()
        for k in ("obs", "reward", "terminated", "truncated", "n_elements"):
            assert torch.equal(torch.cat([o_parts[0][k], o_parts[1][k]]), o_full[k]), (t, k)
    with pytest.raises(MeshgenError):
        parts[0].step_host_end()
    for e in [full] + parts:
        e.close()
