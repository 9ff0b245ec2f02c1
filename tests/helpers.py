"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOW = np.array([-1.0, -1.5, 0.0], dtype=np.float32)
HIGH = np.array([1.0, 1.5, 1.5], dtype=np.float32)
TRACES = ["boundary0", "boundary16", "boundary15", "test1", "dolphine3", "easy1_1", "half_wheel", "star"]


def load_domains():
    z = np.load(os.path.join(GOLDEN, "domains.npz"))
    doms = {k: z[k] for k in z.files if not k.startswith("area__")}
    areas = {k[6:]: float(z[k]) for k in z.files if k.startswith("area__")}
    return doms, areas


def load_trace(name):
    z = np.load(os.path.join(GOLDEN, f"trace_{name}.npz"))
    return {k: z[k] for k in z.files}


def action_stream(seed, T):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(LOW, HIGH).astype(np.float32) for _ in range(T)])


def rel_close(a, b, tol=1e-9):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))


def assert_rollout_matches(got, exp, what, reward_tol=1e-9, reward_exact=False):
    """got/exp: dicts with obs, reward, terminated, truncated, n_elements, terminal_obs (+ optional n_boundary,
    ref_index, success).  Discrete outputs and float32 observations must be bit-exact."""
    T = len(exp["reward"])
    for key in ("terminated", "truncated", "n_elements", "success", "n_boundary", "ref_index"):
        if key in got and key in exp:
            g, e = np.asarray(got[key]).astype(np.int64), np.asarray(exp[key]).astype(np.int64)
            bad = np.nonzero(g != e)[0]
            assert bad.size == 0, f"{what}: {key} differs first at step {bad[0]}: got {g[bad[0]]} expected {e[bad[0]]}"
    done = (np.asarray(exp["terminated"]).astype(bool) | np.asarray(exp["truncated"]).astype(bool))
    for key in ("obs", "terminal_obs"):
        if key in got and key in exp:
            g, e = np.asarray(got[key], np.float32), np.asarray(exp[key], np.float32)
            bad = (g.view(np.uint32) != e.view(np.uint32)).any(axis=1) & ~((g == e).all(axis=1))
            if key == "terminal_obs":
                bad &= done            # the terminal observation is only defined where the episode ended
            bad = np.nonzero(bad)[0]
            assert bad.size == 0, f"{what}: {key} differs first at step {bad[0]}:\n got {g[bad[0]]}\n exp {e[bad[0]]}"
    g, e = np.asarray(got["reward"], np.float64), np.asarray(exp["reward"], np.float64)
    ok = (g == e) if reward_exact else rel_close(g, e, reward_tol)
    bad = np.nonzero(~ok)[0]
    assert bad.size == 0, f"{what}: reward differs first at step {bad[0]}: got {g[bad[0]]!r} expected {e[bad[0]]!r}"
    return T
