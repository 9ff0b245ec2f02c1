"""CPU: the C oracle (oracle/boundary_env_oracle.c) reproduces the golden traces recorded from the
live reference env (oracle/record_golden.py) -- bit-exact, rewards included."""
import numpy as np
import pytest

from helpers import TRACES, assert_rollout_matches, load_trace
from oracle.c_oracle import OracleEnv, lib


def test_round_helpers_against_python():
    L = lib()
    assert L.oracle_selftest_round(2026, 300000) == 0
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-20, 20, 20000), np.round(rng.uniform(-7, 7, 20000), 5),
                         (np.round(rng.uniform(-7, 7, 20000), 4) * 1e4 + 0.5) / 1e4])
    for x in xs:
        x = float(x)
        assert L.oracle_py_round4(x) == round(x, 4)
        assert L.oracle_np_round4(x) == float(round(np.float64(x), 4))
    for x in rng.uniform(-7, 7, 20000).astype(np.float32):
        assert np.float32(L.oracle_np_round4f(float(x))) == round(np.float32(x), 4)


def test_py_sum_matches_builtin_sum():
    import ctypes as C
    import sys
    L = lib()
    L.oracle_py_sum.restype = C.c_double
    L.oracle_py_sum.argtypes = [C.c_void_p, C.c_int]
    if sys.version_info < (3, 12):
        pytest.skip("builtin sum() is only compensated from CPython 3.12")
    rng = np.random.default_rng(1)
    for _ in range(2000):
        k = int(rng.integers(1, 9))
        x = rng.uniform(0, 3, k)
        assert L.oracle_py_sum(x.ctypes.data, k) == sum([float(v) for v in x])


@pytest.mark.parametrize("name", TRACES)
def test_oracle_reproduces_golden_trace(name):
    tr = load_trace(name)
    env = OracleEnv(tr["xy0"], original_area=float(tr["original_area"]))
    assert np.array_equal(env.obs(), tr["reset_obs"])
    assert env.ref_index == int(tr["reset_ref_index"])
    assert env.base_length == float(tr["reset_base_length"])
    assert tuple(env.area_range()) == tuple(tr["area_range"])
    T = len(tr["reward"])
    ids_ok = True
    # step one by one so that the pre-reset state (boundary ids, reference index) can be checked
    got = dict(obs=np.zeros((T, 18), np.float32), terminal_obs=np.zeros((T, 18), np.float32), reward=np.zeros(T),
               terminated=np.zeros(T, np.uint8), truncated=np.zeros(T, np.uint8), n_elements=np.zeros(T, np.int32),
               success=np.zeros(T, np.uint8), n_boundary=np.zeros(T, np.int32), ref_index=np.zeros(T, np.int32))
    for i in range(T):
        obs, r, te, tr_, info = env.step(tr["actions"][i])
        got["reward"][i], got["terminated"][i], got["truncated"][i] = r, te, tr_
        got["n_elements"][i] = env.n_elements
        got["success"][i] = env.last_info()["success"]
        got["n_boundary"][i] = env.n
        got["ref_index"][i] = env.ref_index
        ids, xy = env.boundary()
        exp_ids = tr["ids"][i][: tr["n_boundary"][i]]
        ids_ok = ids_ok and np.array_equal(ids, exp_ids)
        assert ids_ok, f"{name}: boundary vertex ids differ at step {i}"
        assert env.base_length == tr["base_length"][i] and env.current_area == tr["current_area"][i]
        if not np.isnan(tr["new_xy"][i, 0]):
            j = int(np.argmax(ids))
            assert np.array_equal(xy[j], tr["new_xy"][i]), f"{name}: inserted vertex coords differ at step {i}"
        if te or tr_:
            got["terminal_obs"][i] = 0 if obs is None else obs
            assert (obs is None) == bool(tr["obs_none"][i])
            obs = env.reset()
        got["obs"][i] = obs
    assert_rollout_matches(got, tr, f"oracle vs golden[{name}]", reward_exact=True)


def test_oracle_rollout_equals_stepwise():
    tr = load_trace("boundary0")
    env = OracleEnv(tr["xy0"], original_area=float(tr["original_area"]))
    out = env.rollout(tr["actions"])
    assert_rollout_matches(out, tr, "oracle rollout vs golden[boundary0]", reward_exact=True)
