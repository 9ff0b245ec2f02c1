"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper over oracle/liboracle.so (the C restatement of
the reference BoudaryEnv, oracle/boundary_env_oracle.c).  Used by tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "boundary_env_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, f64 = C.c_void_p, C.c_int, C.c_double
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [vp, i32, f64]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_reset.argtypes = [vp]
        L.oracle_step.argtypes = [vp, vp, vp, vp, vp]
        for name in ("oracle_n", "oracle_n_elements", "oracle_ref_index", "oracle_failed_num",
                     "oracle_obs_none", "oracle_crashed", "oracle_n_candidates", "oracle_n_vertices"):
            getattr(L, name).restype = i32
            getattr(L, name).argtypes = [vp]
        for name in ("oracle_base_length", "oracle_current_area", "oracle_original_area"):
            getattr(L, name).restype = f64
            getattr(L, name).argtypes = [vp]
        L.oracle_area_range.argtypes = [vp, vp]
        L.oracle_obs.argtypes = [vp, vp]
        L.oracle_last_info.argtypes = [vp, vp]
        L.oracle_boundary.argtypes = [vp, vp, vp]
        L.oracle_candidates.argtypes = [vp, vp, vp]
        L.oracle_elements.argtypes = [vp, vp]
        L.oracle_vertex_xy.argtypes = [vp, vp]
        L.oracle_move.argtypes = [vp, vp, f64, vp, vp, vp]
        L.oracle_n_excluded.restype = i32
        L.oracle_n_excluded.argtypes = [vp]
        L.oracle_set_smoothing.argtypes = [vp, i32]
        L.oracle_n_smoothings.restype = i32
        L.oracle_n_smoothings.argtypes = [vp]
        L.oracle_rollout.argtypes = [vp, vp, i32] + [vp] * 9
        L.oracle_run_random.restype = C.c_long
        L.oracle_run_random.argtypes = [vp, C.c_uint64, C.c_long, vp, vp, vp]
        L.oracle_selftest_round.restype = C.c_long
        L.oracle_selftest_round.argtypes = [C.c_uint64, C.c_long]
        L.oracle_py_round4.restype = f64
        L.oracle_py_round4.argtypes = [f64]
        L.oracle_np_round4.restype = f64
        L.oracle_np_round4.argtypes = [f64]
        L.oracle_np_round4f.restype = C.c_float
        L.oracle_np_round4f.argtypes = [C.c_float]
        L.oracle_cw_angle.restype = f64
        L.oracle_cw_angle.argtypes = [vp]
        L.oracle_is_cross.restype = i32
        L.oracle_is_cross.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def numpy_poly_area(xy: np.ndarray) -> float:
    """Boundary2D.poly_area exactly as the reference evaluates it
    (v2/src/mesh_rl/components_core.py:485-487), including its np.dot summation order."""
    xy = np.asarray(xy, dtype=np.float64)
    return float(0.5 * np.abs(np.dot(xy[:, 0], np.roll(xy[:, 1], 1)) - np.dot(xy[:, 1], np.roll(xy[:, 0], 1))))


class OracleEnv:
    """Single-environment CPU oracle with the reference's reset/step semantics."""

    def __init__(self, xy, original_area: float | None = None):
        self.xy0 = np.ascontiguousarray(np.asarray(xy, dtype=np.float64))
        self.n0 = len(self.xy0)
        if original_area is None:
            original_area = numpy_poly_area(self.xy0)
        self._h = lib().oracle_create(_p(self.xy0), self.n0, float(original_area))

    def __del__(self):
        try:
            if self._h:
                lib().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- reference API ------------------------------------------------
    def reset(self):
        lib().oracle_reset(self._h)
        return self.obs()

    def step(self, action):
        a = np.ascontiguousarray(np.asarray(action, dtype=np.float32))
        r = C.c_double()
        te, tr = C.c_int(), C.c_int()
        lib().oracle_step(self._h, _p(a), C.byref(r), C.byref(te), C.byref(tr))
        obs = None if lib().oracle_obs_none(self._h) else self.obs()
        return obs, r.value, bool(te.value), bool(tr.value), {"is_complete": not tr.value}

    def set_smoothing(self, enabled=True):
        """move(): run the restated smooth_pave (M:816-821) where the reference does, instead of stopping there."""
        lib().oracle_set_smoothing(self._h, 1 if enabled else 0)

    @property
    def n_smoothings(self):
        return lib().oracle_n_smoothings(self._h)

    def move(self, new_point, type):
        """E:459-594 move((r, phi), type) -> (obs | None, 0, done, {"is_complete": ...}, needs_smoothing).
        needs_smoothing: the reference called smooth_pave in this move (every candidate was excluded); without
        set_smoothing() the oracle stops there (done)."""
        p = np.ascontiguousarray(np.asarray(new_point, dtype=np.float64))
        d, c, sm = C.c_int(), C.c_int(), C.c_int()
        lib().oracle_move(self._h, _p(p), float(type), C.byref(d), C.byref(c), C.byref(sm))
        obs = None if lib().oracle_obs_none(self._h) else self.obs()
        return obs, 0, bool(d.value), {"is_complete": bool(c.value)}, bool(sm.value)

    @property
    def n_excluded(self):
        return lib().oracle_n_excluded(self._h)

    # ---- state views --------------------------------------------------
    def obs(self):
        o = np.empty(18, dtype=np.float32)
        lib().oracle_obs(self._h, _p(o))
        return o

    @property
    def n(self):
        return lib().oracle_n(self._h)

    @property
    def n_elements(self):
        return lib().oracle_n_elements(self._h)

    @property
    def ref_index(self):
        return lib().oracle_ref_index(self._h)

    @property
    def failed_num(self):
        return lib().oracle_failed_num(self._h)

    @property
    def crashed(self):
        return bool(lib().oracle_crashed(self._h))

    @property
    def base_length(self):
        return lib().oracle_base_length(self._h)

    @property
    def current_area(self):
        return lib().oracle_current_area(self._h)

    @property
    def original_area(self):
        return lib().oracle_original_area(self._h)

    def area_range(self):
        out = np.empty(2, dtype=np.float64)
        lib().oracle_area_range(self._h, _p(out))
        return out

    def last_info(self):
        out = np.empty(4, dtype=np.int32)
        lib().oracle_last_info(self._h, _p(out))
        return dict(rule=int(out[0]), success=int(out[1]), inside=int(out[2]), existing=int(out[3]))

    def boundary(self):
        n = self.n
        ids = np.empty(n, dtype=np.int32)
        xy = np.empty((n, 2), dtype=np.float64)
        lib().oracle_boundary(self._h, _p(ids), _p(xy))
        return ids, xy

    def candidates(self):
        m = lib().oracle_n_candidates(self._h)
        ids = np.empty(m, dtype=np.int32)
        keys = np.empty(m, dtype=np.float64)
        lib().oracle_candidates(self._h, _p(ids), _p(keys))
        return ids, keys

    def elements(self):
        m = self.n_elements
        out = np.empty((m, 4), dtype=np.int32)
        lib().oracle_elements(self._h, _p(out))
        return out

    def vertex_xy(self):
        m = lib().oracle_n_vertices(self._h)
        out = np.empty((m, 2), dtype=np.float64)
        lib().oracle_vertex_xy(self._h, _p(out))
        return out

    # ---- batched helpers ----------------------------------------------
    def rollout(self, actions):
        """T steps with auto-reset; returns a dict of per-step arrays (VecEnv convention)."""
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.float32).reshape(-1, 3))
        T = len(a)
        out = dict(
            obs=np.empty((T, 18), np.float32), reward=np.empty(T, np.float64),
            terminated=np.empty(T, np.uint8), truncated=np.empty(T, np.uint8),
            n_elements=np.empty(T, np.int32), n_boundary=np.empty(T, np.int32),
            ref_index=np.empty(T, np.int32), terminal_obs=np.empty((T, 18), np.float32),
            success=np.empty(T, np.uint8))
        lib().oracle_rollout(self._h, _p(a), T, _p(out["obs"]), _p(out["reward"]), _p(out["terminated"]),
                             _p(out["truncated"]), _p(out["n_elements"]), _p(out["n_boundary"]),
                             _p(out["ref_index"]), _p(out["terminal_obs"]), _p(out["success"]))
        return out

    def run_random(self, seed: int, steps: int):
        ns, ne, sn = C.c_long(), C.c_long(), C.c_double()
        lib().oracle_run_random(self._h, seed, steps, C.byref(ns), C.byref(ne), C.byref(sn))
        return dict(steps=steps, success=ns.value, episodes=ne.value, sum_n=sn.value)
