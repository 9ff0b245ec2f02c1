"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference environment.

Imports the unmodified reference ``BoudaryEnv`` from ``/root/reference`` (v2 copy,
``v2/src/mesh_rl/envs/boundary_env.py:34``) inside this container so that

* ``oracle/record_golden.py`` can record golden traces into ``tests/golden/``;
* ``tests/`` can validate the C restatement (``oracle/boundary_env_oracle.c``) against
  the live reference whenever ``/root/reference`` is present (it is NOT on the GPU box).

Nothing in the product package may import this module.

The reference hard-imports packages that are absent here (gymnasium, matplotlib,
stable_baselines3); they are never *used* by reset/step, so empty stand-in modules
are injected through ``sys.modules`` and the ``mesh_rl`` package ``__init__`` (which
imports the SB3 trainers, ``v2/src/mesh_rl/__init__.py:14-18``) is bypassed by
pre-registering an empty package object with the right ``__path__``.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MESHGEN_REFERENCE_ROOT", "/root/reference")
_V2_SRC = os.path.join(REFERENCE_ROOT, "v2", "src")

LOW = np.array([-1.0, -1.5, 0.0], dtype=np.float32)
HIGH = np.array([1.0, 1.5, 1.5], dtype=np.float32)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_V2_SRC, "mesh_rl", "envs", "boundary_env.py"))


class _Stub(types.ModuleType):
    """Module whose every (non-dunder) attribute is another inert stub/callable."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        value = _Stub(self.__name__ + "." + name)
        setattr(self, name, value)
        return value

    def __call__(self, *args, **kwargs):  # pragma: no cover - never hit by reset/step
        return None


class _Box:
    """Minimal gymnasium.spaces.Box stand-in (same shape as the reference tests' stub,
    v2/tests/mesh_rl/test_boundary_env_equiv.py:31-45)."""

    def __init__(self, low, high, shape=None, dtype=None):
        if shape is not None:
            self.low = np.full(tuple(shape), low, dtype=np.float32)
            self.high = np.full(tuple(shape), high, dtype=np.float32)
        else:
            self.low = np.array(low, dtype=np.float32)
            self.high = np.array(high, dtype=np.float32)
        self.shape = self.low.shape
        self.dtype = dtype or np.float32

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class _Env:
    pass


_loaded = {}


def _install_stubs() -> None:
    def ensure(name, mod=None):
        if name not in sys.modules:
            sys.modules[name] = mod if mod is not None else _Stub(name)
        return sys.modules[name]

    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        try:
            __import__(name)
        except Exception:
            ensure(name)
    for pkg in ("gymnasium", "gym"):
        try:
            __import__(pkg)
        except Exception:
            m = types.ModuleType(pkg)
            sp = types.ModuleType(pkg + ".spaces")
            sp.Box = _Box
            m.Env = _Env
            m.spaces = sp
            ensure(pkg, m)
            ensure(pkg + ".spaces", sp)
    try:
        __import__("stable_baselines3.common.env_checker")
    except Exception:
        ensure("stable_baselines3")
        ensure("stable_baselines3.common")
        ec = types.ModuleType("stable_baselines3.common.env_checker")
        ec.check_env = lambda *a, **k: None
        ensure("stable_baselines3.common.env_checker", ec)


def load_reference():
    """Return a namespace with the reference's BoudaryEnv, Vertex, Segment, Boundary2D,
    read_polygon (all the unmodified reference objects)."""
    if "ns" in _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if _V2_SRC not in sys.path:
        sys.path.insert(0, _V2_SRC)
    if "mesh_rl" not in sys.modules:
        pkg = types.ModuleType("mesh_rl")
        pkg.__path__ = [os.path.join(_V2_SRC, "mesh_rl")]
        sys.modules["mesh_rl"] = pkg
    from mesh_rl.envs.boundary_env import BoudaryEnv  # type: ignore
    from mesh_rl.components_core import Vertex, Segment, Boundary2D, Mesh, PointEnvironment  # type: ignore
    from mesh_rl.geometry import read_polygon  # type: ignore

    ns = types.SimpleNamespace(
        BoudaryEnv=BoudaryEnv, Vertex=Vertex, Segment=Segment, Boundary2D=Boundary2D,
        Mesh=Mesh, PointEnvironment=PointEnvironment, read_polygon=read_polygon,
    )
    _loaded["ns"] = ns
    return ns


def domain_path(name: str) -> str:
    return os.path.join(REFERENCE_ROOT, "ui", "domains", name + ".json")


def load_domain_xy(name: str) -> np.ndarray:
    """(n,2) float64 coordinates exactly as the reference's read_polygon builds them
    (v2/src/mesh_rl/geometry.py:44-46: px / 100.0)."""
    with open(domain_path(name), "r", encoding="utf-8") as fr:
        pts = json.loads(fr.readline())
    return np.array([[p[0] / 100.0, p[1] / 100.0] for p in pts], dtype=np.float64)


# general/polygon.py:79-83 -- boundary(index=0): the 30-vertex config-1 polygon.
BOUNDARY0_XY = np.array(
    [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 6), (2, 6), (3, 6), (4, 6), (5, 6), (6, 6),
     (7, 5), (8, 4), (9, 3), (10, 2), (11, 1), (12, 0), (11, -1), (10, -2), (9, -3), (8, -4), (7, -5),
     (6, -6), (5, -5), (4, -4), (3, -3), (2, -2), (1, -1), (0, 0)], dtype=np.float64)


def make_boundary(xy):
    """Build a reference Boundary2D from an (n,2) array the way read_polygon does
    (geometry.py:46-51). Coordinates are passed as Python floats."""
    ns = load_reference()
    pts = [ns.Vertex(float(p[0]), float(p[1])) for p in np.asarray(xy)]
    for i in range(len(pts)):
        seg = ns.Segment(pts[i - 1], pts[i])
        pts[i - 1].assign_segment(seg)
        pts[i].assign_segment(seg)
    return ns.Boundary2D(pts)


def make_env(xy):
    ns = load_reference()
    return ns.BoudaryEnv(make_boundary(xy))


class TracedEnv:
    """Reference env + vertex-id bookkeeping + auto-reset, producing per-step records.

    Vertex ids: 0..n0-1 for the original polygon, n0+k for the k-th vertex inserted in the
    current episode.
    """

    def __init__(self, xy):
        self.env = make_env(xy)
        self.n0 = len(xy)
        self.obs, _ = self.env.reset()
        self._rebuild_ids()

    def _rebuild_ids(self):
        self.ids = {id(v): k for k, v in enumerate(self.env.updated_boundary.vertices)}
        self.next_id = self.n0

    def boundary_ids(self):
        out = []
        for v in self.env.updated_boundary.vertices:
            key = id(v)
            if key not in self.ids:
                self.ids[key] = self.next_id
                self.next_id += 1
            out.append(self.ids[key])
        return out

    def state(self, have_ref=True):
        env = self.env
        verts = env.updated_boundary.vertices
        ids = self.boundary_ids()
        xy = np.array([[float(v.x), float(v.y)] for v in verts], dtype=np.float64)
        rp = env.current_point_environment.reference_point
        cands = [(self.ids[id(v)], float(k)) for v, k in env.candidate_vertices]
        return dict(ids=ids, xy=xy, ref_index=verts.index(rp) if have_ref else -1, n=len(verts),
                    n_elements=len(env.generated_meshes), candidates=cands,
                    base_length=float(env.current_point_environment.base_length),
                    current_area=float(env.current_area), failed_num=env.failed_num)

    def step(self, action):
        """One reference step with auto-reset. Returns a dict record."""
        env = self.env
        n_el_before = len(env.generated_meshes)
        obs, rew, term, trunc, info = env.step(np.asarray(action, dtype=np.float32))
        n_el = len(env.generated_meshes)
        rec = dict(reward=float(rew), terminated=bool(term), truncated=bool(trunc),
                   is_complete=bool(info["is_complete"]), n_elements=n_el,
                   success=n_el > n_el_before, terminal_obs=None)
        st = self.state(have_ref=obs is not None)
        rec["pre_reset_state"] = st
        rec["obs_none"] = obs is None
        if term or trunc:
            rec["terminal_obs"] = None if obs is None else np.array(obs, dtype=np.float32)
            obs, _ = env.reset()
            self._rebuild_ids()
            st = self.state()
        rec["obs"] = np.array(obs, dtype=np.float32)
        rec["state"] = st
        self.obs = rec["obs"]
        return rec


def action_stream(seed: int, T: int) -> np.ndarray:
    """one ``rng.uniform(low, high)`` draw per step, cast to float32."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(LOW, HIGH).astype(np.float32) for _ in range(T)])
