"""Drop-in mirror of the reference module ``rl.boundary_env`` / ``mesh_rl.envs.boundary_env``:
exports ``BoudaryEnv``, ``read_polygon`` and ``boundary`` (the three names the trainers import,
rl/baselines/RL_Mesh.py:6,24-25; v2 training/train_loop.py:17, evaluation/eval_loop.py:17).

``BoudaryEnv`` is a batch-of-one view over the CUDA ``BatchedBoudaryEnv``: same constructor
(``BoudaryEnv(boundary)``, ``BoudaryEnv.from_domain_file(path)``), Gymnasium 5-tuple ``step``,
``reset(*, seed, static, options) -> (obs, {})``, spaces, ``generated_meshes``.  It returns numpy
like the reference does; use ``BatchedBoudaryEnv`` for device-resident tensors.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, Optional

import numpy as np

from .batched_env import ACTION_HIGH, ACTION_LOW, OBS_DIM, BatchedBoudaryEnv, as_xy

try:  # SB3 2.x requires a gymnasium.Env subclass (stable_baselines3 _patch_env)
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _EnvBase = _gym.Env
except Exception:  # gymnasium is not installed in this image: plain class with a minimal Box
    _gym = None
    _EnvBase = object

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is not None:
                self.low = np.full(tuple(shape), low, dtype=dtype)
                self.high = np.full(tuple(shape), high, dtype=dtype)
            else:
                self.low = np.asarray(low, dtype=dtype)
                self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class _spaces:  # noqa: N801
        Box = _Box


class Vertex:
    """Coordinate holder compatible with the reference's Vertex for the purposes of this path
    (general/components.py:50-53: ``.x``, ``.y``)."""
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = x, y

    def __repr__(self):
        return f"Vertex({self.x}, {self.y})"


class Boundary2D:
    """Clockwise polygon: ``.vertices[i].x/.y`` (general/components.py:206-208)."""

    def __init__(self, vertices):
        self.vertices = list(vertices)

    def __len__(self):
        return len(self.vertices)


def read_polygon(filename) -> Boundary2D:
    """JSON first line ``[[x_px, y_px], ...]`` -> coordinates / 100
    (general/polygon.py:110-117, v2 geometry.py:34-52)."""
    with Path(filename).open("r", encoding="utf-8") as fr:
        vertices = json.loads(fr.readline())
    return Boundary2D([Vertex(p[0] / 100.0, p[1] / 100.0) for p in vertices])


_BUILTIN = {
    0: [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 6), (2, 6), (3, 6), (4, 6), (5, 6), (6, 6), (7, 5), (8, 4),
        (9, 3), (10, 2), (11, 1), (12, 0), (11, -1), (10, -2), (9, -3), (8, -4), (7, -5), (6, -6), (5, -5), (4, -4),
        (3, -3), (2, -2), (1, -1), (0, 0)],
    -1: [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 6), (2, 6), (3, 6), (4, 6), (5, 6), (6, 6), (6, 5), (6, 4),
         (6, 3), (6, 2), (6, 1), (6, 0), (5, 0), (4, 0), (3, 0), (2, 0), (1, 0), (0, 0)],
}


def boundary(index: int = 0) -> Boundary2D:
    """Built-in test polygons of the reference (general/polygon.py:76-108); index 0 is the
    30-vertex domain of BASELINE config 1."""
    if index not in _BUILTIN:
        raise ValueError(f"built-in boundary {index} is not provided (available: {sorted(_BUILTIN)})")
    return Boundary2D([Vertex(x, y) for x, y in _BUILTIN[index]])


class _MeshList:
    """``env.generated_meshes`` stand-in: evaluators only take ``len()`` (eval_loop.py:103) or
    iterate quads; items are (4,2) coordinate arrays fetched from the device element log."""

    def __init__(self, env: "BoudaryEnv"):
        self._env = env

    def _fetch(self):
        quads, vxy, ne = self._env._batched.get_elements(0)
        return quads, vxy, ne

    def __len__(self):
        return int(self._fetch()[2])

    def __iter__(self):
        quads, vxy, _ = self._fetch()
        for q in quads:
            yield vxy[q]

    def __getitem__(self, i):
        quads, vxy, _ = self._fetch()
        return vxy[quads[i]]


class BoudaryEnv(_EnvBase):
    """Single-environment Gymnasium facade (reference: envs/boundary_env.py:34-457)."""

    metadata = {"render_modes": []}
    TYPE_THRESHOLD = 0.3

    @classmethod
    def from_domain_file(cls, filename, *, experiment_version: Optional[str] = None, env_name: Optional[int] = None):
        return cls(read_polygon(filename), experiment_version=experiment_version, env_name=env_name)

    def __init__(self, boundary, experiment_version: Optional[str] = None, env_name: Optional[int] = None, device=None):
        self._xy = as_xy(boundary)
        self._batched = BatchedBoudaryEnv([self._xy], num_envs=1, device=device, auto_reset=False)
        self.action_space = _spaces.Box(ACTION_LOW.copy(), ACTION_HIGH.copy(), dtype=np.float32)
        self.observation_space = _spaces.Box(low=np.full((OBS_DIM,), -999.0, np.float32),
                                             high=np.full((OBS_DIM,), 999.0, np.float32), dtype=np.float32)
        self.neighbor_num, self.radius_num, self.radius, self.max_radius = 6, 3, 4, 2
        self.experiment_version = experiment_version if experiment_version else "test"
        self.env_name = env_name if env_name is not None else 1
        self.generated_meshes = _MeshList(self)
        self._out = None
        self.current_state = None

    # -- Gymnasium API ----------------------------------------------------------------------
    def seed(self, seed: Optional[int] = None) -> None:
        if seed is not None:
            np.random.seed(seed)

    def reset(self, *, seed: Optional[int] = None, static: bool = False, options: Optional[Dict[str, Any]] = None):
        if seed is not None:
            self.seed(seed)
        obs = self._batched.reset()[0].cpu().numpy().copy()
        if static:
            obs[1] = 0.0       # a static point environment reports area ratio 0 (C:1198-1199); only the reset obs is affected
        self.current_state = obs
        return obs, {}

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, 3)
        # auto_reset is off for the single-env facade: like the reference, the env keeps its final
        # state (generated_meshes, boundary) until the caller resets it.
        self._out = self._batched.step_host(a, self._out)
        o = self._out
        terminated, truncated = bool(o["terminated"][0]), bool(o["truncated"][0])
        obs = o["obs"][0].copy()
        self.current_state = obs
        return obs, np.float64(o["reward"][0]), terminated, truncated, {"is_complete": not truncated}

    def close(self) -> None:
        self._batched.close()

    def render(self, mode: str = "human") -> None:
        print(f"Generated elements: {len(self.generated_meshes)}")

    # -- element export (rl/boundary_env.py:648-669, general/mesh.py:1842-1864) ---------------
    def _mesh(self):
        quads, vxy, _ = self._batched.get_elements(0)
        return len(self._xy), quads, vxy

    def write_2_file(self, filename) -> None:
        from .export import write_2_file
        write_2_file(filename, *self._mesh())

    def write_generated_elements_2_file(self, filename, format: str = "inp") -> None:
        from .export import write_inp
        write_inp(filename, *self._mesh())

