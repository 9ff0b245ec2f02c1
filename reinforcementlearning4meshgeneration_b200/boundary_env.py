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


class Mesh:
    """One generated element: ``.vertices`` (4 ``Vertex``), ``.ids`` (vertex ids) -- what the reference's plotting /
    quality helpers read from a ``Mesh`` (general/components.py:730-737)."""
    __slots__ = ("vertices", "ids")

    def __init__(self, xy, ids):
        self.vertices = [Vertex(float(p[0]), float(p[1])) for p in xy]
        self.ids = [int(i) for i in ids]

    def __array__(self, dtype=None, copy=None):
        a = np.array([[v.x, v.y] for v in self.vertices], dtype=np.float64)
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return np.asarray(self)[k]


class _MeshList:
    """``env.generated_meshes`` stand-in.  ``len()`` (what the evaluators read, eval_loop.py:103, testbed.py:182) comes
    from the element count the last step / reset reported -- no device synchronisation; iteration / indexing fetch the
    element log and yield ``Mesh`` objects."""

    def __init__(self, env: "BoudaryEnv"):
        self._env = env

    def _fetch(self):
        quads, vxy, ne = self._env._batched.get_elements(0)
        return quads, vxy, ne

    def __len__(self):
        return int(self._env._n_elements)

    def __iter__(self):
        quads, vxy, _ = self._fetch()
        for q in quads:
            yield Mesh(vxy[q], q)

    def __getitem__(self, i):
        quads, vxy, _ = self._fetch()
        if isinstance(i, slice):
            return [Mesh(vxy[q], q) for q in quads[i]]
        return Mesh(vxy[quads[i]], quads[i])


# backend factory: (xy, device=..., auto_reset=..., log_capacity=...) -> object with the BatchedBoudaryEnv surface used
# below (reset, step_host, get_elements, get_state, close).  The CPU-only tests swap in a fake here; the product
# always uses the CUDA env.
def _default_backend(xy, **kw):
    return BatchedBoudaryEnv([xy], num_envs=1, **kw)


_BACKEND_FACTORY = _default_backend


class BoudaryEnv(_EnvBase):
    """Single-environment facade (reference: envs/boundary_env.py:34-457; legacy rl/boundary_env.py:21-263).

    ``api="gymnasium"`` (default): ``reset(*, seed, static, options) -> (obs, {})`` and the 5-tuple ``step`` of the v2
    env.  ``api="gym"``: the legacy surface the v1 trainers use (rl/boundary_env.py:67-84, :263): ``reset() -> obs``
    and ``step -> (obs, reward, done, info)``.  ``reinforcementlearning4meshgeneration_b200.legacy`` exports the
    legacy flavour under the reference's names for mounting as ``rl.boundary_env``.
    """

    metadata = {"render_modes": []}
    TYPE_THRESHOLD = 0.3
    API = "gymnasium"

    @classmethod
    def from_domain_file(cls, filename, *, experiment_version: Optional[str] = None, env_name: Optional[int] = None):
        return cls(read_polygon(filename), experiment_version=experiment_version, env_name=env_name)

    def __init__(self, boundary, experiment_version: Optional[str] = None, env_name: Optional[int] = None, device=None,
                 api: Optional[str] = None):
        self._xy = as_xy(boundary)
        self.api = api or self.API
        if self.api not in ("gymnasium", "gym"):
            raise ValueError("api must be 'gymnasium' or 'gym'")
        # element counts grow with the domain's area (the reference's evaluation runs report ~5 x n0): a single env can
        # afford a deep log
        self._batched = _BACKEND_FACTORY(self._xy, device=device, auto_reset=False, log_capacity=max(64 * len(self._xy), 1024))
        self.action_space = _spaces.Box(ACTION_LOW.copy(), ACTION_HIGH.copy(), dtype=np.float32)
        self.observation_space = _spaces.Box(low=np.full((OBS_DIM,), -999.0, np.float32),
                                             high=np.full((OBS_DIM,), 999.0, np.float32), dtype=np.float32)
        self.neighbor_num, self.radius_num, self.radius, self.max_radius = 6, 3, 4, 2
        self.experiment_version = experiment_version if experiment_version else "test"
        self.env_name = env_name if env_name is not None else 1
        self.generated_meshes = _MeshList(self)
        self._n_elements = 0
        self._out = None
        self.current_state = None
        self.original_vertices = [Vertex(float(x), float(y)) for x, y in self._xy]      # rl/boundary_env.py:23
        self.boundary = Boundary2D(self.original_vertices)                              # rl/boundary_env.py:22

    # -- Gymnasium / gym API ------------------------------------------------------------------
    def seed(self, seed: Optional[int] = None) -> None:
        if seed is not None:
            np.random.seed(seed)

    def reset(self, *, seed: Optional[int] = None, static: bool = False, options: Optional[Dict[str, Any]] = None):
        if seed is not None:
            self.seed(seed)
        obs = np.array(self._batched.reset()[0].cpu().numpy(), dtype=np.float32, copy=True)
        if static:
            obs[1] = 0.0       # a static point environment reports area ratio 0 (C:1198-1199); only the reset obs is affected
        self.current_state = obs
        self._n_elements = 0
        return obs if self.api == "gym" else (obs, {})

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, 3)
        # auto_reset is off for the single-env facade: like the reference, the env keeps its final
        # state (generated_meshes, boundary) until the caller resets it.
        self._out = self._batched.step_host(a, self._out)
        o = self._out
        terminated, truncated = bool(o["terminated"][0]), bool(o["truncated"][0])
        obs = o["obs"][0].copy()
        self.current_state = obs
        self._n_elements = int(o["n_elements"][0])
        reward, info = np.float64(o["reward"][0]), {"is_complete": not truncated}
        if self.api == "gym":
            return obs, reward, terminated or truncated, info
        return obs, reward, terminated, truncated, info

    def move(self, new_point, type, lr_1=None, lr_2=None):
        """E:459-594 / rl/boundary_env.py:265-432: apply the geometric move ``new_point = (r, phi)`` with element type
        ``type``; returns ``(next_state | None, 0, done, {"is_complete": bool})``.  Where every candidate is on the
        not-valid list the mesh is smoothed like in the reference (smooth_pave; lr_1 / lr_2 are unused there too) and the
        episode goes on; if that is not possible (element log overflown) the episode ends and the info dict carries
        ``"needs_smoothing": True``."""
        r = self._batched.move(np.asarray([[float(new_point[0]), float(new_point[1])]], np.float64), np.asarray([float(type)], np.float64))
        done, complete, exhausted = bool(r["done"][0]), bool(r["is_complete"][0]), bool(r["exhausted"][0])
        self._n_elements = int(r["n_elements"][0])
        obs = np.array(r["obs"][0].cpu().numpy(), dtype=np.float32, copy=True)
        info = {"is_complete": complete}
        if exhausted:
            info["needs_smoothing"] = True
        none = exhausted or not obs.any()
        self.current_state = None if none else obs
        return (None if none else obs), 0, done, info

    def close(self) -> None:
        self._batched.close()

    def render(self, mode: str = "human") -> None:
        print(f"Generated elements: {len(self.generated_meshes)}")

    # -- what the legacy evaluators touch besides reset / step (testbed.py:107-274, CustomizeCallback.py:131-133) ----
    @property
    def updated_boundary(self) -> Boundary2D:
        """Current front as a ``Boundary2D`` (read-only view: ``.vertices[i].x/.y``; general/mesh.py:601-674)."""
        st = self._batched.get_state(0)
        return Boundary2D([Vertex(float(x), float(y)) for x, y in st["xy"]])

    def save_meshes(self, name, meshes=None, quality: bool = False, indexing: bool = False, type: int = 0, dpi: int = 300,
                    style: str = "k.-") -> str:
        """general/mesh.py:1785-1792: draw the generated elements.  With matplotlib the figure is saved like the
        reference does; without it (this image) an SVG with the same content is written next to ``name``.  Returns the
        path written."""
        from .export import save_meshes_figure
        meshes = list(self.generated_meshes) if meshes is None else list(meshes)
        return save_meshes_figure(name, meshes, self._xy, indexing=indexing, dpi=dpi, style=style)

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
