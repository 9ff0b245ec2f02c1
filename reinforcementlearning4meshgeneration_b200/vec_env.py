"""SB3-compatible ``VecEnv`` over the CUDA batched env.

Mirrors the surface of the reference's vendored ``DummyVecEnv``
(rl/baselines/dummy_vec_env.py:12-125): ``num_envs``, ``observation_space``, ``action_space``,
``step_async`` / ``step_wait`` -> ``(obs[N,18], rews f32[N], dones bool[N], infos list[dict])``,
``reset``, ``close``, ``seed``, ``get_attr`` / ``set_attr`` / ``env_method`` / ``env_is_wrapped``.
On done, ``infos[i]["terminal_observation"]`` holds the last observation of the episode and the
env is reset in place (stock SB3 behaviour; the vendored copy comments the reset out, :49 -- pass
``auto_reset=False`` for that variant).  SB3 2.x extras: ``infos[i]["TimeLimit.truncated"]`` and a
Monitor-style ``infos[i]["episode"] = {"r", "l", "t"}``; plus ``infos[i]["is_complete"]`` from the
reference env (envs/boundary_env.py:386) and ``infos[i]["n_elements"]``.

stable_baselines3 / gymnasium are imported lazily: when SB3 is installed the class derives from
``stable_baselines3.common.vec_env.VecEnv`` so it drops into ``SAC('MlpPolicy', env, ...)``
(v2 algorithms/sb3_algos.py:114, legacy RL_Mesh.py:186-196); otherwise it is a plain class with
the same methods (this image has neither package).  The numpy contract of SB3 forces one
host round trip per step; it goes through ``mg_step_host`` with pinned buffers.
"""
from __future__ import annotations

import time
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from .batched_env import ACT_DIM, OBS_DIM, BatchedBoudaryEnv
from .boundary_env import _spaces

try:  # pragma: no cover - SB3 is not installed in the build image
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv
except Exception:  # noqa: BLE001
    _SB3VecEnv = None


def _pinned(shape, dtype):
    t = torch.empty(shape, dtype=dtype)
    try:
        return t.pin_memory()
    except Exception:  # no CUDA runtime (CPU-only tests with a fake backend)
        return t


class _EnvView:
    """``venv.envs[i]``: what the reference's callbacks / evaluators reach through a ``DummyVecEnv`` for one env
    (rl/baselines/CustomizeCallback.py:131-133, rl/baselines/dummy_vec_env.py:21): ``generated_meshes``,
    ``save_meshes``, ``boundary`` / ``original_vertices`` / ``updated_boundary``, ``write_2_file`` -- read-only views
    fetched from the batched env on demand (``mg_get_elements`` / ``mg_get_state``)."""

    def __init__(self, batched, index: int):
        self._b, self._i = batched, int(index)

    def _elements(self):
        quads, vxy, ne = self._b.get_elements(self._i)
        return quads, vxy, ne

    @property
    def generated_meshes(self):
        from .boundary_env import Mesh
        quads, vxy, _ = self._elements()
        return [Mesh(vxy[q], q) for q in quads]

    @property
    def original_vertices(self):
        from .boundary_env import Vertex
        st = self._b.get_state(self._i)
        _, vxy, _ = self._elements()
        return [Vertex(float(x), float(y)) for x, y in vxy[: st["n0"]]]

    @property
    def boundary(self):
        from .boundary_env import Boundary2D
        return Boundary2D(self.original_vertices)

    @property
    def updated_boundary(self):
        from .boundary_env import Boundary2D, Vertex
        return Boundary2D([Vertex(float(x), float(y)) for x, y in self._b.get_state(self._i)["xy"]])

    def save_meshes(self, name, meshes=None, quality=False, indexing=False, type=0, dpi=300, style="k.-"):
        from .export import save_meshes_figure
        _, vxy, _ = self._elements()
        n0 = self._b.get_state(self._i)["n0"]
        return save_meshes_figure(name, self.generated_meshes if meshes is None else list(meshes), vxy[:n0], indexing=indexing,
                                  dpi=dpi, style=style)

    def write_2_file(self, filename):
        from .export import write_2_file
        quads, vxy, _ = self._elements()
        write_2_file(filename, self._b.get_state(self._i)["n0"], quads, vxy)


class _VecEnvCore:
    """Backend-agnostic VecEnv logic (unit-tested on the CPU with a fake batched env)."""

    def _init_core(self, batched, monitor: bool = True):
        self._b = batched
        self.num_envs = int(batched.num_envs)
        from .batched_env import ACTION_HIGH, ACTION_LOW
        self.action_space = _spaces.Box(ACTION_LOW.copy(), ACTION_HIGH.copy(), dtype=np.float32)
        self.observation_space = _spaces.Box(low=np.full((OBS_DIM,), -999.0, np.float32),
                                             high=np.full((OBS_DIM,), 999.0, np.float32), dtype=np.float32)
        N = self.num_envs
        self._act = _pinned((N, ACT_DIM), torch.float32)
        self._out = dict(obs=_pinned((N, OBS_DIM), torch.float32), reward=_pinned((N,), torch.float64),
                         terminated=_pinned((N,), torch.uint8), truncated=_pinned((N,), torch.uint8),
                         terminal_obs=_pinned((N, OBS_DIM), torch.float32), n_elements=_pinned((N,), torch.int32))
        self._monitor = monitor
        self.envs = [_EnvView(batched, i) for i in range(N)]      # DummyVecEnv.envs (dummy_vec_env.py:21)
        if hasattr(batched, "set_obs_delta"):
            batched.set_obs_delta(True)       # the adapter owns its pinned buffers and hands out copies
        self._ep_ret = np.zeros(N, np.float64)
        self._ep_len = np.zeros(N, np.int64)
        self._t0 = time.time()
        self._actions = None
        self.metadata = {"render_modes": []}
        self.render_mode = None

    # -- VecEnv API -------------------------------------------------------------------------
    def reset(self):
        obs = self._b.reset()
        self._ep_ret[:] = 0
        self._ep_len[:] = 0
        return obs.cpu().numpy().copy()

    def step_async(self, actions) -> None:
        # the step is enqueued here (mg_step_host_begin) and runs on the GPU while the caller goes on; step_wait joins it
        self._actions = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, ACT_DIM)
        self._act.copy_(torch.from_numpy(self._actions))
        self._in_flight = hasattr(self._b, "step_host_begin")
        if self._in_flight:
            self._b.step_host_begin(self._act, self._out)

    def step_wait(self):
        if getattr(self, "_in_flight", False):
            self._b.step_host_end()
            self._in_flight = False
            o = self._out
        else:
            o = self._b.step_host(self._act, self._out)
        obs = o["obs"].numpy().copy()
        rew64 = o["reward"].numpy()
        term = o["terminated"].numpy().astype(bool)
        trunc = o["truncated"].numpy().astype(bool)
        dones = term | trunc
        self._ep_ret += rew64
        self._ep_len += 1
        infos: List[dict] = [{} for _ in range(self.num_envs)]
        for i in np.nonzero(dones)[0]:
            info = infos[i]
            info["terminal_observation"] = o["terminal_obs"][i].numpy().copy()
            info["TimeLimit.truncated"] = bool(trunc[i] and not term[i])
            info["is_complete"] = bool(not trunc[i])
            info["n_elements"] = int(o["n_elements"][i])
            if self._monitor:
                info["episode"] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i]),
                                   "t": round(time.time() - self._t0, 6)}
            self._ep_ret[i] = 0
            self._ep_len[i] = 0
        if not getattr(self._b, "auto_reset", True):
            for i in np.nonzero(dones)[0]:
                obs[i] = infos[i]["terminal_observation"]
        return obs, rew64.astype(np.float32), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        self._b.close()

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def get_images(self) -> Sequence[np.ndarray]:
        return []

    def render(self, mode: str = "human"):
        return None

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        """Attribute of the individual envs (generated_meshes, boundary, ...) or, failing that, of the adapter."""
        if attr_name == "render_mode":
            return [None for _ in self._indices(indices)]
        if hasattr(_EnvView, attr_name):
            return [getattr(self.envs[i], attr_name) for i in self._indices(indices)]
        return [getattr(self, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
        if method_name == "generated_meshes_count":
            return [self._b.n_elements_of(i) if hasattr(self._b, "n_elements_of") else self._b.get_state(i)["n_elements"]
                    for i in self._indices(indices)]
        if method_name in ("save_meshes", "write_2_file"):
            return [getattr(self.envs[i], method_name)(*method_args, **method_kwargs) for i in self._indices(indices)]
        raise AttributeError(f"env_method {method_name!r} is not available on the batched CUDA env")

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False for _ in self._indices(indices)]


if _SB3VecEnv is not None:  # pragma: no cover - exercised only where SB3 is installed
    class SB3VecEnv(_VecEnvCore, _SB3VecEnv):
        def __init__(self, domains=None, num_envs: int = 1, monitor: bool = True, **kwargs):
            batched = domains if isinstance(domains, BatchedBoudaryEnv) else BatchedBoudaryEnv(domains, num_envs=num_envs, **kwargs)
            self._init_core(batched, monitor)
            _SB3VecEnv.__init__(self, self.num_envs, self.observation_space, self.action_space)
else:
    class SB3VecEnv(_VecEnvCore):
        def __init__(self, domains=None, num_envs: int = 1, monitor: bool = True, **kwargs):
            batched = domains if hasattr(domains, "step_host") else BatchedBoudaryEnv(domains, num_envs=num_envs, **kwargs)
            self._init_core(batched, monitor)
