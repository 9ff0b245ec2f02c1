"""Native batched API: ``BatchedBoudaryEnv`` -- N reference ``BoudaryEnv`` instances stepped in
lock-step by the sm_100a kernels, observations / rewards / flags returned as device-resident
torch tensors (no host synchronisation on the step path).

Reference behaviour: v2/src/mesh_rl/envs/boundary_env.py:34 (``BoudaryEnv``), reset :136-184,
step :388-457; auto-reset follows the stock SB3 VecEnv convention
(rl/baselines/dummy_vec_env.py:40-52 with the reset the vendored copy comments out).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import EpisodeStats, PolygenCfg, StateView, check

OBS_DIM = 18
ACT_DIM = 3
# E:78-80
ACTION_LOW = np.array([-1.0, -1.5, 0.0], dtype=np.float32)
ACTION_HIGH = np.array([1.0, 1.5, 1.5], dtype=np.float32)


def poly_area(xy: np.ndarray) -> float:
    """Boundary2D.poly_area as the reference's __init__ evaluates it on the host
    (components_core.py:485-487, envs/boundary_env.py:72)."""
    xy = np.asarray(xy, dtype=np.float64)
    return float(0.5 * np.abs(np.dot(xy[:, 0], np.roll(xy[:, 1], 1)) - np.dot(xy[:, 1], np.roll(xy[:, 0], 1))))


def as_xy(boundary) -> np.ndarray:
    """Accept an (n,2) array or any object exposing ``.vertices[i].x/.y`` (the reference's
    Boundary2D, rl/boundary_env.py:21-24)."""
    if hasattr(boundary, "vertices"):
        return np.array([[float(v.x), float(v.y)] for v in boundary.vertices], dtype=np.float64)
    xy = np.asarray(boundary, dtype=np.float64)
    if xy.ndim != 2 or xy.shape[1] != 2:
        raise ValueError("boundary must be (n, 2) or expose .vertices")
    return np.ascontiguousarray(xy)


class StepResult:
    __slots__ = ("obs", "reward", "terminated", "truncated", "terminal_obs", "n_elements")

    def __init__(self, obs, reward, terminated, truncated, terminal_obs, n_elements):
        self.obs, self.reward, self.terminated, self.truncated = obs, reward, terminated, truncated
        self.terminal_obs, self.n_elements = terminal_obs, n_elements

    def __iter__(self):
        info = {"terminal_obs": self.terminal_obs, "n_elements": self.n_elements,
                "is_complete": self.truncated == 0}
        return iter((self.obs, self.reward, self.terminated, self.truncated, info))


class BatchedBoudaryEnv:
    """``num_envs`` BoudaryEnv instances on one CUDA device.

    domains   : a polygon or list of polygons ((n,2) arrays or Boundary2D-like objects), or None
                when ``random_polygons`` is given.
    env_domain: per-env index into ``domains`` (default: round-robin blocks, env e -> e*D//N).
    random_polygons: dict of generator settings (see ``PolygenCfg``) -> every reset draws a fresh
                random star polygon in-kernel (BASELINE configs 3/4).
    obs_delta : the observation tensor returned by ``reset`` / ``step`` is one persistent buffer; with
                ``obs_delta`` (default) a step only rewrites the rows whose env changed (a failed step returns a
                bit-identical observation in the reference too), so do not modify that tensor in place -- copy it,
                or pass ``obs_delta=False``.
    log_capacity: entries of the per-env element / inserted-vertex logs (default 2 x max_verts; element counts are
                exact either way, ``get_elements`` raises when an episode outgrew the log).
    """

    def __init__(self, domains=None, num_envs: int = 1, device: Optional[int | str | torch.device] = None,
                 env_domain: Optional[Sequence[int]] = None, max_verts: Optional[int] = None,
                 random_polygons: Optional[dict] = None, seed: int = 0, env_id_offset: int = 0,
                 auto_reset: bool = True, obs_delta: bool = True, log_capacity: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedBoudaryEnv needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._L = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self._h = C.c_void_p()
        self.random_mode = random_polygons is not None
        self._host_args = {}
        if self.random_mode:
            cfg = dict(ctr_x=250.0, ctr_y=250.0, ave_radius=100.0, irregularity=0.55, spikeyness=0.7,
                       min_coarse=8, max_coarse=24, min_verts=64, max_verts=512)
            cfg.update(random_polygons)
            self.max_verts = int(max_verts or cfg["max_verts"])
            check(self._L.mg_create(C.byref(self._h), self.device.index, self.num_envs, self.max_verts), None, "mg_create")
            c = PolygenCfg(**cfg)
            check(self._L.mg_set_random(self._h, int(seed), C.byref(c), int(env_id_offset)), self._h, "mg_set_random")
            self.domains = None
        else:
            if domains is None:
                raise ValueError("give domains or random_polygons")
            if hasattr(domains, "vertices") or (isinstance(domains, np.ndarray) and domains.ndim == 2):
                domains = [domains]
            self.domains = [as_xy(d) for d in domains]
            D = len(self.domains)
            self.max_verts = int(max_verts or max(len(d) for d in self.domains))
            if env_domain is None:
                env_domain = (np.arange(self.num_envs, dtype=np.int64) * D) // self.num_envs
            self.env_domain = np.ascontiguousarray(np.asarray(env_domain, dtype=np.int32))
            if len(self.env_domain) != self.num_envs:
                raise ValueError("env_domain must have num_envs entries")
            check(self._L.mg_create(C.byref(self._h), self.device.index, self.num_envs, self.max_verts), None, "mg_create")
            offsets = np.zeros(D + 1, dtype=np.int32)
            offsets[1:] = np.cumsum([len(d) for d in self.domains])
            xy = np.ascontiguousarray(np.concatenate(self.domains, axis=0))
            areas = np.array([poly_area(d) for d in self.domains], dtype=np.float64)
            check(self._L.mg_set_domains(self._h, xy.ctypes.data, offsets.ctypes.data, D, self.env_domain.ctypes.data,
                                         areas.ctypes.data), self._h, "mg_set_domains")
        self.auto_reset = bool(auto_reset)
        check(self._L.mg_set_auto_reset(self._h, int(self.auto_reset)), self._h, "mg_set_auto_reset")
        check(self._L.mg_set_obs_delta(self._h, int(bool(obs_delta))), self._h, "mg_set_obs_delta")
        for kv in filter(None, os.environ.get("MESHGEN_OPTIONS", "").split(",")):     # tuning aid: "fuse_decide=0,..."
            k, v = kv.split("=")
            check(self._L.mg_set_option(self._h, k.strip().encode(), int(v)), self._h, "mg_set_option")
        if log_capacity is not None:
            check(self._L.mg_set_log_capacity(self._h, int(log_capacity), int(log_capacity)), self._h, "mg_set_log_capacity")
        N = self.num_envs
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((N, OBS_DIM), dtype=torch.float32, device=self.device)
            self.reward = torch.zeros(N, dtype=torch.float64, device=self.device)
            self.terminated = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self.truncated = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self.terminal_obs = torch.zeros((N, OBS_DIM), dtype=torch.float32, device=self.device)
            self.n_elements = torch.zeros(N, dtype=torch.int32, device=self.device)
            self._act = torch.zeros((N, ACT_DIM), dtype=torch.float32, device=self.device)
            self._stats_dev = torch.zeros(12, dtype=torch.int64, device=self.device)

    # ------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.mg_destroy(self._h)
            self._h = C.c_void_p()
        if getattr(self, "_host_args", None):
            self._host_args.clear()                # (the entries hold the caller's buffers)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset all envs (mask None) or those with mask != 0; returns obs[N,18] (device)."""
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = C.c_void_p(mask.data_ptr())
        check(self._L.mg_reset(self._h, mptr, C.c_void_p(self.obs.data_ptr()), self._stream()), self._h, "mg_reset")
        return self.obs

    def step(self, actions: torch.Tensor) -> StepResult:
        """One transition for every env. ``actions``: float32 [N,3] on this device.
        Returns device tensors (views of internal buffers, overwritten by the next step)."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.shape != (self.num_envs, ACT_DIM):
            raise ValueError(f"actions must have shape ({self.num_envs}, {ACT_DIM})")
        check(self._L.mg_step(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(self.obs.data_ptr()),
                              C.c_void_p(self.reward.data_ptr()), C.c_void_p(self.terminated.data_ptr()),
                              C.c_void_p(self.truncated.data_ptr()), C.c_void_p(self.terminal_obs.data_ptr()),
                              C.c_void_p(self.n_elements.data_ptr()), self._stream()), self._h, "mg_step")
        return StepResult(self.obs, self.reward, self.terminated, self.truncated, self.terminal_obs, self.n_elements)

    def move(self, polar, type) -> dict:
        """``BoudaryEnv.move`` (E:459-594) for every env: ``polar`` float64 [N,2] = (r, phi), ``type`` float64 [N].
        Returns device tensors: obs (static point environment), done, is_complete, exhausted, n_elements.  Reward is always
        0.  Where every reference candidate is on the not-valid list the mesh is smoothed and the episode goes on like in
        the reference (smooth_pave, M:816-821; one host synchronisation per call); ``exhausted`` marks an env
        that needed it and could not get it (``set_option("smooth_pave", 0)``, an overflown element log): such an env is reported done."""
        polar = torch.as_tensor(polar, dtype=torch.float64).to(self.device).contiguous().reshape(self.num_envs, 2)
        type = torch.as_tensor(type, dtype=torch.float64).to(self.device).contiguous().reshape(self.num_envs)
        if not hasattr(self, "_move_flags"):
            self._move_flags = torch.zeros((3, self.num_envs), dtype=torch.uint8, device=self.device)
        f = self._move_flags
        check(self._L.mg_move(self._h, C.c_void_p(polar.data_ptr()), C.c_void_p(type.data_ptr()), C.c_void_p(self.obs.data_ptr()),
                              C.c_void_p(f[0].data_ptr()), C.c_void_p(f[1].data_ptr()), C.c_void_p(f[2].data_ptr()),
                              C.c_void_p(self.n_elements.data_ptr()), self._stream()), self._h, "mg_move")
        return dict(obs=self.obs, done=f[0], is_complete=f[1], exhausted=f[2], n_elements=self.n_elements)

    def step_host(self, actions: np.ndarray, out: Optional[dict] = None, _begin_only: bool = False) -> dict:
        """Same transition through host (numpy / pinned) buffers: H2D + step + D2H + sync inside the
        library (mg_step_host) -- the path a numpy-facing caller such as SB3 pays.  With pinned ``out`` buffers
        (torch ``.pin_memory()``) the GPU writes the results straight into them -- with ``obs_delta`` on only the
        observation rows, rewards, flags and element counts that differ from what the buffers already hold: keep passing
        the same ``out`` buffers and do not modify them in place."""
        # The marshalled pointers of (action tensor, result buffers) are kept for callers that keep passing the same
        # objects -- ~6 us of ctypes conversions per call otherwise; entries hold the objects, so an id cannot be recycled.
        ent = self._host_args.get(id(actions)) if out is not None else None
        if ent is not None and ent[0] is actions and ent[1] is out and tuple(map(id, out.values())) == ent[2] \
                and actions.data_ptr() == ent[3][0].value:
            args = ent[3]
        else:
            is_t = isinstance(actions, torch.Tensor)
            if is_t:
                if actions.dtype != torch.float32 or not actions.is_contiguous() or actions.device.type != "cpu":
                    raise ValueError("actions tensor must be a contiguous float32 CPU tensor (pinned for the fastest path)")
            else:
                a = np.ascontiguousarray(actions, dtype=np.float32)
            N = self.num_envs
            if out is None:
                out = dict(obs=np.empty((N, OBS_DIM), np.float32), reward=np.empty(N, np.float64),
                           terminated=np.empty(N, np.uint8), truncated=np.empty(N, np.uint8),
                           terminal_obs=np.empty((N, OBS_DIM), np.float32), n_elements=np.empty(N, np.int32))
                cacheable = False
            else:
                cacheable = is_t and all(isinstance(v, torch.Tensor) for v in out.values())

            def ptr(x):
                return C.c_void_p(x.data_ptr()) if isinstance(x, torch.Tensor) else C.c_void_p(x.ctypes.data)

            args = (ptr(actions) if is_t else C.c_void_p(a.ctypes.data), ptr(out["obs"]), ptr(out["reward"]), ptr(out["terminated"]),
                    ptr(out["truncated"]), ptr(out["terminal_obs"]), ptr(out["n_elements"]))
            if cacheable:
                if len(self._host_args) >= 128:
                    self._host_args.clear()
                self._host_args[id(actions)] = (actions, out, tuple(map(id, out.values())), args, tuple(out.values()))
        rc = (self._L.mg_step_host_begin if _begin_only else self._L.mg_step_host)(self._h, *args)
        if rc != 0:
            check(rc, self._h, "mg_step_host")
        return out

    def step_host_begin(self, actions, out: dict) -> dict:
        """First half of ``step_host``: enqueue the step and return; ``out`` is defined after ``step_host_end``.  With the
        envs split over two ``BatchedBoudaryEnv`` objects the host works on one half (results, policy, next launch) while
        the other half runs on the GPU -- the step_async / step_wait split of SB3's VecEnv (mg_step_host_begin / _end)."""
        return self.step_host(actions, out, _begin_only=True)

    def step_host_end(self) -> None:
        rc = self._L.mg_step_host_end(self._h)
        if rc != 0:
            check(rc, self._h, "mg_step_host_end")

    def snapshot(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Complete env state as one uint8 device tensor (mg_snapshot_save); ``.cpu()`` it to persist."""
        nbytes = int(self._L.mg_snapshot_bytes(self._h))
        if out is None:
            out = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        if out.numel() != nbytes or out.device != self.device or out.dtype != torch.uint8:
            raise ValueError(f"snapshot buffer must be a uint8 tensor of {nbytes} bytes on {self.device}")
        check(self._L.mg_snapshot_save(self._h, C.c_void_p(out.data_ptr()), self._stream()), self._h, "mg_snapshot_save")
        return out

    def restore(self, blob: torch.Tensor) -> torch.Tensor:
        """Load a ``snapshot()`` (same num_envs / max_verts / domains or generator); returns the current obs."""
        blob = blob.to(device=self.device, dtype=torch.uint8).contiguous()
        check(self._L.mg_snapshot_load(self._h, C.c_void_p(blob.data_ptr()), int(blob.numel()), self._stream()), self._h,
              "mg_snapshot_load")
        return self.reset(torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device))   # empty mask: reads the cached obs

    def set_obs_delta(self, enabled: bool = True) -> None:
        """step / step_host write only the observation rows that changed (see mg_set_obs_delta); the caller
        must then pass the same, unmodified observation buffer on every call."""
        check(self._L.mg_set_obs_delta(self._h, int(enabled)), self._h, "mg_set_obs_delta")

    set_host_delta = set_obs_delta      # round-1 name

    def last_host_bytes(self):
        h2d, d2h = C.c_int64(), C.c_int64()
        check(self._L.mg_last_host_bytes(self._h, C.byref(h2d), C.byref(d2h)), self._h, "mg_last_host_bytes")
        return h2d.value, d2h.value

    def sample_actions(self, seed: int, step_index, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Uniform actions in the action box from the library's Philox stream (synthetic policy).  ``step_index``: an
        int, or a device tensor of two int64 -- [0] the step index, advanced by the launch itself, [1] scratch (zero) --
        for loops that are captured in a CUDA graph (mg_sample_actions_seq)."""
        out = self._act if out is None else out
        if isinstance(step_index, torch.Tensor):
            assert step_index.dtype == torch.int64 and step_index.numel() >= 2 and step_index.device == self.device
            check(self._L.mg_sample_actions_seq(self._h, int(seed), C.c_void_p(step_index.data_ptr()), C.c_void_p(out.data_ptr()),
                                                self._stream()), self._h, "mg_sample_actions_seq")
        else:
            check(self._L.mg_sample_actions(self._h, int(seed), int(step_index), C.c_void_p(out.data_ptr()), self._stream()),
                  self._h, "mg_sample_actions")
        return out

    # ------------------------------------------------------------------
    def get_state(self, env: int) -> dict:
        cap = self.max_verts + 2
        xy = np.zeros((cap, 2), np.float64)
        vid = np.zeros(cap, np.int32)
        key = np.zeros(cap, np.float64)
        stamp = np.zeros(cap, np.int32)
        v = StateView()
        v.xy_host, v.vertex_id_host = xy.ctypes.data, vid.ctypes.data
        v.cand_key_host, v.cand_stamp_host = key.ctypes.data, stamp.ctypes.data
        check(self._L.mg_get_state(self._h, int(env), C.byref(v)), self._h, "mg_get_state")
        n = v.n
        cand = [(int(vid[j]), float(key[j]), int(stamp[j])) for j in range(n) if np.isfinite(key[j])]
        cand.sort(key=lambda t: (t[1], t[2]))
        return dict(n=n, ref_index=v.ref_index, n_elements=v.n_elements, failed_num=v.failed_num, n0=v.n0,
                    base_length=v.base_length, current_area=v.current_area, original_area=v.original_area,
                    area_range=(v.area_min, v.area_crit), memo_flags=int(v.memo_flags), xy=xy[:n].copy(), ids=vid[:n].copy(),
                    cand_key=key[:n].copy(), cand_stamp=stamp[:n].copy(),
                    candidates=[(c[0], c[1]) for c in cand])

    def get_elements(self, env: int, allow_truncated: bool = False):
        """(quads[m,4] int32 vertex ids, vertex_xy[nv,2], element count) of env's current episode
        (generated_meshes).  Raises when the episode outgrew the element / inserted-vertex log unless
        ``allow_truncated`` (the count is exact either way)."""
        ce, ci = C.c_int32(), C.c_int32()
        check(self._L.mg_log_capacity(self._h, C.byref(ce), C.byref(ci)), self._h, "mg_log_capacity")
        cap_e, cap_v = ce.value, self.max_verts + 2 + ci.value
        quads = np.zeros((cap_e, 4), np.int32)
        vxy = np.zeros((cap_v, 2), np.float64)
        ne, nv = C.c_int32(), C.c_int32()
        rc = self._L.mg_get_elements(self._h, int(env), quads.ctypes.data, cap_e, C.byref(ne), vxy.ctypes.data, cap_v, C.byref(nv))
        if not (rc == -4 and allow_truncated):      # MG_ERR_CAPACITY: counts and the returned prefix are valid
            check(rc, self._h, "mg_get_elements")
        return quads[:min(ne.value, cap_e)].copy(), vxy[:min(nv.value, cap_v)].copy(), ne.value

    def debug_polygon(self, env: int, episode: int = -1) -> dict:
        """Random-polygon mode: the polygon of ``episode`` of ``env`` (default: its current episode) regenerated
        from its counter, with the generator's coarse polygon (pixels, clockwise) and densifier spacing."""
        xy = np.zeros((self.max_verts + 2, 2), np.float64)
        coarse = np.zeros(64, np.int32)
        n, k, area, spacing = C.c_int32(), C.c_int32(), C.c_double(), C.c_double()
        check(self._L.mg_debug_polygon(self._h, int(env), int(episode), xy.ctypes.data, self.max_verts + 2, C.byref(n), C.byref(area),
                                       coarse.ctypes.data, C.byref(k), C.byref(spacing)), self._h, "mg_debug_polygon")
        return dict(xy=xy[:n.value].copy(), n=n.value, original_area=area.value,
                    coarse_px=coarse[:2 * k.value].reshape(-1, 2).copy(), spacing_px=spacing.value)

    def set_log_capacity(self, max_elements: int, max_inserted: Optional[int] = None) -> None:
        """Capacity of the per-env element / inserted-vertex logs (default 8 x max_verts each); reset afterwards."""
        check(self._L.mg_set_log_capacity(self._h, int(max_elements), int(max_inserted or max_elements)), self._h,
              "mg_set_log_capacity")

    def n_elements_of(self, env: int) -> int:
        """len(generated_meshes) of one env (a 128-byte read-back, not the whole log)."""
        v = StateView()
        check(self._L.mg_get_state(self._h, int(env), C.byref(v)), self._h, "mg_get_state")
        return int(v.n_elements)

    def stats(self, reset: bool = False) -> dict:
        s = EpisodeStats()
        check(self._L.mg_stats(self._h, C.byref(s), int(reset)), self._h, "mg_stats")
        return s.as_dict()

    def stats_async(self, out: Optional[torch.Tensor] = None, reset: bool = False) -> torch.Tensor:
        """Episode statistics summed on the device into a 12 x int64 tensor laid out like ``mg_episode_stats``
        (entries 0..9 int64 counters, entries 10..11 the bit patterns of two float64 sums: ``out[10:].view(torch.float64)``),
        enqueued on the current stream without any host synchronisation (mg_stats_async)."""
        out = self._stats_dev if out is None else out
        if out.numel() != 12 or out.dtype != torch.int64 or out.device != self.device or not out.is_contiguous():
            raise ValueError("stats buffer must be a contiguous 12 x int64 tensor on the env's device")
        check(self._L.mg_stats_async(self._h, C.c_void_p(out.data_ptr()), int(reset), self._stream()), self._h, "mg_stats_async")
        return out

    def set_option(self, name: str, value: int) -> None:
        """Tuning switch of the library (mg_set_option); results never depend on it."""
        check(self._L.mg_set_option(self._h, name.encode(), int(value)), self._h, "mg_set_option")

    def set_kernel_timing(self, enabled: bool = True) -> None:
        check(self._L.mg_set_kernel_timing(self._h, int(enabled)), self._h, "mg_set_kernel_timing")

    def kernel_times(self) -> dict:
        """Mean device time (ms) of the four step kernels over the steps since the last call."""
        ms, n = (C.c_double * 4)(), C.c_int64()
        check(self._L.mg_kernel_times(self._h, ms, C.byref(n)), self._h, "mg_kernel_times")
        return {"screen_ms": ms[0], "decide_ms": ms[1], "update_ms": ms[2], "observe_ms": ms[3], "steps": n.value}

    @property
    def launch_count(self) -> int:
        return int(self._L.mg_launch_count(self._h))
