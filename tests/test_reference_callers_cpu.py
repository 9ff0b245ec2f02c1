"""CPU: the reference's own, unmodified callers run on the drop-in facade (SURVEY 8b "import points that must resolve
unchanged").  No GPU here, so the facade's backend factory is swapped for a fake ``BatchedBoudaryEnv`` driven by the C
oracle; everything above it (module surface, API flavours, generated_meshes, save_meshes, the VecEnv ``.envs`` shim) is
the product code.

* the v2 evaluation loop ``mesh_rl.evaluation.eval_loop.evaluate_models`` (eval_loop.py:48-113) is imported from the
  reference tree with the facade mounted as ``mesh_rl.envs.boundary_env`` and a scripted stand-in for the SB3 model;
  its summary must equal the one the same loop produces on the reference's own env;
* the legacy flavour (``rl.boundary_env``: ``reset() -> obs``, 4-tuple ``step``; rl/boundary_env.py:67-84, :263) is
  driven the way rl/baselines/testbed.py:160-195 drives it.
"""
import importlib
import json
import sys

import numpy as np
import pytest
import torch

from helpers import HIGH, LOW, load_domains
from oracle import ref_loader
from oracle.c_oracle import OracleEnv


class OracleBackend:
    """The slice of BatchedBoudaryEnv's surface the single-env facade uses, answered by the CPU oracle."""
    auto_reset = False
    num_envs = 1

    def __init__(self, xy, device=None, auto_reset=False, log_capacity=None, **kw):
        self.o = OracleEnv(xy)
        self.closed = False

    def reset(self, mask=None):
        return torch.from_numpy(self.o.reset()[None].copy())

    def step_host(self, a, out=None):
        a = np.asarray(a, np.float32).reshape(3)
        obs, r, te, tr, _ = self.o.step(a)
        if out is None:
            out = dict(obs=np.zeros((1, 18), np.float32), reward=np.zeros(1), terminated=np.zeros(1, np.uint8),
                       truncated=np.zeros(1, np.uint8), terminal_obs=np.zeros((1, 18), np.float32), n_elements=np.zeros(1, np.int32))
        out["obs"][0] = 0 if obs is None else obs
        out["reward"][0], out["terminated"][0], out["truncated"][0] = r, te, tr
        out["n_elements"][0] = self.o.n_elements
        return out

    def get_elements(self, e, allow_truncated=False):
        return self.o.elements(), self.o.vertex_xy(), self.o.n_elements

    def n_elements_of(self, e):
        return self.o.n_elements

    def get_state(self, e):
        ids, xy = self.o.boundary()
        return dict(n=self.o.n, n0=self.o.n0, xy=xy, ids=ids, ref_index=self.o.ref_index, n_elements=self.o.n_elements)

    def close(self):
        self.closed = True


class ScriptedModel:
    """Stand-in for an SB3 model: ``predict`` replays a seeded uniform action stream (one draw per call)."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def predict(self, obs, deterministic=False):
        assert np.asarray(obs).shape == (18,) and np.asarray(obs).dtype == np.float32
        return self.rng.uniform(LOW, HIGH).astype(np.float32), None


@pytest.fixture
def facade(monkeypatch):
    import reinforcementlearning4meshgeneration_b200.boundary_env as m
    monkeypatch.setattr(m, "_BACKEND_FACTORY", lambda xy, **kw: OracleBackend(xy, **kw))
    return m


@pytest.mark.skipif(not ref_loader.reference_available(), reason="the reference tree is only present in the build container")
def test_reference_evaluate_models_runs_unmodified_on_the_facade(facade, monkeypatch, tmp_path):
    from pathlib import Path
    ref_loader.load_reference()                       # stubs for gymnasium / matplotlib / SB3, mesh_rl package path
    sb3 = sys.modules["stable_baselines3"]

    class Loader:
        @staticmethod
        def load(model_path, env=None):
            assert hasattr(env, "reset") and hasattr(env, "step")
            return ScriptedModel(int(Path(model_path).stem.split("_")[-1]))

    for algo in ("A2C", "DDPG", "PPO", "SAC", "TD3"):
        monkeypatch.setattr(sb3, algo, Loader, raising=False)
    sys.modules.pop("mesh_rl.evaluation.eval_loop", None)
    sys.modules.pop("mesh_rl.evaluation", None)
    ev_ref = importlib.import_module("mesh_rl.evaluation.eval_loop")
    cfg_mod = importlib.import_module("mesh_rl.config")
    paths = cfg_mod.PathConfig(project_root=Path(ref_loader.REFERENCE_ROOT), outputs_root=tmp_path / "ref_out")
    cfg = ev_ref.EvalConfig(algo="sac", model_paths=[Path("model_11.zip"), Path("model_12.zip")],
                            domains=["half_wheel", "star", "tool"], version="t")
    res_ref = ev_ref.evaluate_models(cfg, paths=paths, save_summary=False)
    assert sum(map(sum, (r["n_elements"] for r in res_ref.values()))) > 10

    # the same loop, unmodified, with the facade mounted where the reference's env module lives
    monkeypatch.setitem(sys.modules, "mesh_rl.envs.boundary_env", facade)
    sys.modules.pop("mesh_rl.evaluation.eval_loop", None)
    ev_new = importlib.import_module("mesh_rl.evaluation.eval_loop")
    assert ev_new.BoudaryEnv is facade.BoudaryEnv
    paths2 = cfg_mod.PathConfig(project_root=Path(ref_loader.REFERENCE_ROOT), outputs_root=tmp_path / "new_out")
    res_new = ev_new.evaluate_models(cfg, paths=paths2, save_summary=True)
    assert res_new == res_ref
    assert json.load(open(tmp_path / "new_out" / "evaluation" / "t" / "evaluation_summary.json")) == res_ref
    sys.modules.pop("mesh_rl.evaluation.eval_loop", None)


def test_legacy_gym_flavour_like_testbed(facade, tmp_path):
    """rl/baselines/testbed.py:160-195 on ``rl.boundary_env`` = the legacy flavour: reset() -> obs, 4-tuple step,
    len(env.generated_meshes), env.save_meshes(..., meshes=env.generated_meshes, ...), env.boundary / original_vertices /
    updated_boundary."""
    import reinforcementlearning4meshgeneration_b200.legacy as legacy
    assert {"BoudaryEnv", "read_polygon", "boundary"} <= set(dir(legacy))
    doms, areas = load_domains()
    env = legacy.BoudaryEnv(legacy.boundary())
    o = OracleEnv(np.array([[v.x, v.y] for v in legacy.boundary().vertices], np.float64))
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (18,) and np.array_equal(obs, o.reset())
    model = ScriptedModel(3)
    steps = 0
    while True:
        action, _ = model.predict(obs)
        out = env.step(action)
        assert len(out) == 4
        obs, reward, done, info = out
        eo, er, te, tr, _ = o.step(action)
        assert reward == er and done == (te or tr) and info == {"is_complete": not tr}
        steps += 1
        if done:
            break
        assert np.array_equal(obs, eo)
    assert len(env.generated_meshes) == o.n_elements > 3
    assert len(env.original_vertices) == 30 and len(env.boundary.vertices) == 30
    ids, xy = o.boundary()
    assert [(v.x, v.y) for v in env.updated_boundary.vertices] == [tuple(p) for p in xy.tolist()]
    m0 = env.generated_meshes[0]
    assert len(m0.vertices) == 4 and np.asarray(m0).shape == (4, 2)
    path = env.save_meshes(tmp_path / "mesh.png", meshes=env.generated_meshes, indexing=True, style="k-", dpi=30)
    txt = open(path).read() if str(path).endswith(".svg") else None
    if txt is not None:
        assert txt.count("<polygon") == 1 + o.n_elements and txt.count("<text") == o.n_elements
    env.close()
    assert env._batched.closed


def test_vecenv_envs_shim(facade, tmp_path):
    """CustomizeCallback.py:131-133: ``env.envs[0].save_meshes(path, meshes=env.envs[0].generated_meshes, ...)``."""
    from reinforcementlearning4meshgeneration_b200.vec_env import SB3VecEnv

    class FakeMany(OracleBackend):
        auto_reset = True
        num_envs = 2

        def __init__(self, xy):
            self.envs_ = [OracleEnv(xy), OracleEnv(xy)]
            self.closed = False

        def reset(self):
            return torch.from_numpy(np.stack([e.reset() for e in self.envs_]))

        def step_host(self, act, out):
            a = act.numpy()
            for i, e in enumerate(self.envs_):
                obs, r, te, tr, _ = e.step(a[i])
                out["terminal_obs"][i] = torch.from_numpy(np.zeros(18, np.float32) if obs is None else obs)
                out["n_elements"][i] = e.n_elements
                if te or tr:
                    obs = e.reset()
                out["obs"][i] = torch.from_numpy(obs)
                out["reward"][i], out["terminated"][i], out["truncated"][i] = r, int(te), int(tr)
            return out

        def get_elements(self, i, allow_truncated=False):
            e = self.envs_[i]
            return e.elements(), e.vertex_xy(), e.n_elements

        def get_state(self, i):
            e = self.envs_[i]
            ids, xy = e.boundary()
            return dict(n=e.n, n0=e.n0, xy=xy, ids=ids, n_elements=e.n_elements)

        def n_elements_of(self, i):
            return self.envs_[i].n_elements

    doms, _ = load_domains()
    venv = SB3VecEnv(FakeMany(doms["boundary0"]))
    assert len(venv.envs) == 2
    venv.reset()
    rng = np.random.default_rng(0)
    for _ in range(40):
        venv.step(rng.uniform(LOW, HIGH, size=(2, 3)).astype(np.float32))
    e0 = venv.envs[0]
    n = len(e0.generated_meshes)
    assert n == venv.env_method("generated_meshes_count", indices=0)[0] == len(venv.get_attr("generated_meshes", 0)[0])
    assert len(e0.original_vertices) == 30 and len(e0.updated_boundary.vertices) == venv._b.envs_[0].n
    path = e0.save_meshes(tmp_path / "cb.png", meshes=e0.generated_meshes, indexing=True, style="k-", dpi=30)
    assert str(path).endswith((".png", ".svg"))
