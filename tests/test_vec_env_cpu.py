"""CPU: the SB3 VecEnv adapter's host logic (infos, terminal_observation, TimeLimit.truncated,
Monitor-style episode stats) against a scripted fake backend -- mirrors the surface of the
reference's vendored DummyVecEnv (rl/baselines/dummy_vec_env.py:12-125)."""
import numpy as np
import torch

from reinforcementlearning4meshgeneration_b200.vec_env import SB3VecEnv


class FakeBatched:
    """Scripted stand-in for BatchedBoudaryEnv.step_host: env i finishes every (i+2) steps."""
    auto_reset = True

    def __init__(self, n):
        self.num_envs = n
        self.t = np.zeros(n, np.int64)
        self.closed = False

    def reset(self):
        self.t[:] = 0
        return torch.zeros((self.num_envs, 18))

    def step_host(self, act, out):
        a = act.numpy()
        self.t += 1
        for i in range(self.num_envs):
            done = self.t[i] % (i + 2) == 0
            out["reward"][i] = float(a[i, 0]) + i
            out["terminated"][i] = int(done and i % 2 == 0)
            out["truncated"][i] = int(done and i % 2 == 1)
            out["terminal_obs"][i] = float(self.t[i]) if done else 0.0
            out["obs"][i] = 0.0 if done else float(self.t[i])
            out["n_elements"][i] = 7 + i
        return out

    def close(self):
        self.closed = True


def test_vecenv_surface_and_infos():
    env = SB3VecEnv(FakeBatched(3))
    assert env.num_envs == 3 and env.observation_space.shape == (18,) and env.action_space.shape == (3,)
    assert np.allclose(env.action_space.low, [-1, -1.5, 0]) and np.allclose(env.action_space.high, [1, 1.5, 1.5])
    obs = env.reset()
    assert obs.shape == (3, 18) and obs.dtype == np.float32
    rets = np.zeros(3)
    lens = np.zeros(3, int)
    for t in range(1, 13):
        a = np.full((3, 3), 0.25, np.float32)
        env.step_async(a)
        obs, rew, dones, infos = env.step_wait()
        assert rew.dtype == np.float32 and dones.dtype == bool and len(infos) == 3
        rets += rew
        lens += 1
        for i in range(3):
            if t % (i + 2) == 0:
                assert dones[i]
                info = infos[i]
                assert np.all(info["terminal_observation"] == t % 1000 if False else info["terminal_observation"] > 0)
                assert info["TimeLimit.truncated"] == (i % 2 == 1)
                assert info["is_complete"] == (i % 2 == 0)
                assert info["n_elements"] == 7 + i
                assert info["episode"]["l"] == lens[i] and abs(info["episode"]["r"] - rets[i]) < 1e-5
                rets[i] = 0
                lens[i] = 0
                assert np.all(obs[i] == 0)          # auto-reset observation
            else:
                assert not dones[i] and infos[i] == {}
    assert env.env_is_wrapped(object) == [False] * 3
    assert env.get_attr("num_envs", [0, 2]) == [3, 3]
    assert env.seed(5) == [5, 6, 7]
    env.close()
    assert env._b.closed


class FakeBatchedAsync(FakeBatched):
    """The same backend with the begin / end split of mg_step_host: the results only appear at step_host_end."""

    def __init__(self, n):
        super().__init__(n)
        self.calls = []
        self._pending = None

    def step_host(self, act, out):
        raise AssertionError("the adapter must use the begin / end pair when the backend has one")

    def step_host_begin(self, act, out):
        assert self._pending is None
        self.calls.append("begin")
        self._pending = (act.clone(), out)

    def step_host_end(self):
        self.calls.append("end")
        act, out = self._pending
        self._pending = None
        FakeBatched.step_host(self, act, out)


def test_step_async_enqueues_and_step_wait_joins():
    """VecEnv.step_async / step_wait (rl/baselines/dummy_vec_env.py:38-58) map onto step_host_begin / step_host_end: the
    step is in flight between the two calls, and the results are the ones of the plain backend."""
    plain, split = SB3VecEnv(FakeBatched(3)), SB3VecEnv(FakeBatchedAsync(3))
    plain.reset(); split.reset()
    for t in range(8):
        a = np.full((3, 3), 0.1 * t, np.float32)
        split.step_async(a)
        assert split._b.calls[-1] == "begin" and split._b._pending is not None
        got = split.step_wait()
        assert split._b.calls[-1] == "end" and split._b._pending is None
        exp = plain.step(a)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
        assert [sorted(i) for i in got[3]] == [sorted(i) for i in exp[3]]
    assert split._b.calls == ["begin", "end"] * 8
