"""CPU: the batched evaluation loop's host logic (summary shape of eval_loop.py:48-113) on a scripted fake env."""
import json
import types

import torch

from reinforcementlearning4meshgeneration_b200.evaluation import evaluate_models


class FakeEnv:
    """env i finishes at step 3 + i; even envs complete, odd ones are truncated; element count = 10 * (i + 1)."""
    auto_reset = False

    def __init__(self, n):
        self.num_envs = n
        self.t = 0
        self.n_elements = torch.zeros(n, dtype=torch.int32)
        self.closed = False

    def reset(self):
        self.t = 0
        return torch.zeros((self.num_envs, 18))

    def step(self, act):
        assert act.shape == (self.num_envs, 3) and act.dtype == torch.float32
        self.t += 1
        idx = torch.arange(self.num_envs)
        fin = self.t >= 3 + idx                      # stays "done" afterwards, like a finished reference env
        term = (fin & (idx % 2 == 0)).to(torch.uint8)
        trunc = (fin & (idx % 2 == 1)).to(torch.uint8)
        self.n_elements = torch.where(self.t >= 3 + idx, 10 * (idx + 1), self.t * torch.ones_like(idx)).to(torch.int32)
        return types.SimpleNamespace(obs=torch.full((self.num_envs, 18), float(self.t)), terminated=term, truncated=trunc,
                                     n_elements=self.n_elements)

    def close(self):
        self.closed = True


def test_summary_matches_the_reference_shape(tmp_path):
    envs = []

    def make_env():
        envs.append(FakeEnv(4))
        return envs[-1]

    calls = []

    def policy(obs):
        calls.append(float(obs[0, 0]))
        return torch.zeros((obs.shape[0], 3))

    out = evaluate_models(make_env, {"1200": policy, "1201": lambda o: [[0.0, 0.0, 0.0]] * o.shape[0]},
                          save_summary=str(tmp_path / "evaluation_summary.json"))
    assert out == {"1200": {"completed": [1, 0, 1, 0], "n_elements": [10, 20, 30, 40]},
                   "1201": {"completed": [1, 0, 1, 0], "n_elements": [10, 20, 30, 40]}}
    assert json.load(open(tmp_path / "evaluation_summary.json")) == out
    assert len(calls) == 6 and calls[:3] == [0.0, 1.0, 2.0]      # stops when the slowest env (3 + 3 steps) is done
    assert all(e.closed for e in envs)


def test_max_steps_cuts_unfinished_episodes():
    out = evaluate_models(lambda: FakeEnv(3), {"m": lambda o: torch.zeros((o.shape[0], 3))}, max_steps=4)
    assert out["m"]["completed"] == [1, 0, 0] and out["m"]["n_elements"] == [10, 20, 4]
