"""Device-resident replay buffer for the batched BoudaryEnv.

The reference trains with SB3's off-policy algorithms (rl/baselines/RL_Mesh.py:186-197,
v2 training/train_loop.py:132): ``OffPolicyAlgorithm._store_transition`` swaps the terminal observation
in for finished envs and ``ReplayBuffer.add`` copies the step into numpy rings on the host.  Here the N
transitions of a step are written by ONE kernel launch (``mg_replay_add``) into torch tensors that
never leave the device; ``sample`` draws uniformly over the stored transitions like
``ReplayBuffer.sample`` / ``_get_samples`` (dones are masked by the timeout flag:
``handle_timeout_termination``)."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch

from ._lib import check
from .batched_env import ACT_DIM, OBS_DIM, BatchedBoudaryEnv


class ReplaySamples(NamedTuple):
    observations: torch.Tensor        # [B, 18] f32
    actions: torch.Tensor             # [B, 3]  f32 (env action box, not squashed)
    next_observations: torch.Tensor   # [B, 18] f32 (terminal observation where the episode ended)
    dones: torch.Tensor               # [B, 1]  f32: done and not a time-limit truncation
    rewards: torch.Tensor             # [B, 1]  f32


class DeviceReplayBuffer:
    """Ring of ``capacity_steps`` slots x ``env.num_envs`` transitions, all on ``env.device``."""

    def __init__(self, env: BatchedBoudaryEnv, capacity_steps: int, handle_timeout_termination: bool = True):
        self.env = env
        self.N = env.num_envs
        self.capacity_steps = int(capacity_steps)
        if self.capacity_steps <= 0:
            raise ValueError("capacity_steps must be positive")
        dev, S, N = env.device, self.capacity_steps, self.N
        self.obs = torch.zeros((S, N, OBS_DIM), dtype=torch.float32, device=dev)
        self.next_obs = torch.zeros((S, N, OBS_DIM), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((S, N, ACT_DIM), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((S, N), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((S, N), dtype=torch.uint8, device=dev)
        self.timeouts = torch.zeros((S, N), dtype=torch.uint8, device=dev)
        self.handle_timeout_termination = handle_timeout_termination
        self.pos = 0
        self.full = False

    def __len__(self) -> int:
        return (self.capacity_steps if self.full else self.pos) * self.N

    def add(self, prev_obs: torch.Tensor, actions: torch.Tensor, step_result) -> None:
        """Store the step ``prev_obs --actions--> step_result`` (a ``StepResult`` of ``env.step``).
        ``prev_obs`` must be a tensor that the step did not overwrite (clone ``env.obs`` before stepping)."""
        r = step_result
        for t, dt in ((prev_obs, torch.float32), (actions, torch.float32)):
            if t.device != self.env.device or t.dtype != dt or not t.is_contiguous():
                raise ValueError("prev_obs / actions must be contiguous float32 tensors on the env's device")
        p = lambda t: C.c_void_p(t.data_ptr())
        env = self.env
        check(env._L.mg_replay_add(env._h, self.capacity_steps, self.pos, p(self.obs), p(self.next_obs), p(self.actions),
                                   p(self.rewards), p(self.dones), p(self.timeouts), p(prev_obs), p(actions), p(r.obs),
                                   p(r.reward), p(r.terminated), p(r.truncated), p(r.terminal_obs), env._stream()),
              env._h, "mg_replay_add")
        self.pos += 1
        if self.pos == self.capacity_steps:
            self.full, self.pos = True, 0

    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None) -> ReplaySamples:
        n = len(self)
        if n == 0:
            raise RuntimeError("the replay buffer is empty")
        idx = torch.randint(0, n, (batch_size,), device=self.env.device, generator=generator)
        obs = self.obs.view(-1, OBS_DIM)[idx]
        nxt = self.next_obs.view(-1, OBS_DIM)[idx]
        act = self.actions.view(-1, ACT_DIM)[idx]
        rew = self.rewards.view(-1)[idx].unsqueeze(1)
        done = self.dones.view(-1)[idx].to(torch.float32)
        if self.handle_timeout_termination:
            done = done * (1.0 - self.timeouts.view(-1)[idx].to(torch.float32))
        return ReplaySamples(obs, act, nxt, done.unsqueeze(1), rew)
