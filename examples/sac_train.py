#!/usr/bin/env python
"""BASELINE config 5 end to end on the device: SAC on the batched BoudaryEnv with a device-resident
replay buffer.  Nothing crosses PCIe inside the loop: observations, actions, rewards,
the replay ring (``mg_replay_add``) and the learner all live on the GPU.

Network shapes and hyper-parameters are the reference's SAC settings
(v2/src/mesh_rl/algorithms/sb3_algos.py:56-68, v2/config/train_sac_basic_legacy.yaml:24-33): MlpPolicy
net_arch [128,128,128] ReLU, lr 3e-4, batch 100, learning_starts 10 000, gamma 0.5; SB3 defaults for the rest
(tau 0.005, automatic entropy coefficient, target entropy -|A|, one gradient step per vector step).

    python examples/sac_train.py --domain tests/golden/domains.npz:random1_1 --envs 4096 --steps 2000
    python examples/sac_train.py --random --envs 16384 --steps 1000 --updates-per-step 4 --batch 4096
"""
import argparse
import copy
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning4meshgeneration_b200 import ACTION_HIGH, ACTION_LOW, BatchedBoudaryEnv  # noqa: E402
from reinforcementlearning4meshgeneration_b200.replay import DeviceReplayBuffer  # noqa: E402


def mlp(i, o, hidden=(128, 128, 128)):
    layers, d = [], i
    for h in hidden:
        layers += [nn.Linear(d, h), nn.ReLU()]
        d = h
    return nn.Sequential(*layers, nn.Linear(d, o))


class Actor(nn.Module):
    def __init__(self):
        super().__init__()
        self.net = mlp(18, 6)

    def forward(self, obs):
        mu, log_std = self.net(obs).chunk(2, dim=-1)
        log_std = log_std.clamp(-20, 2)
        u = mu + log_std.exp() * torch.randn_like(mu)
        a = torch.tanh(u)
        logp = (-0.5 * ((u - mu) / log_std.exp()) ** 2 - log_std - 0.5 * np.log(2 * np.pi)).sum(-1, keepdim=True)
        logp = logp - torch.log(1 - a.pow(2) + 1e-6).sum(-1, keepdim=True)
        return a, logp


class SAC:
    """Soft actor-critic learner with the reference's settings (twin critics, polyak targets, automatic entropy
    coefficient with target entropy -|A|); ``update`` takes one batch of the device replay buffer."""

    def __init__(self, dev, gamma=0.5, lr=3e-4, tau=0.005):
        self.gamma, self.tau = gamma, tau
        self.actor, self.q1, self.q2 = Actor().to(dev), mlp(21, 1).to(dev), mlp(21, 1).to(dev)
        self.q1_t, self.q2_t = copy.deepcopy(self.q1), copy.deepcopy(self.q2)
        self.log_alpha = torch.zeros(1, device=dev, requires_grad=True)
        self.opt_a = torch.optim.Adam(self.actor.parameters(), lr=lr)
        self.opt_q = torch.optim.Adam(list(self.q1.parameters()) + list(self.q2.parameters()), lr=lr)
        self.opt_al = torch.optim.Adam([self.log_alpha], lr=lr)
        self.target_entropy = -3.0
        self.low = torch.from_numpy(ACTION_LOW.copy()).to(dev)
        self.high = torch.from_numpy(ACTION_HIGH.copy()).to(dev)

    def update(self, b):
        """One gradient step on critics, actor and entropy coefficient; returns (critic loss, actor loss) tensors."""
        a_unit = (b.actions - self.low) / (self.high - self.low) * 2 - 1
        alpha = self.log_alpha.exp().detach()
        with torch.no_grad():
            na, nlogp = self.actor(b.next_observations)
            nx = torch.cat([b.next_observations, na], 1)
            target = b.rewards + (1 - b.dones) * self.gamma * (torch.min(self.q1_t(nx), self.q2_t(nx)) - alpha * nlogp)
        x = torch.cat([b.observations, a_unit], 1)
        loss_q = F.mse_loss(self.q1(x), target) + F.mse_loss(self.q2(x), target)
        self.opt_q.zero_grad(set_to_none=True); loss_q.backward(); self.opt_q.step()
        pa, plogp = self.actor(b.observations)
        px = torch.cat([b.observations, pa], 1)
        loss_a = (alpha * plogp - torch.min(self.q1(px), self.q2(px))).mean()
        self.opt_a.zero_grad(set_to_none=True); loss_a.backward(); self.opt_a.step()
        loss_al = -(self.log_alpha * (plogp.detach() + self.target_entropy).mean())
        self.opt_al.zero_grad(set_to_none=True); loss_al.backward(); self.opt_al.step()
        with torch.no_grad():
            for net, tgt in ((self.q1, self.q1_t), (self.q2, self.q2_t)):
                for p_, tp in zip(net.parameters(), tgt.parameters()):
                    tp.lerp_(p_, self.tau)
        return loss_q.detach(), loss_a.detach()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--domain", default=None, help="<npz>:<key> or a reference-style domain .json file")
    ap.add_argument("--random", action="store_true", help="random star polygons (BASELINE config 3 generator)")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=1000, help="vector steps")
    ap.add_argument("--buffer-steps", type=int, default=256, help="replay capacity in vector steps")
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--updates-per-step", type=int, default=1)
    ap.add_argument("--learning-starts", type=int, default=10_000, help="transitions collected with uniform actions first")
    ap.add_argument("--gamma", type=float, default=0.5)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--tau", type=float, default=0.005)
    ap.add_argument("--log-every", type=int, default=100)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if args.random or args.domain is None:
        env = BatchedBoudaryEnv(None, num_envs=args.envs, device=dev, random_polygons=dict(min_verts=64, max_verts=256), seed=1)
    elif ".npz:" in args.domain:
        path, key = args.domain.split(":")
        env = BatchedBoudaryEnv([np.load(path)[key]], num_envs=args.envs, device=dev)
    else:
        from reinforcementlearning4meshgeneration_b200.domains import load_domain
        env = BatchedBoudaryEnv([load_domain(args.domain)], num_envs=args.envs, device=dev)
    low, high = torch.from_numpy(ACTION_LOW.copy()).to(dev), torch.from_numpy(ACTION_HIGH.copy()).to(dev)
    to_env = lambda a: (low + (a + 1) * 0.5 * (high - low)).contiguous()

    learner = SAC(dev, gamma=args.gamma, lr=args.lr, tau=args.tau)
    actor = learner.actor

    buf = DeviceReplayBuffer(env, args.buffer_steps)
    obs = env.reset().clone()
    t0 = time.perf_counter()
    n_updates = 0
    for t in range(args.steps):
        with torch.no_grad():
            if len(buf) < args.learning_starts:
                act = env.sample_actions(123, t)
            else:
                act = to_env(actor(obs)[0])
        r = env.step(act)
        buf.add(obs, act, r)
        obs.copy_(r.obs)
        if len(buf) >= args.learning_starts:
            for _ in range(args.updates_per_step):
                learner.update(buf.sample(args.batch))
                n_updates += 1
        if (t + 1) % args.log_every == 0:
            s = env.stats(reset=True)            # the only host sync of the loop
            if s["episodes"]:
                print(f"step {t+1}: episodes {s['episodes']} completed {s['completed']} mean return "
                      f"{s['sum_return']/s['episodes']:.3f} mean elements {s['elements']/s['episodes']:.1f} "
                      f"success rate {s['successes']/max(1, s['steps']):.3f} updates {n_updates} alpha {learner.log_alpha.exp().item():.3f}", flush=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{args.envs} envs x {args.steps} steps in {dt:.1f} s: {args.envs*args.steps/dt:.3e} env-steps/s with SAC in the loop "
          f"({n_updates} gradient steps, replay {len(buf)} transitions on the device)")


if __name__ == "__main__":
    main()
