#!/usr/bin/env python
"""Differential soak of the CUDA path against the CPU oracle (test infrastructure; run by hand on a GPU box):

    python tests/soak.py [--per-domain 256] [--steps 1000] [--random-envs 8192] [--random-steps 400] [--seed 99]

(a) every committed domain (or, with --all-domains, all 58 usable domains of the reference) x `per-domain` envs x `steps` steps with auto-reset, device Philox actions with a
    share of dyadic action components (exact rounding ties of the new vertex, E:202-210);
(b) `random-envs` random star polygons (BASELINE config 3 generator), first episode of each env.
Every env is replayed through oracle/liboracle.so on all host threads; every mismatch is listed (not just the
first) and written to gpurun_out/soak_report.json so that it can be reproduced on a CPU-only machine (polygon + the float32 action bits of each failing env)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from helpers import assert_rollout_matches, load_domains  # noqa: E402

LOW_A = np.array([-1.0, -1.5, 0.0], np.float32)
HIGH_A = np.array([1.0, 1.5, 1.5], np.float32)


def record(env, T, seed, dyadic_share, rng):
    import torch
    dev, N = env.device, env.num_envs
    rec = dict(act=torch.zeros((T, N, 3), device=dev), obs=torch.zeros((T, N, 18), device=dev),
               tobs=torch.zeros((T, N, 18), device=dev), rew=torch.zeros((T, N), dtype=torch.float64, device=dev),
               te=torch.zeros((T, N), dtype=torch.uint8, device=dev), tr=torch.zeros((T, N), dtype=torch.uint8, device=dev),
               ne=torch.zeros((T, N), dtype=torch.int32, device=dev))
    for t in range(T):
        a = env.sample_actions(seed, t)
        if dyadic_share > 0:
            # snap a share of the (x, y) action components to multiples of 1/16 (exact in float32 and in 4 decimals)
            m = torch.rand((N, 1), device=dev) < dyadic_share
            a = torch.where(m.expand(-1, 3) & torch.tensor([False, True, True], device=dev), torch.round(a * 16) / 16, a)
        rec["act"][t] = a
        r = env.step(a)
        rec["obs"][t] = r.obs; rec["tobs"][t] = r.terminal_obs; rec["rew"][t] = r.reward
        rec["te"][t] = r.terminated; rec["tr"][t] = r.truncated; rec["ne"][t] = r.n_elements
    return {k: v.cpu().numpy() for k, v in rec.items()}


def replay(rec, e, xy, area, what, first_episode_only=False):
    from oracle.c_oracle import OracleEnv
    o = OracleEnv(xy, original_area=area)
    exp = o.rollout(rec["act"][:, e])
    T = len(exp["reward"])
    L = T
    if first_episode_only:
        d = np.nonzero(exp["terminated"] | exp["truncated"])[0]
        L = int(d[0]) + 1 if d.size else T
    got = dict(obs=rec["obs"][:L, e].copy(), terminal_obs=rec["tobs"][:L, e], reward=rec["rew"][:L, e], terminated=rec["te"][:L, e],
               truncated=rec["tr"][:L, e], n_elements=rec["ne"][:L, e])
    exp = {k: v[:L].copy() for k, v in exp.items()}
    if first_episode_only and L <= T and (exp["terminated"][L - 1] or exp["truncated"][L - 1]):
        got["obs"][L - 1] = exp["obs"][L - 1]          # the reset observation belongs to a fresh polygon
    try:
        assert_rollout_matches(got, exp, what, reward_tol=1e-9)
    except AssertionError as ex:
        return dict(what=what, env=int(e), error=str(ex)[:600], xy=np.asarray(xy).tolist(), area=float(area),
                    actions=rec["act"][:L, e].view(np.uint32).tolist()), L, int(exp["success"].sum())
    return None, L, int(exp["success"].sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-domain", type=int, default=256)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--random-envs", type=int, default=8192)
    ap.add_argument("--random-steps", type=int, default=400)
    ap.add_argument("--seed", type=int, default=99)
    ap.add_argument("--dyadic", type=float, default=0.15)
    ap.add_argument("--all-domains", action="store_true", help="all 58 usable reference domains (tests/golden/domains_all.npz)")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "gpurun_out", "soak_report.json"))
    args = ap.parse_args()
    import torch
    from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
    torch.manual_seed(args.seed)
    rng = np.random.default_rng(args.seed)
    report = dict(args=vars(args), failures=[], parts=[])

    if args.per_domain > 0:
        doms, areas = load_domains()
        if args.all_domains:
            z = np.load(os.path.join(HERE, "golden", "domains_all.npz"))
            doms = {k: z[k] for k in z.files if not k.startswith("area__")}
            areas = {k[6:]: float(z[k]) for k in z.files if k.startswith("area__")}
        names = sorted(doms)
        N = args.per_domain * len(names)
        env_domain = np.repeat(np.arange(len(names)), args.per_domain)
        env = BatchedBoudaryEnv([doms[k] for k in names], num_envs=N, env_domain=env_domain)
        env.reset()
        t0 = time.time()
        rec = record(env, args.steps, args.seed, args.dyadic, rng)
        t1 = time.time()
        with ThreadPoolExecutor(max_workers=args.threads) as ex:
            res = list(ex.map(lambda e: replay(rec, e, doms[names[env_domain[e]]], areas[names[env_domain[e]]],
                                               f"domain {names[env_domain[e]]} env {e} seed {args.seed}"), range(N)))
        fails = [r[0] for r in res if r[0]]
        part = dict(part="domains", envs=N, env_steps=int(sum(r[1] for r in res)), elements=int(sum(r[2] for r in res)),
                    failures=len(fails), gpu_s=round(t1 - t0, 1), oracle_s=round(time.time() - t1, 1))
        report["parts"].append(part); report["failures"] += fails
        print(json.dumps(part), flush=True)
        env.close()

    if args.random_envs > 0:
        N = args.random_envs
        env = BatchedBoudaryEnv(None, num_envs=N, random_polygons=dict(min_verts=64, max_verts=512), seed=args.seed, auto_reset=False)
        env.reset()
        states = [env.get_state(e) for e in range(N)]
        t0 = time.time()
        rec = record(env, args.random_steps, args.seed + 1, args.dyadic, rng)
        t1 = time.time()
        with ThreadPoolExecutor(max_workers=args.threads) as ex:
            res = list(ex.map(lambda e: replay(rec, e, states[e]["xy"], states[e]["original_area"],
                                               f"random polygon env {e} seed {args.seed}", first_episode_only=True), range(N)))
        fails = [r[0] for r in res if r[0]]
        part = dict(part="random polygons (first episodes)", envs=N, env_steps=int(sum(r[1] for r in res)),
                    elements=int(sum(r[2] for r in res)), failures=len(fails), gpu_s=round(t1 - t0, 1), oracle_s=round(time.time() - t1, 1))
        report["parts"].append(part); report["failures"] += fails
        print(json.dumps(part), flush=True)
        env.close()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)
    for f in report["failures"][:20]:
        print("FAIL", f["what"], "\n   ", f["error"].replace("\n", "\n    "), flush=True)
    print("soak:", "OK" if not report["failures"] else f"{len(report['failures'])} env(s) differ")
    return 1 if report["failures"] else 0


if __name__ == "__main__":
    sys.exit(main())
