"""GPU-side helpers for the parity tests: run the batched CUDA env over per-env action streams."""
import numpy as np
import torch


def run_gpu(env, actions_tn3, state_every=0, state_envs=()):
    """actions_tn3: np.float32 [T, N, 3].  Returns dict of np arrays [T, N, ...] (+ sampled states)."""
    T, N, _ = actions_tn3.shape
    dev = env.device
    acts = torch.from_numpy(np.ascontiguousarray(actions_tn3)).to(dev)
    out = dict(obs=torch.zeros((T, N, 18), dtype=torch.float32, device=dev),
               terminal_obs=torch.zeros((T, N, 18), dtype=torch.float32, device=dev),
               reward=torch.zeros((T, N), dtype=torch.float64, device=dev),
               terminated=torch.zeros((T, N), dtype=torch.uint8, device=dev),
               truncated=torch.zeros((T, N), dtype=torch.uint8, device=dev),
               n_elements=torch.zeros((T, N), dtype=torch.int32, device=dev))
    states = {}
    for t in range(T):
        r = env.step(acts[t])
        out["obs"][t].copy_(r.obs)
        out["terminal_obs"][t].copy_(r.terminal_obs)
        out["reward"][t].copy_(r.reward)
        out["terminated"][t].copy_(r.terminated)
        out["truncated"][t].copy_(r.truncated)
        out["n_elements"][t].copy_(r.n_elements)
        if state_every and t % state_every == 0:
            for e in state_envs:
                states[(t, e)] = env.get_state(e)
    torch.cuda.synchronize(dev)
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res["states"] = states
    return res


def per_env(res, e):
    return {k: v[:, e] for k, v in res.items() if k != "states"}
