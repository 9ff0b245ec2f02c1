"""Batched counterpart of the reference's evaluation loop (v2/src/mesh_rl/evaluation/eval_loop.py:48-113, legacy
rl/baselines/testbed.py:107-229): one environment per domain, all stepped together, one episode each per model.

The reference loads an SB3 model per domain and loops ``model.predict -> env.step`` until done, recording
``info["is_complete"]`` and ``len(env.generated_meshes)``; the summary has the same shape here:
``{model_id: {"completed": [0/1 per domain], "n_elements": [count per domain]}}``.  ``predict`` is any callable
``obs[N,18] float32 tensor -> actions[N,3]`` (an SB3 policy's ``predict`` on ``obs.cpu().numpy()`` works too); the
single-env ``BoudaryEnv`` facade remains a drop-in for the reference's own ``evaluate_models``."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch


def evaluate_models(make_env: Callable[[], "object"], models: Mapping[str, Callable], *, max_steps: int = 100_000,
                    save_summary: Optional[str] = None, mesh_dir: Optional[str] = None,
                    domain_names: Optional[Sequence[str]] = None) -> Dict[str, Dict[str, List[int]]]:
    """``make_env()`` -> a ``BatchedBoudaryEnv`` with one env per domain and ``auto_reset=False`` (so that each env
    keeps its final mesh, like the reference env does until ``reset``).  Returns the reference's summary dict and
    optionally writes it (``evaluation_summary.json``, eval_loop.py:105-111) and the generated meshes
    (``write_2_file`` JSON per domain, testbed.py:182-191)."""
    results: Dict[str, Dict[str, List[int]]] = {}
    for model_id, predict in models.items():
        env = make_env()
        if getattr(env, "auto_reset", False):
            raise ValueError("evaluation needs auto_reset=False: a finished env must keep its mesh")
        n = env.num_envs
        obs = env.reset()
        done = torch.zeros(n, dtype=torch.bool, device=obs.device)
        completed = torch.zeros(n, dtype=torch.bool, device=obs.device)
        n_el = torch.zeros(n, dtype=torch.int32, device=obs.device)
        def save_mesh(e):
            from .export import write_2_file
            out = Path(mesh_dir)
            out.mkdir(parents=True, exist_ok=True)
            quads, vxy, _ = env.get_elements(e)
            n0 = len(env.domains[env.env_domain[e]]) if getattr(env, "domains", None) else 0
            name = domain_names[e] if domain_names else f"domain{e}"
            write_2_file(out / f"{model_id}_{name}.json", n0, quads, vxy)

        for _ in range(max_steps):
            act = predict(obs)
            if not isinstance(act, torch.Tensor):
                act = torch.as_tensor(np.asarray(act, dtype=np.float32))
            r = env.step(act.to(device=obs.device, dtype=torch.float32).contiguous())
            fin = (r.terminated | r.truncated).bool() & ~done
            completed |= fin & r.terminated.bool()
            n_el = torch.where(fin, r.n_elements.to(torch.int32), n_el)
            done |= fin
            obs = r.obs
            if mesh_dir is not None and bool(fin.any()):
                # the batch keeps stepping finished envs (a truncated env would go on meshing): the episode's mesh is
                # the one at the step it finished
                for e in fin.nonzero().flatten().cpu().tolist():
                    save_mesh(e)
            if bool(done.all()):
                break
        n_el = torch.where(done, n_el, env.n_elements.to(torch.int32))        # episodes cut by max_steps
        results[model_id] = {"completed": [int(x) for x in completed.cpu().tolist()],
                             "n_elements": [int(x) for x in n_el.cpu().tolist()]}
        if mesh_dir is not None:
            for e in (~done).nonzero().flatten().cpu().tolist():       # episodes cut by max_steps
                save_mesh(e)
        env.close()
    if save_summary is not None:
        Path(save_summary).parent.mkdir(parents=True, exist_ok=True)
        with open(save_summary, "w", encoding="utf-8") as f:
            json.dump(results, f)
    return results
