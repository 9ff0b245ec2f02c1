"""TEST INFRASTRUCTURE: record the polygon sample of bench.py's CPU arm (workloads c3 / c4).

Run once on a GPU box (gpurun): reads back the first-episode polygons of global env ids 0..63 of the device
generator (seed 2026, 64..512 vertices -- the bench configuration) and writes gpurun_out/c3_polys.npz, which is
committed as tests/golden/c3_polys.npz.  bench.py --impl reference then needs neither a GPU nor the product
package to run the same polygons through the CPU oracle / the Python reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv
    n = 64
    env = BatchedBoudaryEnv(None, num_envs=n, random_polygons=dict(min_verts=64, max_verts=512), seed=2026)
    env.reset()
    polys = {f"p{e}": env.get_state(e)["xy"] for e in range(n)}
    for e in range(n):
        assert np.array_equal(polys[f"p{e}"], env.debug_polygon(e, 0)["xy"])
    out = os.path.join(ROOT, "gpurun_out", "c3_polys.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **polys)
    print("wrote", out, "vertex counts", sorted(len(p) for p in polys.values())[::8])


if __name__ == "__main__":
    main()
