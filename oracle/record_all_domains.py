"""TEST INFRASTRUCTURE (build container only): every domain the reference ships (ui/domains/*.json, first JSON line,
pixel / 100 -- geometry.py:34-52) plus its np.dot-based original area (components_core.py:485-487), for the
differential soak (tests/soak.py --all-domains).  Domains on which the reference itself fails at reset or within
a short random rollout (e.g. easy1: ZeroDivisionError) are listed and skipped.  Writes tests/golden/domains_all.npz."""
import glob
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.c_oracle import OracleEnv, numpy_poly_area  # noqa: E402

REF = "/root/reference/ui/domains"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "domains_all.npz")


def main():
    out, skipped = {}, []
    for f in sorted(glob.glob(os.path.join(REF, "*.json"))):
        name = os.path.basename(f)[:-5]
        try:
            pts = json.loads(open(f).readline())
            xy = np.array([(p[0] / 100.0, p[1] / 100.0) for p in pts], dtype=np.float64)
            if len(xy) < 6:
                raise ValueError("fewer than 6 vertices")
            o = OracleEnv(xy)
            o.reset()
            if o.crashed or o.ref_index < 0:
                raise ValueError("reset fails (the reference raises on this domain)")
            o.run_random(1, 400)
            if o.crashed:
                raise ValueError("random rollout hits behaviour the reference leaves undefined")
        except Exception as ex:  # noqa: BLE001
            skipped.append((name, str(ex)))
            continue
        out[name] = xy
        out["area__" + name] = np.float64(numpy_poly_area(xy))
    np.savez_compressed(OUT, **out)
    print(len(out) // 2, "domains,", os.path.getsize(OUT), "bytes; skipped:", skipped)


if __name__ == "__main__":
    main()
