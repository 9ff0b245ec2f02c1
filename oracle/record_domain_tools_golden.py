"""TEST INFRASTRUCTURE (build container only): golden vectors for reinforcementlearning4meshgeneration_b200/domains.py
recorded from the reference's own functions -- ui/GenerateRandomPolygon.py:generatePolygon and
ui/tk-ui.py:{clockwise_angle, calculate_density, check_clockwise}.  Both files run GUI / PIL code at import time, so
the functions are lifted out of the source with `ast` and executed unmodified.
Writes tests/golden/domain_tools.json."""
import ast
import json
import math
import os
import random
import sys
import types

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "domain_tools.json")


def lift(path, names):
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"math": math, "random": random, "clip": out.get("clip"), "clockwise_angle": out.get("clockwise_angle"), "print": lambda *a, **k: None}
            exec(compile(mod, path, "exec"), ns)
            out[node.name] = ns[node.name]
            # later functions may call earlier ones
            for f in out.values():
                f.__globals__.update(out)
    return out


def main():
    gen = lift(os.path.join(REF, "ui", "GenerateRandomPolygon.py"), {"generatePolygon", "clip"})
    tk = lift(os.path.join(REF, "ui", "tk-ui.py"), {"clockwise_angle", "calculate_density", "distance", "check_clockwise"})
    cases = {"generate": [], "densify": [], "clockwise": []}
    for seed, nv in [(1, 16), (2, 8), (3, 24), (4, 12), (5, 20)]:
        random.seed(seed)
        pts = gen["generatePolygon"](ctrX=250, ctrY=250, aveRadius=100, irregularity=0.55, spikeyness=0.7, numVerts=nv)
        cases["generate"].append({"seed": seed, "numVerts": nv, "points": pts})

    class Entry:
        def __init__(self, v):
            self.v = v

        def get(self):
            return str(self.v)

    class Frame:
        points = None

        def _create_circle(self, *a, **k):
            pass

    for case in cases["generate"]:
        pts = [tuple(p) for p in reversed(case["points"])]
        rng = random.Random(case["seed"] + 100)
        for base, dens in [(20.0, [1.0] * len(pts)), (12.5, [rng.choice([0.5, 1.0, 1.5, 2.0]) for _ in pts])]:
            self = types.SimpleNamespace(points=pts, base_entry=Entry(base), density_entries=[Entry(d) for d in dens], base_frame=Frame())
            self.distance = lambda a, b, _f=tk["distance"]: _f(self, a, b)
            try:
                tk["calculate_density"](self, None)
            except ZeroDivisionError:
                continue
            cases["densify"].append({"points": pts, "densities": dens, "base_length": base, "result": self.base_frame.points})
        self = types.SimpleNamespace(points=pts)
        cases["clockwise"].append({"points": pts, "clockwise": bool(tk["check_clockwise"](self))})
        self = types.SimpleNamespace(points=list(reversed(pts)))
        cases["clockwise"].append({"points": list(reversed(pts)), "clockwise": bool(tk["check_clockwise"](self))})
    json.dump(cases, open(OUT, "w"))
    print({k: len(v) for k, v in cases.items()})


if __name__ == "__main__":
    sys.exit(main())
