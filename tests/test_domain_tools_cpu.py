"""Domain tooling against golden vectors recorded from the reference's own functions
(oracle/record_domain_tools_golden.py): random star polygons, the densifier, the clockwise test, the
domain file format, the curriculum stages."""
import json
import os
import random

import numpy as np

from helpers import GOLDEN
from reinforcementlearning4meshgeneration_b200 import domains as D

G = json.load(open(os.path.join(GOLDEN, "domain_tools.json")))


def test_generate_polygon_matches_reference_stream():
    for c in G["generate"]:
        got = D.generate_polygon(numVerts=c["numVerts"], rng=random.Random(c["seed"]))
        assert [list(p) for p in got] == c["points"], f"seed {c['seed']}"


def test_densifier_matches_reference():
    assert len(G["densify"]) >= 3
    for c in G["densify"]:
        got = D.densify([tuple(p) for p in c["points"]], c["densities"], c["base_length"])
        assert len(got) == len(c["result"]) and len(got) % 2 == 0
        assert np.array_equal(np.array(got), np.array(c["result"])), "densified coordinates differ"


def test_check_clockwise_and_file_round_trip(tmp_path):
    for c in G["clockwise"]:
        pts = [tuple(p) for p in c["points"]]
        assert D.check_clockwise(pts) == c["clockwise"]
        f = tmp_path / "dom.json"
        written = D.save_domain_pixels(f, pts)
        assert D.check_clockwise(written)
        assert D.load_domain_pixels(f) == [tuple(p) for p in written]
        assert np.array_equal(D.load_domain(f), np.array(written, dtype=np.float64) / 100.0)


def test_random_domain_is_a_valid_env_polygon():
    from oracle.c_oracle import OracleEnv
    for seed in range(12):
        xy = D.random_domain(seed)
        n = len(xy)
        assert 64 <= n <= 512 and n % 2 == 0
        d = np.linalg.norm(xy - np.roll(xy, 1, axis=0), axis=1)
        assert d.min() > 1e-4
        shoelace = np.sum(np.roll(xy[:, 0], 1) * xy[:, 1] - np.roll(xy[:, 1], 1) * xy[:, 0])
        assert shoelace < 0, "clockwise in the env's frame"
        o = OracleEnv(xy)
        o.reset()
        r = o.run_random(seed, 300)
        assert not o.crashed and r["success"] > 0


def test_curriculum_runner(tmp_path):
    st = D.default_curriculum("sac")
    assert len(st) == 1 and st[0].domain == "random1_1" and st[0].timesteps == 1_500_000 and st[0].index == 0
    for name in ("a", "b"):
        D.save_domain_pixels(tmp_path / f"{name}.json", [(0, 0), (0, 100), (100, 100), (100, 0)])
    seen = []

    class Env:
        def __init__(self, xy):
            self.xy = xy
            self.closed = False

        def close(self):
            self.closed = True

    envs = []

    def make_env(xy):
        envs.append(Env(xy))
        return envs[-1]

    def train(env, stage, model):
        seen.append((stage.domain, stage.timesteps, model))
        return (model or 0) + stage.timesteps

    out = D.run_curriculum([D.CurriculumStage(0, "a", 10), D.CurriculumStage(1, "b", 5)], tmp_path, make_env, train)
    assert out == 15 and seen == [("a", 10, None), ("b", 5, 10)] and all(e.closed for e in envs)
    assert envs[0].xy.shape == (4, 2) and envs[0].xy.max() == 1.0
