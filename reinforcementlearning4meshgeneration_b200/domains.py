"""Domain tooling either side of the hot path: the on-disk polygon format, the
coarse-polygon densifier, the random star-polygon generator and the curriculum over domains.

Host-side Python like the reference's own tooling; nothing here is on the step path.  Citations are
relative to the reference tree:
  ui/tk-ui.py:64-101      load / save of a domain (first JSON line ``[[x_px, y_px], ...]``, clockwise)
  ui/tk-ui.py:169-175     check_clockwise
  ui/tk-ui.py:187-193     clockwise_angle
  ui/tk-ui.py:252-276     calculate_density (the densifier)
  ui/GenerateRandomPolygon.py:5-60   generatePolygon / clip
  general/polygon.py:110-117, v2/src/mesh_rl/geometry.py:34-52   read_polygon (pixel / 100)
  v2/src/mesh_rl/training/curriculum.py:18-39, 42-112           CurriculumStage / default_curriculum / train_curriculum
"""
from __future__ import annotations

import json
import math
import random as _random
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

Point = Tuple[float, float]


# ---------------------------------------------------------------------------------------------
# on-disk format
# ---------------------------------------------------------------------------------------------
def check_clockwise(points: Sequence[Point]) -> bool:
    """ui/tk-ui.py:169-175 (screen coordinates, y down: a negative shoelace sum is "clockwise")."""
    return sum(points[i - 1][0] * p[1] - points[i - 1][1] * p[0] for i, p in enumerate(points)) < 0


def load_domain_pixels(filename) -> List[Point]:
    """First JSON line of a domain file, in pixels (ui/tk-ui.py:64-72)."""
    with Path(filename).open("r", encoding="utf-8") as fr:
        res = json.loads(fr.readline())
    return [(p[0], p[1]) for p in res]


def load_domain(filename) -> np.ndarray:
    """Domain file -> (n, 2) float64 array in env units (pixel / 100; geometry.py:34-52)."""
    return np.array([(p[0] / 100.0, p[1] / 100.0) for p in load_domain_pixels(filename)], dtype=np.float64)


def save_domain_pixels(filename, points: Sequence[Point]) -> List[Point]:
    """Write a domain file the way the reference GUI does (ui/tk-ui.py:84-101): the vertex order is
    reversed when check_clockwise() says it is not clockwise.  Returns the points as written."""
    pts = [(p[0], p[1]) for p in points]
    if not check_clockwise(pts):
        pts = list(reversed(pts))
    with Path(filename).open("w", encoding="utf-8") as f:
        f.write(json.dumps(pts))
    return pts


# ---------------------------------------------------------------------------------------------
# densifier
# ---------------------------------------------------------------------------------------------
def clockwise_angle(A: Point, B: Point) -> float:
    """ui/tk-ui.py:187-193."""
    theta = -math.atan2(-(B[1] - A[1]), B[0] - A[0])
    return theta if math.copysign(1, theta) >= 0 else 2 * math.pi + theta


def densify(points: Sequence[Point], densities: Sequence[float], base_length: float) -> List[Point]:
    """ui/tk-ui.py:252-276 calculate_density: walk the coarse polygon; the edge prev -> cur gets x interior
    points whose spacing grows linearly from A = density[prev] * base to B = density[cur] * base, then cur
    itself; if the total would be odd the middle point of the last edge is dropped."""
    if len(points) != len(densities):
        raise ValueError("one density per vertex")
    res: List[Point] = []
    n = len(points)
    for i in range(n):
        k, v = points[i], densities[i]
        pk, pv = points[i - 1], densities[i - 1]
        B = v * base_length
        A = pv * base_length
        L = math.sqrt((pk[0] - k[0]) ** 2 + (pk[1] - k[1]) ** 2)
        x = round((2 * L - A - B) / (A + B))
        interpolations: List[float] = []
        if x > 0:                       # the reference divides by x (ZeroDivisionError when x == 0)
            e = (B - A) / x
            interpolations = [A * (j + 1) + e * (j ** 2 + j) / 2 for j in range(x)]
        interpolations.append(L)
        if i == n - 1 and (len(res) + len(interpolations)) % 2 == 1:
            interpolations.pop(int(len(interpolations) / 2))
        angle = clockwise_angle(pk, k)
        res.extend((pk[0] + t * math.cos(angle), pk[1] + t * math.sin(angle)) for t in interpolations)
    return res


def densify_uniform(points: Sequence[Point], target_vertices: int) -> List[Point]:
    """The BASELINE config-3 use of the densifier: one spacing A = perimeter / target for every vertex."""
    per = sum(math.dist(points[i - 1], points[i]) for i in range(len(points)))
    return densify(points, [1.0] * len(points), per / float(target_vertices))


# ---------------------------------------------------------------------------------------------
# random star polygons
# ---------------------------------------------------------------------------------------------
def _clip(x, lo, hi):
    """ui/GenerateRandomPolygon.py:52-60."""
    if lo > hi:
        return x
    return lo if x < lo else hi if x > hi else x


def generate_polygon(ctrX=250, ctrY=250, aveRadius=100, irregularity=0.55, spikeyness=0.7, numVerts=16,
                     rng: Optional[_random.Random] = None) -> List[Tuple[int, int]]:
    """ui/GenerateRandomPolygon.py:5-49 with the same draw order from a ``random.Random`` (defaults :63);
    integer pixel vertices, counter-clockwise in a y-up frame."""
    rng = rng or _random
    irregularity = _clip(irregularity, 0, 1) * 2 * math.pi / numVerts
    spikeyness = _clip(spikeyness, 0, 1) * aveRadius
    lower = (2 * math.pi / numVerts) - irregularity
    upper = (2 * math.pi / numVerts) + irregularity
    steps, total = [], 0
    for _ in range(numVerts):
        tmp = rng.uniform(lower, upper)
        steps.append(tmp)
        total = total + tmp
    k = total / (2 * math.pi)
    steps = [s / k for s in steps]
    points = []
    angle = rng.uniform(0, 2 * math.pi)
    for i in range(numVerts):
        r_i = _clip(rng.gauss(aveRadius, spikeyness), 0, 2 * aveRadius)
        points.append((int(ctrX + r_i * math.cos(angle)), int(ctrY + r_i * math.sin(angle))))
        angle = angle + steps[i]
    return points


def random_domain(seed: int, min_verts: int = 64, max_verts: int = 512, **gen) -> np.ndarray:
    """Host twin of the in-kernel workload generator (BASELINE config 3): a star polygon with 8..24
    coarse vertices, duplicate consecutive vertices nudged apart (zero-length edges crash the reference,
    SURVEY App. D), reversed to clockwise, densified to an even vertex count in [min_verts, max_verts],
    pixel / 100.  (Same distribution as the device generator, not the same stream: Philox vs Mersenne.)"""
    rng = _random.Random(seed)
    K = rng.randint(8, 24)
    pts = generate_polygon(numVerts=K, rng=rng, **gen)
    for i in range(K):
        if pts[i] == pts[i - 1]:
            pts[i] = (pts[i][0] + 1, pts[i][1])
    pts = list(reversed(pts))
    target = rng.randint(min_verts, max_verts)
    dense = densify_uniform(pts, target)
    while len(dense) > max_verts or len(dense) < min_verts:
        target = max(min_verts, min(max_verts, int(target * (max_verts if len(dense) > max_verts else min_verts + 2) / len(dense))))
        dense = densify_uniform(pts, target)
        if len(dense) > max_verts:
            target -= 2
    return np.array([(x / 100.0, y / 100.0) for x, y in dense], dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# curriculum (training/curriculum.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class CurriculumStage:
    """Single curriculum stage: domain + timesteps (curriculum.py:18-24)."""
    index: int
    domain: str
    timesteps: int


def default_curriculum(algo: str = "sac") -> List[CurriculumStage]:
    """curriculum.py:27-39: the one active legacy stage, 1.5 M steps on ``random1_1``."""
    return [CurriculumStage(index=0, domain="random1_1", timesteps=1_500_000)]


def run_curriculum(stages: Sequence[CurriculumStage], domain_dir, make_env: Callable, train_stage: Callable,
                   model=None):
    """curriculum.py:42-112 without the SB3 specifics: for every stage build the env of its domain
    (``make_env(xy) -> env``; e.g. ``lambda xy: SB3VecEnv([xy], num_envs=4096)``), hand it with the
    previous stage's model to ``train_stage(env, stage, model) -> model`` and carry the model on."""
    if not stages:
        raise RuntimeError("Curriculum contained no stages.")
    for stage in stages:
        xy = load_domain(Path(domain_dir) / f"{stage.domain}.json")
        env = make_env(xy)
        try:
            model = train_stage(env, stage, model)
        finally:
            close = getattr(env, "close", None)
            if close:
                close()
    return model
