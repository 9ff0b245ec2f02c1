"""CPU: numerical checks of the exact-safe filters of csrc/mg_math.cuh / mg_kernels.cu.

The kernels decide most angle / distance *classes* from cross and dot products (or squared distances) and only
evaluate the reference's exact expression inside a narrow undecided band.  Each filter is a claim of the form
"outside the band the exact expression has this class"; here the claims are restated with NumPy / CPython (the same
arithmetic the reference runs) and hammered with random and adversarial inputs placed right at the band edges.
This tests the *proofs*, not the device code (which the GPU parity tests and tests/soak.py cover)."""
import math

import numpy as np

PI = math.pi
RNG = np.random.default_rng(12345)


def quantised_angle(cr, dt):
    """C:99-108 Vertex.to_find_clockwise_angle on precomputed cross / dot (vectorised)."""
    th = -np.arctan2(cr, dt)
    th = np.where(np.signbit(th), 2 * PI + th, th)
    return np.round(th, 4)


def adversarial_pairs(ratio, n=400_000):
    """(cross, dot) pairs whose |cross| / |dot| sits within a few ulps .. 1e-3 relative of `ratio`, all sign
    combinations, magnitudes over 12 decades."""
    dt = RNG.uniform(0.1, 1.0, n) * 10.0 ** RNG.integers(-6, 6, n) * RNG.choice([-1.0, 1.0], n)
    eps = RNG.choice([0.0, 1e-16, 1e-15, 1e-13, 1e-10, 1e-7, 1e-5, 1e-3], n) * RNG.choice([-1.0, 1.0], n)
    cr = np.abs(dt) * ratio * (1.0 + eps) * RNG.choice([-1.0, 1.0], n)
    return cr, dt


def random_pairs(n=600_000):
    return RNG.normal(size=n) * 10.0 ** RNG.integers(-5, 3, n), RNG.normal(size=n) * 10.0 ** RNG.integers(-5, 3, n)


def test_zero_class_filter():
    """angle_zero_class / angle_is_zero: |cross| > 1e-4 |dot| => the quantised angle is none of 0, 3.1416, 6.2832."""
    for cr, dt in (adversarial_pairs(1e-4), random_pairs()):
        decided = np.abs(cr) > 1e-4 * np.abs(dt)
        a = quantised_angle(cr[decided], dt[decided])
        assert not np.any((a == 0.0) | (a == 3.1416) | (a == 6.2832))


def test_corner_angle_filter():
    """corner_angle_invalid: cross >= 0 => invalid; cross < 0 and |cross| > 0.0318 |dot| => inside [0.01 pi, 0.99 pi];
    cross < 0 and |cross| < 0.0310 |dot| => outside."""
    for ratio in (0.0318, 0.0310):
        for cr, dt in (adversarial_pairs(ratio), random_pairs()):
            a = quantised_angle(cr, dt)
            invalid = (a > 0.99 * PI) | (a < 0.01 * PI)
            nonneg = ~(cr < 0)
            assert np.all(invalid[nonneg])
            A, B = np.abs(cr), np.abs(dt)
            sure_valid = (cr < 0) & (A > 0.0318 * B)
            sure_invalid = (cr < 0) & (A < 0.0310 * B)
            assert not np.any(invalid[sure_valid])
            assert np.all(invalid[sure_invalid])


def test_not_candidate_filter():
    """surely_not_candidate (M:249): cross >= 0, or dot < 0 and |cross| < 0.0875 |dot| => first angle >= 0.972 pi or == 0."""
    for cr, dt in (adversarial_pairs(0.0875), random_pairs()):
        a = quantised_angle(cr, dt)
        not_cand = (a >= PI * 0.972) | (a == 0)
        sure = ~(cr < 0) | ((dt < 0) & (np.abs(cr) < 0.0875 * np.abs(dt)))
        assert np.all(not_cand[sure])


def test_squared_distance_band():
    """dist_less: s < lo => sqrt(s) < r and s > hi => sqrt(s) > r, with lo / hi the float32 values rounded outwards
    from r^2 (1 -+ 1e-6)."""
    n = 500_000
    r = RNG.uniform(0.5, 2.0, n) * 10.0 ** RNG.integers(-4, 2, n)
    r2 = r * r
    lo = np.nextafter((r2 * (1.0 - 1e-6)).astype(np.float32), np.float32(0)).astype(np.float64)        # at or below the rd rounding
    hi = np.nextafter((r2 * (1.0 + 1e-6)).astype(np.float32), np.float32(np.inf)).astype(np.float64)   # at or above the ru rounding
    lo_rd = np.where((r2 * (1.0 - 1e-6)).astype(np.float32).astype(np.float64) <= r2 * (1.0 - 1e-6),
                     (r2 * (1.0 - 1e-6)).astype(np.float32).astype(np.float64), lo)
    hi_ru = np.where((r2 * (1.0 + 1e-6)).astype(np.float32).astype(np.float64) >= r2 * (1.0 + 1e-6),
                     (r2 * (1.0 + 1e-6)).astype(np.float32).astype(np.float64), hi)
    for edge, below in ((lo_rd, True), (hi_ru, False)):
        s = np.nextafter(edge, 0.0 if below else np.inf)          # the closest decided value to the band
        d = np.sqrt(s)
        assert np.all(d < r) if below else np.all(d > r)
    assert np.all(lo_rd < r2) and np.all(hi_ru > r2)


def test_tie_band_of_the_new_vertex():
    """near_round4_tie: a coordinate whose distance to the nearest k + 0.5 (in units of 1e-4) exceeds the band rounds
    to the same 4 decimals under any perturbation 1000x larger than the frame estimate's error (<= 1e-14 relative)."""
    n = 400_000
    x = RNG.uniform(-20, 20, n)
    scale = 3 * RNG.uniform(0.05, 3.0, n) + np.abs(x) + 1
    tol = 1e-8 * scale
    p = x * 1e4
    outside = 0.5 - np.abs(p - np.rint(p)) >= tol
    pert = (1e-11 * scale) * RNG.choice([-1.0, 1.0], n)       # 1000x the estimate's error bound
    same = np.rint((x + pert) * 1e4) == np.rint(p)
    assert np.all(same[outside])


def test_point_in_polygon_prune():
    """point_inside early-out: an edge with both endpoints strictly on one side of the ray's line (by 1e-9) and one of
    them more than 1e-4 rad off the ray axis as seen from P cannot satisfy ray.straddle(edge) (C:499-524): the
    collinearity pre-test needs both quantised angles in {0, pi, 2 pi}, and otherwise the two cross products have
    the same sign."""
    n = 400_000
    P = RNG.uniform(-5, 5, (n, 2))
    side = RNG.choice([-1.0, 1.0], n)
    mag = 10.0 ** RNG.uniform(-8.9, 1, (n, 2))
    a = np.stack([RNG.uniform(-20, 20, n), P[:, 1] + side * mag[:, 0]], axis=1)
    b = np.stack([RNG.uniform(-20, 20, n), P[:, 1] + side * mag[:, 1]], axis=1)
    dya, dyb = a[:, 1] - P[:, 1], b[:, 1] - P[:, 1]
    same_side = ((dya > 1e-9) & (dyb > 1e-9)) | ((dya < -1e-9) & (dyb < -1e-9))
    off_axis = (np.abs(dya) > 1e-4 * np.abs(a[:, 0] - P[:, 0])) | (np.abs(dyb) > 1e-4 * np.abs(b[:, 0] - P[:, 0]))
    pruned = same_side & off_axis & (P[:, 0] < 9000.0)
    # ray.straddle(edge): s1 = P, s2 = (10000, P.y); o1 = a, o2 = b
    vmx, vmy = 10000.0 - P[:, 0], np.zeros(n)
    v1x, v1y = a[:, 0] - P[:, 0], a[:, 1] - P[:, 1]
    v2x, v2y = b[:, 0] - P[:, 0], b[:, 1] - P[:, 1]
    ang1 = quantised_angle(v1x * vmy - v1y * vmx, v1x * vmx + v1y * vmy)
    ang2 = quantised_angle(v2x * vmy - v2y * vmx, v2x * vmx + v2y * vmy)
    zero = lambda t: (t == 0.0) | (t == 3.1416) | (t == 6.2832)          # round(sin(angle), 4) == 0
    collinear_branch = zero(ang1) & zero(ang2)
    straddles = (v1x * vmy - vmx * v1y) * (v2x * vmy - vmx * v2y) <= 0
    assert not np.any(collinear_branch[pruned])
    assert not np.any(straddles[pruned])
    assert pruned.mean() > 0.5


def test_sin_rounds_to_zero_classes():
    """sin_rounds_to_zero (C:506-508): over all 62 833 quantised angles, round(sin(a), 4) == 0 exactly for
    a in {0, 3.1416, 6.2832}."""
    hits = [k for k in range(62833) if round(math.sin(k / 1e4), 4) == 0]
    assert hits == [0, 31416, 62832]
