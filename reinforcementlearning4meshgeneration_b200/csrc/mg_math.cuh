// Geometry primitives of the BoudaryEnv hot path, restated for one-warp-per-env execution on
// sm_100a.  Everything here is FP64 (FP32 only where the reference itself is float32) and is
// compiled with -fmad=false: CPython / NumPy never contract a*b+c, and the 4-decimal
// quantisation of every angle (C:108) makes discrete decisions depend on exact IEEE results.
//
// The expensive primitive is the quantised clockwise angle (one atan2).  Most call sites only
// need a *classification* of that angle (is it 0 / pi / 2pi?  is it inside [0.01pi, 0.99pi]?  is
// it >= 0.972pi?).  Those are answered from the cross / dot products with exact-safe filters
// (margins far wider than the 5e-5 rad quantisation) and fall back to the exact atan2 path only
// inside the narrow undecided band, so the result is always the reference's.
//
// Citations: C = v2/src/mesh_rl/components_core.py (legacy twin general/components.py, -8 lines)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

constexpr double PI = 3.141592653589793;
constexpr unsigned FULL = 0xffffffffu;

struct P2 {
    double x, y;
};

__device__ __forceinline__ P2 mk(double x, double y) {
    P2 p;
    p.x = x;
    p.y = y;
    return p;
}

// ---- rounding ---------------------------------------------------------------------------
// NumPy scalar round(x, 4) (np.float64): rint(x * 1e4) / 1e4.
__device__ __forceinline__ double np_round4(double x) { return rint(x * 1e4) / 1e4; }

// rint(x * 1e4) as CPython's round(x, 4) would pick it: correctly rounded on the exact product,
// exact ties to even.  x * 1e4 = p + e exactly (one explicit FMA recovers e); the two roundings
// can only differ when p lands exactly on k + 0.5.
__device__ __forceinline__ double py_rint4(double x) {
    double p = x * 1e4;
    double r = rint(p);
    if (fabs(p - r) == 0.5) {
        double e = __fma_rn(x, 1e4, -p);
        if (e > 0) r = floor(p) + 1.0;
        else if (e < 0) r = floor(p);
    }
    return r;
}
__device__ __forceinline__ double py_round4(double x) { return py_rint4(x) / 1e4; }

// NumPy scalar round(x, 4) for np.float32: everything stays float32 (C:1290).
__device__ __forceinline__ float np_round4f(float x) { return __fdiv_rn(rintf(__fmul_rn(x, 1e4f)), 1e4f); }

// out-of-line wrappers: one copy of each libdevice routine per kernel image (the fused step
// kernel was instruction-fetch bound when these were inlined at every call site)
// {sin, cos} returned by value: pointer outputs would live in local memory, and with the shared-memory carve-out
// at its maximum L1 is tiny, so every local access is an L2 round trip
__device__ __noinline__ double2 mg_sincos(double x) {
    double s, c;
    sincos(x, &s, &c);
    return make_double2(s, c);
}
__device__ __noinline__ double mg_sin(double x) { return sin(x); }
__device__ __noinline__ double mg_pow(double x, double y) { return pow(x, y); }

// ---- distance / angle -------------------------------------------------------------------
// C:25-26 Point2D.distance_to.  (CPython evaluates dx ** 2 through libm pow, which is not always
// the correctly rounded dx*dx; the difference is <= 1 ulp of the distance and never reaches a
// discrete decision except by exact coincidence -- see DESIGN.md "numerics".)
__device__ __forceinline__ double pdist(P2 a, P2 b) {
    double dx = a.x - b.x, dy = a.y - b.y;
    return sqrt(dx * dx + dy * dy);
}

// pdist(a, b) < r decided from the squared distance: sqrt is monotonic and correctly rounded, so outside a narrow
// relative band around r^2 the comparison is already decided; the exact sqrt only runs inside the band.  The band
// edges are kept as floats rounded outwards (1e-6 relative: two registers instead of four across the O(n) scans).
struct DistBound {
    double r;
    float lo, hi;
};
__device__ __forceinline__ DistBound dist_bound(double r) {
    DistBound b;
    b.r = r;
    b.lo = __double2float_rd((r * r) * (1.0 - 1e-6));
    b.hi = __double2float_ru((r * r) * (1.0 + 1e-6));
    return b;
}
__device__ __forceinline__ bool dist_less(P2 a, P2 b, const DistBound &B) {
    double dx = a.x - b.x, dy = a.y - b.y;
    double s = dx * dx + dy * dy;
    if (s < (double)B.lo) return true;
    if (s > (double)B.hi) return false;
    return sqrt(s) < B.r;
}

// cross / dot of (p1 - c) and (p2 - c) exactly as C:100-106 forms them
__device__ __forceinline__ void cross_dot(P2 c, P2 p1, P2 p2, double &cr, double &dt) {
    double v1x = p1.x - c.x, v1y = p1.y - c.y;
    double v2x = p2.x - c.x, v2y = p2.y - c.y;
    cr = v1x * v2y - v1y * v2x;
    dt = v1x * v2x + v1y * v2y;
}

// C:106-108: theta = -atan2(cross, dot); round(theta, 4) if copysign(1, theta) >= 0 else
// round(2 pi + theta, 4).  Quantised to 1e-4 rad, in [0, 6.2832]; -0.0 maps to 6.2832.
__device__ __noinline__ double cw_angle_crdt(double cr, double dt) {
    double th = -atan2(cr, dt);
    if (signbit(th)) th = 2 * PI + th;
    return py_round4(th);
}

// C:99-108 Vertex.to_find_clockwise_angle(self = c, point1, point2)
__device__ __forceinline__ double cw_angle(P2 c, P2 p1, P2 p2) {
    double cr, dt;
    cross_dot(c, p1, p2, cr, dt);
    return cw_angle_crdt(cr, dt);
}

// round(math.sin(angle), 4) == 0 for a quantised angle  <=>  angle in {0, 3.1416, 6.2832}
// (sin(3.1415) = 9.3e-5 and sin(6.2831) = -8.5e-5 round to +-0.0001).  Used by C:506-508.
__device__ __forceinline__ bool sin_rounds_to_zero(double a) { return a == 0.0 || a == 3.1416 || a == 6.2832; }

// Is the quantised angle one of {0, 3.1416, 6.2832}?  Those classes cover |theta - k pi| < 5.8e-5,
// i.e. |cross| < 5.8e-5 |dot|; outside |cross| <= 1e-4 |dot| the answer is No without an atan2.
__device__ __forceinline__ bool angle_zero_class(double cr, double dt) {
    if (fabs(cr) > 1e-4 * fabs(dt)) return false;
    return sin_rounds_to_zero(cw_angle_crdt(cr, dt));
}
// Is the quantised angle exactly 0 (C:1255)?
__device__ __forceinline__ bool angle_is_zero(double cr, double dt) {
    if (fabs(cr) > 1e-4 * fabs(dt)) return false;
    return cw_angle_crdt(cr, dt) == 0.0;
}

// C:752-757: quad corner angle outside [0.01 pi, 0.99 pi]?  theta in (0, pi) needs cross < 0;
// 0.01 pi = 0.031416: |cross| < 0.0310 |dot| is surely outside, |cross| > 0.0318 |dot| surely inside.
__device__ __forceinline__ bool corner_angle_invalid(double cr, double dt) {
    if (!(cr < 0)) return true;                       // theta = 0 or theta in [pi, 2 pi]
    double a = fabs(cr), b = fabs(dt);
    if (a > 0.0318 * b) return false;
    if (a < 0.0310 * b) return true;
    double ang = cw_angle_crdt(cr, dt);
    return ang > 0.99 * PI || ang < 0.01 * PI;
}

// M:249: not a reference candidate when the first angle is >= 0.972 pi or == 0.
// 0.972 pi = pi - 0.08796: cross >= 0 gives theta = 0 or >= pi; for cross < 0 and dot < 0,
// |cross| < 0.0875 |dot| puts theta within 0.0873 rad of pi.  Otherwise undecided (exact path).
__device__ __forceinline__ bool surely_not_candidate(double cr, double dt) {
    if (!(cr < 0)) return true;
    return dt < 0 && fabs(cr) < 0.0875 * fabs(dt);
}

// C:490-491
__device__ __forceinline__ double cross_product(double v1x, double v1y, double v2x, double v2y) {
    return v1x * v2y - v2x * v1y;
}

// C:509-519: the collinear branch of Segment.straddle (rare)
__device__ __noinline__ bool straddle_collinear(P2 s1, P2 s2, P2 o1, P2 o2) {
    double l1 = pdist(s1, s2), l2 = pdist(o1, o2);
    if (l1 > l2) {
        P2 m = mk((s2.x + s1.x) / 2, (s2.y + s1.y) / 2);
        return fmin(pdist(m, o2), pdist(m, o1)) <= l1 / 2;
    }
    P2 m = mk((o2.x + o1.x) / 2, (o2.y + o1.y) / 2);
    return fmin(pdist(m, s2), pdist(m, s1)) <= l2 / 2;
}

// C:499-524 Segment.straddle(self = (s1, s2), another = (o1, o2))
__device__ __forceinline__ bool straddle(P2 s1, P2 s2, P2 o1, P2 o2) {
    double v1x = o1.x - s1.x, v1y = o1.y - s1.y;
    double v2x = o2.x - s1.x, v2y = o2.y - s1.y;
    double vmx = s2.x - s1.x, vmy = s2.y - s1.y;
    // angles at s1: (o1, s2) and (o2, s2); cross/dot in the operand order of C:100-106
    double cr1 = v1x * vmy - v1y * vmx, dt1 = v1x * vmx + v1y * vmy;
    double cr2 = v2x * vmy - v2y * vmx, dt2 = v2x * vmx + v2y * vmy;
    if (angle_zero_class(cr1, dt1) && angle_zero_class(cr2, dt2)) return straddle_collinear(s1, s2, o1, o2);
    return cross_product(v1x, v1y, vmx, vmy) * cross_product(v2x, v2y, vmx, vmy) <= 0;
}

// C:526-541 Segment.is_cross(self = (a1, a2), another = (b1, b2)), one lane.
__device__ __forceinline__ bool is_cross(P2 a1, P2 a2, P2 b1, P2 b2) {
    return straddle(a1, a2, b1, b2) && straddle(b1, b2, a1, a2);
}

// C:678-692 Segment(p1, p2).distance(point a)
__device__ __forceinline__ double seg_point_distance(P2 p1, P2 p2, P2 a) {
    double A = p2.x - p1.x, B = p2.y - p1.y;
    double s = (A * a.x + B * a.y - B * p1.y - A * p1.x) / (A * A + B * B);
    if (0 <= s && s <= 1) return pdist(a, mk(p1.x + s * A, p1.y + s * B));
    if (s < 0) return pdist(a, p1);
    return pdist(a, p2);
}

// Python builtin sum() over floats under CPython >= 3.12 (Neumaier compensated summation,
// Python/bltinmodule.c).  K is tiny (4..6) on this path.
template <int K>
__device__ __forceinline__ double py_sum(const double (&x)[K]) {
    double f = x[0], c = 0;
#pragma unroll
    for (int i = 1; i < K; i++) {
        double t = f + x[i];
        if (fabs(f) >= fabs(x[i])) c += (f - t) + x[i];
        else c += (x[i] - t) + f;
        f = t;
    }
    if (c != 0 && isfinite(c)) f += c;
    return f;
}

// ---- warp helpers ----------------------------------------------------------------------
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(FULL, v, o);
        v = w < v ? w : v;
    }
    return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// order-preserving map double -> uint64 (for lexicographic (value, order) arg-min reductions)
__device__ __forceinline__ unsigned long long d2key(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

}  // namespace mg
