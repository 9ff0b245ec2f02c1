// Geometry primitives of the BoudaryEnv hot path, restated for one-warp-per-env execution on
// sm_100a.  Everything here is FP64 (FP32 only where the reference itself is float32) and is
// compiled with -fmad=false: CPython / NumPy never contract a*b+c, and the 4-decimal
// quantisation of every angle (C:108) makes discrete decisions depend on exact IEEE results.
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

// out-of-line wrappers: one copy of each libdevice routine in the kernel image (the step kernel is
// instruction-fetch bound when these are inlined at every call site)
__device__ __noinline__ void mg_sincos(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __noinline__ double mg_sin(double x) { return sin(x); }
__device__ __noinline__ double mg_pow(double x, double y) { return pow(x, y); }
__device__ __noinline__ double mg_atan2(double y, double x) { return atan2(y, x); }

// ---- distance / angle -------------------------------------------------------------------
// C:25-26 Point2D.distance_to.  (CPython evaluates dx ** 2 through libm pow, which is not always
// the correctly rounded dx*dx; the difference is <= 1 ulp of the distance and never reaches a
// discrete decision except by exact coincidence -- see DESIGN.md "numerics".)
__device__ __forceinline__ double pdist(P2 a, P2 b) {
    double dx = a.x - b.x, dy = a.y - b.y;
    return sqrt(dx * dx + dy * dy);
}

// C:99-108 Vertex.to_find_clockwise_angle(self = c, point1, point2): clockwise angle p1 -> p2
// about c, quantised to 1e-4 rad, in [0, 6.2832]; -0.0 maps to 6.2832.
__device__ __noinline__ double cw_angle(P2 c, P2 p1, P2 p2) {
    double v1x = p1.x - c.x, v1y = p1.y - c.y;
    double v2x = p2.x - c.x, v2y = p2.y - c.y;
    double cr = v1x * v2y - v1y * v2x;
    double dt = v1x * v2x + v1y * v2y;
    double th = -mg_atan2(cr, dt);
    if (signbit(th)) th = 2 * PI + th;
    return py_round4(th);
}

// round(math.sin(angle), 4) == 0 for a quantised angle  <=>  angle in {0, 3.1416, 6.2832}
// (sin(3.1415) = 9.3e-5 and sin(6.2831) = -8.5e-5 round to +-0.0001).  Used by C:506-508.
__device__ __forceinline__ bool sin_rounds_to_zero(double a) { return a == 0.0 || a == 3.1416 || a == 6.2832; }

// C:490-491
__device__ __forceinline__ double cross_product(double v1x, double v1y, double v2x, double v2y) {
    return v1x * v2y - v2x * v1y;
}

// C:499-524 Segment.straddle(self = (s1, s2), another = (o1, o2)) once the collinearity pre-test
// (both quantised angles at s1 have sin rounding to 0) is known.
__device__ __noinline__ bool straddle_decide(P2 s1, P2 s2, P2 o1, P2 o2, bool collinear) {
    if (collinear) {
        double l1 = pdist(s1, s2), l2 = pdist(o1, o2);
        if (l1 > l2) {
            P2 m = mk((s2.x + s1.x) / 2, (s2.y + s1.y) / 2);
            return fmin(pdist(m, o2), pdist(m, o1)) <= l1 / 2;
        }
        P2 m = mk((o2.x + o1.x) / 2, (o2.y + o1.y) / 2);
        return fmin(pdist(m, s2), pdist(m, s1)) <= l2 / 2;
    }
    double v1x = o1.x - s1.x, v1y = o1.y - s1.y;
    double v2x = o2.x - s1.x, v2y = o2.y - s1.y;
    double vmx = s2.x - s1.x, vmy = s2.y - s1.y;
    return cross_product(v1x, v1y, vmx, vmy) * cross_product(v2x, v2y, vmx, vmy) <= 0;
}

// Scalar (one-lane) C:526-541 Segment.is_cross(self = (a1, a2), another = (b1, b2)).
__device__ __forceinline__ bool is_cross_scalar(P2 a1, P2 a2, P2 b1, P2 b2) {
    bool z0 = sin_rounds_to_zero(cw_angle(a1, b1, a2));
    bool z1 = sin_rounds_to_zero(cw_angle(a1, b2, a2));
    if (!straddle_decide(a1, a2, b1, b2, z0 && z1)) return false;
    bool z2 = sin_rounds_to_zero(cw_angle(b1, a1, b2));
    bool z3 = sin_rounds_to_zero(cw_angle(b1, a2, b2));
    return straddle_decide(b1, b2, a1, a2, z2 && z3);
}

// Quad-lane is_cross: the four lanes of an aligned lane quad each evaluate ONE of the four
// quantised angles of is_cross(A, B) (the atan2 is the expensive part) and exchange the
// collinearity bits with a ballot.  Must be called by all 32 lanes; `active` says whether this
// quad holds a real segment pair.  Returns the predicate in every lane of the quad.
// The reference short-circuits `straddle(A,B) and straddle(B,A)`; both operands are pure, so
// evaluating all four angles is equivalent.
__device__ __forceinline__ bool is_cross_quad(P2 a1, P2 a2, P2 b1, P2 b2, bool active, int lane) {
    int sub = lane & 3;
    bool z = false;
    if (active) {
        P2 c = sub < 2 ? a1 : b1;
        P2 p1 = sub == 0 ? b1 : (sub == 1 ? b2 : (sub == 2 ? a1 : a2));
        P2 p2 = sub < 2 ? a2 : b2;
        z = sin_rounds_to_zero(cw_angle(c, p1, p2));
    }
    unsigned zb = (__ballot_sync(FULL, z) >> (lane & ~3)) & 0xFu;
    if (!active) return false;
    bool sab = straddle_decide(a1, a2, b1, b2, (zb & 3u) == 3u);
    bool sba = straddle_decide(b1, b2, a1, a2, (zb & 12u) == 12u);
    return sab && sba;
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
__device__ __forceinline__ P2 shfl_p(P2 v, int src) { return mk(__shfl_sync(FULL, v.x, src), __shfl_sync(FULL, v.y, src)); }

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
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
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
