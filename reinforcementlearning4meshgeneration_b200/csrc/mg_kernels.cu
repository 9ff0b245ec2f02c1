// Batched BoudaryEnv reset/step for sm_100a.
//
// A step is a handful of kernels (see the block comment above mg_step_screen_kernel): a screen over all envs, one thread
// per env, that settles every step whose outcome follows from the env's 128-byte record (memoised rule -1 / +1
// verdicts, Mesh.is_valid of the new-vertex quad from the neighbour fan), and warp-per-item kernels for the steps that
// need the whole boundary.  There the active boundary (updated_boundary.vertices) is staged once per item from HBM
// into a per-warp shared-memory ring with a single 1-D bulk async copy (cp.async.bulk + mbarrier, UBLKCP in SASS);
// every O(n) predicate then strides the ring with 32 lanes (double2 loads) and is resolved with ballots / shuffle
// reductions.  Angle classifications use the exact-safe filters of mg_math.cuh, so atan2 only runs where the quantised
// value itself is needed (candidate keys, observation, element quality).
//
// Citations: E = v2/src/mesh_rl/envs/boundary_env.py, M = v2/src/mesh_rl/mesh_core.py,
// C = v2/src/mesh_rl/components_core.py, D = v2/src/mesh_rl/data_core.py (reference tree).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "mg_math.cuh"
#include "mg_state.cuh"
#include "mg_smooth.cuh"

namespace mg {

#ifndef MG_UNROLL_OBS
#define MG_UNROLL_OBS 1
#endif
#ifndef MG_UNROLL_BQ
#define MG_UNROLL_BQ 1
#endif
#ifndef MG_UNROLL_PIP
#define MG_UNROLL_PIP 1
#endif
constexpr int UNROLL_OBS = MG_UNROLL_OBS, UNROLL_BQ = MG_UNROLL_BQ, UNROLL_PIP = MG_UNROLL_PIP;   // tuning knobs (profiles/README.md)
constexpr int QCAP = 128 + 256;  // per-warp scratch: 4x32 ints + 32x4 doubles (coarse polygon of the generator)

// Item timeline of the warp-per-item kernels (profiling variant only, -DMG_TRACE; tests/trace_items.py): one record per
// item = {kernel, kind, n, SM, start ns, end ns} appended to a device buffer armed with mg_set_option("trace", capacity).
#ifdef MG_TRACE
struct TraceRec { int kernel, kind, n, sm; unsigned long long t0, t1; };
__device__ TraceRec *g_trace = nullptr;
__device__ unsigned g_trace_cap = 0, g_trace_n = 0;
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_item(int kernel, int kind, int n, unsigned long long t0, int lane) {
    if (lane == 0 && g_trace != nullptr) {
        const unsigned i = atomicAdd(&g_trace_n, 1u);
        if (i < g_trace_cap) {
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            TraceRec r; r.kernel = kernel; r.kind = kind; r.n = n; r.sm = (int)sm; r.t0 = t0; r.t1 = trace_now();
            g_trace[i] = r;
        }
    }
}
#define MG_TRACE_T0 const unsigned long long trace_t0_ = trace_now();
#define MG_TRACE_BLOCK trace_item(0, -1, (int)blockIdx.x, trace_now(), lane);
#define MG_TRACE_ITEM(kernel, kind, n) trace_item(kernel, kind, n, trace_t0_, lane);
#else
#define MG_TRACE_T0
#define MG_TRACE_BLOCK
#define MG_TRACE_ITEM(kernel, kind, n)
#endif

// ---------------------------------------------------------------------------------------------
// per-warp view of the environment
// ---------------------------------------------------------------------------------------------
// i in [-2n, 3n) -> [0, n)
__device__ __forceinline__ int wrapn(int i, int n) {
    if (i < 0) i += n;
    if (i < 0) i += n;
    if (i >= n) i -= n;
    if (i >= n) i -= n;
    return i;
}

struct Quad {
    double x[4], y[4];        // candidate quad, passed by value (registers) to the out-of-line predicates
    __device__ __forceinline__ P2 at(int k) const { return mk(x[k], y[k]); }
};

struct Warp {
    double2 *ring;   // shared memory, this warp's vertex ring
    int *queue;      // shared memory, this warp's integer scratch
    int n;           // live boundary size
    int lane;
    __device__ __forceinline__ int wrap(int i) const {
        if (i < 0) i += n;
        else if (i >= n) i -= n;
        return i;
    }
    // Python list indexing B[i] for i in [-n, 2n)
    __device__ __forceinline__ P2 at(int i) const {
        double2 v = ring[wrap(i)];
        return mk(v.x, v.y);
    }
};

// {sin, cos} of a quantised angle (or of its half with tab = sc_half) from the host-libm table; any other
// argument falls back to the device routines
__device__ __forceinline__ void sincos_quantised(const double2 *tab, double angle, bool half, double &s, double &c) {
    const double kf = rint(angle * 1e4);
    if (tab != nullptr && kf >= 0.0 && kf <= (double)(ANGLE_TAB_N - 1) && kf / 1e4 == angle) {
        const double2 v = __ldg(tab + (int)kf);
        s = v.x; c = v.y;
    } else {
        const double2 v = mg_sincos(half ? angle / 2 : angle);
        s = v.x; c = v.y;
    }
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Stage `n` vertices (16 B each) from global memory into the warp's ring with one bulk async
// copy issued by lane 0 and completed on a per-warp mbarrier.
__device__ __forceinline__ void stage_issue(double2 *ring, unsigned long long *mbar, const double2 *src, int n, int lane) {
    unsigned bar = smem_u32(mbar);
    unsigned bytes = (unsigned)n * 16u;
    // order this warp's earlier generic-proxy accesses to the ring before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#ifdef MG_RING_EVICT_FIRST
        // (round 1 streamed every env's ring through L2 once per step and marked the copies evict-first; with the screen
        // kernel only ~8 % of the rings are touched per step -- 25 MB -- and the observe kernel re-reads the ones the
        // update kernel staged, so they are left to the normal L2 policy)
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                smem_u32(ring)),
            "l"(src), "r"(bytes), "r"(bar), "l"(pol)
            : "memory");
#else
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring)),
            "l"(src), "r"(bytes), "r"(bar)
            : "memory");
#endif
    }
}
// 16-byte load that asks L2 to keep the line (the 48-byte hot records, 3 MB at 65 536 envs, are re-read every step)
__device__ __forceinline__ int4 ldg_keep(const void *p) {
    int4 v;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stage_wait(unsigned long long *mbar, unsigned parity) {
    unsigned bar = smem_u32(mbar);
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
// Stage `n` vertices (16 B each) from global memory into the warp's ring with one bulk async
// copy issued by lane 0 and completed on a per-warp mbarrier.
__device__ __forceinline__ void stage_ring(double2 *ring, unsigned long long *mbar, const double2 *src, int n, int lane,
                                           unsigned parity) {
    stage_issue(ring, mbar, src, n, lane);
    stage_wait(mbar, parity);
}

__device__ __forceinline__ void init_mbar(unsigned long long *mbar, int lane) {
#ifndef MG_NO_BULK_COPY
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
#endif
}

// ---------------------------------------------------------------------------------------------
// candidate keys (M:228-257 check_boundary_point) and reference point (M:295-316)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cand_key_from_angles(double a0, double a1) {
    if (a0 >= PI * 0.972 || a0 == 0) return CUDART_INF;
    const double lam = 0.618;
    double sum = a0 * lam;
    sum += a1 * (1 - lam);
    return sum * (180.0 / PI);
}

// Full rebuild (M:259-287): key for every vertex, stamp = list index (stable sort order).
// Two passes so that the atan2 work runs on full warps: pass 1 classifies every vertex from its
// cross/dot products (most vertices of a densified polygon are surely not candidates) and compacts
// the undecided ones into a small queue; the queue is drained 32 entries at a time.
__device__ __noinline__ void rebuild_candidates(const Warp w, double *key, int32_t *stamp) {
    const int n = w.n, lane = w.lane;
    int qn = 0;
    auto drain = [&](int count) {           // exact keys for queue[qn - count .. qn), one vertex per lane
        int j = lane < count ? w.queue[qn - count + lane] : -1;
        if (j >= 0) {
            P2 c = w.at(j);
            double k = CUDART_INF;
            double a0 = cw_angle(c, w.at(j + 1), w.at(j - 1));
            if (!(a0 >= PI * 0.972 || a0 == 0)) k = cand_key_from_angles(a0, cw_angle(c, w.at(j + 2), w.at(j - 2)));
            key[j] = k;
        }
        qn -= count;
        __syncwarp();
    };
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int j = base + lane;
        bool undecided = false;
        if (j < n) {
            double cr, dt;
            cross_dot(w.at(j), w.at(j + 1), w.at(j - 1), cr, dt);
            undecided = !surely_not_candidate(cr, dt);
            if (!undecided) key[j] = CUDART_INF;
            stamp[j] = j;
        }
        unsigned m = __ballot_sync(FULL, undecided);
        if (undecided) w.queue[qn + __popc(m & ((1u << lane) - 1))] = j;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) drain(32);
    }
    if (qn > 0) drain(qn);
}

// arg-min of (key, stamp) over the n live vertices; -1 when the candidate list is empty.
__device__ __noinline__ int find_reference_index(const Warp w, const double *key, const int32_t *stamp) {
    double bk = CUDART_INF;
    int bs = 0x7fffffff, bj = -1;
#pragma unroll 4
    for (int j = w.lane; j < w.n; j += 32) {
        double k = __ldcg(key + j);
        int s = __ldcg(stamp + j);
        if (k < bk || (k == bk && k != CUDART_INF && s < bs)) {
            bk = k;
            bs = s;
            bj = j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ok = __shfl_xor_sync(FULL, bk, o);
        int os = __shfl_xor_sync(FULL, bs, o);
        int oj = __shfl_xor_sync(FULL, bj, o);
        if (ok < bk || (ok == bk && ok != CUDART_INF && os < bs)) {
            bk = ok;
            bs = os;
            bj = oj;
        }
    }
    return bk == CUDART_INF ? -1 : bj;
}

// ---------------------------------------------------------------------------------------------
// observation (C:1059-1090 PointEnvironment, C:1192-1290 get_radius_points, E:665-738)
// returns obs[lane] for lane < 18 and the base length (by value: reference outputs of an out-of-line function live
// in local memory, an L2 round trip with the shared-memory carve-out at its maximum).
// ---------------------------------------------------------------------------------------------
struct ObsOut {
    double base;
    float obs;
};
__device__ __noinline__ ObsOut compute_obs(const Warp w, const double2 *sc, int idx, double area_ratio) {
    const int lane = w.lane, n = w.n;
    const double inv_radius = 0.25;   // x / 4 == x * 0.25 exactly
    P2 ref = w.at(idx), right_p = w.at(idx - 1), left_p = w.at(idx + 1 >= n ? idx + 1 - n : idx + 1);

    // --- batch 1: 6 fan distances for base_length, 6 fan distances to ref, 6 angles ---------
    // lanes 0..5 : |N[k+1] N[k]|  with N = [i+3, i+2, i+1, i, i-1, i-2, i-3]           (C:1087-1090)
    // lanes 8..13: angles  theta, a(i-2), a(i-3), a(i+2), a(i+3), rot                      (C:1201-1235)
    // lanes 16..21: |ref B[i-1]|, |ref B[i+1]|, |ref B[i-2]|, |ref B[i-3]|, |ref B[i+2]|, |ref B[i+3]|
    double val = 0;
    if (lane < 6) {
        int a = idx + 2 - lane, b = idx + 3 - lane;   // N[k] = at(idx+3-k); dist(N[k+1], N[k])
        val = pdist(w.at(wrapn(a, n)), w.at(wrapn(b, n)));
    } else if (lane >= 8 && lane < 14) {
        int t = lane - 8;
        P2 p1, p2 = right_p;
        if (t == 0) p1 = left_p;
        else if (t == 1) p1 = w.at(wrapn(idx - 2, n));
        else if (t == 2) p1 = w.at(wrapn(idx - 3, n));
        else if (t == 3) p1 = w.at(wrapn(idx + 2, n));
        else if (t == 4) p1 = w.at(wrapn(idx + 3, n));
        else {
            p1 = right_p;
            p2 = mk(ref.x + 1, ref.y + 0);
        }
        val = cw_angle(ref, p1, p2);
    } else if (lane >= 16 && lane < 22) {
        int t = lane - 16;
        const int o6 = t == 0 ? -1 : (t == 1 ? 1 : (t == 2 ? -2 : (t == 3 ? -3 : (t == 4 ? 2 : 3))));
        val = pdist(ref, w.at(wrapn(idx + o6, n)));
    }
    double dl[6];
#pragma unroll
    for (int k = 0; k < 6; k++) dl[k] = shfl_d(val, k);
    double base = py_round4(py_sum<6>(dl) / 6);
    const double theta = shfl_d(val, 8), rot = shfl_d(val, 13);
    const double T = base * 4;
    const double clip = theta + PI / 2;

    // r_points float32[9][2] (C:1193): every lane that owns a distance / angle turns it into its float32 entry
    // itself (one division for the six distances instead of six lane-uniform ones) and the entry is shuffled to the
    // lane that outputs it; only the three sector entries, which the scan below updates, stay lane-uniform.
    //   lanes 16..21 = d_r, d_l, d_r1, d_r2, d_l1, d_l2    lanes 9..12 = a_r1, a_r2, a_l1, a_l2    lane 8 = theta
    float tv = 0.0f;
    if (lane >= 16 && lane < 22) tv = (float)((val * inv_radius) / base);
    else if (lane == 8) tv = (float)val;
    else if (lane == 9 || lane == 10) tv = (float)(val < PI ? val : fmax(val, 1.5 * PI) - 2 * PI);
    else if (lane == 11 || lane == 12) tv = (float)fmin(val, clip);
    // output lane L = 2 i (+1) holds r[i][0] (r[i][1]): source lanes of r[0], r[1], r[2], -, -, -, r[6], r[7], r[8]
    const int src_of[18] = {16, 0, 18, 9, 19, 10, 0, 0, 0, 0, 0, 0, 21, 12, 20, 11, 17, 8};
    int src = 0;
#pragma unroll
    for (int i = 0; i < 18; i++)
        if (lane == i) src = src_of[i];
    float out = __shfl_sync(FULL, tv, src);
    if (lane == 1) out = (float)area_ratio;
    float s0[3], s1[3];          // sector entries r[3..5]
#pragma unroll
    for (int j = 0; j < 3; j++) { s0[j] = 1.0f; s1[j] = (float)fmin((2 * j + 1) * theta / 6, clip); }

    // p_s = ref + rotate((T cos(theta/2), T sin(theta/2)), rot)                      (C:154-168, C:1243)
    double s_h, c_h, s_r, c_r;
    sincos_quantised(sc ? sc + ANGLE_TAB_N : nullptr, theta, true, s_h, c_h);      // sc = Params::sc_full; the half table follows
    sincos_quantised(sc, rot, false, s_r, c_r);
    double px = T * c_h, py = T * s_h;
    double qx = c_r * px - s_r * py;
    double qy = s_r * px + c_r * py;
    P2 ps = mk(ref.x + qx, ref.y + qy);
    double ux = ps.x - ref.x, uy = ps.y - ref.y;

    // --- scan of the other n-1 vertices in the reference's order o = 1 .. n-1  (m = idx - o) --
    const double sector = theta / 3;
    unsigned long long bs0 = ~0ull, bs1 = ~0ull, bs2 = ~0ull;  // per sector: (float bits of cand, order)
    float ma0 = 0, ma1 = 0, ma2 = 0;                           // (float)min(angle, clip) of the lane's best candidate
    double best_ray = CUDART_INF;                             // f64 value of the nearest bisector hit
    int best_ray_o = 0x7fffffff;
    // Two passes (like rebuild_candidates): pass 1 is a light filter that keeps the few vertices that can matter --
    // inside the radius (sector candidates, C:1255-1263) or with an edge the bisector's line may cross (C:657-676) --
    // and compacts their order numbers into the warp's queue; pass 2 runs the reference's per-vertex body on full
    // warps of queued entries.  A vertex is dropped only when it is surely outside the radius (1e-12 relative
    // margin on the squared distance) AND its edge passes the exact-safe early-out below: the body would do
    // nothing for it.  Both per-lane accumulators are (value, order) minima, so the evaluation order is free.
    auto body = [&](const int o) {
        int j = idx - o;
        if (j < 0) j += n;
        P2 q = w.at(j);
        double d = pdist(ref, q);
        double cr, dt;
        cross_dot(ref, q, right_p, cr, dt);
        // the quantised angle itself is only needed inside the radius; outside, only "angle == 0"
        double angle = 1.0;
        if (d < T) {
            angle = cw_angle_crdt(cr, dt);
            if (angle == 0) return;                         // C:1255
        } else if (angle_is_zero(cr, dt)) return;
        double kk = angle / sector;
        int k = (kk < 3.0) ? (int)kk : 3;                     // int() truncation; NaN/inf -> no sector
        if (k < 3 && d < T) {
            float cand = (float)((d * inv_radius) / base);
            if (cand < 1.0f) {
                unsigned long long keyv = ((unsigned long long)__float_as_uint(cand) << 32) | (unsigned)o;
                const float fa = (float)fmin(angle, clip);
                if (k == 0) { if (keyv < bs0) { bs0 = keyv; ma0 = fa; } }
                else if (k == 1) { if (keyv < bs1) { bs1 = keyv; ma1 = fa; } }
                else { if (keyv < bs2) { bs2 = keyv; ma2 = fa; } }
            }
        }
        // C:657-676 ll.intersection_vertex(seg) with ll = (ref, p_s), seg = (B[m], B[m+1])
        P2 q2 = w.at(j + 1 >= n ? j + 1 - n : j + 1);
        double wx = q2.x - q.x, wy = q2.y - q.y;
        {
            // Exact-safe early-out: a hit needs the ray's line to cross the edge (0 < h < 1).  With
            // sa, sb = signed offsets of the edge's endpoints from that line, h = sa / (sa - sb); both on
            // the same side by a 1e-6 relative margin puts h at least ~5e-7 outside [0, 1], far beyond
            // the rounding error of C:663-674 for an edge that is not degenerate in x or y.
            double ax = q.x - ref.x, ay = q.y - ref.y;
            double sa = ux * ay - uy * ax;
            double sb = ux * (q2.y - ref.y) - uy * (q2.x - ref.x);
            double tol = 1e-6 * (fabs(ux) + fabs(uy)) * (fabs(ax) + fabs(ay) + fabs(wx) + fabs(wy));
            bool conditioned = (wx == 0 || fabs(wx) > 1e-6 * fabs(wy)) && (wy == 0 || fabs(wy) > 1e-6 * fabs(wx));
            if (conditioned && ((sa > tol && sb > tol) || (sa < -tol && sb < -tol))) return;
        }
        double ss, hh;
        if (wy == 0) {
            if (uy == 0) return;
            ss = (q.y - ref.y) / uy;
            hh = (ref.x - q.x + ss * ux) / wx;
        } else if (wx == 0) {
            if (ux == 0) return;
            ss = (q.x - ref.x) / ux;
            hh = (ref.y - q.y + ss * uy) / wy;
        } else {
            ss = ((ref.x - q.x) / wx - (ref.y - q.y) / wy) / (uy / wy - ux / wx);
            hh = (ref.x - q.x + ss * ux) / wx;
        }
        if (0 < ss && ss < 1 && 0 < hh && hh < 1) {
            double v = (pdist(ref, mk(ref.x + ss * ux, ref.y + ss * uy)) * inv_radius) / base;
            if (v < 1.0 && (v < best_ray || (v == best_ray && o < best_ray_o))) {
                best_ray = v;
                best_ray_o = o;
            }
        }
    };
    {
        const double T2_far = (T * T) * (1.0 + 1e-12);
        const double T2_beyond = 2.0 * (T * T) * (1.0 + 1e-6);
        const double ur = 1e-6 * (fabs(ux) + fabs(uy));
        int qn = 0, base_o = 1;
#pragma unroll 1
        while (true) {
            if (base_o < n) {
                const int o = base_o + lane;
                bool keep = false;
                const bool valid = o < n && o != 1 && o != n - 1;   // right_p / left_p (C:1249)
                double ax = 0, ay = 0, wx = 0, wy = 0, d2 = 0;
                P2 q2 = ref;
                bool conditioned = false, open = false;
                if (valid) {
                    int j = idx - o;
                    if (j < 0) j += n;
                    const P2 q = w.at(j);
                    q2 = w.at(j + 1 >= n ? j + 1 - n : j + 1);
                    ax = q.x - ref.x; ay = q.y - ref.y; wx = q2.x - q.x; wy = q2.y - q.y;
                    d2 = ax * ax + ay * ay;
                    conditioned = (wx == 0 || fabs(wx) > 1e-6 * fabs(wy)) && (wy == 0 || fabs(wy) > 1e-6 * fabs(wx));
                    // Cheap first cut: every point of the edge is within |w| of B[m], so with
                    // |ref B[m]|^2 > 2 T^2 + 2 |w|^2 >= (T + |w|)^2 the whole edge is farther than T from ref.  The vertex
                    // is then outside the radius and its edge cannot meet the bisector segment (length T) -- the body
                    // would do nothing.  (Edges that are degenerate in x or y stay in: the reference's intersection
                    // formulas are chaotic for them, C:657-676.)  Boundary vertices are spatially coherent, so most
                    // 32-vertex chunks of a large polygon are cut as a whole and skip the orientation tests below.
                    open = !(conditioned && d2 > T2_beyond + 2.0 * (1.0 + 1e-6) * (wx * wx + wy * wy));
                }
                if (__any_sync(FULL, open)) {
                    if (open) {
                        const bool far = d2 > T2_far;
                        const double sa = ux * ay - uy * ax;
                        const double sb = ux * (q2.y - ref.y) - uy * (q2.x - ref.x);
                        const double tol = ur * (fabs(ax) + fabs(ay) + fabs(wx) + fabs(wy));
                        const bool off_line = conditioned && ((sa > tol && sb > tol) || (sa < -tol && sb < -tol));
                        keep = !(far && off_line);
                    }
                    const unsigned m = __ballot_sync(FULL, keep);
                    if (keep) w.queue[qn + __popc(m & ((1u << lane) - 1))] = o;
                    qn += __popc(m);
                    __syncwarp();
                }
                base_o += 32;
            }
            const bool scanning = base_o < n;
            if (qn >= 32 || (!scanning && qn > 0)) {          // one call site: the body is big
                const int cnt = qn < 32 ? qn : 32;
                if (lane < cnt) body(w.queue[qn - cnt + lane]);
                qn -= cnt;
                __syncwarp();
            }
            if (!scanning && qn == 0) break;
        }
    }
    // sector minima: first (in order o) strictly smaller float32 value wins (C:1259-1263)
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const unsigned long long mine = k == 0 ? bs0 : (k == 1 ? bs1 : bs2);
        const float mine_ang = k == 0 ? ma0 : (k == 1 ? ma1 : ma2);
        unsigned long long m = warp_min_u64(mine);
        if (m != ~0ull) {
            unsigned src = __ffs(__ballot_sync(FULL, mine == m)) - 1;
            s0[k] = __uint_as_float((unsigned)(m >> 32));
            s1[k] = __shfl_sync(FULL, mine_ang, src);
        }
    }
    // nearest bisector hit (C:1266-1287)
    {
        unsigned long long kv = d2key(best_ray);
        unsigned long long m = warp_min_u64(kv);
        unsigned cand_mask = __ballot_sync(FULL, kv == m && best_ray < 1.0);
        if (cand_mask) {
            // among lanes with the same value pick the smallest order
            unsigned oo = (kv == m && best_ray < 1.0) ? (unsigned)best_ray_o : 0xffffffffu;
            unsigned omin = __reduce_min_sync(FULL, oo);
            double sv2 = shfl_d(best_ray, __ffs(cand_mask) - 1);
            if ((float)sv2 < s0[1]) {                         // float32 comparison with r[4][0] (NEP 50)
                int mi = idx - (int)omin;                      // _i (may be negative like in Python)
                float v2 = 0.0f;
                if (lane < 3) {
                    int jj = wrapn(lane - 1 + mi, n);
                    v2 = (float)((pdist(ref, w.at(jj)) * inv_radius) / base);
                } else if (lane >= 4 && lane < 7) {
                    int jj = wrapn(lane - 4 - 1 + mi, n);
                    v2 = (float)cw_angle(ref, w.at(jj), right_p);
                }
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    s0[j] = __shfl_sync(FULL, v2, j);
                    s1[j] = __shfl_sync(FULL, v2, 4 + j);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (lane == 6 + 2 * j) out = s0[j];
        if (lane == 7 + 2 * j) out = s1[j];
    }
    ObsOut R;
    R.base = base;
    R.obs = np_round4f(out);
    return R;
}

// ---------------------------------------------------------------------------------------------
// estimated_area_range (M:705-718): needs the mean, the 2nd smallest and 2nd largest edge.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ double2 estimate_area_range(const Warp w) {   // {area_min, area_crit}
    double s = 0, lo1 = CUDART_INF, lo2 = CUDART_INF, hi1 = -CUDART_INF, hi2 = -CUDART_INF;
#pragma unroll 1
    for (int j = w.lane; j < w.n; j += 32) {
        double l = pdist(w.at(j - 1), w.at(j));
        s += l;
        if (l < lo1) { lo2 = lo1; lo1 = l; } else if (l < lo2) lo2 = l;
        if (l > hi1) { hi2 = hi1; hi1 = l; } else if (l > hi2) hi2 = l;
    }
    s = warp_sum_d(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double a1 = __shfl_xor_sync(FULL, lo1, o), a2 = __shfl_xor_sync(FULL, lo2, o);
        double b1 = __shfl_xor_sync(FULL, hi1, o), b2 = __shfl_xor_sync(FULL, hi2, o);
        double n1 = fmin(lo1, a1), n2 = fmin(fmax(lo1, a1), fmin(lo2, a2));
        lo1 = n1; lo2 = n2;
        double m1 = fmax(hi1, b1), m2 = fmax(fmin(hi1, b1), fmax(hi2, b2));
        hi1 = m1; hi2 = m2;
    }
    double L = s / w.n;
    double max_L = fmin(hi2, 2 * L);
    double min_L = fmin(L / sqrt(2.0), lo2);
    return make_double2(min_L, (max_L + 3 * min_L) / 4);
}

// sequential shoelace (C:485-487 up to np.dot's BLAS summation order)
__device__ __noinline__ double shoelace_area(const Warp w) {
    double s1 = 0, s2 = 0;
#pragma unroll 1
    for (int j = w.lane; j < w.n; j += 32) {
        P2 a = w.at(j), p = w.at(j - 1);
        s1 += a.x * p.y;
        s2 += a.y * p.x;
    }
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    return 0.5 * fabs(s1 - s2);
}

// ---------------------------------------------------------------------------------------------
// point-in-polygon (M:74-128 calculate_crossing_segments, M:176-187, M:565-572)
// ---------------------------------------------------------------------------------------------
// rint(1e4 * dy) with the reference's rounding flavour: NumPy's when either operand is an
// inserted vertex (np.float64), CPython's otherwise.  They only differ on exact ties after the
// multiply, so the vertex-id lookup is deferred to that (practically unreachable) case.
__device__ __forceinline__ double rint4_mixed(double dy, const int32_t *vid, int ja, int jb, int n0) {
    double p = dy * 1e4;
    double r = rint(p);
    if (fabs(p - r) == 0.5) {
        bool is_np = vid[ja] >= n0 || vid[jb] >= n0;
        if (!is_np) r = py_rint4(dy);
    }
    return r;
}

__device__ __forceinline__ bool point_inside(const Warp w, P2 P, const int32_t *vid, int n0) {
    const int n = w.n;
    int hits = 0;
    const P2 ray2 = mk(10000, P.y);
    const bool can_prune = P.x < 9000.0;
#pragma unroll UNROLL_PIP
    for (int base = 0; base < n; base += 32) {
        int j = base + w.lane;
        bool hit = false;
        if (j < n) {
            int jb = j == 0 ? n - 1 : j - 1;
            P2 a = w.at(j), b = w.at(jb);
            // Exact early-out, evaluated first because it removes ~97 % of the edges: an edge is only
            // counted if edge.is_cross(ray) (M:90).  If both endpoints lie strictly on the same side of
            // the ray's line and at least one of them is more than 1e-4 rad off that line as seen from
            // P, ray.straddle(edge) is False (the collinearity pre-test C:506-508 needs both quantised
            // angles in {0, pi, 2 pi}; the cross products then have equal signs).
            double dya = a.y - P.y, dyb = b.y - P.y;
            bool same_side = (dya > 1e-9 && dyb > 1e-9) || (dya < -1e-9 && dyb < -1e-9);
            bool off_axis = fabs(dya) > 1e-4 * fabs(a.x - P.x) || fabs(dyb) > 1e-4 * fabs(b.x - P.x);
            if (!(can_prune && same_side && off_axis)) {
                double ro = rint4_mixed(a.y - b.y, vid, j, jb, n0);
                if (ro != 0) {
                    // would this edge be counted if it crosses the ray?  (M:91-118)
                    bool counted;
                    if (rint(dya * 1e4) == 0) {
                        int jc = j + 1 == n ? 0 : j + 1;
                        double rn = rint4_mixed(w.at(jc).y - a.y, vid, jc, j, n0);
                        counted = !(rn == 0 || rn * ro < 0) && ro < 0;
                    } else if (rint(dyb * 1e4) == 0) {
                        int jc = jb == 0 ? n - 1 : jb - 1;
                        double rp = rint4_mixed(b.y - w.at(jc).y, vid, jb, jc, n0);
                        counted = !(rp == 0 || rp * ro < 0) && !(ro < 0);
                    } else counted = true;
                    if (counted) hit = is_cross(a, b, P, ray2);
                }
            }
        }
        hits += __popc(__ballot_sync(FULL, hit));
    }
    return (hits & 1) != 0;
}

// E:766-769 find_same_point: any boundary vertex within 0.001 of P
__device__ __forceinline__ bool find_same_point(const Warp w, P2 P) {
    bool f = false;
    const DistBound B = dist_bound(0.001);
#pragma unroll 4
    for (int j = w.lane; j < w.n; j += 32) f |= dist_less(w.at(j), P, B);
    return __any_sync(FULL, f);
}

// ---------------------------------------------------------------------------------------------
// candidate quad: validity (C:738-757, C:814-826) + corner angles
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mesh_is_valid(const Warp w, const Quad Q) {
    const int lane = w.lane;
    const P2 m[4] = {Q.at(0), Q.at(1), Q.at(2), Q.at(3)};
    // lanes 0-3: corner i = m[i].angle(m[i+1], m[i-1]) outside [0.01 pi, 0.99 pi]?   (C:746-757)
    // lanes 4-5: is_cross((m0,m1),(m2,m3)), is_cross((m0,m3),(m1,m2))                  (C:814-826)
    // Mesh.is_valid is a pure conjunction, so the evaluation order does not matter.
    bool bad = false;
    if (lane < 4) {
        P2 c = m[0], p1 = m[1], p2 = m[3];
        if (lane == 1) { c = m[1]; p1 = m[2]; p2 = m[0]; }
        if (lane == 2) { c = m[2]; p1 = m[3]; p2 = m[1]; }
        if (lane == 3) { c = m[3]; p1 = m[0]; p2 = m[2]; }
        double cr, dt;
        cross_dot(c, p1, p2, cr, dt);
        bad = corner_angle_invalid(cr, dt);
    }
    if (__any_sync(FULL, bad)) return false;
    if (lane == 4) bad = is_cross(m[0], m[1], m[2], m[3]);
    if (lane == 5) bad = is_cross(m[0], m[3], m[1], m[2]);
    return !__any_sync(FULL, bad);
}

// M:536-556 check_intersection_with_boundary.  qi[] = boundary indices of the quad vertices
// (-1 for the not-yet-inserted new vertex), ri = position of the reference point in the quad.
__device__ __forceinline__ bool intersects_boundary(const Warp w, const Quad Q, const int4 qv, int ri, P2 ref) {
    const int n = w.n;
    const P2 m[4] = {Q.at(0), Q.at(1), Q.at(2), Q.at(3)};
    const int qi[4] = {qv.x, qv.y, qv.z, qv.w};
    double max_dist = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k != ri) max_dist = fmax(max_dist, pdist(ref, m[k]));
    // ri is 1 or 2 (quad_indices); selects instead of m[(ri + k) & 3] keep the quad in registers
    const bool r1 = ri == 1;
    const P2 c1a = r1 ? m[0] : m[1], c1b = r1 ? m[3] : m[0], c2a = c1b, c2b = r1 ? m[2] : m[3];
    const DistBound MD = dist_bound(max_dist);
    auto in_mesh = [&](int j) { return j == qi[0] || j == qi[1] || j == qi[2] || j == qi[3]; };
    // Compact code on purpose (one copy of the crossing test, no unrolling): the update kernel is bound by instruction
    // fetch (profiles/r2e_*: stall_no_instruction 4.9 of 13 cycles per issue), and only the few vertices inside the
    // distance cut ever reach the test.
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int j = base + w.lane;
        bool hit = false;
        if (j < n && !in_mesh(j)) {
            P2 v = w.at(j);
            if (dist_less(ref, v, MD)) {
                // any of (c1,prev) (c1,next) (c2,prev) (c2,next) crossing -> True (order irrelevant)
#pragma unroll 1
                for (int t = 0; t < 4 && !hit; t++) {
                    const int jq = (t & 1) ? (j + 1 == n ? 0 : j + 1) : (j == 0 ? n - 1 : j - 1);
                    if (in_mesh(jq)) continue;
                    const P2 q = w.at(jq);
                    hit = (t & 2) ? is_cross(c2a, c2b, v, q) : is_cross(c1a, c1b, v, q);
                }
            }
        }
        if (__any_sync(FULL, hit)) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// boundary quality of a freshly inserted vertex (M:355-408 compute_boundary_quality)
// a_next / a_prev: interior angles at B[idx+1] and B[idx-1] (already evaluated for the candidates)
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ double boundary_quality_new_vertex(const Warp w, int idx, double a_next, double a_prev) {
    const int n = w.n;
    P2 add_v = w.at(idx);
    double amin = CUDART_INF;
    if (a_next < PI / 3) amin = a_next;
    if (a_prev < PI / 3) amin = fmin(amin, a_prev);
    double q1 = amin != CUDART_INF ? 3 * amin / PI : 1;
    int e1 = idx + 1 >= n ? idx + 1 - n : idx + 1, e2 = wrapn(idx + 2, n), e3 = idx - 1 < 0 ? idx - 1 + n : idx - 1,
        e4 = wrapn(idx - 2, n);
    double dist = pdist(add_v, w.at(e1)) + pdist(add_v, w.at(e3));
    const DistBound DB = dist_bound(dist);
    // close_vs: not excluded, nearer than `dist`, and not directly after an accepted index
    double m_d = CUDART_INF;
    unsigned carry = 0;   // parity of the run of "close" flags reaching the end of the previous chunk
#pragma unroll UNROLL_BQ
    for (int base = 0; base < n; base += 32) {
        int j = base + w.lane;
        bool c = false;
        if (j < n && j != idx && j != e1 && j != e2 && j != e3 && j != e4) c = dist_less(add_v, w.at(j), DB);
        unsigned wbits = __ballot_sync(FULL, c);
        if (c) {
            unsigned below = ~wbits & ((1u << w.lane) - 1);        // zero bits below me
            int off = below ? w.lane - (32 - __clz(below)) : w.lane + (int)carry;
            if ((off & 1) == 0) {
                int jn = j + 1 == n ? 0 : j + 1;
                m_d = fmin(m_d, seg_point_distance(w.at(jn), w.at(j), add_v));
            }
        }
        if (wbits == 0xffffffffu) carry = carry;                   // 32 more: parity unchanged
        else carry = (unsigned)__clz(~wbits) & 1u;                  // length of the run of ones at the top
    }
    m_d = warp_min_d(m_d);
    double targt_len = dist / 2;
    double dl[4];
#pragma unroll
    for (int k = -2; k < 2; k++) {
        int a = wrapn(idx + k, n), b = wrapn(idx + k + 1, n);
        dl[k + 2] = pdist(w.at(a), w.at(b));
    }
    double mean_dist = py_sum<4>(dl) / 4;
    double smoothness = fmin(mean_dist, targt_len) / fmax(mean_dist, targt_len);
    double q2 = 1;
    if (m_d != CUDART_INF) q2 = m_d < 0.5 * dist ? m_d / (0.5 * dist) : 1;
    return mg_pow(smoothness * q1 * q2, 1.0 / 3);
}

// M:418-452: element without a new vertex; t0,t1 = new indices of the two surviving quad vertices
__device__ __noinline__ double boundary_quality_no_new(const Warp w, int t0, int t1, double ang0, double ang1) {
    const int n = w.n;
    double amin = CUDART_INF;
    if (ang0 < PI / 3) amin = ang0;
    if (ang1 < PI / 3) amin = fmin(amin, ang1);
    int index = t0 < t1 ? t0 : t1;
    double targt_len = pdist(w.at(t0), w.at(t1));
    double dl[5];
#pragma unroll
    for (int k = -2; k < 3; k++) {
        int a = wrapn(index + k, n), b = wrapn(index + k + 1, n);
        dl[k + 2] = pdist(w.at(a), w.at(b));
    }
    double mean_dist = py_sum<5>(dl) / 5;
    double smoothness = fmin(mean_dist, targt_len) / fmax(mean_dist, targt_len);
    double angle_quality = amin != CUDART_INF ? 3 * amin / PI : 1;
    return mg_pow(angle_quality * smoothness, 1.0 / 2);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG for the synthetic policy and the polygon generator)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ double u01(unsigned a, unsigned b) {   // 53-bit uniform in [0,1)
    return (double)((((unsigned long long)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------
// random star polygon (ui/GenerateRandomPolygon.py:5-49) + densifier (ui/tk-ui.py:252-276)
// Written by the warp into ring[0..n); returns n (even, min_verts <= n <= max_verts).
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ int generate_polygon(const Params &P, const Warp w, long long global_env, int episode, double *dbg = nullptr);

// ---------------------------------------------------------------------------------------------
// shared-memory layout of the one-warp blocks (decide / update / observe / reset / template kernels)
// ---------------------------------------------------------------------------------------------
struct SmemLayout {
    double2 *ring;
    int *queue;
    int4 *stash;                 // 12 x 16 B: the env's EnvHot (chunks 0..7) and EnvCold (chunks 8..11)
    unsigned long long *mbar;
};
constexpr int STASH_BYTES = 256;
// with_queue = false: kernels that never touch Warp::queue (decide, update)
__device__ __forceinline__ SmemLayout carve(unsigned char *raw, int cap, bool with_queue = true) {
    SmemLayout L;
    L.ring = reinterpret_cast<double2 *>(raw);
    L.queue = reinterpret_cast<int *>(L.ring + cap);
    L.stash = reinterpret_cast<int4 *>(L.queue + (with_queue ? QCAP : 0));
    L.mbar = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(L.stash) + STASH_BYTES);
    return L;
}
size_t smem_bytes(int cap, bool with_queue = true) { return (size_t)cap * 16 + (with_queue ? (size_t)QCAP * 4 : 0) + STASH_BYTES + 16; }

// words / doubles of the stash: EnvHot = {n, ref_index, n_elements, flags | base_length(d2), failed_num(w6), ep_len(w7) |
// ep_return(d4), current_area(d5) | fan d6..15} (base_length is double 2); EnvCold at byte 128 = {original_area(d16), area_min(d17) | area_crit(d18),
// n0(w38), next_vid(w39) | stamp_ctr(w40), domain(w41), episode(w42)}
enum { W_N = 0, W_REF = 1, W_NEL = 2, W_FLAGS = 3, W_FAILED = 6, W_EP_LEN = 7, W_N0 = 38, W_NEXT_VID = 39, W_STAMP_CTR = 40 };
enum { D_BASE = 2, D_EP_RETURN = 4, D_CUR_AREA = 5, D_ORIGINAL_AREA = 16, D_AREA_MIN = 17, D_AREA_CRIT = 18 };
__device__ __forceinline__ void stash_records(const Params &P, int4 *stash, int env, int lane) {
    if (lane < 8) stash[lane] = __ldcg(reinterpret_cast<const int4 *>(P.hot + env) + lane);
    else if (lane < 12) stash[lane] = __ldcg(reinterpret_cast<const int4 *>(P.cold + env) + lane - 8);
}

// ---------------------------------------------------------------------------------------------
// candidate quads (E:236-283) and the memoised verdicts of the rule -1 / +1 quads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void quad_indices(int rule, bool new_vertex, int idx, int n, int (&qi)[4], int &ri) {
    const int ip1 = idx + 1 >= n ? idx + 1 - n : idx + 1, im1 = idx - 1 < 0 ? idx - 1 + n : idx - 1;
    if (new_vertex) {
        qi[0] = -1; qi[1] = im1; qi[2] = idx; qi[3] = ip1; ri = 2;
    } else if (rule == -1) {
        qi[0] = im1; qi[1] = idx; qi[2] = ip1; qi[3] = wrapn(idx + 2, n); ri = 1;
    } else {
        qi[0] = wrapn(idx - 2, n); qi[1] = im1; qi[2] = idx; qi[3] = ip1; ri = 2;
    }
}

// EnvHot::flags of the state in the ring.  A rule -1 / rule +1 element consists of boundary vertices only, so whether
// such a step is accepted (E:319-323: Mesh.is_valid(0) and not check_intersection_with_boundary) is a function of
// the state.  The O(1) half (Mesh.is_valid) is evaluated here, when the state changes; a valid quad is marked
// pending and the O(n) half runs in the decide kernel the first time a rule action actually asks for it.
__device__ __noinline__ int memo_flags(const Warp w, int idx) {
    const int n = w.n, lane = w.lane;
    if (idx < 0 || n < 6) return 0;
    // lanes 0..5: the six tests of Mesh.is_valid (4 corners, 2 edge crossings; a pure conjunction, C:738-757 / C:814-826)
    // on the rule -1 quad [B[i-1], B[i], B[i+1], B[i+2]]; lanes 8..13: the same on the rule +1 quad [B[i-2] .. B[i+1]]
    const int g = (lane >> 3) & 1, t = lane & 7, first = idx - 1 - g;
    const P2 m0 = w.at(wrapn(first, n)), m1 = w.at(wrapn(first + 1, n)), m2 = w.at(wrapn(first + 2, n)), m3 = w.at(wrapn(first + 3, n));
    bool bad = false;
    if (lane < 16) {
        if (t < 4) {
            P2 c = m0, p1 = m1, p2 = m3;
            if (t == 1) { c = m1; p1 = m2; p2 = m0; }
            if (t == 2) { c = m2; p1 = m3; p2 = m1; }
            if (t == 3) { c = m3; p1 = m0; p2 = m2; }
            double cr, dt;
            cross_dot(c, p1, p2, cr, dt);
            bad = corner_angle_invalid(cr, dt);
        } else if (t == 4) bad = is_cross(m0, m1, m2, m3);
        else if (t == 5) bad = is_cross(m0, m3, m1, m2);
    }
    const unsigned bm = __ballot_sync(FULL, bad);
    return ((bm & 0x003Fu) == 0 ? HOT_PEND_M1 : 0) | ((bm & 0x3F00u) == 0 ? HOT_PEND_P1 : 0);
}

// Mesh.is_valid(0) (C:738-757, C:814-826) of one quad by one thread (the screen kernel): same predicates as
// mesh_is_valid, evaluated in sequence; the conjunction does not depend on the order.
__device__ __forceinline__ bool quad_valid_serial(P2 m0, P2 m1, P2 m2, P2 m3) {
    double cr, dt;
    cross_dot(m0, m1, m3, cr, dt);
    if (corner_angle_invalid(cr, dt)) return false;
    cross_dot(m1, m2, m0, cr, dt);
    if (corner_angle_invalid(cr, dt)) return false;
    cross_dot(m2, m3, m1, cr, dt);
    if (corner_angle_invalid(cr, dt)) return false;
    cross_dot(m3, m0, m2, cr, dt);
    if (corner_angle_invalid(cr, dt)) return false;
    if (is_cross(m0, m1, m2, m3)) return false;
    if (is_cross(m0, m3, m1, m2)) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// record stores (by the env's warp; the ring holds the env's current boundary)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_hot(EnvHot *dst_rec, const Warp &w, int ref_index, int n_elements, int flags,
                                          double base_length, int failed_num, int ep_len, double ep_return, double current_area) {
    int4 *dst = reinterpret_cast<int4 *>(dst_rec);
    if (w.lane == 0) {
        dst[0] = make_int4(w.n, ref_index, n_elements, flags);
        dst[1] = make_int4(__double2loint(base_length), __double2hiint(base_length), failed_num, ep_len);
        dst[2] = make_int4(__double2loint(ep_return), __double2hiint(ep_return), __double2loint(current_area),
                           __double2hiint(current_area));
    }
    if (w.lane < 5) {                       // fan = B[i-2 .. i+2]
        const int j = w.n > 0 ? wrapn((ref_index < 0 ? 0 : ref_index) - 2 + w.lane, w.n) : 0;
        reinterpret_cast<double2 *>(dst)[3 + w.lane] = w.ring[j];
    }
}

// ---------------------------------------------------------------------------------------------
// reset of one env (E:136-184): restore the polygon, rebuild candidates, first observation.  Writes the env's
// rings and records; returns the observation (lane < 18).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float reset_env(const Params &P, Warp &w, int env, int domain, int episode) {
    const int lane = w.lane;
    const size_t off = (size_t)env * P.cap;
    float obs;
    __syncwarp();
    if (!P.random_mode) {
        const DomainProto *T = P.t_proto + domain;
        const int n0 = T->cold.n0;
        const size_t toff = (size_t)domain * P.cap;
#pragma unroll 1
        for (int j = lane; j < n0; j += 32) {
            P.xy[off + j] = P.t_xy[toff + j];
            P.key[off + j] = P.t_key[toff + j];
            P.stamp[off + j] = P.t_stamp[toff + j];
            P.vid[off + j] = j;
        }
        if (lane < 8) reinterpret_cast<int4 *>(P.hot + env)[lane] = reinterpret_cast<const int4 *>(&T->hot)[lane];
        if (lane == 8) {
            EnvCold C = T->cold;
            C.domain = domain; C.episode = episode;
            P.cold[env] = C;
        }
        obs = lane < MG_OBS_DIM ? P.t_obs[domain * MG_OBS_DIM + lane] : 0.0f;
    } else {
        const int n = generate_polygon(P, w, P.env_id_offset + env, episode);
        w.n = n;
        __syncwarp();
#pragma unroll 1
        for (int j = lane; j < n; j += 32) {
            P.xy[off + j] = w.ring[j];
            P.vid[off + j] = j;
        }
        rebuild_candidates(w, P.key + off, P.stamp + off);
        __syncwarp();
        EnvCold C;
        C.original_area = shoelace_area(w);
        { const double2 ar = estimate_area_range(w); C.area_min = ar.x; C.area_crit = ar.y; }
        C.n0 = n; C.next_vid = n; C.stamp_ctr = 0; C.domain = domain; C.episode = episode;
#pragma unroll
        for (int k = 0; k < 5; k++) C.pad[k] = 0;
        const int ref_index = find_reference_index(w, P.key + off, P.stamp + off);
        obs = 0.0f;
        double base = 0;
        if (ref_index >= 0) {
            const ObsOut R = compute_obs(w, P.sc_full, ref_index, C.original_area / C.original_area);
            obs = R.obs; base = R.base;
        }
        const int flags = memo_flags(w, ref_index);
        store_hot(P.hot + env, w, ref_index, 0, flags, base, 0, 0, 0.0, C.original_area);
        if (lane == 0) P.cold[env] = C;
    }
    return obs;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// Per-domain reset template: one warp per domain.
__global__ void __launch_bounds__(32) mg_template_kernel(const __grid_constant__ Params P, double2 *t_xy, double *t_key, int32_t *t_stamp,
                                                       DomainProto *t_proto, float *t_obs, const int32_t *n0s, const double *areas) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int d = blockIdx.x;
    if (d >= P.n_domains) return;
    SmemLayout L = carve(smem_raw, P.cap);
    Warp w;
    w.ring = L.ring; w.queue = L.queue; w.lane = lane; w.n = n0s[d];
    const size_t toff = (size_t)d * P.cap;
#pragma unroll 1
    for (int j = lane; j < w.n; j += 32) w.ring[j] = t_xy[toff + j];
    __syncwarp();
    rebuild_candidates(w, t_key + toff, t_stamp + toff);
    __syncwarp();
    EnvCold C;
    C.original_area = areas ? areas[d] : shoelace_area(w);
    { const double2 ar = estimate_area_range(w); C.area_min = ar.x; C.area_crit = ar.y; }
    C.n0 = w.n; C.next_vid = w.n; C.stamp_ctr = 0; C.domain = d; C.episode = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) C.pad[k] = 0;
    const int ref_index = find_reference_index(w, t_key + toff, t_stamp + toff);
    float obs = 0.0f;
    double base = 0;
    if (ref_index >= 0) {
        const ObsOut R = compute_obs(w, P.sc_full, ref_index, C.original_area / C.original_area);
        obs = R.obs; base = R.base;
    }
    const int flags = memo_flags(w, ref_index);
    store_hot(&t_proto[d].hot, w, ref_index, 0, flags, base, 0, 0, 0.0, C.original_area);
    if (lane == 0) t_proto[d].cold = C;
    if (lane < MG_OBS_DIM) t_obs[d * MG_OBS_DIM + lane] = obs;
}

__global__ void __launch_bounds__(32) mg_reset_kernel(const __grid_constant__ Params P, const uint8_t *mask, float *obs_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int env = blockIdx.x;
    if (env >= P.num_envs) return;
    SmemLayout L = carve(smem_raw, P.cap);
    Warp w;
    w.ring = L.ring; w.queue = L.queue; w.lane = lane; w.n = 0;
    float obs;
    if (mask == nullptr || mask[env]) {
        const int domain = P.cold[env].domain, episode = P.cold[env].episode;
        obs = reset_env(P, w, env, domain, episode);
        if (lane < MG_OBS_DIM) P.obs_cache[(size_t)env * MG_OBS_DIM + lane] = obs;
    } else {
        obs = lane < MG_OBS_DIM ? P.obs_cache[(size_t)env * MG_OBS_DIM + lane] : 0.0f;
    }
    if (obs_out && lane < MG_OBS_DIM) obs_out[(size_t)env * MG_OBS_DIM + lane] = obs;
}

// ---------------------------------------------------------------------------------------------
// One environment transition (E:388-457) + VecEnv auto-reset, as a few kernels with SMALL, HOMOGENEOUS code paths.
//
//   1  mg_step_screen_kernel   all envs, ONE THREAD per env.  Reads the 128-byte EnvHot record and the action and
//      settles every step whose outcome follows from them: a rule -1 / +1 action against the memoised verdict
//      (flags), a rule-0 action whose new-vertex quad [P, B[i-1], B[i], B[i+1]] is not a valid element (Mesh.is_valid
//      needs only the neighbour fan) while the rule -1 fallback (E:258-262) is known to fail too.  A failed step
//      leaves the state untouched and returns the cached observation (verified bit-identical), so
//      these envs never touch their boundary ring: reward, flags, counters and the 32 changed bytes of the record
//      are all that moves.  Under a uniform random policy ~91 % of the steps end here.
//   The others run ONE WARP per work item over compacted lists; the boundary is staged once per item into a
//   shared-memory ring with one cp.async.bulk:
//   2  mg_step_decide_kernel   (only with the option fuse_decide = 0; by default this work is the first half of kernel 3)
//      rule-0 candidates whose new vertex still needs the whole boundary (point-in-polygon M:74-128, find_same_point
//      E:766-769, boundary intersection M:536-556) and rule -1 / +1 actions whose verdict is still pending (valid
//      quad, intersection test not done yet: evaluated once per state, then memoised).  Failed steps end here.
//   3  mg_step_update_kernel   accepted elements: element log, boundary update, candidate keys, area, quality,
//      reward, termination (M:601-674, C:943-958, C:881-892, M:355-452, E:590-607, E:345-351).
//   4  mg_step_observe_kernel  every env whose state changed: reference point (M:295-316), next observation
//      (C:1192-1290), new memo; and the in-place resets of the envs that finished inside kernel 3.
//   5  mg_step_reset_kernel    in-place resets (template copy or fresh random polygon) of the envs kernel 1 truncated,
//      on a side stream next to kernels 3 / 4.
//
// Why several launches and why the lists are sorted by item kind: a warp-per-item kernel is bound by instruction
// fetch as soon as the warps of an SM are at unrelated places of a large image (profiles/README.md: one 248 KB ring
// kernel spent 57 % of its stall samples in stall_no_instruction, round 1's 186 KB apply kernel 25 %, round 2's fused
// decide + update kernel 38 % until its list was served kind by kind).  Kernels hand items over through 32-byte work
// records.
//
// Work lists are appended with one atomicAdd per warp and list segment.  Two sets of counters alternate: the screen
// kernel of step s uses set (CNT_STEP & 1) and records it in CNT_CUR; the later kernels read CNT_CUR, and the first
// thread of the observe kernel clears the other (idle) set and advances CNT_STEP for the next step.  The parity lives in
// device memory, so no memset sits between launches and any sequence of mg_step calls can be captured in a CUDA graph.
// ---------------------------------------------------------------------------------------------
struct StepIO {
    const float *act;
    float *obs_out;
    double *rew_out;
    uint8_t *term_out;
    uint8_t *trunc_out;
    float *term_obs_out;
    int32_t *n_elem_out;
    int obs_full;     // 1: every row of obs_out is written.  0: obs_out still holds the previous step's observations
                      // (same unmodified buffer, mg_set_obs_delta): only the rows that change are written
    // Result delta (mg_step_host with pinned result buffers in delta mode): device-resident copies of what the caller's
    // reward / flag / element-count arrays hold.  A value is written to the caller's array -- a posted PCIe write --
    // only when it differs from that copy (a failed step mostly repeats the env's previous reward, flags and count:
    // ~85 % of the result writes of a step); res_full = 1 writes everything once (new buffers).  All nullptr otherwise.
    double *rew_sh;
    uint8_t *term_sh, *trunc_sh;
    int32_t *nel_sh;
    int res_full;
};

// reward, flags and element count of one env to the caller's arrays (see StepIO::rew_sh); returns the bytes written in
// delta mode (0 otherwise: every env's 10 or 14 bytes are then accounted on the host)
__device__ __forceinline__ int emit_results(const StepIO &io, int env, double reward, bool term, bool trunc, int n_elements) {
    if (io.rew_sh == nullptr) {
        io.rew_out[env] = reward;
        io.term_out[env] = term;
        io.trunc_out[env] = trunc;
        if (io.n_elem_out) io.n_elem_out[env] = n_elements;
        return 0;
    }
    const bool full = io.res_full != 0;
    const double r0 = io.rew_sh[env];
    const uint8_t t0 = io.term_sh[env], u0 = io.trunc_sh[env];
    int bytes = 0;
    if (full || __double_as_longlong(r0) != __double_as_longlong(reward)) { io.rew_out[env] = reward; io.rew_sh[env] = reward; bytes += 8; }
    if (full || t0 != (uint8_t)term) { io.term_out[env] = term; io.term_sh[env] = term; bytes += 1; }
    if (full || u0 != (uint8_t)trunc) { io.trunc_out[env] = trunc; io.trunc_sh[env] = trunc; bytes += 1; }
    if (io.n_elem_out) {
        const int32_t n0 = io.nel_sh[env];
        if (full || n0 != n_elements) { io.n_elem_out[env] = n_elements; io.nel_sh[env] = n_elements; bytes += 4; }
    }
    return bytes;
}

// |x * 1e4 - (k + 0.5)| < tol for some integer k: np_round4(x) could flip under a perturbation of x
__device__ __forceinline__ bool near_round4_tie(double x, double tol) {
    double p = x * 1e4;
    return 0.5 - fabs(p - rint(p)) < tol;
}
// E:202-210 / D:112-137 exactly as written: theta = 2 pi - atan2(dy, dx), cos/sin of that double
__device__ __noinline__ double2 action_frame_exact(double ax, double ay, double dx, double dy, double base, P2 ref) {
    double th = 2 * PI - atan2(dy, dx);
    const double2 sc = mg_sincos(th);
    const double s = sc.x, c = sc.y;
    double ox = c * ax + s * ay;
    double oy = -s * ax + c * ay;
    ox *= base; oy *= base;
    ox += ref.x; oy += ref.y;
    return make_double2(ox, oy);
}

// action -> candidate vertex (E:783-792, E:202-210, D:112-137).
// theta = 2 pi - atan2(dy, dx)  =>  cos(theta) ~ dx / r, sin(theta) ~ -dy / r.  The frame is first evaluated from
// the edge vector (no atan2 / sincos); that value is within ~1e-14 of the reference's, so both round to the same
// 4-decimal vertex unless a coordinate sits next to a rounding tie.  Ties are NOT rare on axis-aligned domains
// (dyadic action components times a short base length land exactly on k + 0.5, and the reference's
// sin(fl(2 pi)) = -2.4e-16 then decides): in that band the reference's own expression is evaluated (exact-safe
// filter, like the angle classes of mg_math.cuh).
__device__ __forceinline__ P2 action_to_point(float a1, float a2, P2 ref, P2 right_p, double base_length) {
    double ax = (double)np_round4f(a1), ay = (double)np_round4f(a2);
    double dx = right_p.x - ref.x, dy = right_p.y - ref.y;
    double r2 = dx * dx + dy * dy;
    double rinv = rsqrt(r2);                        // estimate only: the tie band below is > 1000x wider than its error
    double c = r2 > 0 ? dx * rinv : 1.0, s = r2 > 0 ? -(dy * rinv) : 0.0;
    double ox = c * ax + s * ay;
    double oy = -s * ax + c * ay;
    ox *= base_length; oy *= base_length;
    ox += ref.x; oy += ref.y;
    const double tol = 1e-8 * (3 * fabs(base_length) + fabs(ref.x) + fabs(ref.y) + 1);   // in units of 1e-4
    if (near_round4_tie(ox, tol) || near_round4_tie(oy, tol)) {
        const double2 o = action_frame_exact(ax, ay, dx, dy, base_length, ref);
        ox = o.x; oy = o.y;
    }
    return mk(np_round4(ox), np_round4(oy));
}

// Work lists are binned by boundary size (largest first): the item kernels hand out long items before short ones, so
// that the tail of a launch is made of short items (longest-processing-time-first scheduling).
__device__ __forceinline__ int size_bin(int n, int cap) {
    return n > (3 * cap) / 4 ? 0 : (n > cap / 2 ? 1 : (n > cap / 4 ? 2 : 3));
}
// one work record appended by one warp (lane 0)
__device__ __forceinline__ void push_item(WorkItem *list, int *counters, const Params &P, int bin, const WorkItem &W, int lane) {
    if (lane == 0) {
        const int i = atomicAdd(&counters[bin], 1);
        if (i < P.num_envs) list[(size_t)bin * P.num_envs + i] = W;
    }
}
__device__ __forceinline__ WorkItem make_item(int env, int n, int kind, int rule, int flag, int done, double x, double y) {
    WorkItem W;
    W.newx = x; W.newy = y; W.env = env; W.n = n;
    W.kind = (int8_t)kind; W.rule = (int8_t)rule; W.flag = (int8_t)flag; W.done = (int8_t)done; W.pad = 0;
    return W;
}
// item t of a segmented list (segments in order); counts[] = the list's segment counters, already clamped
template <int NSEG>
__device__ __forceinline__ WorkItem fetch_item(const WorkItem *list, const int (&counts)[NSEG], int num_envs, int t) {
    int seg = 0;
#pragma unroll
    for (int b = 0; b < NSEG - 1; b++)
        if (seg == b && t >= counts[b]) { t -= counts[b]; seg = b + 1; }
    return list[(size_t)seg * num_envs + t];
}

// ---- kernel 1: screen ------------------------------------------------------------------------
constexpr int SCREEN_THREADS = 128;
__global__ void __launch_bounds__(SCREEN_THREADS) mg_step_screen_kernel(const __grid_constant__ Params P, const __grid_constant__ StepIO io) {
    __shared__ uint8_t s_copy[SCREEN_THREADS];     // bit 0: cached observation -> obs_out, bit 1: -> term_obs_out
    // The block's 128 records (16 KB) come in with coalesced 16-byte loads, eight in flight per thread -- one DRAM round
    // trip instead of two dependent ones (scalars, then the fan) -- and are read back from shared memory; chunk c of
    // record r sits at r * 8 + (c ^ (r & 7)) so that neither the stores nor the per-thread reads conflict.
    __shared__ int4 s_rec[SCREEN_THREADS * 8];
    const int env = blockIdx.x * SCREEN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = env < P.num_envs;
    asm volatile("griddepcontrol.launch_dependents;");      // the update kernel may be placed while this grid runs (see "pdl")
    float a0 = 0, a1 = 0, a2 = 0;
    int set;
    {
        const int4 *src = reinterpret_cast<const int4 *>(P.hot) + (size_t)blockIdx.x * SCREEN_THREADS * 8;
        const int limit = (P.num_envs - blockIdx.x * SCREEN_THREADS) * 8;        // chunks of this block that exist
        int4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int g = k * SCREEN_THREADS + threadIdx.x;
            v[k] = g < limit ? src[g] : make_int4(0, 0, 0, 0);
        }
        // the action and the step parity are requested before the first use of the records
        if (active) { a0 = io.act[(size_t)env * 3 + 0]; a1 = io.act[(size_t)env * 3 + 1]; a2 = io.act[(size_t)env * 3 + 2]; }
        set = P.counters[CNT_STEP] & 1;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int g = k * SCREEN_THREADS + threadIdx.x, r = g >> 3, c = g & 7;
            s_rec[r * 8 + (c ^ (r & 7))] = v[k];
        }
    }
    (void)a2;
    if (env == 0) P.counters[CNT_CUR] = set;          // the set this step's lists use
    int *cnt = P.counters + CNT_SET * set;
    __syncthreads();
    auto rec_chunk = [&](int c) { return s_rec[threadIdx.x * 8 + (c ^ (threadIdx.x & 7))]; };

    // where the step goes: 0 = settled here, 1 = decide list, 2 = accept list
    int route = 0;
    bool settled_done = false, trunc_flag = false, term_flag = false;
    WorkItem W = make_item(env, 0, WORK_APPLY, 0, 0, 0, 0.0, 0.0);
    int n = 0, n_elements = 0, ep_len = 0;
    double ep_return = 0;
    uint8_t copy_code = 0;
    int res_bytes = 0;
    if (active) {
        const int4 c0 = rec_chunk(0), c1 = rec_chunk(1), c2 = rec_chunk(2);
        n = c0.x;
        const int ref_index = c0.y, flags = c0.w;
        n_elements = c0.z;
        const double base_length = __hiloint2double(c1.y, c1.x);
        int failed_num = c1.z;
        ep_len = c1.w;
        ep_return = __hiloint2double(c2.y, c2.x);
        W.n = n;
        // An env without a reference point (empty candidate list, E:736-738 returns None) has no defined
        // continuation in the reference (its next step raises): it is reported truncated.
        const bool dead = ref_index < 0 || n < 3;
        double reward = 0;
        bool done = false, fail_penalty = false;
        if (dead) {
        } else if (n <= 5) {                              // E:428-430
            reward = 10; done = true;
        } else if (a0 <= -0.5f || a0 >= 0.5f) {           // rule -1 / +1: memoised verdict
            const bool m1 = a0 <= -0.5f;
            W.rule = m1 ? -1 : 1;
            if (flags & (m1 ? HOT_OK_M1 : HOT_OK_P1)) route = 2;
            else if (flags & (m1 ? HOT_PEND_M1 : HOT_PEND_P1)) { route = 1; W.kind = WORK_DECIDE_RULE; }
            else fail_penalty = true;
        } else {
            const int4 q1 = rec_chunk(4), q2 = rec_chunk(5), q3 = rec_chunk(6);          // fan[1], fan[2], fan[3]
            const P2 ref = mk(__hiloint2double(q2.y, q2.x), __hiloint2double(q2.w, q2.z)),
                     right_p = mk(__hiloint2double(q1.y, q1.x), __hiloint2double(q1.w, q1.z)),
                     left_p = mk(__hiloint2double(q3.y, q3.x), __hiloint2double(q3.w, q3.z));
            const P2 newp = action_to_point(a1, a2, ref, right_p, base_length);
            // the new-vertex element (E:264-270) must pass Mesh.is_valid (E:319); if the candidate vertex coincides
            // with a boundary vertex the rule -1 element is used instead (E:258-262).  When neither can be accepted
            // the step fails whatever the point-in-polygon test says (same reward, E:279 / E:357).
            const bool qvalid = quad_valid_serial(newp, right_p, ref, left_p);
            if (qvalid || (flags & (HOT_OK_M1 | HOT_PEND_M1))) {
                route = 1;
                W.kind = WORK_DECIDE_NEW; W.flag = qvalid ? 1 : 0; W.newx = newp.x; W.newy = newp.y;
            } else fail_penalty = true;
        }
        if (route == 0) {
            // failed step (or the n <= 5 sentinel): nothing changed, the observation is the cached one
            if (fail_penalty) reward += n_elements ? -1.0 / n_elements : -1;          // E:279 / E:357
            failed_num++;
            bool is_complete = true;
            if (failed_num >= 100) { done = true; is_complete = false; }              // E:382-384
            term_flag = done && is_complete;
            trunc_flag = done && !is_complete;
            if (dead && !done) { done = true; trunc_flag = true; }                    // sentinel, see DESIGN.md
            ep_return += reward;
            ep_len++;
            settled_done = done;
            res_bytes = emit_results(io, env, reward, term_flag, trunc_flag, n_elements);
            int4 *dst = reinterpret_cast<int4 *>(P.hot + env);
            dst[1] = make_int4(c1.x, c1.y, failed_num, ep_len);
            dst[2] = make_int4(__double2loint(ep_return), __double2hiint(ep_return), c2.z, c2.w);
            copy_code = (io.obs_full ? 1 : 0) | ((done && io.term_obs_out) ? 2 : 0);
        }
    }
    // ---- work lists: the lanes of a warp that feed the same (list, size bin) share one atomicAdd, and the atomics of the
    // up to nine lists are issued together (one round trip to L2, not one per list) --------------------------------
    {
        // envs whose episode ended in a step settled here are reset by mg_step_reset_kernel, which runs next to the
        // update kernel (a reset is the longest item of a step and depends on nothing the other kernels produce)
        const bool to_reset = settled_done && P.auto_reset;
        // destinations: decide list segments (new-vertex candidates 0..3, pending rule verdicts 4..7), accept list bins
        // (8..11), reset list (12)
        const int bin = size_bin(n, P.cap);
        const int dest = route == 1 ? (W.kind == WORK_DECIDE_RULE ? NBINS : 0) + bin
                                    : (route == 2 ? DECIDE_SEGS + bin : (to_reset ? DECIDE_SEGS + NBINS : -1));
        const unsigned grp = __match_any_sync(FULL, dest);
        const int leader = __ffs(grp) - 1;
        int base = 0;
        if (dest >= 0 && lane == leader) {
            int *ctr = dest < DECIDE_SEGS ? cnt + CNT_DECIDE + dest
                                          : (dest < DECIDE_SEGS + NBINS ? cnt + CNT_ACCEPT + dest - DECIDE_SEGS : cnt + CNT_RESET);
            base = atomicAdd(ctr, __popc(grp));
        }
        base = __shfl_sync(FULL, base, leader);
        const int slot = base + __popc(grp & ((1u << lane) - 1));
        if (dest >= 0 && slot < P.num_envs) {
            if (dest < DECIDE_SEGS) P.decide_list[(size_t)dest * P.num_envs + slot] = W;
            else if (dest < DECIDE_SEGS + NBINS) P.accept_list[(size_t)(dest - DECIDE_SEGS) * P.num_envs + slot] = W;
            else P.reset_list[slot] = env;
        }
    }
    // ---- statistics: one set of atomics per warp --------------------------------------------
    {
        const unsigned steps = __popc(__ballot_sync(FULL, active));
        const unsigned sum_n = __reduce_add_sync(FULL, active ? (unsigned)n : 0u);
        const unsigned dm = __ballot_sync(FULL, settled_done);
        StatsAcc *T = P.stats + ((blockIdx.x * (SCREEN_THREADS / 32) + (threadIdx.x >> 5)) & (STAT_SLOTS - 1));
        const unsigned rm = __ballot_sync(FULL, route != 0);
        const unsigned ring_n = __reduce_add_sync(FULL, route != 0 ? (unsigned)n : 0u);
        if (io.rew_sh != nullptr) {
            const int wb = __reduce_add_sync(FULL, res_bytes);
            if (lane == 0 && wb) atomicAdd(&cnt[CNT_RESBYTES], wb);
        }
        if (lane == 0 && steps) {
            atomicAdd(&T->steps, (unsigned long long)steps);
            atomicAdd(&T->sum_n, (unsigned long long)sum_n);
            if (rm) {
                atomicAdd(&T->ring_items, (unsigned long long)__popc(rm));
                atomicAdd(&T->sum_n_ring, (unsigned long long)ring_n);
            }
        }
        if (dm) {
            const unsigned n_term = __popc(__ballot_sync(FULL, settled_done && term_flag));
            const unsigned n_trunc = __popc(__ballot_sync(FULL, settled_done && trunc_flag));
            const unsigned elems = __reduce_add_sync(FULL, settled_done ? (unsigned)n_elements : 0u);
            const double ret = warp_sum_d(settled_done ? ep_return : 0.0);
            const unsigned len = __reduce_add_sync(FULL, settled_done ? (unsigned)ep_len : 0u);
            if (lane == 0) {
                atomicAdd(&cnt[CNT_DONE], __popc(dm));
                atomicAdd(&T->episodes, (unsigned long long)__popc(dm));
                atomicAdd(&T->completed, (unsigned long long)n_term);
                atomicAdd(&T->truncated, (unsigned long long)n_trunc);
                atomicAdd(&T->elements, (unsigned long long)elems);
                atomicAdd(&T->sum_return, ret);
                atomicAdd(&T->sum_length, (double)len);
            }
        }
    }
    // ---- observation rows of the settled envs (cached: nothing changed) ------------------------
    s_copy[threadIdx.x] = copy_code;
    if (__syncthreads_or(copy_code != 0)) {
        const size_t row0 = (size_t)blockIdx.x * SCREEN_THREADS * MG_OBS_DIM;
        for (int e = threadIdx.x; e < SCREEN_THREADS * MG_OBS_DIM; e += SCREEN_THREADS) {
            const uint8_t code = s_copy[e / MG_OBS_DIM];
            if (code) {
                const float v = P.obs_cache[row0 + e];
                if (code & 1) io.obs_out[row0 + e] = v;
                if (code & 2) io.term_obs_out[row0 + e] = v;
            }
        }
    }
}

// ---- the warp-per-item kernels ------------------------------------------------------------------
// Item loop shared by them: every item comes from a ticket counter -- also a block's first one, so that an item never
// waits for one particular block to become resident (blocks of the reset kernel on the side stream can hold slots when
// a grid starts: a statically assigned first item then started tens of microseconds late).  The next ticket is
// requested shortly before the END of the current item (MG_ITEM_TICKET): its latency hides behind the item's last
// stores, and an item is only bound to a block when that block is about to be free.  (Requested at the start of an
// item, every second-wave item was bound at t = 0 to an arbitrary block and waited for that block's first item, however
// long -- profiles/r2_trace_c3_before.txt.)
#define MG_ITEM_LOOP_BEGIN(total, ticket)                                                   \
    int *const ticket_ = (ticket);                                                          \
    int t_ = 0;                                                                             \
    if (lane == 0) t_ = atomicAdd(ticket_, 1);                                              \
    t_ = __shfl_sync(FULL, t_, 0);                                                          \
    MG_TRACE_BLOCK                                                                          \
    if (t_ < (total)) init_mbar(L.mbar, lane);                                              \
    [[maybe_unused]] unsigned phase = 0;                                                    \
    _Pragma("unroll 1") while (t_ < (total)) {                                              \
        int next_t_ = 0;                                                                    \
        bool have_next_ = false;
#define MG_ITEM_TICKET                                                                      \
        if (!have_next_) {                                                                  \
            if (lane == 0) next_t_ = atomicAdd(ticket_, 1);                                 \
            have_next_ = true;                                                              \
        }
#define MG_ITEM_LOOP_END                          \
        MG_ITEM_TICKET                            \
        t_ = __shfl_sync(FULL, next_t_, 0);       \
    }

template <int NSEG>
__device__ __forceinline__ int load_counts(const int *ctr, int num_envs, int (&counts)[NSEG]) {
    int total = 0;
#pragma unroll
    for (int b = 0; b < NSEG; b++) { counts[b] = min(ctr[b], num_envs); total += counts[b]; }
    return total;
}

struct Stash {
    int4 *p;
    __device__ __forceinline__ int i(int word) const { return reinterpret_cast<const int32_t *>(p)[word]; }
    __device__ __forceinline__ double d(int dword) const { return reinterpret_cast<const double *>(p)[dword]; }
};

// E:319-323 for a quad whose Mesh.is_valid(0) is already known to hold: not check_intersection_with_boundary
__device__ __noinline__ bool quad_clear_of_boundary(const Warp w, const Quad Q, const int4 qv, int ri, P2 ref) {
    return !intersects_boundary(w, Q, qv, ri, ref);
}
// the pending half of a memoised rule -1 / +1 verdict (see memo_flags)
__device__ __forceinline__ bool rule_quad_clear(const Warp w, int rule, int idx) {
    int qi[4], ri;
    quad_indices(rule, false, idx, w.n, qi, ri);
    Quad Q;
#pragma unroll
    for (int k = 0; k < 4; k++) { const P2 p = w.at(qi[k]); Q.x[k] = p.x; Q.y[k] = p.y; }
    return quad_clear_of_boundary(w, Q, make_int4(qi[0], qi[1], qi[2], qi[3]), ri, w.at(idx));
}

// The verdict on one decide item (E:236-283, E:319-323); the boundary is staged in w.ring and the records in the
// stash.  Returns true when an element is accepted (rule / new_vertex say which); a failed step is finished here
// (nothing changed: cached observation, same tail as the screen kernel).
__device__ __forceinline__ bool decide_item(const Params &P, const StepIO &io, int *cnt, const Warp &w, const WorkItem &W,
                                            const Stash S, int &rule, bool &new_vertex, int *ticket, int &next_t) {
    const int lane = w.lane, env = W.env;
    const size_t off = (size_t)env * P.cap;
    const int idx = S.i(W_REF);
    int flags = S.i(W_FLAGS);
    rule = W.rule;
    new_vertex = false;
    bool accepted = false, flags_changed = false;
    const P2 newp = mk(W.newx, W.newy);
    auto rule_verdict = [&](int r) {          // memoised verdict of the rule r element, resolving a pending one
        const int ok_bit = r == -1 ? HOT_OK_M1 : HOT_OK_P1, pend_bit = r == -1 ? HOT_PEND_M1 : HOT_PEND_P1;
        if (flags & ok_bit) return true;
        if (!(flags & pend_bit)) return false;
        const bool ok = rule_quad_clear(w, r, idx);
        flags = (flags & ~pend_bit) | (ok ? ok_bit : 0);
        flags_changed = true;
        return ok;
    };
    if (W.kind == WORK_DECIDE_RULE) {
        accepted = rule_verdict(rule);
    } else if (point_inside(w, newp, P.vid + off, S.i(W_N0))) {
        if (find_same_point(w, newp)) {                    // E:258-262: an existing vertex: the rule -1 element
            rule = -1;
            accepted = rule_verdict(-1);
        } else if (W.flag) {                               // Mesh.is_valid of the new-vertex quad (screen kernel)
            int qi[4], ri;
            quad_indices(0, true, idx, w.n, qi, ri);
            Quad Q;
            Q.x[0] = newp.x; Q.y[0] = newp.y;
#pragma unroll
            for (int k = 1; k < 4; k++) { const P2 p = w.at(qi[k]); Q.x[k] = p.x; Q.y[k] = p.y; }
            accepted = quad_clear_of_boundary(w, Q, make_int4(qi[0], qi[1], qi[2], qi[3]), ri, w.at(idx));
            new_vertex = accepted;
        }
    }
    if (flags_changed && lane == 0) P.hot[env].flags = flags;
    if (accepted) return true;
    // ---- failed step ----------------------------------------------------------------------------
    if (lane == 0) next_t = atomicAdd(ticket, 1);                        // the block's next item (see MG_ITEM_TICKET)
    const int n_el = S.i(W_NEL);
    const double reward = n_el ? -1.0 / n_el : -1;                      // E:279 / E:357
    const int failed_num = S.i(W_FAILED) + 1;
    const bool done = failed_num >= 100;                                 // E:382-384
    const double ep_return = S.d(D_EP_RETURN) + reward;
    const int ep_len = S.i(W_EP_LEN) + 1;
    if (lane == 0) {
        const int wb = emit_results(io, env, reward, false, done, n_el);
        if (wb) atomicAdd(&cnt[CNT_RESBYTES], wb);
        int4 *dst = reinterpret_cast<int4 *>(P.hot + env);
        const int4 c1 = S.p[1], c2 = S.p[2];
        dst[1] = make_int4(c1.x, c1.y, failed_num, ep_len);
        dst[2] = make_int4(__double2loint(ep_return), __double2hiint(ep_return), c2.z, c2.w);
        if (done) {
            StatsAcc *T = P.stats + (env & (STAT_SLOTS - 1));
            atomicAdd(&cnt[CNT_DONE], 1);
            atomicAdd(&T->episodes, 1ull);
            atomicAdd(&T->truncated, 1ull);
            atomicAdd(&T->elements, (unsigned long long)n_el);
            atomicAdd(&T->sum_return, ep_return);
            atomicAdd(&T->sum_length, (double)ep_len);
        }
    }
    if ((io.obs_full || (done && io.term_obs_out)) && lane < MG_OBS_DIM) {
        const float o = P.obs_cache[(size_t)env * MG_OBS_DIM + lane];
        if (io.obs_full) io.obs_out[(size_t)env * MG_OBS_DIM + lane] = o;
        if (done && io.term_obs_out) io.term_obs_out[(size_t)env * MG_OBS_DIM + lane] = o;
    }
    if (done && P.auto_reset) push_item(P.observe_list, cnt + CNT_OBSERVE, P, 0, make_item(env, 0, WORK_RESET, 0, 0, 0, 0.0, 0.0), lane);
    return false;
}

// ---- kernel 2: decide (only launched when the decide and update kernels are not fused) ----------------------------
#ifndef MG_MINB_DECIDE
#define MG_MINB_DECIDE 24
#endif
__global__ void __launch_bounds__(32, MG_MINB_DECIDE) mg_step_decide_kernel(const __grid_constant__ Params P, const __grid_constant__ StepIO io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const SmemLayout L = carve(smem_raw, P.cap, false);
    int *cnt = P.counters + CNT_SET * P.counters[CNT_CUR];
    int counts[DECIDE_SEGS];
    const int total = load_counts(cnt + CNT_DECIDE, P.num_envs, counts);
    const Stash S{L.stash};
    MG_ITEM_LOOP_BEGIN(total, &cnt[CNT_TICKET_DECIDE])
        const WorkItem W = fetch_item(P.decide_list, counts, P.num_envs, t_);
        __syncwarp();
        stage_issue(L.ring, L.mbar, P.xy + (size_t)W.env * P.cap, W.n, lane);
        stash_records(P, L.stash, W.env, lane);
        __syncwarp();
        Warp w;
        w.ring = L.ring; w.queue = nullptr; w.lane = lane; w.n = W.n;
        stage_wait(L.mbar, phase);
        phase ^= 1u;
        int rule; bool new_vertex;
        if (decide_item(P, io, cnt, w, W, S, rule, new_vertex, ticket_, next_t_))
            push_item(P.accept_list, cnt + CNT_ACCEPT, P, size_bin(W.n, P.cap),
                      make_item(W.env, W.n, WORK_APPLY, rule, new_vertex ? 1 : 0, 0, W.newx, W.newy), lane);
        else have_next_ = true;
    MG_ITEM_LOOP_END
}

// One accepted element (E:319-351 / move E:520-535): element log, update_boundary (M:601-674), candidate keys of the
// four neighbours (M:206-226) and -- for step(), REWARD = true -- area (C:943-958), robust quality (C:881-892),
// boundary quality (M:355-452), speed penalty (E:590-607).  The boundary is staged in w.ring (updated in place together
// with the global rows), the records in the stash.
struct ApplyOut {
    int nn, n_elements, next_vid0;
    double reward, current_area;
    bool done;
};
template <bool REWARD>
__device__ __forceinline__ ApplyOut apply_element(const Params &P, Warp &w, const Stash S, int env, int n, int idx, int rule,
                                                  bool new_vertex, P2 newp, int fan_vid) {
    const int lane = w.lane;
    const size_t off = (size_t)env * P.cap;
    int n_elements = S.i(W_NEL);
    double current_area = S.d(D_CUR_AREA);
    P2 m[4]; int qi[4]; int ri;
    quad_indices(rule, new_vertex, idx, n, qi, ri);
#pragma unroll
    for (int k = 0; k < 4; k++) m[k] = qi[k] < 0 ? newp : w.at(qi[k]);
    // the quad's four quantised corner angles (C:752, C:888, C:946-947), one per lane
    double corner[4];
    {
        double ca = 0;
        if (lane < 4) {
            P2 c = m[0], p1 = m[1], p2 = m[3];
            if (lane == 1) { c = m[1]; p1 = m[2]; p2 = m[0]; }
            if (lane == 2) { c = m[2]; p1 = m[3]; p2 = m[1]; }
            if (lane == 3) { c = m[3]; p1 = m[0]; p2 = m[2]; }
            ca = cw_angle(c, p1, p2);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) corner[k] = shfl_d(ca, k);
    }
    const int ip1 = qi[3], im1 = new_vertex ? qi[1] : (rule == -1 ? qi[0] : qi[1]);
    double reward = 0;
    bool done = false;

    // ---- update_boundary (M:601-674) ---------------------------------------------------
    int nb[4];          // the four neighbours whose candidate keys are re-evaluated, in order
    int t0 = 0, t1 = 0; // surviving quad vertices (no-new-vertex case), new indices
    const int next_vid0 = S.i(W_NEXT_VID);
    // ---- element log, area and robust quality first: they only need the quad (old ring + new vertex), and doing
    // them here ends the live ranges of the quad, its corner angles and the vertex ids before the boundary update
    {
        // quad vertex k is B[first + k] (the new vertex, id next_vid0, takes slot 0 of a new-vertex quad)
        const int first = new_vertex ? -2 : (rule == -1 ? -1 : -2);       // offset of quad slot 0 from the reference point
        const int my_id = __shfl_sync(FULL, fan_vid, (first + 2 + lane) & 7);
        if (P.elem && lane < 4 && n_elements < P.elem_cap)
            P.elem[((size_t)env * P.elem_cap + n_elements) * 4 + lane] = (new_vertex && lane == 0) ? next_vid0 : my_id;
    }
    n_elements++;
    // ---- area (C:943-958), robust quality (C:881-892) -----------------------------------
    double mesh_area = 0, e_reward = 0;
    if (REWARD) {
        double e0 = pdist(m[0], m[3]), e1 = pdist(m[1], m[0]), e2 = pdist(m[2], m[1]), e3 = pdist(m[3], m[2]);
        double sn0, sn2, cs_unused;
        sincos_quantised(P.sc_full, corner[0], false, sn0, cs_unused);
        sincos_quantised(P.sc_full, corner[2], false, sn2, cs_unused);
        mesh_area = 0.5 * e0 * e1 * sn0 + 0.5 * e2 * e3 * sn2;
        current_area -= mesh_area;
        double mn = fmin(fmin(e0, e1), fmin(e2, e3));
        double q1 = sqrt(2.0) * mn / fmax(pdist(m[0], m[2]), pdist(m[1], m[3]));
        double amin = fmin(fmin(corner[0], corner[1]), fmin(corner[2], corner[3]));
        double amax = fmax(fmax(corner[0], corner[1]), fmax(corner[2], corner[3]));
        e_reward = sqrt(q1 * (amin / amax));
    }
    __syncwarp();
    if (new_vertex) {
        // insert P at index(ref) and remove ref: the slot is replaced in place
        if (lane == 0) {
            w.ring[idx] = make_double2(newp.x, newp.y);
            P.xy[off + idx] = make_double2(newp.x, newp.y);
            P.vid[off + idx] = next_vid0;
            P.key[off + idx] = CUDART_INF;
            const int ins = next_vid0 - S.i(W_N0);
            if (P.ins_xy && ins < P.ins_cap) P.ins_xy[(size_t)env * P.ins_cap + ins] = make_double2(newp.x, newp.y);
        }
        __syncwarp();
        nb[0] = ip1; nb[1] = im1; nb[2] = wrapn(idx + 2, n); nb[3] = wrapn(idx - 2, n);
    } else {
        // remove the two middle quad vertices; compact ring + key/stamp/vid
        const int r0 = qi[1], r1 = qi[2];
        const int lo = r0 < r1 ? r0 : r1, hi = r0 < r1 ? r1 : r0;
        auto newpos = [&](int j) { return j - (j > lo ? 1 : 0) - (j > hi ? 1 : 0); };
#pragma unroll 1
        for (int base = lo; base < n; base += 128) {
            // every element moves left by at most 2, so a group only overwrites slots that it (or an
            // earlier group) has already read; 4 chunks of loads are in flight per DRAM round trip
            double2 v[4]; double k[4]; int st[4], id[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int j = base + 32 * c + lane;
                bool mv = j < n && j != lo && j != hi;
                v[c] = make_double2(0, 0); k[c] = 0; st[c] = 0; id[c] = 0;
                if (mv) { v[c] = w.ring[j]; k[c] = P.key[off + j]; st[c] = P.stamp[off + j]; id[c] = P.vid[off + j]; }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int j = base + 32 * c + lane;
                if (j < n && j != lo && j != hi) {
                    int q = newpos(j);
                    w.ring[q] = v[c]; P.xy[off + q] = v[c]; P.key[off + q] = k[c]; P.stamp[off + q] = st[c]; P.vid[off + q] = id[c];
                }
            }
            __syncwarp();
        }
        t0 = newpos(qi[0]); t1 = newpos(qi[3]);
        w.n = n - 2;
        const int nn = n - 2;
        const int id = t0 > t1 ? t0 : t1;
        nb[0] = id; nb[1] = id - 1 < 0 ? id - 1 + nn : id - 1; nb[2] = wrapn(id + 1, nn);
        nb[3] = wrapn(id - 2, nn);
    }
    const int nn = w.n;
    // ---- candidate keys of the four neighbours (lanes 2k, 2k+1 -> angles a0, a1 of nb[k]) ----
    double ang = 0;
    if (lane < 8) {
        int k = lane >> 1, v = nb[k], d = 1 + (lane & 1);
        ang = cw_angle(w.at(v), w.at(wrapn(v + d, nn)), w.at(wrapn(v - d, nn)));
    }
    double nb_a0[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        nb_a0[k] = shfl_d(ang, 2 * k);
        double a1v = shfl_d(ang, 2 * k + 1);
        double kv = cand_key_from_angles(nb_a0[k], a1v);
        bool later_dup = false;
#pragma unroll
        for (int k2 = k + 1; k2 < 4; k2++) later_dup |= nb[k2] == nb[k];
        if (lane == k && !later_dup) {
            P.key[off + nb[k]] = kv;
            P.stamp[off + nb[k]] = S.i(W_STAMP_CTR) - 1 - k;
        }
    }
    __syncwarp();
    if (REWARD) {
        // ---- boundary quality (M:410-452) ---------------------------------------------------
        double b_reward;
        if (new_vertex) b_reward = boundary_quality_new_vertex(w, idx, nb_a0[0], nb_a0[1]);
        else {
            double g0 = 0, g1 = 0;            // interior angles at the two survivors
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (nb[k] == t0) g0 = nb_a0[k];
                if (nb[k] == t1) g1 = nb_a0[k];
            }
            b_reward = boundary_quality_no_new(w, t0, t1, g0, g1);
        }
        double quality = e_reward + 1 * (b_reward - 1);              // M:1754-1766
        // ---- speed penalty (E:590-607) ------------------------------------------------------
        const double a_min = S.d(D_AREA_MIN), a_crit = S.d(D_AREA_CRIT);
        double min_area = a_min * a_min, crit = a_crit * a_crit, pen;
        if (min_area <= mesh_area && mesh_area < crit) pen = (mesh_area - crit) / (crit - min_area);
        else if (mesh_area < min_area) pen = -1;
        else pen = 0;
        reward += quality + pen;
    }
    if (nn <= 5) {                                       // E:345-351
        reward += 10; done = true;
        if (nn == 4) {
            if (P.elem && lane < 4 && n_elements < P.elem_cap)
                P.elem[((size_t)env * P.elem_cap + n_elements) * 4 + lane] = P.vid[off + lane];
            n_elements++;
        }
    }
    ApplyOut R;
    R.nn = nn; R.n_elements = n_elements; R.next_vid0 = next_vid0; R.reward = reward; R.current_area = current_area; R.done = done;
    return R;
}

// ---- kernel 3: update (fused mode: decide + update) ---------------------------------------------
#ifndef MG_MINB_UPDATE
#define MG_MINB_UPDATE 20
#endif
__global__ void __launch_bounds__(32, MG_MINB_UPDATE) mg_step_update_kernel(const __grid_constant__ Params P, const __grid_constant__ StepIO io, int fused) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    // Programmatic dependent launches (mg_set_option "pdl"): the blocks of a step kernel are placed while its predecessor
    // is still running -- as blocks of that grid exit -- and wait here for it to complete, instead of being launched after
    // it; both instructions are no-ops under a plain launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const SmemLayout L = carve(smem_raw, P.cap, false);
    int *cnt = P.counters + CNT_SET * P.counters[CNT_CUR];
    // fused mode: the decide list is processed here too (a warp that accepts an element applies it with the boundary it
    // has already staged); items are taken bin by bin, accepted elements before open decisions
    // Items are served kind by kind -- new-vertex candidates (the longest: decision + update), pending rule verdicts,
    // accepted rule elements -- and largest first within a kind: the warps that share an SM start on the same code.
    int counts_a[NBINS], counts_d[DECIDE_SEGS];
    const int total_a = load_counts(cnt + CNT_ACCEPT, P.num_envs, counts_a);
    int total_d = 0;
#pragma unroll
    for (int b = 0; b < DECIDE_SEGS; b++) counts_d[b] = 0;
    if (fused) total_d = load_counts(cnt + CNT_DECIDE, P.num_envs, counts_d);
    const int total = total_d + total_a;
    const Stash S{L.stash};
    MG_ITEM_LOOP_BEGIN(total, &cnt[CNT_TICKET_UPDATE])
        MG_TRACE_T0
        const WorkItem W = t_ < total_d ? fetch_item(P.decide_list, counts_d, P.num_envs, t_)
                                        : fetch_item(P.accept_list, counts_a, P.num_envs, t_ - total_d);
        const int env = W.env;
        const size_t off = (size_t)env * P.cap;
        __syncwarp();
        stage_issue(L.ring, L.mbar, P.xy + off, W.n, lane);
        // Both records are parked in this warp's shared stash and only the fields the item works on are kept in
        // registers; everything else comes back from shared memory where it is used (carried in registers it was
        // spilled to local memory, an L2 round trip with the shared-memory carve-out at its maximum).
        stash_records(P, L.stash, env, lane);
#ifndef MG_NO_ROW_PREFETCH
        // the key / stamp / id rows are only touched on accepted elements, so they come from DRAM: pull them into L2
        // now, behind the ring copy, instead of paying the round trips in the compaction loop
        {
            const char *kp = reinterpret_cast<const char *>(P.key + off), *sp = reinterpret_cast<const char *>(P.stamp + off),
                       *vp = reinterpret_cast<const char *>(P.vid + off);
            for (int b = lane * 128; b < W.n * 8; b += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + b));
            for (int b = lane * 128; b < W.n * 4; b += 32 * 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + b));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + b));
            }
        }
#endif
        __syncwarp();
        Warp w;
        w.ring = L.ring; w.queue = nullptr; w.lane = lane; w.n = W.n;
        const int n = W.n, idx = S.i(W_REF);
        // ids of the five vertices any of the three quads is made of (B[i-2..i+2]), requested now so that the DRAM
        // round trip hides behind the ring copy and the decision instead of sitting in front of the element log
        const int fan_vid = lane < 5 ? P.vid[off + wrapn(idx - 2 + lane, n)] : 0;
        stage_wait(L.mbar, phase);
        phase ^= 1u;
        int rule = W.rule;
        bool new_vertex = W.flag != 0;
        if (W.kind != WORK_APPLY) {
            if (!decide_item(P, io, cnt, w, W, S, rule, new_vertex, ticket_, next_t_)) {
                MG_TRACE_ITEM(1, 10 + W.kind, n)
                t_ = __shfl_sync(FULL, next_t_, 0);
                continue;
            }
        }
        const ApplyOut A = apply_element<true>(P, w, S, env, n, idx, rule, new_vertex, mk(W.newx, W.newy), fan_vid);
        MG_ITEM_TICKET
        const int nn = A.nn, n_elements = A.n_elements, next_vid0 = A.next_vid0;
        const double reward = A.reward, current_area = A.current_area;
        const bool done = A.done;
        // ---- results of the step; the next observation follows in the observe kernel -----------------------
        const double ep_return = S.d(D_EP_RETURN) + reward;
        const int ep_len = S.i(W_EP_LEN) + 1;
        if (lane == 0) {
            StatsAcc *T = P.stats + (env & (STAT_SLOTS - 1));
            atomicAdd(&T->successes, 1ull);
            atomicAdd(&T->sum_n_success, (unsigned long long)n);
            if (done) {
                atomicAdd(&cnt[CNT_DONE], 1);
                atomicAdd(&T->episodes, 1ull);
                atomicAdd(&T->completed, 1ull);
                atomicAdd(&T->elements, (unsigned long long)n_elements);
                atomicAdd(&T->sum_return, ep_return);
                atomicAdd(&T->sum_length, (double)ep_len);
            }
            const int wb = emit_results(io, env, reward, done, false, n_elements);
            if (wb) atomicAdd(&cnt[CNT_RESBYTES], wb);
            // the part of the records this kernel owns (reference index, base length, flags and fan follow in observe)
            int4 *dst = reinterpret_cast<int4 *>(P.hot + env);
            const int4 c1 = S.p[1];
            dst[0] = make_int4(nn, -1, n_elements, 0);
            dst[1] = make_int4(c1.x, c1.y, 0, ep_len);
            dst[2] = make_int4(__double2loint(ep_return), __double2hiint(ep_return), __double2loint(current_area),
                               __double2hiint(current_area));
            P.cold[env].next_vid = next_vid0 + (new_vertex ? 1 : 0);
            P.cold[env].stamp_ctr = S.i(W_STAMP_CTR) - 4;
        }
        // (a completed episode is followed by the env's in-place reset in the same observe item: the longest kind, first bin)
        push_item(P.observe_list, cnt + CNT_OBSERVE, P, done && P.auto_reset ? 0 : size_bin(nn, P.cap),
                  make_item(env, nn, WORK_OBSERVE, 0, 0, done ? 1 : 0, 0.0, 0.0), lane);
        MG_TRACE_ITEM(1, W.kind, n)
    MG_ITEM_LOOP_END
}

// ---- kernel 4: observe (+ resets) ---------------------------------------------------------------
// in-place reset of one finished env (V:40-52): new episode, first observation to the caller
__device__ __noinline__ void reset_in_place(const Params &P, const StepIO &io, int env, Warp w) {
    const int domain = P.cold[env].domain, episode = P.cold[env].episode + 1;
    const float obs = reset_env(P, w, env, domain, episode);
    if (w.lane < MG_OBS_DIM) {
        P.obs_cache[(size_t)env * MG_OBS_DIM + w.lane] = obs;
        io.obs_out[(size_t)env * MG_OBS_DIM + w.lane] = obs;
    }
    __syncwarp();
}

#ifndef MG_MINB_OBSERVE
#define MG_MINB_OBSERVE 20
#endif
// One observe / reset item; the env's records are (re)loaded here.
__device__ __forceinline__ void observe_item(const Params &P, const StepIO &io, int *cnt, const SmemLayout &L, const WorkItem &W, unsigned &phase,
                                             int *ticket, int &next_t, int lane) {
    const Stash S{L.stash};
    const int env = W.env;
    Warp w;
    w.ring = L.ring; w.queue = L.queue; w.lane = lane; w.n = 0;
    if (W.kind == WORK_RESET) {
        reset_in_place(P, io, env, w);
        if (lane == 0) next_t = atomicAdd(ticket, 1);
        return;
    }
    const size_t off = (size_t)env * P.cap;
    __syncwarp();
    stage_issue(L.ring, L.mbar, P.xy + off, W.n, lane);
    stash_records(P, L.stash, env, lane);
    __syncwarp();
    w.n = W.n;
    stage_wait(L.mbar, phase);
    phase ^= 1u;
    // ---- next state (E:361-386) -----------------------------------------------------
    const int ref_index = find_reference_index(w, P.key + off, P.stamp + off);
    float obs = 0.0f;
    double base = S.d(D_BASE);
    if (ref_index >= 0) {
        const ObsOut R = compute_obs(w, P.sc_full, ref_index, S.d(D_CUR_AREA) / S.d(D_ORIGINAL_AREA));
        obs = R.obs; base = R.base;
    }
    // An env without a reference point (empty candidate list, E:736-738) is reported truncated (sentinel)
    const bool truncated = !W.done && ref_index < 0;
    const bool done = W.done || truncated;
    const bool reset_follows = done && P.auto_reset;
    if (!reset_follows && lane == 0) next_t = atomicAdd(ticket, 1);      // the block's next item (see MG_ITEM_TICKET)
    if (truncated && lane == 0) {
        io.trunc_out[env] = 1;
        if (io.trunc_sh) { io.trunc_sh[env] = 1; atomicAdd(&cnt[CNT_RESBYTES], 1); }
        StatsAcc *T = P.stats + (env & (STAT_SLOTS - 1));
        atomicAdd(&cnt[CNT_DONE], 1);
        atomicAdd(&T->episodes, 1ull);
        atomicAdd(&T->truncated, 1ull);
        atomicAdd(&T->elements, (unsigned long long)S.i(W_NEL));
        atomicAdd(&T->sum_return, S.d(D_EP_RETURN));
        atomicAdd(&T->sum_length, (double)S.i(W_EP_LEN));
    }
    if (lane < MG_OBS_DIM) {
        if (io.term_obs_out && done) io.term_obs_out[(size_t)env * MG_OBS_DIM + lane] = obs;
        P.obs_cache[(size_t)env * MG_OBS_DIM + lane] = obs;
        io.obs_out[(size_t)env * MG_OBS_DIM + lane] = obs;
    }
    if (reset_follows) {
        __syncwarp();
        reset_in_place(P, io, env, w);
        if (lane == 0) next_t = atomicAdd(ticket, 1);
    } else {
        // the state changed: new memo (rule -1 / +1 verdicts, neighbour fan) and the rest of the record
        const int flags = done ? 0 : memo_flags(w, ref_index);
        store_hot(P.hot + env, w, ref_index, S.i(W_NEL), flags, base, 0, S.i(W_EP_LEN), S.d(D_EP_RETURN), S.d(D_CUR_AREA));
    }
}

__global__ void __launch_bounds__(32, MG_MINB_OBSERVE) mg_step_observe_kernel(const __grid_constant__ Params P, const __grid_constant__ StepIO io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const SmemLayout L = carve(smem_raw, P.cap, true);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // no-op unless launched as a programmatic dependent
    const int set = P.counters[CNT_CUR];                   // written by the screen kernel of this step
    int *cnt = P.counters + CNT_SET * set;
    if (blockIdx.x == 0 && lane == 0) {                    // the other counter set is idle during this step
        int *idle = P.counters + CNT_SET * (set ^ 1);
#pragma unroll
        for (int k = 0; k < CNT_SET; k++) idle[k] = 0;
        P.counters[CNT_STEP] = (set ^ 1);                  // parity of the next step; only the screen kernel reads it
    }
    int counts[NBINS];
    const int total = load_counts(cnt + CNT_OBSERVE, P.num_envs, counts);
    MG_ITEM_LOOP_BEGIN(total, &cnt[CNT_TICKET_OBSERVE])
        MG_TRACE_T0
        const WorkItem W = fetch_item(P.observe_list, counts, P.num_envs, t_);
        observe_item(P, io, cnt, L, W, phase, ticket_, next_t_, lane);
        have_next_ = true;
        MG_TRACE_ITEM(2, W.kind, W.n)
    MG_ITEM_LOOP_END
}

// ---- kernel 5: resets of the envs the screen kernel truncated ------------------------------------------------------
// Runs on a side stream next to the update kernel: a reset in random-polygon mode (generator, candidate rebuild, first
// observation) is the longest item of a step (~40 us), touches only its own env and needs nothing the update / observe
// kernels produce.  (Inside the observe kernel the resets alone stretched that launch to ~52 us for ~30 us of work.)
__global__ void __launch_bounds__(32, MG_MINB_OBSERVE) mg_step_reset_kernel(const __grid_constant__ Params P, const __grid_constant__ StepIO io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const SmemLayout L = carve(smem_raw, P.cap, true);
    int *cnt = P.counters + CNT_SET * P.counters[CNT_CUR];
    const int total = min(cnt[CNT_RESET], P.num_envs);
    MG_ITEM_LOOP_BEGIN(total, &cnt[CNT_TICKET_RESET])
        MG_TRACE_T0
        const int env = P.reset_list[t_];
        Warp w;
        w.ring = L.ring; w.queue = L.queue; w.lane = lane; w.n = 0;
        reset_in_place(P, io, env, w);
        MG_TRACE_ITEM(3, WORK_RESET, 0)
    MG_ITEM_LOOP_END
}

// ---------------------------------------------------------------------------------------------
// BoudaryEnv.move() (E:459-594; legacy rl/boundary_env.py:265-432): the deterministic "apply this geometric move"
// entry point of the data-generation utilities (general/EBRD.py, FNN_evaluation.py).  Differences from step():
// polar action (r, phi) in units of radius * base_length rounded to 6 decimals by CPython's round, rule selection by
// TYPE_THRESHOLD = 0.3, no find_same_point fallback, reward 0, no failed-step counter; a failed move puts the
// reference point on the not-valid list and the next reference point is the first candidate that is not within 0.001
// of a listed point (M:310-314, M:428-433); point environments are static (area-ratio slot of the observation = 0,
// C:1209-1214).  One warp per env, one launch: this is not a throughput path.
// When every candidate is excluded the reference smooths the whole mesh (smooth_pave, general/mesh.py:790-1067) and
// goes on: this kernel reports such an env as done with `exhausted` set, and mg_move then runs mg_smooth_kernel on it.
// ---------------------------------------------------------------------------------------------
// rint(x * 1e6) / 1e6 as CPython's round(x, 6) picks it (see py_rint4)
__device__ __forceinline__ double py_round6(double x) {
    double p = x * 1e6;
    double r = rint(p);
    if (fabs(p - r) == 0.5) {
        double e = __fma_rn(x, 1e6, -p);
        if (e > 0) r = floor(p) + 1.0;
        else if (e < 0) r = floor(p);
    }
    return r / 1e6;
}

// arg-min of (key, stamp) over the candidates that are not within 0.001 of a not-valid point; -1 when none is left
__device__ __noinline__ int find_reference_index_excl(const Warp w, const double *key, const int32_t *stamp, const double2 *excl, int nexcl) {
    double bk = CUDART_INF;
    int bs = 0x7fffffff, bj = -1;
#pragma unroll 1
    for (int j = w.lane; j < w.n; j += 32) {
        const double k = key[j];
        if (k == CUDART_INF) continue;
        const P2 v = w.at(j);
        bool hit = false;
#pragma unroll 1
        for (int q = 0; q < nexcl && !hit; q++) {
            const double2 e = excl[q];
            hit = pdist(mk(e.x, e.y), v) < 0.001;
        }
        if (hit) continue;
        const int s = stamp[j];
        if (k < bk || (k == bk && s < bs)) { bk = k; bs = s; bj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ok = __shfl_xor_sync(FULL, bk, o);
        int os = __shfl_xor_sync(FULL, bs, o);
        int oj = __shfl_xor_sync(FULL, bj, o);
        if (ok < bk || (ok == bk && ok != CUDART_INF && os < bs)) { bk = ok; bs = os; bj = oj; }
    }
    return bk == CUDART_INF ? -1 : bj;
}

struct MoveIO {
    const double *polar;     // [N][2] (r, phi)
    const double *type;      // [N]
    float *obs_out;          // [N][18]
    uint8_t *done_out, *complete_out, *exhausted_out;
    int32_t *n_elem_out;
};

__global__ void __launch_bounds__(32) mg_move_kernel(const __grid_constant__ Params P, const __grid_constant__ MoveIO io, double2 *excl_all,
                                                     int32_t *excl_id_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int env = blockIdx.x;
    if (env >= P.num_envs) return;
    const SmemLayout L = carve(smem_raw, P.cap, true);
    const Stash S{L.stash};
    const size_t off = (size_t)env * P.cap;
    double2 *excl = excl_all + off;
    stash_records(P, L.stash, env, lane);
    __syncwarp();
    Warp w;
    w.ring = L.ring; w.queue = L.queue; w.lane = lane; w.n = S.i(W_N);
    const int n = w.n, idx = S.i(W_REF);
    int nexcl = reinterpret_cast<const int32_t *>(L.stash)[43];          // EnvCold::pad[0]
    int n_elements = S.i(W_NEL);
    auto finish = [&](float obs, bool done, bool complete, bool exhausted) {
        if (lane < MG_OBS_DIM) io.obs_out[(size_t)env * MG_OBS_DIM + lane] = obs;
        if (lane == 0) {
            io.done_out[env] = done; io.complete_out[env] = complete; io.exhausted_out[env] = exhausted;
            if (io.n_elem_out) io.n_elem_out[env] = n_elements;
        }
    };
    if (idx < 0 || n < 3) {                 // no reference point: the reference would have raised before this call
        finish(0.0f, true, false, true);
        return;
    }
    if (n <= 5) {                           // E:484-486 (the reference leaves is_complete unbound here and raises)
        float o = lane < MG_OBS_DIM ? P.obs_cache[(size_t)env * MG_OBS_DIM + lane] : 0.0f;
        if (lane == 1) o = 0.0f;
        finish(o, true, n <= 4, false);
        return;
    }
    for (int j = lane; j < n; j += 32) w.ring[j] = P.xy[off + j];
    __syncwarp();
    const P2 ref = w.at(idx), right_p = w.at(idx - 1);
    // ---- polar action -> candidate vertex (E:467-478, E:202-210 with dist = 1, D:112-137) -------------------------
    const double base_length = S.d(D_BASE);
    const double pr = io.polar[2 * (size_t)env], phi = io.polar[2 * (size_t)env + 1], type = io.type[env];
    P2 newp;
    {
        const double2 sc = mg_sincos(phi);
        const double px = py_round6(base_length * 4 * pr * sc.y), py = py_round6(base_length * 4 * pr * sc.x);
        const double2 o = action_frame_exact(px, py, right_p.x - ref.x, right_p.y - ref.y, 1.0, ref);
        newp = mk(np_round4(o.x), np_round4(o.y));
    }
    // ---- the element (E:490-520) --------------------------------------------------------------------------------
    int rule = 0;
    bool new_vertex = false, have_mesh = true;
    if (type <= 0.3) rule = -1;
    else if (type >= 1 - 0.3) rule = 1;
    else if (point_inside(w, newp, P.vid + off, S.i(W_N0))) new_vertex = true;
    else have_mesh = false;
    bool accepted = false;
    if (have_mesh) {
        int qi[4], ri;
        quad_indices(rule, new_vertex, idx, n, qi, ri);
        Quad Q;
#pragma unroll
        for (int k = 0; k < 4; k++) { const P2 p = qi[k] < 0 ? newp : w.at(qi[k]); Q.x[k] = p.x; Q.y[k] = p.y; }
        accepted = mesh_is_valid(w, Q) && !intersects_boundary(w, Q, make_int4(qi[0], qi[1], qi[2], qi[3]), ri, ref);
    }
    bool done = false;
    int next_vid = S.i(W_NEXT_VID), stamp_ctr = S.i(W_STAMP_CTR);
    if (accepted) {
        const int fan_vid = lane < 5 ? P.vid[off + wrapn(idx - 2 + lane, n)] : 0;
        const ApplyOut A = apply_element<false>(P, w, S, env, n, idx, rule, new_vertex, newp, fan_vid);
        n_elements = A.n_elements;
        done = A.done;
        next_vid += new_vertex ? 1 : 0;
        stamp_ctr -= 4;
    } else {
        if (lane == 0 && nexcl < P.cap) {                                               // E:538-540
            excl[nexcl] = make_double2(ref.x, ref.y);
            excl_id_all[off + nexcl] = P.vid[off + idx];                                // the vertex itself (E:570-576 compares identities)
        }
        nexcl++;
    }
    __syncwarp();
    __threadfence_block();
    // ---- next state with the not-valid points collected so far (E:527 / E:541), static point environment --------
    const int ref_index = find_reference_index_excl(w, P.key + off, P.stamp + off, excl, nexcl < P.cap ? nexcl : P.cap);
    if (accepted) nexcl = 0;                                                             // E:542-543
    float obs = 0.0f, obs_cache = 0.0f;
    double base = base_length;
    if (ref_index >= 0) {
        const ObsOut R = compute_obs(w, P.sc_full, ref_index, S.d(D_CUR_AREA) / S.d(D_ORIGINAL_AREA));
        obs_cache = R.obs; base = R.base;
        obs = lane == 1 ? 0.0f : R.obs;                                                  // static: area-ratio slot = 0
    }
    const bool exhausted = ref_index < 0 && w.n > 4;          // the reference would call smooth_pave here (E:548-575)
    done = done || exhausted;
    if (lane < MG_OBS_DIM) P.obs_cache[(size_t)env * MG_OBS_DIM + lane] = obs_cache;    // step() afterwards: non-static
    const int flags = done ? 0 : memo_flags(w, ref_index);
    store_hot(P.hot + env, w, ref_index, n_elements, flags, base, S.i(W_FAILED), S.i(W_EP_LEN), S.d(D_EP_RETURN), S.d(D_CUR_AREA));
    if (lane == 0) {
        P.cold[env].next_vid = next_vid;
        P.cold[env].stamp_ctr = stamp_ctr;
        P.cold[env].pad[0] = nexcl;
    }
    finish(obs, done, w.n <= 4, exhausted);
}

// The rest of move() for the envs mg_move_kernel reported as exhausted (E:548-583): smooth_pave (mg_smooth.cuh), the
// last_not_valid_points rule, an empty not-valid list, the next state.  One warp per listed env.  An env this kernel cannot smooth (log overflow, a vertex with more
// than SM_MAXDEG segments, a construction on which the reference raises) keeps its exhausted / done flags.
__global__ void __launch_bounds__(32) mg_smooth_kernel(const __grid_constant__ Params P, const __grid_constant__ MoveIO io, const int32_t *env_list,
                                                       unsigned char *scratch_all, size_t scratch_bytes, int32_t *excl_id_all,
                                                       int32_t *last_id_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int env = env_list[blockIdx.x];
    const SmemLayout L = carve(smem_raw, P.cap, true);
    const Stash S{L.stash};
    const size_t off = (size_t)env * P.cap;
    stash_records(P, L.stash, env, lane);
    __syncwarp();
    const int n = S.i(W_N), n0 = S.i(W_N0), nv = S.i(W_NEXT_VID), n_elements = S.i(W_NEL);
    const int domain = reinterpret_cast<const int32_t *>(L.stash)[41];     // EnvCold::domain
    int nexcl = reinterpret_cast<const int32_t *>(L.stash)[43], nlast = reinterpret_cast<const int32_t *>(L.stash)[44];   // pad[0], pad[1]
    int failed = 0;
    // random-polygon mode: the episode's original polygon is not stored -- regenerate it from the generator's counter
    // (as mg_get_elements does) into the warp's ring; lane 0 copies it into the vertex pool below
    int n_gen = n0;
    if (P.random_mode) {
        Warp g;
        g.ring = L.ring; g.queue = L.queue; g.lane = lane; g.n = 0;
        n_gen = generate_polygon(P, g, P.env_id_offset + env, reinterpret_cast<const int32_t *>(L.stash)[42]);   // EnvCold::episode
        __syncwarp();
    }
    if (lane == 0) {
        SmoothScratch M;
        const size_t V = (size_t)P.cap + P.ins_cap;
        unsigned char *base = scratch_all + (size_t)blockIdx.x * scratch_bytes;
        M.pool = reinterpret_cast<double2 *>(base);
        M.deg = reinterpret_cast<int *>(base + V * 16);
        M.adj = M.deg + V;
        M.front = M.adj + V * SM_MAXDEG;
        M.onfront = reinterpret_cast<uint8_t *>(M.front + P.cap);
        M.n = n; M.n0 = n0; M.nv = nv; M.failed = false;
        if (n_gen != n0 || nv - n0 > P.ins_cap || n_elements > P.elem_cap || P.elem == nullptr || P.ins_xy == nullptr || n < 5) M.failed = true;
        if (!M.failed) {
            // vertex pool: the episode's original polygon (never moved) + the inserted vertices at their current positions
            for (int id = 0; id < n0; id++) M.pool[id] = P.random_mode ? L.ring[id] : P.t_xy[(size_t)domain * P.cap + id];
            for (int id = n0; id < nv; id++) M.pool[id] = P.ins_xy[(size_t)env * P.ins_cap + id - n0];
            for (int id = 0; id < nv; id++) { M.deg[id] = 0; M.onfront[id] = 0; }
            // Segments in the reference's creation order: Boundary2D.deep_copy (C:221-228), then Mesh.connect_vertices of
            // every element (C:840-845)
            for (int i = 0; i < n0; i++) sm_connect(M, (i + n0 - 1) % n0, i);
            for (int e = 0; e < n_elements && !M.failed; e++) {
                const int32_t *q = P.elem + ((size_t)env * P.elem_cap + e) * 4;
                for (int i = 0; i < 4; i++) {
                    const int a = q[i], b = q[(i + 3) & 3];
                    if (a < 0 || a >= nv || b < 0 || b >= nv) { M.failed = true; break; }
                    if (!sm_has(M, a, b)) sm_connect(M, a, b);
                }
            }
            for (int i = 0; i < n; i++) {
                const int id = P.vid[off + i];
                M.front[i] = id;
                if (id < 0 || id >= nv) M.failed = true; else M.onfront[id] = 1;
            }
        }
        if (!M.failed) sm_smooth_front(M);
        if (!M.failed) sm_smooth_interior(M, 400);
        if (!M.failed) {
            for (int i = 0; i < n; i++) P.xy[off + i] = M.pool[M.front[i]];
            for (int id = n0; id < nv; id++) P.ins_xy[(size_t)env * P.ins_cap + id - n0] = M.pool[id];
        }
        failed = M.failed ? 1 : 0;
        __threadfence_block();
    }
    failed = __shfl_sync(FULL, failed, 0);
    if (failed) return;                                     // exhausted / done stay set
    // ---- find_reference_candidates(0), the last_not_valid_points rule, the next state (E:570-583) ----------------------
    Warp w;
    w.ring = L.ring; w.queue = L.queue; w.lane = lane; w.n = n;
    for (int j = lane; j < n; j += 32) w.ring[j] = P.xy[off + j];
    __syncwarp();
    rebuild_candidates(w, P.key + off, P.stamp + off);
    __syncwarp();
    __threadfence_block();
    bool done = false;
    const int32_t *cur = excl_id_all + off;
    int32_t *last = last_id_all + off;
    const int ne = nexcl < P.cap ? nexcl : P.cap, nl = nlast < P.cap ? nlast : P.cap;
    if (nl > 0 && ne > 0 && last[0] == cur[0] && last[nl - 1] == cur[ne - 1] && nexcl == nlast) done = true;
    __syncwarp();
    for (int j = lane; j < ne; j += 32) last[j] = cur[j];
    nlast = nexcl;
    nexcl = 0;
    const int ref_index = find_reference_index(w, P.key + off, P.stamp + off);
    float obs = 0.0f, obs_cache = 0.0f;
    double base = S.d(D_BASE);
    if (ref_index >= 0) {
        const ObsOut R = compute_obs(w, P.sc_full, ref_index, S.d(D_CUR_AREA) / S.d(D_ORIGINAL_AREA));
        obs_cache = R.obs; base = R.base;
        obs = lane == 1 ? 0.0f : R.obs;                      // static point environment
    } else done = true;                                      // E:582-583
    if (lane < MG_OBS_DIM) {
        P.obs_cache[(size_t)env * MG_OBS_DIM + lane] = obs_cache;
        io.obs_out[(size_t)env * MG_OBS_DIM + lane] = obs;
    }
    const int flags = done ? 0 : memo_flags(w, ref_index);
    store_hot(P.hot + env, w, ref_index, n_elements, flags, base, S.i(W_FAILED), S.i(W_EP_LEN), S.d(D_EP_RETURN), S.d(D_CUR_AREA));
    if (lane == 0) {
        P.cold[env].stamp_ctr = 0;
        P.cold[env].pad[0] = nexcl;
        P.cold[env].pad[1] = nlast;
        io.done_out[env] = done; io.complete_out[env] = 0; io.exhausted_out[env] = 0;
    }
}

// Uniform actions in Box([-1,-1.5,0],[1,1.5,1.5]) -- the synthetic policy of the benchmarks.
// step_ctr != nullptr: the step index is step_ctr[0] in device memory and the last block to finish advances it
// (step_ctr[1] counts the blocks that are done), so that the launch can be replayed from a CUDA graph.
__global__ void mg_sample_actions_kernel(int num_envs, uint64_t seed, uint64_t step, int64_t env_id_offset, float *act,
                                         unsigned long long *step_ctr) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (step_ctr) {
        step = *reinterpret_cast<volatile unsigned long long *>(step_ctr);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(step_ctr + 1, 1ull) == gridDim.x - 1) {
                step_ctr[1] = 0;
                step_ctr[0] = step + 1;
            }
        }
    }
    if (e >= num_envs) return;
    unsigned long long g = (unsigned long long)(env_id_offset + e);
    uint4 r = philox4x32(make_uint4((unsigned)step, (unsigned)(step >> 32), (unsigned)g, (unsigned)(g >> 32)),
                         make_uint2((unsigned)seed, (unsigned)(seed >> 32) ^ 0xA5A5A5A5u));
    const float lo[3] = {-1.0f, -1.5f, 0.0f}, hi[3] = {1.0f, 1.5f, 1.5f};
    unsigned rr[3] = {r.x, r.y, r.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float u = (float)(rr[k] >> 8) * (1.0f / 16777216.0f);
        act[(size_t)e * 3 + k] = lo[k] + (hi[k] - lo[k]) * u;
    }
}

// mg_replay_add: the N transitions of one step into slot `slot` of the replay ring (coalesced: one thread per
// observation element, the first threads of every env also move the action / reward / flags).
__global__ void mg_replay_add_kernel(int num_envs, float *__restrict__ b_obs, float *__restrict__ b_next, float *__restrict__ b_act,
                                     float *__restrict__ b_rew, uint8_t *__restrict__ b_done, uint8_t *__restrict__ b_to,
                                     const float *__restrict__ prev_obs, const float *__restrict__ act,
                                     const float *__restrict__ new_obs, const double *__restrict__ rew,
                                     const uint8_t *__restrict__ term, const uint8_t *__restrict__ trunc,
                                     const float *__restrict__ term_obs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_envs * MG_OBS_DIM) return;
    const int e = t / MG_OBS_DIM, k = t - e * MG_OBS_DIM;
    const bool te = term[e] != 0, tr = trunc[e] != 0, done = te || tr;
    b_obs[t] = prev_obs[t];
    b_next[t] = done ? term_obs[t] : new_obs[t];
    if (k < MG_ACT_DIM) b_act[e * MG_ACT_DIM + k] = act[e * MG_ACT_DIM + k];
    if (k == 3) b_rew[e] = (float)rew[e];
    if (k == 4) b_done[e] = done ? 1 : 0;
    if (k == 5) b_to[e] = tr ? 1 : 0;
}

// Sum of the accumulator slots -> one mg_episode_stats (one warp).
__global__ void mg_stats_kernel(StatsAcc *stats, mg_episode_stats *out, int reset) {
    unsigned long long a[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double r = 0, l = 0;
    for (int e = threadIdx.x; e < STAT_SLOTS; e += 32) {
        StatsAcc T = stats[e];
        a[0] += T.episodes; a[1] += T.completed; a[2] += T.truncated; a[3] += T.steps; a[4] += T.successes;
        a[5] += T.elements; a[6] += T.sum_n; a[7] += T.sum_n_success; a[8] += T.ring_items; a[9] += T.sum_n_ring;
        r += T.sum_return; l += T.sum_length;
        if (reset) { StatsAcc Z = {}; stats[e] = Z; }
    }
#pragma unroll
    for (int k = 0; k < 10; k++)
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(FULL, a[k], o);
    r = warp_sum_d(r); l = warp_sum_d(l);
    if (threadIdx.x == 0) {
        out->episodes = (long long)a[0]; out->completed = (long long)a[1]; out->truncated = (long long)a[2];
        out->steps = (long long)a[3]; out->successes = (long long)a[4]; out->elements = (long long)a[5];
        out->sum_n = (long long)a[6]; out->sum_n_success = (long long)a[7];
        out->ring_items = (long long)a[8]; out->sum_n_ring = (long long)a[9];
        out->sum_return = r; out->sum_length = l;
    }
}

// ---------------------------------------------------------------------------------------------
// random star polygon generator (BASELINE configs 3/4)
// ---------------------------------------------------------------------------------------------
// Semantics of ui/GenerateRandomPolygon.py:5-49 (defaults :63) followed by the uniform-density
// densifier of ui/tk-ui.py:252-276; the Python RNG stream is not reproducible on a GPU
// (SURVEY.md row 7), the distribution is:
//   K ~ U{min_coarse..max_coarse} coarse vertices, angular steps ~ U(2pi/K -+ irr*2pi/K)
//   normalised to 2pi, start angle ~ U(0, 2pi), radius ~ clip(N(ave, spike*ave), 0.2 ave, 2 ave)
//   (floor 0.2 ave instead of 0: the reference crashes on zero-length edges),
//   integer pixel coordinates by int() truncation, consecutive duplicates nudged apart, order
//   reversed to clockwise.  Densifier with spacing A = perimeter / target, target ~ U{min..max}:
//   coarse edge prev->cur of length L gets x = round((2L - 2A) / 2A) interior points at
//   prev + A (j+1) dir, j < x, followed by cur; if the total is odd the middle point of the last
//   edge is dropped (tk-ui.py:267-269).  Coordinates / 100 (geometry.py:46).
// One coarse vertex per lane (max_coarse <= 32).  The ring is written to w.ring[0..n).
__device__ __noinline__ int generate_polygon(const Params &P, const Warp w, long long global_env, int episode, double *dbg) {
    const mg_polygen_cfg &G = P.gen;
    const int lane = w.lane;
    int *cx = w.queue, *cy = w.queue + 32, *cnt = w.queue + 64, *offs = w.queue + 96;
    const uint2 key = make_uint2((unsigned)P.seed, (unsigned)(P.seed >> 32));
    const unsigned long long g = (unsigned long long)global_env;
    auto draw = [&](unsigned slot) {
        return philox4x32(make_uint4((unsigned)g, (unsigned)(g >> 32), (unsigned)episode, slot), key);
    };
    const uint4 r0 = draw(0);
    int K = G.min_coarse + (int)(u01(r0.x, r0.y) * (double)(G.max_coarse - G.min_coarse + 1));
    K = min(K, G.max_coarse);
    const double start = 2 * PI * u01(r0.z, r0.w);
    const double irr = fmin(fmax(G.irregularity, 0.0), 1.0) * 2 * PI / K;
    const double spike = fmin(fmax(G.spikeyness, 0.0), 1.0) * G.ave_radius;
    const double lower = 2 * PI / K - irr, upper = 2 * PI / K + irr;
    double step = 0, radius = 0;
    if (lane < K) {
        uint4 r = draw(1 + lane);
        step = lower + (upper - lower) * u01(r.x, r.y);
        uint4 r2 = draw(65 + lane);
        double u1 = u01(r2.x, r2.y), u2 = u01(r2.z, r2.w);
        double gs = sqrt(-2.0 * log(1.0 - u1)) * cos(2 * PI * u2);      // Box-Muller
        radius = fmin(fmax(G.ave_radius + gs * spike, 0.2 * G.ave_radius), 2 * G.ave_radius);
    }
    const double ksum = warp_sum_d(step) / (2 * PI);
    step = step / ksum;
    double incl = step;                                // inclusive prefix of the normalised steps
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane < K) {
        double ang = start + (incl - step);
        const double2 sc = mg_sincos(ang);
        const double s = sc.x, c = sc.y;
        cx[lane] = (int)(G.ctr_x + radius * c);
        cy[lane] = (int)(G.ctr_y + radius * s);
    }
    __syncwarp();
    if (lane == 0) {                                   // distinct consecutive vertices
        for (int i = 0; i < K; i++) {
            int p = i == 0 ? K - 1 : i - 1;
            if (cx[i] == cx[p] && cy[i] == cy[p]) cx[i] += 1;
        }
        if (cx[0] == cx[K - 1] && cy[0] == cy[K - 1]) cy[0] += 1;
    }
    __syncwarp();
    // clockwise order: c(e) = coarse[K-1-e]; edge e runs from c(e-1) to c(e)
    double L = 0, dx = 0, dy = 0, pxv = 0, pyv = 0, cxv = 0, cyv = 0;
    if (lane < K) {
        int cur = K - 1 - lane, prv = lane == 0 ? 0 : K - lane;          // c(e-1) = coarse[K-e], c(-1) = coarse[0]
        pxv = cx[prv]; pyv = cy[prv]; cxv = cx[cur]; cyv = cy[cur];
        dx = cxv - pxv; dy = cyv - pyv;
        L = sqrt(dx * dx + dy * dy);
    }
    const double perim = warp_sum_d(L);
    const uint4 r3 = draw(200);
    const int maxv = min(G.max_verts, P.cap);
    int target = G.min_verts + (int)(u01(r3.x, r3.y) * (double)(maxv - G.min_verts + 1));
    target = min(target, maxv);
    double A = perim / target;
    int c = 0, total = 0;
    for (int it = 0; it < 16; it++) {
        c = 0;
        if (lane < K) {
            double x = rint((2 * L - A - A) / (A + A));
            c = (x > 0 ? (int)x : 0) + 1;
        }
        total = __reduce_add_sync(FULL, c);
        if (total > maxv) A *= 1.01 * (double)total / maxv;
        else if (total < G.min_verts + 1) A *= 0.99 * (double)total / (G.min_verts + 1);
        else break;
    }
    // exclusive prefix of the per-edge counts
    int incl_c = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULL, incl_c, o);
        if (lane >= o) incl_c += y;
    }
    if (lane < K) { cnt[lane] = c; offs[lane] = incl_c - c; }
    __syncwarp();
    if (dbg != nullptr) {                              // parity read-back (mg_debug_polygon): K, spacing, clockwise coarse polygon in pixels
        if (lane == 0) { dbg[0] = K; dbg[1] = A; }
        if (lane < K) { dbg[2 + 2 * lane] = cx[K - 1 - lane]; dbg[3 + 2 * lane] = cy[K - 1 - lane]; }
    }
    const bool odd = (total & 1) != 0;
    const int c_last = cnt[K - 1];
    const int drop = odd ? offs[K - 1] + c_last / 2 : -1;   // flat index popped from the last edge (tk-ui.py:267-269)
    // coarse edge data in shared scratch so that any lane can emit any point
    double *edge = reinterpret_cast<double *>(w.queue + 128);            // [K][4]: prev x, prev y, unit x, unit y
    if (lane < K) {
        edge[4 * lane + 0] = pxv; edge[4 * lane + 1] = pyv;
        edge[4 * lane + 2] = (cxv - pxv) / L; edge[4 * lane + 3] = (cyv - pyv) / L;
    }
    __syncwarp();
#pragma unroll 1
    for (int f = lane; f < total; f += 32) {
        if (f == drop) continue;
        int lo = 0, hi = K - 1;                       // last edge e with offs[e] <= f
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (offs[mid] <= f) lo = mid; else hi = mid - 1;
        }
        const int e = lo, j = f - offs[e], ce = cnt[e];
        double X, Y;
        if (j == ce - 1) {                            // the coarse vertex itself
            int cur = K - 1 - e;
            X = cx[cur]; Y = cy[cur];
        } else {
            double d = A * (j + 1);
            X = edge[4 * e + 0] + d * edge[4 * e + 2];
            Y = edge[4 * e + 1] + d * edge[4 * e + 3];
        }
        int pos = (drop >= 0 && f > drop) ? f - 1 : f;
        if (pos < P.cap) w.ring[pos] = make_double2(X / 100.0, Y / 100.0);
    }
    __syncwarp();
    int n = odd ? total - 1 : total;
    return n <= P.cap ? n : (P.cap & ~1);
}

// The polygon of episode `episode` of env `env` in random-polygon mode, regenerated from (seed, global env id,
// episode): the generator is a pure function of its counter, so the original vertices of an episode need not be
// kept (mg_get_elements), and the parity tests can read back any episode's polygon (mg_debug_polygon).
// dbg (optional) = {K, spacing, coarse clockwise polygon in pixels (2 K values)}, area_out = the shoelace area the
// reset computes for this ring.
__global__ void __launch_bounds__(32) mg_regen_polygon_kernel(const __grid_constant__ Params P, int env, int episode, double2 *out_xy,
                                                            int32_t *out_n, double *dbg, double *area_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = carve(smem_raw, P.cap);
    Warp w;
    w.ring = L.ring; w.queue = L.queue; w.lane = threadIdx.x; w.n = 0;
    const int n = generate_polygon(P, w, P.env_id_offset + env, episode < 0 ? P.cold[env].episode : episode, dbg);
    w.n = n;
    __syncwarp();
    for (int j = threadIdx.x; j < n; j += 32) out_xy[j] = w.ring[j];
    const double area = shoelace_area(w);
    if (threadIdx.x == 0) {
        *out_n = n;
        if (area_out) *area_out = area;
    }
}

}  // namespace mg
