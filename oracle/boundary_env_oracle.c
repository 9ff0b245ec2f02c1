/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, scalar, sequential) of the reference
 * BoudaryEnv reset / step / move path, smooth_pave included.  It exists so that the CUDA product path can be checked on a
 * machine where the Python reference is absent (the GPU box), and as the "port" CPU baseline
 * of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product package never does.
 *
 * Parity pin: this file is validated bit-for-bit (obs, flags, element counts, boundary vertex
 * ids + coordinates, candidate list; rewards exactly equal) against the *live* reference env
 * (v2/src/mesh_rl/envs/boundary_env.py) by tests/test_oracle_vs_live_reference_cpu.py and the live halves of
 * tests/test_oracle_move_cpu.py whenever /root/reference is present (and by oracle/sweep_vs_reference.py,
 * sweep_move_vs_reference.py, sweep_smooth_vs_reference.py run by hand), and against the golden traces in
 * tests/golden/ (recorded from the reference by oracle/record_golden.py and the sweep scripts' --record)
 * everywhere else.  The reference ships no golden vectors
 * of its own for this path.
 *
 * It deliberately keeps the reference's *data-structure semantics* (Python list with index 0,
 * identity comparison of vertices, per-vertex segment lists, an explicit sorted candidate list)
 * instead of the warp-parallel (key, stamp) formulation used by the CUDA kernels, so the two
 * are independent statements of the same behaviour.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-builtin  (see oracle/Makefile).
 *   -ffp-contract=off : CPython/NumPy never fuse a*b+c.
 *   -fno-builtin      : CPython evaluates x**2 as libm pow(x, 2.0), which differs from x*x in
 *                       ~0.08 % of inputs with glibc 2.39; gcc would otherwise fold it to x*x.
 *
 * Citations "C:" = v2/src/mesh_rl/components_core.py, "M:" = v2/src/mesh_rl/mesh_core.py,
 * "E:" = v2/src/mesh_rl/envs/boundary_env.py, "D:" = v2/src/mesh_rl/data_core.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.141592653589793

typedef struct {
    double x, y;
    int is_np;      /* coordinates are np.float64 in the reference (inserted vertices, E:210) */
    int *adj;       /* partner vertex ids, one per Segment in Vertex.segments (C:75-79) */
    int nadj, capadj;
} OVertex;

typedef struct {
    int n0;
    double *xy0;            /* original polygon, 2*n0 */
    double original_area;   /* C:485-487 (passed in: np.dot summation order is BLAS-specific) */
    OVertex *v; int nv, capv;        /* vertex pool; pool index == vertex id */
    int *B; int n; int capB;         /* updated_boundary.vertices (list of vertex ids) */
    int *cand_id; double *cand_key; int ncand; int capcand; int cand_none; /* candidate_vertices */
    int *elements; int n_elements; int cap_elements;   /* generated_meshes (4 ids each) */
    double current_area;
    double area_min, area_crit;      /* estimated_area_range (M:705-718) */
    int failed_num;
    int ref_id;                      /* current_point_environment.reference_point; -1 = None */
    double base_length;
    float obs[18];
    int obs_none;
    int crashed;                     /* the Python reference would have raised here */
    /* episode statistics (not part of the reference; bookkeeping for the harness) */
    double ep_return; int ep_len;
    int last_rule; int last_success; int last_inside; int last_existing;
    double dbg[8];                   /* last successful step: b_reward, e_reward, area, penalty, ... */
    /* move() (E:459-594): not_valid_points as coordinates (the reference excludes candidates within 0.001 of one of
     * them, M:428-433), static point environments (area-ratio slot of the observation = 0, C:1209-1214) */
    double *excl; int *excl_id; int nexcl, capexcl;
    int *last_excl_id; int nlast, caplast;   /* last_not_valid_points (E:103, E:570-578): vertex identities */
    int static_obs;
    int needs_smoothing;             /* move(): every candidate is excluded -> the reference calls smooth_pave (M:816-821) */
    int smooth_enabled;              /* 0: stop there and report needs_smoothing (round-1/2 tests); 1: run smooth_pave */
    int n_smoothings;
} OEnv;

/* ---------------------------------------------------------------- rounding ------------- */

/* CPython round(x, 4) on a float: correctly rounded to 4 decimals on the exact binary value,
 * ties (exact) to even (Objects/floatobject.c double_round).  x*1e4 = p + e exactly (fma). */
static double py_round4(double x) {
    double p = x * 1e4;
    double e = fma(x, 1e4, -p);
    double r = nearbyint(p);
    double d = p - r;
    if (d == 0.5 || d == -0.5) {
        if (e > 0) r = floor(p) + 1.0;
        else if (e < 0) r = floor(p);
    }
    return r / 1e4;
}
/* exact (slow) version used by the self-test: printf/strtod are correctly rounded in glibc */
static double py_round4_slow(double x) {
    char buf[512];
    snprintf(buf, sizeof buf, "%.4f", x);
    return strtod(buf, NULL);
}
/* NumPy scalar round(x, 4) for np.float64: rint(x * 1e4) / 1e4 */
static double np_round4(double x) { return nearbyint(x * 1e4) / 1e4; }
/* NumPy scalar round(x, 4) for np.float32, all in float32 */
static float np_round4f(float x) {
    volatile float p = x * 1e4f;
    volatile float r = nearbyintf(p);
    return r / 1e4f;
}
static double round4_mixed(double x, int is_np) { return is_np ? np_round4(x) : py_round4(x); }

/* ---------------------------------------------------------------- primitives ----------- */

typedef struct { double x, y; } P2;

/* C:25-26 Point2D.distance_to: math.sqrt((dx) ** 2 + (dy) ** 2); ** 2 is libm pow */
static double dist2(double ax, double ay, double bx, double by) {
    return sqrt(pow(ax - bx, 2.0) + pow(ay - by, 2.0));
}
static double pdist(P2 a, P2 b) { return dist2(a.x, a.y, b.x, b.y); }

/* Python builtin sum() over floats, starting from int 0.  The oracle is pinned against the
 * reference executed under CPython 3.12, whose sum() uses Neumaier compensated summation
 * (Python/bltinmodule.c builtin_sum_impl); CPython <= 3.11 adds left to right.  The two differ
 * by at most an ulp of the result, far inside the 1e-9 reward tolerance. */
static double py_sum(const double *x, int n) {
    if (n == 0) return 0;
    double f = x[0], c = 0;
    for (int i = 1; i < n; i++) {
        double t = f + x[i];
        if (fabs(f) >= fabs(x[i])) c += (f - t) + x[i];
        else c += (x[i] - t) + f;
        f = t;
    }
    if (c != 0 && isfinite(c)) f += c;
    return f;
}

/* C:99-108 Vertex.to_find_clockwise_angle(self=c, point1, point2) */
static double cw_angle(P2 c, P2 p1, P2 p2) {
    double v1x = p1.x - c.x, v1y = p1.y - c.y;
    double v2x = p2.x - c.x, v2y = p2.y - c.y;
    double a = v1x * v2y, b = v1y * v2x;
    double cr = a - b;
    double d1 = v1x * v2x, d2 = v1y * v2y;
    double dt = d1 + d2;
    double theta = -atan2(cr, dt);
    if (!signbit(theta)) return py_round4(theta);
    return py_round4(2 * PI + theta);
}

/* C:490-491 */
static double cross_product(double v1x, double v1y, double v2x, double v2y) {
    double a = v1x * v2y, b = v2x * v1y;
    return a - b;
}

/* C:499-524 Segment.straddle(self=(s1,s2), another=(o1,o2)) */
static int straddle(P2 s1, P2 s2, P2 o1, P2 o2) {
    double v1x = o1.x - s1.x, v1y = o1.y - s1.y;
    double v2x = o2.x - s1.x, v2y = o2.y - s1.y;
    double vmx = s2.x - s1.x, vmy = s2.y - s1.y;
    double sa = py_round4(sin(cw_angle(s1, o1, s2)));
    double sb = py_round4(sin(cw_angle(s1, o2, s2)));
    if (sa == sb && sb == 0) {
        double l1 = pdist(s1, s2), l2 = pdist(o1, o2);
        if (l1 > l2) {
            P2 m = { (s2.x + s1.x) / 2, (s2.y + s1.y) / 2 };
            double da = pdist(m, o2), db = pdist(m, o1);
            double mn = db < da ? db : da;
            if (mn <= l1 / 2) return 1;
        } else {
            P2 m = { (o2.x + o1.x) / 2, (o2.y + o1.y) / 2 };
            double da = pdist(m, s2), db = pdist(m, s1);
            double mn = db < da ? db : da;
            if (mn <= l2 / 2) return 1;
        }
        return 0;
    }
    return cross_product(v1x, v1y, vmx, vmy) * cross_product(v2x, v2y, vmx, vmy) <= 0;
}
/* C:526-541 */
static int is_cross(P2 s1, P2 s2, P2 o1, P2 o2) {
    return straddle(s1, s2, o1, o2) && straddle(o1, o2, s1, s2);
}

/* ---------------------------------------------------------------- env helpers ---------- */

static P2 VP(const OEnv *e, int id) { P2 p = { e->v[id].x, e->v[id].y }; return p; }
/* Python list indexing with negative wrap (only one wrap is ever needed) */
static int BI(const OEnv *e, int i) { return e->B[i < 0 ? i + e->n : i]; }
static P2 BP(const OEnv *e, int i) { return VP(e, BI(e, i)); }
static int B_index(const OEnv *e, int id) {
    for (int i = 0; i < e->n; i++) if (e->B[i] == id) return i;
    return -1;
}
static int in_B(const OEnv *e, int id) { return B_index(e, id) >= 0; }

static int new_vertex(OEnv *e, double x, double y, int is_np) {
    if (e->nv == e->capv) {
        e->capv = e->capv ? 2 * e->capv : 64;
        e->v = (OVertex *)realloc(e->v, sizeof(OVertex) * e->capv);
    }
    OVertex *v = &e->v[e->nv];
    v->x = x; v->y = y; v->is_np = is_np; v->adj = NULL; v->nadj = 0; v->capadj = 0;
    return e->nv++;
}
static void add_adj(OEnv *e, int a, int b) {
    OVertex *v = &e->v[a];
    if (v->nadj == v->capadj) {
        v->capadj = v->capadj ? 2 * v->capadj : 4;
        v->adj = (int *)realloc(v->adj, sizeof(int) * v->capadj);
    }
    v->adj[v->nadj++] = b;
}
/* C:110-116 */
static int has_segment_with(const OEnv *e, int a, int b) {
    const OVertex *v = &e->v[a];
    for (int k = 0; k < v->nadj; k++) if (v->adj[k] == b) return 1;
    return 0;
}
/* C:123-132: number of distinct partners */
static int n_connected(const OEnv *e, int a) {
    const OVertex *v = &e->v[a];
    int cnt = 0;
    for (int k = 0; k < v->nadj; k++) {
        int dup = 0;
        for (int j = 0; j < k; j++) if (v->adj[j] == v->adj[k]) { dup = 1; break; }
        if (!dup && v->adj[k] != a) cnt++;
    }
    return cnt;
}
/* M:500-505 */
static int count_segts_in_boundary(const OEnv *e, int a) {
    const OVertex *v = &e->v[a];
    int cnt = 0;
    if (!in_B(e, a)) return 0;
    for (int k = 0; k < v->nadj; k++) if (in_B(e, v->adj[k])) cnt++;
    return cnt;
}

static void free_pool(OEnv *e) {
    for (int i = 0; i < e->nv; i++) free(e->v[i].adj);
    e->nv = 0;
}

/* ---------------------------------------------------------------- candidates ----------- */

/* M:228-257 check_boundary_point; returns 0 for None */
static int check_boundary_point(const OEnv *e, int index, double *key) {
    int n = e->n;
    P2 c = BP(e, index);
    double a0 = cw_angle(c, BP(e, (index + 1) % n), BP(e, index - 1));
    if (a0 >= PI * 0.972 || a0 == 0) return 0;
    double lam = 0.618;
    double sum = a0 * lam;
    double a1 = cw_angle(c, BP(e, (index + 2) % n), BP(e, index - 2));
    sum += a1 * (1 - lam);
    *key = sum * (180.0 / PI);      /* math.degrees */
    return 1;
}

static void cand_reserve(OEnv *e, int need) {
    if (need > e->capcand) {
        e->capcand = need * 2 + 16;
        e->cand_id = (int *)realloc(e->cand_id, sizeof(int) * e->capcand);
        e->cand_key = (double *)realloc(e->cand_key, sizeof(double) * e->capcand);
    }
}

/* M:259-287 find_reference_candidates: stable sort by key */
static void find_reference_candidates(OEnv *e) {
    cand_reserve(e, e->n + 8);
    int m = 0;
    for (int i = 0; i < e->n; i++) {
        double key;
        if (check_boundary_point(e, i, &key)) { e->cand_id[m] = e->B[i]; e->cand_key[m] = key; m++; }
    }
    /* stable merge sort on fabs(key - 0) */
    int *tid = (int *)malloc(sizeof(int) * (m + 1));
    double *tk = (double *)malloc(sizeof(double) * (m + 1));
    for (int w = 1; w < m; w *= 2) {
        for (int lo = 0; lo < m; lo += 2 * w) {
            int mid = lo + w < m ? lo + w : m, hi = lo + 2 * w < m ? lo + 2 * w : m;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                if (fabs(e->cand_key[j]) < fabs(e->cand_key[i])) { tid[k] = e->cand_id[j]; tk[k++] = e->cand_key[j++]; }
                else { tid[k] = e->cand_id[i]; tk[k++] = e->cand_key[i++]; }
            }
            while (i < mid) { tid[k] = e->cand_id[i]; tk[k++] = e->cand_key[i++]; }
            while (j < hi) { tid[k] = e->cand_id[j]; tk[k++] = e->cand_key[j++]; }
        }
        memcpy(e->cand_id, tid, sizeof(int) * m);
        memcpy(e->cand_key, tk, sizeof(double) * m);
    }
    free(tid); free(tk);
    e->ncand = m;
    e->cand_none = 0;
}

/* M:188-204 */
static void remove_reference_candidates(OEnv *e, const int *ids, int k) {
    for (int q = 0; q < k; q++) {
        int w = 0;
        for (int i = 0; i < e->ncand; i++)
            if (e->cand_id[i] != ids[q]) { e->cand_id[w] = e->cand_id[i]; e->cand_key[w] = e->cand_key[i]; w++; }
        e->ncand = w;
    }
}
/* M:206-226 */
static void add_reference_candidates(OEnv *e, const int *ids, int k) {
    for (int q = 0; q < k; q++) {
        double key;
        int index = B_index(e, ids[q]);
        if (!check_boundary_point(e, index, &key)) continue;
        cand_reserve(e, e->ncand + 2);
        int pos = e->ncand;
        for (int i = 0; i < e->ncand; i++) if (key <= e->cand_key[i]) { pos = i; break; }
        memmove(e->cand_id + pos + 1, e->cand_id + pos, sizeof(int) * (e->ncand - pos));
        memmove(e->cand_key + pos + 1, e->cand_key + pos, sizeof(double) * (e->ncand - pos));
        e->cand_id[pos] = ids[q]; e->cand_key[pos] = key; e->ncand++;
    }
}

/* ---------------------------------------------------------------- observation ---------- */

static float f32(double x) { return (float)x; }

/* C:1059-1090 PointEnvironment + C:1192-1290 get_radius_points; E:665-738 find_next_state */
static void find_next_state(OEnv *e) {
    if (e->cand_none) find_reference_candidates(e);          /* M:297-298 */
    if (e->ncand == 0) { e->ref_id = -1; e->obs_none = 1; return; }   /* returns None */
    int rid = e->cand_id[0];
    if (e->nexcl > 0) {                                      /* M:310-314: first candidate that is not a not-valid point */
        rid = -1;
        for (int c = 0; c < e->ncand && rid < 0; c++) {
            P2 v = VP(e, e->cand_id[c]);
            int hit = 0;
            for (int k = 0; k < e->nexcl && !hit; k++) {
                P2 q = { e->excl[2 * k], e->excl[2 * k + 1] };
                if (pdist(q, v) < 0.001) hit = 1;            /* M:428-433 is_vertex_inside_list */
            }
            if (!hit) rid = e->cand_id[c];
        }
        if (rid < 0) { e->ref_id = -1; e->obs_none = 1; return; }
    }
    e->ref_id = rid;
    e->obs_none = 0;
    int n = e->n;
    int index = B_index(e, rid);
    P2 ref = VP(e, rid);

    /* C:1081-1090 neighbours [i+3, i+2, i+1, i, i-1, i-2, i-3] and base_length */
    P2 N[7];
    for (int k = 0; k < 4; k++) N[k] = BP(e, (index + (3 - k)) % n);
    for (int k = 1; k <= 3; k++) N[3 + k] = BP(e, index - k);
    double dl[6];
    for (int k = 1; k < 7; k++) dl[k - 1] = pdist(N[k], N[k - 1]);
    double base = py_round4(py_sum(dl, 6) / 6);
    e->base_length = base;

    float r[9][2];
    for (int i = 0; i < 9; i++) { r[i][0] = 1.0f; r[i][1] = 1.0f; }
    P2 right_p = BP(e, index - 1);
    int right_id = BI(e, index - 1);
    P2 left_p = BP(e, (index + 1) % n);
    int left_id = BI(e, (index + 1) % n);
    const double radius = 4;
    double T = base * radius;
    double theta = cw_angle(ref, left_p, right_p);
    double area_ratio = e->static_obs ? 0.0 : e->current_area / e->original_area;     /* C:1209-1214 */

    r[0][0] = f32((pdist(ref, right_p) / radius) / base); r[0][1] = f32(area_ratio);
    r[8][0] = f32((pdist(ref, left_p) / radius) / base);  r[8][1] = f32(theta);
    for (int i = 1; i < 3; i++) {
        P2 q = BP(e, index - i - 1);
        double a = cw_angle(ref, q, right_p);
        r[i][0] = f32((pdist(ref, q) / radius) / base);
        r[i][1] = f32(a < PI ? a : (a > 1.5 * PI ? a : 1.5 * PI) - 2 * PI);
        q = BP(e, (index + 1 + i) % n);
        a = cw_angle(ref, q, right_p);
        r[8 - i][0] = f32((pdist(ref, q) / radius) / base);
        double lim = theta + PI / 2;
        r[8 - i][1] = f32(a < lim ? a : lim);            /* min(a, lim): returns a on ties */
    }
    P2 refp1 = { ref.x + 1, ref.y + 0 };
    double rot = cw_angle(ref, right_p, refp1);
    double clip = theta + PI / 2;
    for (int j = 0; j < 3; j++) {
        double a = (2 * j + 1) * theta / 6;
        r[3 + j][1] = f32(a < clip ? a : clip);
    }
    /* p_s = ref + rotate((T cos(theta/2), T sin(theta/2)), rot)   C:154-168, C:1243 */
    double px = T * cos(theta / 2), py = T * sin(theta / 2);
    double qx = cos(rot) * px - sin(rot) * py;
    double qy = sin(rot) * px + cos(rot) * py;
    P2 ps = { ref.x + qx, ref.y + qy };
    double ux = ps.x - ref.x, uy = ps.y - ref.y;

    double shortest = 1; int shortest_i = 0;
    for (int i = index - 1; i > index - n; i--) {
        int vid = BI(e, i);
        P2 q = VP(e, vid);
        double d = pdist(ref, q);
        if (vid == right_id || vid == left_id) continue;
        double angle = cw_angle(ref, q, right_p);
        if (angle == 0) continue;
        double sector = theta / 3;
        double kk = angle / sector;
        if (sector == 0) e->crashed = 1;                 /* ZeroDivisionError in the reference */
        int k = (kk >= 3.0 || kk != kk) ? 3 : (int)kk;
        if (k < 3 && d < T) {
            float cand = f32((d / radius) / base);
            if (r[k + 3][0] > cand) {                    /* float32 comparison (NEP 50) */
                r[k + 3][0] = cand;
                r[k + 3][1] = f32(angle < clip ? angle : clip);
            }
        }
        /* C:657-676 ll.intersection_vertex(seg), ll = (ref, ps), seg = (B[i], B[i+1]) */
        P2 q2 = BP(e, i + 1);
        double wx = q2.x - q.x, wy = q2.y - q.y;
        double ss, hh;
        if (wy == 0) {
            if (uy == 0) continue;
            ss = (q.y - ref.y) / uy;
            if (wx == 0) e->crashed = 1;
            hh = (ref.x - q.x + ss * ux) / wx;
        } else if (wx == 0) {
            if (ux == 0) continue;
            ss = (q.x - ref.x) / ux;
            hh = (ref.y - q.y + ss * uy) / wy;
        } else {
            double den = uy / wy - ux / wx;
            if (den == 0) e->crashed = 1;
            ss = ((ref.x - q.x) / wx - (ref.y - q.y) / wy) / den;
            hh = (ref.x - q.x + ss * ux) / wx;
        }
        if (0 < ss && ss < 1 && 0 < hh && hh < 1) {
            P2 vv = { ref.x + ss * ux, ref.y + ss * uy };
            double _d = pdist(ref, vv);
            double val = (_d / radius) / base;
            if (shortest > val) { shortest = val; shortest_i = i; }
        }
    }
    if (shortest != 1 && f32(shortest) < r[4][0]) {
        for (int j = 0; j < 3; j++) {
            P2 q = BP(e, j - 1 + shortest_i);
            r[3 + j][0] = f32((pdist(ref, q) / radius) / base);
            r[3 + j][1] = f32(cw_angle(ref, q, right_p));
        }
    }
    for (int i = 0; i < 9; i++) { e->obs[2 * i] = np_round4f(r[i][0]); e->obs[2 * i + 1] = np_round4f(r[i][1]); }
}

/* ---------------------------------------------------------------- reset ---------------- */

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* M:705-718 on the freshly deep-copied boundary */
static void estimate_area_range(OEnv *e) {
    int n = e->n0;
    double *len = (double *)malloc(sizeof(double) * n);
    /* deep_copy segments: Segment(points[i-1], points[i]).length() = points[i-1].distance_to(points[i]) */
    for (int i = 0; i < n; i++) {
        int a = (i + n - 1) % n;
        len[i] = dist2(e->xy0[2 * a], e->xy0[2 * a + 1], e->xy0[2 * i], e->xy0[2 * i + 1]);
    }
    qsort(len, n, sizeof(double), cmp_double);
    double L = py_sum(len, n) / n;
    double max_L = len[n - 2] < 2 * L ? len[n - 2] : 2 * L;       /* min(lengths[-2], 2 L) */
    double a = L / sqrt(2.0);
    double min_L = len[1] < a ? len[1] : a;                        /* min(L/sqrt2, lengths[1]) */
    e->area_min = min_L;
    e->area_crit = (max_L + 3 * min_L) / 4;
    free(len);
}

void oracle_reset(OEnv *e) {
    free_pool(e);
    e->n = e->n0;
    for (int i = 0; i < e->n0; i++) { new_vertex(e, e->xy0[2 * i], e->xy0[2 * i + 1], 0); e->B[i] = i; }
    for (int i = 0; i < e->n0; i++) {            /* C:221-228 deep_copy segments */
        int a = (i + e->n0 - 1) % e->n0;
        add_adj(e, a, i); add_adj(e, i, a);
    }
    e->n_elements = 0;
    e->current_area = e->original_area;
    e->cand_none = 1; e->ncand = 0;
    e->failed_num = 0;
    e->ep_return = 0; e->ep_len = 0;
    e->nexcl = 0; e->static_obs = 0; e->needs_smoothing = 0;
    e->nlast = 0;      /* last_not_valid_points survives reset() in the reference, but holds vertex objects of the old episode */
    find_next_state(e);
    estimate_area_range(e);
}

/* sequential shoelace, used only when the caller passes original_area <= 0 */
static double shoelace(const double *xy, int n) {
    double s1 = 0, s2 = 0;
    for (int i = 0; i < n; i++) {
        int p = (i + n - 1) % n;
        s1 += xy[2 * i] * xy[2 * p + 1];
        s2 += xy[2 * i + 1] * xy[2 * p];
    }
    return 0.5 * fabs(s1 - s2);
}

OEnv *oracle_create(const double *xy, int n, double original_area) {
    OEnv *e = (OEnv *)calloc(1, sizeof(OEnv));
    e->n0 = n;
    e->xy0 = (double *)malloc(sizeof(double) * 2 * n);
    memcpy(e->xy0, xy, sizeof(double) * 2 * n);
    e->capB = n + 8;
    e->B = (int *)malloc(sizeof(int) * e->capB);
    e->original_area = original_area > 0 ? original_area : shoelace(xy, n);
    oracle_reset(e);
    return e;
}

void oracle_destroy(OEnv *e) {
    if (e) { free(e->excl); free(e->excl_id); free(e->last_excl_id); }
    if (!e) return;
    free_pool(e);
    free(e->v); free(e->B); free(e->xy0); free(e->cand_id); free(e->cand_key); free(e->elements);
    free(e);
}

/* ---------------------------------------------------------------- step pieces ---------- */

/* M:74-128 calculate_crossing_segments over updated_boundary, ray = (P, (10000, P.y)) */
static int point_inside(const OEnv *e, P2 P) {
    int n = e->n, count = 0;
    P2 ray2 = { 10000, P.y };
    for (int i = 0; i < n; i++) {
        int ia = e->B[i], ib = BI(e, i - 1);
        P2 a = VP(e, ia), b = VP(e, ib);
        double orientation = round4_mixed(a.y - b.y, e->v[ia].is_np || e->v[ib].is_np);
        if (orientation == 0) continue;
        if (!is_cross(a, b, P, ray2)) continue;
        if (np_round4(a.y - ray2.y) == 0) {
            int ic = e->B[(i + 1) % n];
            double nxt = round4_mixed(e->v[ic].y - a.y, e->v[ic].is_np || e->v[ia].is_np);
            if (nxt == 0) continue;
            else if (nxt * orientation < 0) continue;
            else { if (orientation < 0) count++; else continue; }
        } else {
            if (np_round4(b.y - ray2.y) == 0) {
                int ic = BI(e, i - 2 < -n ? i - 2 + n : i - 2);
                double pre = round4_mixed(b.y - e->v[ic].y, e->v[ib].is_np || e->v[ic].is_np);
                if (pre == 0) continue;
                else if (pre * orientation < 0) continue;
                else { if (orientation < 0) continue; else count++; }
            } else count++;
        }
    }
    return count % 2 != 0;
}

/* E:766-769 */
static int find_same_point(const OEnv *e, P2 P) {
    for (int i = 0; i < e->n; i++) if (pdist(BP(e, i), P) < 0.001) return 1;
    return 0;
}

/* C:738-757 + C:814-826 Mesh.is_valid(0) */
static int mesh_is_valid(const P2 *m) {
    if (is_cross(m[0], m[1], m[2], m[3])) return 0;
    if (is_cross(m[0], m[3], m[1], m[2])) return 0;
    double max_degree = 0.99 * PI, min_degree = 0.01 * PI;
    for (int i = 0; i < 4; i++) {
        double deg = cw_angle(m[i], m[(i + 1) % 4], m[(i + 3) % 4]);
        if (deg > max_degree || deg < min_degree) return 0;
    }
    return 1;
}

static int in_mesh(const int *mid, int id) { return mid[0] == id || mid[1] == id || mid[2] == id || mid[3] == id; }

/* M:536-556; mid[] holds vertex ids (the not-yet-inserted new vertex has id -2) */
static int check_intersection_with_boundary(const OEnv *e, const int *mid, const P2 *m, int ref_id) {
    P2 ref = VP(e, ref_id);
    int n = e->n;
    double max_dist = -1; int first = 1;
    int _index = -1;
    for (int k = 0; k < 4; k++) {
        if (mid[k] == ref_id) { _index = k; continue; }
        double d = pdist(ref, m[k]);
        if (first || d > max_dist) { max_dist = d; first = 0; }
    }
    P2 c1a = m[(_index + 3) % 4], c1b = m[(_index + 2) % 4];      /* [_index-1], [_index-2] */
    P2 c2a = m[(_index + 2) % 4], c2b = m[(_index + 1) % 4];      /* [_index-2], [_index-3] */
    for (int index = 0; index < n; index++) {
        int vid = e->B[index];
        P2 v = VP(e, vid);
        if (!(pdist(ref, v) < max_dist) || in_mesh(mid, vid)) continue;
        int pid = BI(e, index - 1), nid = e->B[(index + 1) % n];
        for (int c = 0; c < 2; c++) {
            P2 ga = c ? c2a : c1a, gb = c ? c2b : c1b;
            if (!in_mesh(mid, pid)) if (is_cross(ga, gb, v, VP(e, pid))) return 1;
            if (!in_mesh(mid, nid)) if (is_cross(ga, gb, v, VP(e, nid))) return 1;
        }
    }
    return 0;
}

static void B_insert(OEnv *e, int pos, int id) {
    memmove(e->B + pos + 1, e->B + pos, sizeof(int) * (e->n - pos));
    e->B[pos] = id; e->n++;
}
static void B_remove(OEnv *e, int id) {
    int pos = B_index(e, id);
    memmove(e->B + pos, e->B + pos + 1, sizeof(int) * (e->n - pos - 1));
    e->n--;
}

/* M:601-674 */
static void update_boundary(OEnv *e, const int *mid, int n_new, int new_pos) {
    int n;
    if (n_new == 1) {
        int target = mid[(new_pos + 2) % 4];                     /* mesh.vertices[index(new) - 2] */
        int id = B_index(e, target);
        B_insert(e, id, mid[new_pos]);
        B_remove(e, target);
        n = e->n;
        int nb[5];
        for (int i = 0; i < 2; i++) { nb[2 * i] = e->B[(id + i + 1) % n]; nb[2 * i + 1] = BI(e, id - i - 1); }
        nb[4] = target;
        remove_reference_candidates(e, nb, 5);
        add_reference_candidates(e, nb, 4);
    } else if (n_new == 0) {
        int removable[4], nr = 0;
        for (int k = 0; k < 4; k++) if (count_segts_in_boundary(e, mid[k]) < 3) removable[nr++] = mid[k];
        for (int k = 0; k < nr; k++) B_remove(e, removable[k]);
        n = e->n;
        int id = -1;
        for (int k = 0; k < 4; k++) {
            int rem = 0;
            for (int q = 0; q < nr; q++) if (removable[q] == mid[k]) rem = 1;
            if (rem) continue;
            int ix = B_index(e, mid[k]);
            if (ix > id) id = ix;
        }
        int lst[8], nl = 0;
        for (int k = 0; k < nr; k++) lst[nl++] = removable[k];
        int nb[4];
        for (int i = 0; i < 2; i++) { nb[2 * i] = e->B[(id + i) % n]; nb[2 * i + 1] = BI(e, id - i - 1); }
        for (int k = 0; k < 4; k++) lst[nl++] = nb[k];
        remove_reference_candidates(e, lst, nl);
        add_reference_candidates(e, nb, 4);
    }
}

/* C:943-958 */
static double mesh_compute_area(const P2 *m) {
    double L[4];
    for (int i = 0; i < 4; i++) L[i] = pdist(m[i], m[(i + 3) % 4]);
    double c1 = cw_angle(m[0], m[1], m[3]);
    double c3 = cw_angle(m[2], m[3], m[1]);
    return 0.5 * L[0] * L[1] * sin(c1) + 0.5 * L[2] * L[3] * sin(c3);
}

/* C:881-892 */
static double mesh_quality_robust(const P2 *m) {
    double mn = 0;
    for (int i = 0; i < 4; i++) {
        double l = pdist(m[(i + 3) % 4], m[i]);
        if (i == 0 || l < mn) mn = l;
    }
    double d1 = pdist(m[0], m[2]), d2 = pdist(m[1], m[3]);
    double q1 = sqrt(2.0) * mn / (d2 > d1 ? d2 : d1);
    double amin = 0, amax = 0;
    for (int i = 0; i < 4; i++) {
        double a = cw_angle(m[i], m[(i + 1) % 4], m[(i + 3) % 4]);
        if (i == 0 || a < amin) amin = a;
        if (i == 0 || a > amax) amax = a;
    }
    double q2 = amin / amax;
    return sqrt(q1 * q2);
}

/* C:678-692 Segment(p1, p2).distance(point) */
static double seg_point_distance(P2 p1, P2 p2, P2 a) {
    double A = p2.x - p1.x, Bv = p2.y - p1.y;
    double s = (A * a.x + Bv * a.y - Bv * p1.y - A * p1.x) / (pow(A, 2.0) + pow(Bv, 2.0));
    if (0 <= s && s <= 1) {
        P2 t = { p1.x + s * A, p1.y + s * Bv };
        return pdist(a, t);
    } else if (s < 0) return pdist(a, p1);
    return pdist(a, p2);
}

/* M:355-408 */
static double compute_boundary_quality(const OEnv *e, int add_id) {
    int n = e->n;
    int index = B_index(e, add_id);
    P2 add_v = VP(e, add_id);
    double amin = 0; int na = 0;
    for (int s = 0; s < 2; s++) {
        int i = s == 0 ? 1 : -1;
        double angle = cw_angle(BP(e, (((index + i) % n) + n) % n), BP(e, (index + i + 1) % n), BP(e, index + i - 1));
        if (angle < PI / 3) { if (na == 0 || angle < amin) amin = angle; na++; }
    }
    double q1 = na ? 3 * amin / PI : 1;
    double dist = pdist(add_v, BP(e, (index + 1) % n)) + pdist(add_v, BP(e, index - 1));
    int ex[5] = { e->B[index], e->B[(index + 1) % n], e->B[(index + 2) % n], BI(e, index - 1), BI(e, index - 2) };
    double m_d = 0; int nd = 0; int last_close = -2;
    for (int i = 0; i < n; i++) {
        int vid = e->B[i], skip = 0;
        for (int k = 0; k < 5; k++) if (ex[k] == vid) skip = 1;
        if (skip) continue;
        if (pdist(add_v, VP(e, vid)) < dist) {
            if (last_close == i - 1) continue;               /* "i - 1 in close_vs" */
            last_close = i;
            double d = seg_point_distance(BP(e, (i + 1) % n), BP(e, i), add_v);
            if (nd == 0 || d < m_d) m_d = d;
            nd++;
        }
    }
    double targt_len = dist / 2;
    double dl[4];
    for (int k = -2; k < 2; k++) {
        int a = ((index + k) % n + n) % n, b = ((index + k + 1) % n + n) % n;
        dl[k + 2] = pdist(BP(e, a), BP(e, b));
    }
    double mean_dist = py_sum(dl, 4) / 4;
    double smoothness = (mean_dist < targt_len ? mean_dist : targt_len) / (mean_dist > targt_len ? mean_dist : targt_len);
    /* Python: min(a, b) / max(a, b) -- when equal both return a; value identical */
    double q2 = 1;
    if (nd) q2 = m_d < 0.5 * dist ? m_d / (0.5 * dist) : 1;
    ((OEnv *)e)->dbg[4] = smoothness; ((OEnv *)e)->dbg[5] = q1; ((OEnv *)e)->dbg[6] = q2; ((OEnv *)e)->dbg[7] = dist;
    return pow(smoothness * q1 * q2, 1.0 / 3);
}

/* M:410-452 */
static double compute_ele_boundary_quality(const OEnv *e, const int *mid) {
    for (int k = 0; k < 4; k++)
        if (n_connected(e, mid[k]) == 2 && in_B(e, mid[k])) return compute_boundary_quality(e, mid[k]);
    int t[4], nt = 0;
    for (int k = 0; k < 4; k++) if (in_B(e, mid[k])) t[nt++] = mid[k];
    if (nt == 0) return 1;
    int n = e->n;
    double amin = 0; int na = 0;
    for (int k = 0; k < nt; k++) {
        int index = B_index(e, t[k]);
        double angle = cw_angle(VP(e, t[k]), BP(e, (index + 1) % n), BP(e, index - 1));
        if (angle < PI / 3) { if (na == 0 || angle < amin) amin = angle; na++; }
    }
    int i1 = B_index(e, t[0]), ir = B_index(e, t[1]);
    int index = i1 < ir ? i1 : ir;
    double targt_len = pdist(VP(e, t[0]), VP(e, t[1]));
    double dl[5];
    for (int k = -2; k < 3; k++) {
        int a = ((index + k) % n + n) % n, b = ((index + k + 1) % n + n) % n;
        dl[k + 2] = pdist(BP(e, a), BP(e, b));
    }
    double mean_dist = py_sum(dl, 5) / 5;
    double smoothness = (mean_dist < targt_len ? mean_dist : targt_len) / (mean_dist > targt_len ? mean_dist : targt_len);
    double angle_quality = na ? 3 * amin / PI : 1;
    return pow(angle_quality * smoothness, 1.0 / 2);
}

/* E:590-607 */
static double get_speed_penalty(const OEnv *e, double area) {
    double min_area = pow(e->area_min, 2.0), crit = pow(e->area_crit, 2.0);
    if (min_area <= area && area < crit) return (area - crit) / (crit - min_area);
    else if (area < min_area) return -1;
    return 0;
}

static void append_element(OEnv *e, const int *mid) {
    if (e->n_elements == e->cap_elements) {
        e->cap_elements = e->cap_elements ? 2 * e->cap_elements : 64;
        e->elements = (int *)realloc(e->elements, sizeof(int) * 4 * e->cap_elements);
    }
    memcpy(e->elements + 4 * e->n_elements, mid, sizeof(int) * 4);
    e->n_elements++;
}

/* C:840-845 */
static void connect_vertices(OEnv *e, const int *mid) {
    for (int i = 0; i < 4; i++) {
        int a = mid[i], b = mid[(i + 3) % 4];
        if (!has_segment_with(e, a, b)) { add_adj(e, a, b); add_adj(e, b, a); }
    }
}

/* E:388-457 step (no auto-reset).  out: reward, terminated, truncated. */
void oracle_step(OEnv *e, const float *action, double *reward_out, int *terminated, int *truncated) {
    int done = 0, failed = 1;
    double reward = 0;
    float rule_type = action[0];
    e->last_rule = 2; e->last_success = 0; e->last_inside = -1; e->last_existing = 0;

    int rid = e->ref_id;
    int index = B_index(e, rid);
    int n = e->n;
    P2 ref = VP(e, rid);

    /* E:783-792 action_2_point + E:202-210 + D:112-137 */
    double ax = (double)np_round4f(action[1]), ay = (double)np_round4f(action[2]);
    P2 p1 = BP(e, index - 1);
    double theta = 2 * PI - atan2(p1.y - ref.y, p1.x - ref.x);
    double ox = cos(theta) * ax + sin(theta) * ay;
    double oy = -sin(theta) * ax + cos(theta) * ay;
    ox *= e->base_length; oy *= e->base_length;
    ox += ref.x; oy += ref.y;
    P2 newp = { np_round4(ox), np_round4(oy) };

    if (n <= 5) {                                   /* E:428-430 */
        reward = 10; done = 1;
    } else {
        int mid[4]; P2 m[4]; int have_mesh = 1; int n_new = 0, new_pos = -1;
        int use_rule_m1 = 0;
        if (rule_type <= -0.5f) { use_rule_m1 = 1; e->last_rule = -1; }
        else if (rule_type >= 0.5f) {
            e->last_rule = 1;
            mid[0] = BI(e, index - 2); mid[1] = BI(e, index - 1); mid[2] = e->B[index]; mid[3] = e->B[(index + 1) % n];
        } else {
            e->last_rule = 0;
            int inside = point_inside(e, newp);
            e->last_inside = inside;
            if (inside) {
                if (find_same_point(e, newp)) { use_rule_m1 = 1; e->last_existing = 1; }
                else {
                    mid[0] = -2; mid[1] = BI(e, index - 1); mid[2] = e->B[index]; mid[3] = e->B[(index + 1) % n];
                    n_new = 1; new_pos = 0;
                }
            } else {
                reward += e->n_elements ? -1.0 / e->n_elements : -1;      /* E:279 */
                have_mesh = 0;
            }
        }
        if (use_rule_m1) {
            mid[0] = BI(e, index - 1); mid[1] = e->B[index]; mid[2] = e->B[(index + 1) % n]; mid[3] = e->B[(index + 2) % n];
        }
        if (have_mesh) {
            for (int k = 0; k < 4; k++) m[k] = mid[k] == -2 ? newp : VP(e, mid[k]);
            if (mesh_is_valid(m) && !check_intersection_with_boundary(e, mid, m, rid)) {
                if (n_new) mid[new_pos] = new_vertex(e, newp.x, newp.y, 1);
                connect_vertices(e, mid);
                append_element(e, mid);
                update_boundary(e, mid, n_new, new_pos);
                double mesh_area = mesh_compute_area(m);
                e->current_area -= mesh_area;
                double b_reward = compute_ele_boundary_quality(e, mid);
                double e_reward = mesh_quality_robust(m);
                double quality = e_reward + 1 * (b_reward - 1);           /* M:1754-1766 */
                double speed_penalty = get_speed_penalty(e, mesh_area);
                reward += quality + speed_penalty;
                e->dbg[0] = b_reward; e->dbg[1] = e_reward; e->dbg[2] = mesh_area; e->dbg[3] = speed_penalty;
                failed = 0;
                e->last_success = 1;
                if (e->n <= 5) {
                    reward += 10; done = 1;
                    if (e->n == 4) { int fm[4] = { e->B[0], e->B[1], e->B[2], e->B[3] }; connect_vertices(e, fm); append_element(e, fm); }
                } else done = 0;
            } else {
                reward += e->n_elements ? -1.0 / e->n_elements : -1;      /* E:357 */
            }
        }
    }
    /* E:361-386 */
    int is_complete = 1;
    e->static_obs = 0;                        /* step() builds a non-static point environment (E:361); it passes
                                               * self.not_valid_points, which only move() ever fills */
    find_next_state(e);
    if (!failed) e->failed_num = 0;
    else {
        e->failed_num++;
        if (e->failed_num >= 100) { done = 1; is_complete = 0; }
    }
    *reward_out = reward;
    *terminated = done && is_complete;
    *truncated = done && !is_complete;
    e->ep_return += reward; e->ep_len++;
}


/* CPython round(x, 6) on a float (same construction as py_round4) */
static double py_round6(double x) {
    double p = x * 1e6;
    double r = nearbyint(p);
    if (fabs(p - r) == 0.5) {
        double err = fma(x, 1e6, -p);
        if (err > 0) r = floor(p) + 1.0;
        else if (err < 0) r = floor(p);
    }
    return r / 1e6;
}

static void excl_push(OEnv *e, double x, double y, int id) {
    if (e->nexcl == e->capexcl) {
        e->capexcl = e->capexcl ? 2 * e->capexcl : 32;
        e->excl = (double *)realloc(e->excl, sizeof(double) * 2 * e->capexcl);
        e->excl_id = (int *)realloc(e->excl_id, sizeof(int) * e->capexcl);
    }
    e->excl[2 * e->nexcl] = x; e->excl[2 * e->nexcl + 1] = y; e->excl_id[e->nexcl] = id; e->nexcl++;
}

/* ---------------------------------------------------------------- smooth_pave ----------
 * M:816-821 smooth_pave(self.boundary.vertices, self.updated_boundary.vertices, iteration=400):
 *   smooth_current_boundary_3()                      re-position every inserted vertex of the front     (M:965-1060)
 *   smooth_fixed_vertices(interior vertices, 400)    Gauss-Seidel averaging over Vertex.segments        (M:1284-1316)
 *   find_reference_candidates(0)                     full candidate rebuild                              (M:259-287)
 * Vertex identity is Python object identity (no __eq__): vertex ids here.  x ** 2 is libm pow (see dist2). */
#define DEG(x) ((x) * (180.0 / PI))          /* math.degrees: x * (180 / pi) */
#define RAD(x) ((x) * (PI / 180.0))          /* math.radians: x * (pi / 180) */
static double sq(double x) { return pow(x, 2.0); }

/* C:123-132 get_connected_vertices: partners in the order their Segment was assigned, without duplicates */
static int connected_vertices(const OEnv *e, int a, int *out) {
    const OVertex *v = &e->v[a];
    int m = 0;
    for (int k = 0; k < v->nadj; k++) {
        int dup = 0;
        for (int j = 0; j < m; j++) if (out[j] == v->adj[k]) { dup = 1; break; }
        if (!dup && v->adj[k] != a) out[m++] = v->adj[k];
    }
    return m;
}
static int max_degree(const OEnv *e) {
    int m = 0;
    for (int i = 0; i < e->nv; i++) if (e->v[i].nadj > m) m = e->v[i].nadj;
    return m;
}

/* M:1106-1127 clockwise_vertices(inner_v, vertices): selection sort by clockwise angle from the previous entry, then the
 * common neighbour of consecutive entries (other than inner_v) is spliced in.  Returns the length of `fin`. */
static int clockwise_vertices(const OEnv *e, int inner, int *vs, int k, int *fin, int *tmp_a, int *tmp_b) {
    P2 c = VP(e, inner);
    for (int i = 1; i < k; i++) {
        double max_angle = -1;
        int flag = i;
        for (int j = i; j < k; j++) {
            double ang = cw_angle(c, VP(e, vs[j]), VP(e, vs[i - 1]));
            if (ang > max_angle) { max_angle = ang; flag = j; }
        }
        if (flag != i) { int t = vs[i]; vs[i] = vs[flag]; vs[flag] = t; }
    }
    int m = 0;
    for (int i = 0; i < k; i++) {
        int prev = vs[(i + k - 1) % k];
        int na = connected_vertices(e, vs[i], tmp_a), nb = connected_vertices(e, prev, tmp_b);
        int inter = -1;
        for (int x = 0; x < na && inter < 0; x++) {
            if (tmp_a[x] == inner) continue;
            for (int y = 0; y < nb; y++) if (tmp_b[y] == tmp_a[x]) { inter = tmp_a[x]; break; }
        }
        fin[m++] = prev;
        if (inter >= 0) fin[m++] = inter;
    }
    return m;
}

/* M:1095-1104 is_inside_boundary(original_v, vertex, boundary, left_v, right_v) */
static int is_inside_boundary(const OEnv *e, P2 orig, P2 cand, const int *bd, int m, int left, int right) {
    for (int i = 0; i < m; i++) {
        int a = bd[i], b = bd[(i + m - 1) % m];
        if ((left == a || left == b) && (right == a || right == b)) continue;
        int s1 = cw_angle(cand, VP(e, a), VP(e, b)) < PI;
        int s2 = cw_angle(orig, VP(e, a), VP(e, b)) < PI;
        if (s1 != s2) return 0;
    }
    return 1;
}

/* the candidate position keeps the vertex on the same side of every edge of its one-ring (M:984-997 and its twins) */
static int ring_test(const OEnv *e, int vid, P2 cand, int left, int right, int *w1, int *w2, int *w3, int *w4) {
    int k = connected_vertices(e, vid, w1);
    int m = clockwise_vertices(e, vid, w1, k, w2, w3, w4);
    return is_inside_boundary(e, VP(e, vid), cand, w2, m, left, right);
}

/* the two intersections of the circle |p - (a, b)| = r with the line through (a, b)-offsets used by M:856-866 / M:876-891
 * / M:921-934: x from the quadratic (M^2 + 1) x^2 - t x + c = 0 exactly as the reference writes it */
static void quad_roots(double M, double N, double t, double c4, double *x1, double *x2) {
    double den = 2 * (sq(M) + 1);
    double disc = sqrt(fabs(sq(t) - 4 * (sq(M) + 1) * c4));
    *x1 = (t + disc) / den;
    *x2 = (t - disc) / den;
}

/* M:832-864 middle_vertex(vertex, left_v, right_v, target_angle) */
static P2 middle_vertex(P2 vertex, P2 left, P2 right, double target_angle) {
    P2 m = { (left.x + right.x) / 2, (left.y + right.y) / 2 };
    double A = right.x - left.x, B = right.y - left.y;
    double D = pdist(left, m) / tan(RAD(target_angle / 2));
    double x1, x2, y1, y2;
    if (B == 0) { x1 = m.x; x2 = m.x; y1 = m.y + D; y2 = m.y - D; }
    else if (A == 0) { x1 = m.x + D; x2 = m.x - D; y1 = m.y; y2 = m.y; }
    else {
        double M = -A / B;
        double N = A * m.x / B + m.y;
        double t = -2 * M * N + 2 * m.x + 2 * M * m.y;
        quad_roots(M, N, t, sq(N - m.y) + sq(m.x) - sq(D), &x1, &x2);
        y1 = M * x1 + N; y2 = M * x2 + N;
    }
    P2 V1 = { x1, y1 }, V2 = { x2, y2 };
    return pdist(V1, vertex) < pdist(V2, vertex) ? V1 : V2;
}

/* shared by side_vertex (M:866-893) and indention_vertex (M:908-935): the point at distance `dist` from (a, b) whose
 * projection on (A, B) is W; *nan_out is set when the reference would raise (math.sqrt of a negative number) */
static void circle_line(double a, double b, double A, double B, double W, double dist, P2 *V1, P2 *V2, int *nan_out) {
    double x1, x2, y1, y2;
    if (B == 0) {
        double r = sq(dist) - sq(W / A);
        if (r < 0) *nan_out = 1;
        x1 = W / A + a; x2 = W / A + a;
        y1 = b + sqrt(r); y2 = b - sqrt(r);
    } else if (A == 0) {
        double r = sq(dist) - sq(W / B);
        if (r < 0) *nan_out = 1;
        x1 = a + sqrt(r); x2 = a - sqrt(r);
        y1 = W / B + b; y2 = W / B + b;
    } else {
        double M = -A / B;
        double N = (W + A * a + B * b) / B;
        double t = 2 * M * b - 2 * M * N + 2 * a;
        quad_roots(M, N, t, sq(N - b) + sq(a) - sq(dist), &x1, &x2);
        y1 = M * x1 + N; y2 = M * x2 + N;
    }
    V1->x = x1; V1->y = y1; V2->x = x2; V2->y = y2;
}
/* M:866-893 side_vertex(vertex, next_v, nn_v, angle, dist) */
static P2 side_vertex(P2 vertex, P2 next, P2 nn, double angle, double dist, int *nan_out) {
    P2 V1, V2;
    double W = dist * pdist(next, nn) * cos(RAD(angle));
    circle_line(next.x, next.y, nn.x - next.x, nn.y - next.y, W, dist, &V1, &V2, nan_out);
    return pdist(V1, vertex) < pdist(V2, vertex) ? V1 : V2;
}
/* M:908-935 indention_vertex(vertex, left_v, right_v, angle, dist) */
static P2 indention_vertex(P2 vertex, P2 left, P2 right, double angle, double dist, int *nan_out) {
    P2 V1, V2;
    double W = dist * pdist(vertex, left) * cos(RAD(angle));
    circle_line(vertex.x, vertex.y, left.x - vertex.x, left.y - vertex.y, W, dist, &V1, &V2, nan_out);
    return cw_angle(V1, left, right) < cw_angle(V2, left, right) ? V1 : V2;
}

/* M:937-963 find_side_vertex(vertex, _next_v, next_v, nn_v, v_angle) */
static P2 find_side_vertex(OEnv *e, int vid, int _next, int next, int nn, double v_angle, int *w1, int *w2, int *w3, int *w4) {
    P2 vertex = VP(e, vid);
    double dist = (pdist(vertex, VP(e, _next)) + pdist(vertex, VP(e, next)) + pdist(VP(e, next), VP(e, nn))) / 3;
    double target_angle = 45;
    P2 n_v;
    for (;;) {
        int bad = 0;
        n_v = side_vertex(vertex, VP(e, next), VP(e, nn), target_angle, dist, &bad);
        if (bad) { e->crashed = 1; return vertex; }
        if (target_angle <= v_angle) return vertex;              /* failed */
        if (ring_test(e, vid, n_v, _next, next, w1, w2, w3, w4)) return n_v;
        target_angle -= 5;
    }
}

/* M:1062-1093 find_indention_vertex(vertex, v_angle) */
static P2 find_indention_vertex(OEnv *e, int vid, double v_angle, int *w1, int *w2, int *w3, int *w4) {
    int index = B_index(e, vid), n = e->n;
    int left = e->B[(index + 1) % n], right = BI(e, index - 1);
    P2 vertex = VP(e, vid), lp = VP(e, left), rp = VP(e, right);
    double dist = (pdist(vertex, lp) + pdist(vertex, rp)) / 2;
    /* C:396-413 get_closet_points(..., exclusion = [B[i-2], right, left, B[i+2]], S_T = dist): only whether it is empty
     * matters; a vertex at distance exactly S_T counts (<=) */
    int ex[4] = { BI(e, index - 2), right, left, e->B[(index + 2) % n] };
    int any = 0;
    for (int i = 0; i < n && !any; i++) {
        int id = e->B[i];
        if (id == vid || id == ex[0] || id == ex[1] || id == ex[2] || id == ex[3]) continue;
        if (pdist(vertex, VP(e, id)) <= dist) any = 1;
    }
    /* M:1129-1138 find_closest_segments(updated_boundary, vertex, dist) */
    for (int i = 0; i < n && !any; i++) {
        int p1 = BI(e, i - 1), p2 = e->B[i];
        if (p1 == vid || p2 == vid) continue;
        P2 a = VP(e, p1), b = VP(e, p2);
        double A = b.x - a.x, B = b.y - a.y;                      /* C:642-649 perpendicular_point */
        double s = (A * vertex.x + B * vertex.y - B * a.y - A * a.x) / (sq(A) + sq(B));
        P2 target = { a.x + s * A, a.y + s * B };
        if (0 <= s && s <= 1 && pdist(vertex, target) <= dist) any = 1;
    }
    if (!any) return vertex;
    int times = 4;
    for (;;) {
        int bad = 0;
        P2 n_v = indention_vertex(vertex, lp, rp, (360 - v_angle) / 2, dist / times, &bad);
        if (bad) { e->crashed = 1; return vertex; }
        if (times >= 10) return vertex;                          /* failed */
        if (ring_test(e, vid, n_v, left, right, w1, w2, w3, w4)) return n_v;
        times += 1;
    }
}

/* M:895-906 inner_vertex(vertex, angle) */
static P2 inner_vertex(const OEnv *e, int vid, double angle) {
    int index = B_index(e, vid), n = e->n;
    P2 left = VP(e, e->B[(index + 1) % n]), right = VP(e, BI(e, index - 1)), vertex = VP(e, vid);
    P2 m = { (left.x + right.x) / 2, (left.y + right.y) / 2 };
    double d = pdist(m, right) * tan(RAD(angle));
    double A = vertex.x - m.x, B = vertex.y - m.y;
    double s = sqrt(sq(d) / (sq(A) + sq(B)));
    P2 r = { m.x + s * A, m.y + s * B };
    return r;
}

/* C:475-481 Boundary2D.compute_boundary_angle */
static double boundary_angle_deg(const OEnv *e, int index) {
    int n = e->n;
    index = ((index % n) + n) % n;
    return DEG(cw_angle(BP(e, index), VP(e, e->B[(index + 1) % n]), BP(e, index - 1)));
}

/* M:965-1060 smooth_current_boundary_3 */
static void smooth_current_boundary_3(OEnv *e) {
    int cap = max_degree(e) + 2;
    int *w1 = (int *)malloc(sizeof(int) * cap), *w2 = (int *)malloc(sizeof(int) * 2 * cap), *w3 = (int *)malloc(sizeof(int) * cap),
        *w4 = (int *)malloc(sizeof(int) * cap);
    for (int i = 0; i < e->n && !e->crashed; i++) {
        int n = e->n, vid = e->B[i];
        if (vid < e->n0) continue;                                /* in self.original_vertices */
        int left = e->B[(i + 1) % n], right = BI(e, i - 1);
        double v_angle = DEG(cw_angle(VP(e, vid), VP(e, left), VP(e, right)));
        if (v_angle <= 90) {
            double target_angle = v_angle >= 45 ? v_angle : 45;
            for (;;) {
                P2 new_v = middle_vertex(VP(e, vid), VP(e, left), VP(e, right), target_angle);
                if (target_angle >= 135) break;                   /* failed */
                if (ring_test(e, vid, new_v, left, right, w1, w2, w3, w4)) { e->v[vid].x = new_v.x; e->v[vid].y = new_v.y; break; }
                target_angle += 5;
            }
        } else if (v_angle <= 180) {
            double left_angle = boundary_angle_deg(e, i + 1), right_angle = boundary_angle_deg(e, i - 1);
            P2 n_v;
            if (right_angle < 45) n_v = find_side_vertex(e, vid, left, right, BI(e, i - 2), right_angle, w1, w2, w3, w4);
            else if (left_angle < 45) n_v = find_side_vertex(e, vid, right, left, e->B[(i + 2) % n], left_angle, w1, w2, w3, w4);
            else n_v = find_indention_vertex(e, vid, v_angle, w1, w2, w3, w4);
            e->v[vid].x = n_v.x; e->v[vid].y = n_v.y;
        } else if (v_angle <= 270) {
            P2 n_v = find_indention_vertex(e, vid, v_angle, w1, w2, w3, w4);
            e->v[vid].x = n_v.x; e->v[vid].y = n_v.y;
        } else {
            P2 n_v = inner_vertex(e, vid, 45);
            e->v[vid].x = n_v.x; e->v[vid].y = n_v.y;
            n_v = find_indention_vertex(e, vid, v_angle, w1, w2, w3, w4);
            e->v[vid].x = n_v.x; e->v[vid].y = n_v.y;
        }
    }
    free(w1); free(w2); free(w3); free(w4);
}

/* M:1284-1316 smooth_fixed_vertices(vertices not on the front, in self.boundary.vertices order = id order, 400) */
static void smooth_fixed_vertices(OEnv *e, int iteration) {
    int cap = max_degree(e) + 2;
    int *cv = (int *)malloc(sizeof(int) * cap);
    double sum_coordinates = 0, diffs = 100;
    int it = 0;
    while (diffs > 0.001 && it < iteration) {
        it++;
        double new_sum = 0;
        for (int vid = 0; vid < e->nv; vid++) {
            if (in_B(e, vid) || vid < e->n0) continue;
            double x = 0, y = 0;
            int count = 0;
            int k = connected_vertices(e, vid, cv);
            for (int q = 0; q < k; q++) {
                x += e->v[cv[q]].x + e->v[vid].x;
                y += e->v[cv[q]].y + e->v[vid].y;
                count++;
            }
            if (count == 0) continue;
            e->v[vid].x = x / (2 * count);
            e->v[vid].y = y / (2 * count);
            new_sum += e->v[vid].x + e->v[vid].y;
        }
        diffs = fabs(new_sum - sum_coordinates);
        sum_coordinates = new_sum;
    }
    free(cv);
}

static void smooth_pave(OEnv *e) {
    smooth_current_boundary_3(e);
    smooth_fixed_vertices(e, 400);
    find_reference_candidates(e);
    e->n_smoothings++;
}

/* E:459-594 move(new_point = (r, phi) polar in units of radius * base_length, type).  Returns through the pointers:
 * done, is_complete, needs_smoothing (the reference would call smooth_pave here; that path is not restated: the
 * caller stops comparing).  Reward is always 0.  new_point components are Python floats (CPython round). */
void oracle_move(OEnv *e, const double *polar, double type, int *done_out, int *complete_out, int *smooth_out) {
    int done = 0, not_valid_element = 1;
    e->needs_smoothing = 0;
    e->static_obs = 1;
    int rid = e->ref_id;
    int n = e->n;
    e->last_rule = 2; e->last_success = 0; e->last_inside = -1; e->last_existing = 0;
    if (rid < 0) { e->crashed = 1; *done_out = 1; *complete_out = 0; *smooth_out = 0; return; }
    int index = B_index(e, rid);
    P2 ref = VP(e, rid);
    const double radius = 4;
    double x = e->base_length * radius * polar[0] * cos(polar[1]);
    double y = e->base_length * radius * polar[0] * sin(polar[1]);
    double px = py_round6(x), py = py_round6(y);
    /* E:202-210 detransformation(point, 1, v1, v2) + D:112-137 */
    P2 p1 = BP(e, index - 1);
    double theta = 2 * PI - atan2(p1.y - ref.y, p1.x - ref.x);
    double ox = cos(theta) * px + sin(theta) * py;
    double oy = -sin(theta) * px + cos(theta) * py;
    ox *= 1; oy *= 1;
    ox += ref.x; oy += ref.y;
    P2 newp = { np_round4(ox), np_round4(oy) };

    if (n <= 5) {
        /* the reference leaves is_complete unbound here and raises; sentinel: done, complete iff n <= 4 */
        *done_out = 1; *complete_out = n <= 4; *smooth_out = 0;
        return;
    }
    int mid[4]; P2 m[4]; int have_mesh = 1; int n_new = 0, new_pos = -1;
    if (type <= 0.3) {
        e->last_rule = -1;
        mid[0] = BI(e, index - 1); mid[1] = e->B[index]; mid[2] = e->B[(index + 1) % n]; mid[3] = e->B[(index + 2) % n];
    } else if (type >= 1 - 0.3) {
        e->last_rule = 1;
        mid[0] = BI(e, index - 2); mid[1] = BI(e, index - 1); mid[2] = e->B[index]; mid[3] = e->B[(index + 1) % n];
    } else {
        e->last_rule = 0;
        int inside = point_inside(e, newp);
        e->last_inside = inside;
        if (inside) {
            mid[0] = -2; mid[1] = BI(e, index - 1); mid[2] = e->B[index]; mid[3] = e->B[(index + 1) % n];
            n_new = 1; new_pos = 0;
        } else have_mesh = 0;
    }
    if (have_mesh) {
        for (int k = 0; k < 4; k++) m[k] = mid[k] == -2 ? newp : VP(e, mid[k]);
        if (mesh_is_valid(m) && !check_intersection_with_boundary(e, mid, m, rid)) {
            if (n_new) mid[new_pos] = new_vertex(e, newp.x, newp.y, 1);
            connect_vertices(e, mid);
            append_element(e, mid);
            not_valid_element = 0;
            e->last_success = 1;
            update_boundary(e, mid, n_new, new_pos);
            find_next_state(e);                               /* with the not-valid points collected so far */
            if (e->n <= 5) {
                done = 1;
                if (e->n == 4) { int fm[4] = { e->B[0], e->B[1], e->B[2], e->B[3] }; append_element(e, fm); }
            }
        }
    }
    if (not_valid_element) {
        excl_push(e, ref.x, ref.y, rid);  /* `reference_point not in not_valid_points`: it cannot be selected twice */
        find_next_state(e);
    } else e->nexcl = 0;
    int is_complete;
    if (e->n > 4) {
        is_complete = 0;
        if (e->obs_none) {
            e->needs_smoothing = 1;
            if (!e->smooth_enabled) done = 1;                 /* the caller stops here */
            else {                                            /* E:548-583 */
                smooth_pave(e);
                if (e->nlast > 0 && e->nexcl > 0 && e->last_excl_id[0] == e->excl_id[0] &&
                    e->last_excl_id[e->nlast - 1] == e->excl_id[e->nexcl - 1] && e->nexcl == e->nlast)
                    done = 1;
                if (e->nexcl > e->caplast) { e->caplast = e->nexcl + 16; e->last_excl_id = (int *)realloc(e->last_excl_id, sizeof(int) * e->caplast); }
                memcpy(e->last_excl_id, e->excl_id, sizeof(int) * e->nexcl);
                e->nlast = e->nexcl;
                e->nexcl = 0;
                find_next_state(e);
                if (e->obs_none) done = 1;
            }
        }
    } else is_complete = 1;
    *done_out = done; *complete_out = is_complete; *smooth_out = e->needs_smoothing;
}
int oracle_n_excluded(const OEnv *e) { return e->nexcl; }
void oracle_set_smoothing(OEnv *e, int enabled) { e->smooth_enabled = enabled; }
int oracle_n_smoothings(const OEnv *e) { return e->n_smoothings; }

/* ---------------------------------------------------------------- accessors ------------ */

int oracle_n(const OEnv *e) { return e->n; }
int oracle_n_elements(const OEnv *e) { return e->n_elements; }
int oracle_ref_index(const OEnv *e) { return e->ref_id < 0 ? -1 : B_index(e, e->ref_id); }
int oracle_failed_num(const OEnv *e) { return e->failed_num; }
int oracle_obs_none(const OEnv *e) { return e->obs_none; }
int oracle_crashed(const OEnv *e) { return e->crashed; }
double oracle_base_length(const OEnv *e) { return e->base_length; }
double oracle_current_area(const OEnv *e) { return e->current_area; }
double oracle_original_area(const OEnv *e) { return e->original_area; }
void oracle_area_range(const OEnv *e, double *out) { out[0] = e->area_min; out[1] = e->area_crit; }
void oracle_obs(const OEnv *e, float *out) { memcpy(out, e->obs, sizeof(float) * 18); }
void oracle_last_info(const OEnv *e, int *out) { out[0] = e->last_rule; out[1] = e->last_success; out[2] = e->last_inside; out[3] = e->last_existing; }
void oracle_boundary(const OEnv *e, int *ids, double *xy) {
    for (int i = 0; i < e->n; i++) {
        if (ids) ids[i] = e->B[i];
        if (xy) { xy[2 * i] = e->v[e->B[i]].x; xy[2 * i + 1] = e->v[e->B[i]].y; }
    }
}
void oracle_debug(const OEnv *e, double *out) { memcpy(out, e->dbg, sizeof e->dbg); }
int oracle_n_candidates(const OEnv *e) { return e->ncand; }
void oracle_candidates(const OEnv *e, int *ids, double *keys) {
    for (int i = 0; i < e->ncand; i++) { ids[i] = e->cand_id[i]; keys[i] = e->cand_key[i]; }
}
void oracle_elements(const OEnv *e, int *out) { memcpy(out, e->elements, sizeof(int) * 4 * e->n_elements); }
int oracle_n_vertices(const OEnv *e) { return e->nv; }
void oracle_vertex_xy(const OEnv *e, double *xy) {
    for (int i = 0; i < e->nv; i++) { xy[2 * i] = e->v[i].x; xy[2 * i + 1] = e->v[i].y; }
}

/* T steps with auto-reset (SB3 VecEnv convention: on done the returned obs is the reset obs and
 * the last obs of the episode goes to terminal_obs).  Any output pointer may be NULL. */
void oracle_rollout(OEnv *e, const float *actions, int T, float *obs, double *reward, uint8_t *terminated,
                    uint8_t *truncated, int32_t *n_elements, int32_t *n_boundary, int32_t *ref_index,
                    float *terminal_obs, uint8_t *success) {
    for (int t = 0; t < T; t++) {
        double r; int te, tr;
        oracle_step(e, actions + 3 * t, &r, &te, &tr);
        if (reward) reward[t] = r;
        if (terminated) terminated[t] = (uint8_t)te;
        if (truncated) truncated[t] = (uint8_t)tr;
        if (n_elements) n_elements[t] = e->n_elements;
        if (success) success[t] = (uint8_t)e->last_success;
        if (n_boundary) n_boundary[t] = e->n;                       /* before the auto-reset */
        if (ref_index) ref_index[t] = oracle_ref_index(e);
        if (te || tr || e->obs_none) {
            if (terminal_obs) {
                if (e->obs_none) memset(terminal_obs + 18 * t, 0, sizeof(float) * 18);
                else memcpy(terminal_obs + 18 * t, e->obs, sizeof(float) * 18);
            }
            if (e->obs_none && !(te || tr) && truncated) truncated[t] = 1;   /* sentinel: see DESIGN.md */
            oracle_reset(e);
        } else if (terminal_obs) memset(terminal_obs + 18 * t, 0, sizeof(float) * 18);
        if (obs) memcpy(obs + 18 * t, e->obs, sizeof(float) * 18);
    }
}

/* throughput probe for the CPU baseline: uniform actions from a splitmix/xorshift stream */
static inline uint64_t rng_next(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
long oracle_run_random(OEnv *e, uint64_t seed, long steps, long *n_success, long *n_episodes, double *sum_n) {
    long succ = 0, eps = 0; double sn = 0;
    const float lo[3] = { -1.0f, -1.5f, 0.0f }, hi[3] = { 1.0f, 1.5f, 1.5f };
    for (long t = 0; t < steps; t++) {
        float a[3];
        for (int k = 0; k < 3; k++) {
            double u = (double)(rng_next(&seed) >> 11) / 9007199254740992.0;
            a[k] = (float)(lo[k] + (hi[k] - lo[k]) * u);
        }
        double r; int te, tr;
        sn += e->n;
        oracle_step(e, a, &r, &te, &tr);
        succ += e->last_success;
        if (te || tr || e->obs_none) { eps++; oracle_reset(e); }
    }
    if (n_success) *n_success = succ;
    if (n_episodes) *n_episodes = eps;
    if (sum_n) *sum_n = sn;
    return steps;
}

/* self-test hook: fast py_round4 against the printf/strtod formulation */
long oracle_selftest_round(uint64_t seed, long N) {
    long bad = 0;
    for (long i = 0; i < N; i++) {
        double u = (double)(rng_next(&seed) >> 11) / 9007199254740992.0;
        double x = (u - 0.5) * 40.0;
        if (i % 3 == 0) x = nearbyint(x * 1e5) / 1e5;        /* 5-decimal values: exercise ties */
        if (i % 7 == 0) x = (nearbyint(x * 1e4) + 0.5) / 1e4;
        if (py_round4(x) != py_round4_slow(x)) bad++;
    }
    return bad;
}
double oracle_py_sum(const double *x, int n) { return py_sum(x, n); }
double oracle_py_round4(double x) { return py_round4(x); }
double oracle_np_round4(double x) { return np_round4(x); }
float oracle_np_round4f(float x) { return np_round4f(x); }
double oracle_cw_angle(const double *p) { P2 c = { p[0], p[1] }, a = { p[2], p[3] }, b = { p[4], p[5] }; return cw_angle(c, a, b); }
int oracle_is_cross(const double *p) {
    P2 a = { p[0], p[1] }, b = { p[2], p[3] }, c = { p[4], p[5] }, d = { p[6], p[7] };
    return is_cross(a, b, c, d);
}
