// smooth_pave (M:816-821; legacy general/mesh.py:790-1067, 1258-1288), reached from BoudaryEnv.move() when every
// reference candidate is on the not-valid list (E:548-583):
//   smooth_current_boundary_3()                    every inserted vertex of the front is re-positioned          (M:965-1060)
//   smooth_fixed_vertices(interior vertices, 400)  Gauss-Seidel averaging over Vertex.segments                  (M:1284-1316)
//   find_reference_candidates(0)                   full candidate rebuild                                        (M:259-287)
// One warp per env that needs it; the smoothing itself is a serial algorithm over the whole mesh graph (a vertex moves
// on the coordinates its neighbours have at that moment) and runs on lane 0 over a scratch slab in global memory --
// this is the data-generation API (general/EBRD.py), not a throughput path.  The mesh graph is not kept by the step
// kernels: it is rebuilt here from the element log in the order the reference creates its Segments
// (Boundary2D.deep_copy C:221-228, Mesh.connect_vertices C:840-845), because Vertex.get_connected_vertices (C:123-132)
// -- and with it clockwise_vertices and the summation order of the averaging -- follows that order.
//
// Tolerance contract: tan / cos / sqrt / atan2 of non-quantised arguments come from the device
// libm here and from glibc in the reference, and the reference's x ** 2 (libm pow) is x * x here: smoothed coordinates
// agree to <= 1e-9 (observed <= 8.2e-13 over 589 smoothings); every discrete outcome (which construction a vertex gets, accepted or not,
// iteration counts, done flags, element counts, next reference point) is compared exactly, as are the 4-decimal
// observations that follow.
#pragma once

namespace mg {

constexpr int SM_MAXDEG = 16;          // Segments per vertex the scratch slab holds (a paved quad mesh has 3..6)

// scratch slab of one smoothing (global memory); V = vertices of the episode (original + inserted)
struct SmoothScratch {
    double2 *pool;      // [V] current coordinates by vertex id
    int *deg;           // [V]
    int *adj;           // [V][SM_MAXDEG] partner ids in Segment-creation order
    int *front;         // [cap] ids of the front (updated_boundary.vertices)
    uint8_t *onfront;   // [V]
    int n, n0, nv;
    bool failed;        // degree overflow / a construction the reference would raise on
};
__host__ __device__ inline size_t smooth_scratch_bytes(int cap, int ins_cap) {
    const size_t V = (size_t)cap + ins_cap;
    return ((V * 16 + V * 4 + V * SM_MAXDEG * 4 + (size_t)cap * 4 + V + 255) / 256) * 256;
}

__device__ __forceinline__ P2 sm_p(const SmoothScratch &S, int id) { const double2 v = S.pool[id]; return mk(v.x, v.y); }
__device__ __forceinline__ double sm_deg(double x) { return x * (180.0 / PI); }      // math.degrees
__device__ __forceinline__ double sm_rad(double x) { return x * (PI / 180.0); }      // math.radians
__device__ __forceinline__ double sm_sq(double x) { return x * x; }                  // x ** 2 (see the header comment)
__device__ __forceinline__ int sm_front(const SmoothScratch &S, int i) {             // Python list indexing, i in [-n, 2n)
    return S.front[i < 0 ? i + S.n : (i >= S.n ? i - S.n : i)];
}
__device__ __forceinline__ int sm_front_index(const SmoothScratch &S, int id) {
    for (int i = 0; i < S.n; i++)
        if (S.front[i] == id) return i;
    return -1;
}
__device__ __forceinline__ bool sm_has(const SmoothScratch &S, int a, int b) {
    for (int k = 0; k < S.deg[a]; k++)
        if (S.adj[a * SM_MAXDEG + k] == b) return true;
    return false;
}
__device__ __forceinline__ void sm_connect(SmoothScratch &S, int a, int b) {        // one Segment(a, b), assigned to a then b
    if (S.deg[a] >= SM_MAXDEG || S.deg[b] >= SM_MAXDEG) { S.failed = true; return; }
    S.adj[a * SM_MAXDEG + S.deg[a]++] = b;
    S.adj[b * SM_MAXDEG + S.deg[b]++] = a;
}

// M:1106-1127 clockwise_vertices(inner_v, connected_vs) -> fin (at most 2 k entries)
__device__ __noinline__ int sm_clockwise_vertices(const SmoothScratch &S, int inner, int *vs, int k, int *fin) {
    const P2 c = sm_p(S, inner);
    for (int i = 1; i < k; i++) {
        double max_angle = -1;
        int flag = i;
        for (int j = i; j < k; j++) {
            const double ang = cw_angle(c, sm_p(S, vs[j]), sm_p(S, vs[i - 1]));
            if (ang > max_angle) { max_angle = ang; flag = j; }
        }
        if (flag != i) { const int t = vs[i]; vs[i] = vs[flag]; vs[flag] = t; }
    }
    int m = 0;
    for (int i = 0; i < k; i++) {
        const int cur = vs[i], prev = vs[(i + k - 1) % k];
        int inter = -1;
        for (int x = 0; x < S.deg[cur] && inter < 0; x++) {
            const int cand = S.adj[cur * SM_MAXDEG + x];
            if (cand != inner && sm_has(S, prev, cand)) inter = cand;
        }
        fin[m++] = prev;
        if (inter >= 0) fin[m++] = inter;
    }
    return m;
}

// M:1095-1104 is_inside_boundary on the one-ring of `vid` (M:984-997 and its twins): does the candidate position keep the
// vertex on the same side of every edge of its one-ring?
__device__ __noinline__ bool sm_ring_test(const SmoothScratch &S, int vid, P2 cand, int left, int right) {
    int vs[SM_MAXDEG], fin[2 * SM_MAXDEG];
    const int k = S.deg[vid];
    for (int q = 0; q < k; q++) vs[q] = S.adj[vid * SM_MAXDEG + q];
    const int m = sm_clockwise_vertices(S, vid, vs, k, fin);
    const P2 orig = sm_p(S, vid);
    for (int i = 0; i < m; i++) {
        const int a = fin[i], b = fin[(i + m - 1) % m];
        if ((left == a || left == b) && (right == a || right == b)) continue;
        const bool s1 = cw_angle(cand, sm_p(S, a), sm_p(S, b)) < PI;
        const bool s2 = cw_angle(orig, sm_p(S, a), sm_p(S, b)) < PI;
        if (s1 != s2) return false;
    }
    return true;
}

__device__ __forceinline__ void sm_quad_roots(double M, double t, double c4, double &x1, double &x2) {
    const double den = 2 * (sm_sq(M) + 1);
    const double disc = sqrt(fabs(sm_sq(t) - 4 * (sm_sq(M) + 1) * c4));
    x1 = (t + disc) / den;
    x2 = (t - disc) / den;
}

// M:832-864 middle_vertex(vertex, left_v, right_v, target_angle)
__device__ __noinline__ P2 sm_middle_vertex(P2 vertex, P2 left, P2 right, double target_angle) {
    const P2 m = mk((left.x + right.x) / 2, (left.y + right.y) / 2);
    const double A = right.x - left.x, B = right.y - left.y;
    const double D = pdist(left, m) / tan(sm_rad(target_angle / 2));
    double x1, x2, y1, y2;
    if (B == 0) { x1 = m.x; x2 = m.x; y1 = m.y + D; y2 = m.y - D; }
    else if (A == 0) { x1 = m.x + D; x2 = m.x - D; y1 = m.y; y2 = m.y; }
    else {
        const double M = -A / B;
        const double N = A * m.x / B + m.y;
        const double t = -2 * M * N + 2 * m.x + 2 * M * m.y;
        sm_quad_roots(M, t, sm_sq(N - m.y) + sm_sq(m.x) - sm_sq(D), x1, x2);
        y1 = M * x1 + N; y2 = M * x2 + N;
    }
    const P2 V1 = mk(x1, y1), V2 = mk(x2, y2);
    return pdist(V1, vertex) < pdist(V2, vertex) ? V1 : V2;
}

// shared by side_vertex (M:866-893) and indention_vertex (M:908-935); bad = the reference would raise (sqrt of a negative)
__device__ __noinline__ void sm_circle_line(double a, double b, double A, double B, double W, double dist, P2 &V1, P2 &V2, bool &bad) {
    double x1, x2, y1, y2;
    if (B == 0) {
        const double r = sm_sq(dist) - sm_sq(W / A);
        if (r < 0) bad = true;
        x1 = W / A + a; x2 = W / A + a;
        y1 = b + sqrt(r); y2 = b - sqrt(r);
    } else if (A == 0) {
        const double r = sm_sq(dist) - sm_sq(W / B);
        if (r < 0) bad = true;
        x1 = a + sqrt(r); x2 = a - sqrt(r);
        y1 = W / B + b; y2 = W / B + b;
    } else {
        const double M = -A / B;
        const double N = (W + A * a + B * b) / B;
        const double t = 2 * M * b - 2 * M * N + 2 * a;
        sm_quad_roots(M, t, sm_sq(N - b) + sm_sq(a) - sm_sq(dist), x1, x2);
        y1 = M * x1 + N; y2 = M * x2 + N;
    }
    V1 = mk(x1, y1); V2 = mk(x2, y2);
}

// M:937-963 find_side_vertex(vertex, _next_v, next_v, nn_v, v_angle)
__device__ __noinline__ P2 sm_find_side_vertex(SmoothScratch &S, int vid, int _next, int next, int nn, double v_angle) {
    const P2 vertex = sm_p(S, vid), pn = sm_p(S, next), pnn = sm_p(S, nn);
    const double dist = (pdist(vertex, sm_p(S, _next)) + pdist(vertex, pn) + pdist(pn, pnn)) / 3;
    double target_angle = 45;
    for (;;) {
        bool bad = false;
        P2 V1, V2;
        const double W = dist * pdist(pn, pnn) * cos(sm_rad(target_angle));                     // M:866-893 side_vertex
        sm_circle_line(pn.x, pn.y, pnn.x - pn.x, pnn.y - pn.y, W, dist, V1, V2, bad);
        if (bad) { S.failed = true; return vertex; }
        const P2 n_v = pdist(V1, vertex) < pdist(V2, vertex) ? V1 : V2;
        if (target_angle <= v_angle) return vertex;                                            // failed
        if (sm_ring_test(S, vid, n_v, _next, next)) return n_v;
        target_angle -= 5;
    }
}

// M:1062-1093 find_indention_vertex(vertex, v_angle)
__device__ __noinline__ P2 sm_find_indention_vertex(SmoothScratch &S, int vid, double v_angle) {
    const int index = sm_front_index(S, vid), n = S.n;
    const int left = sm_front(S, index + 1), right = sm_front(S, index - 1);
    const P2 vertex = sm_p(S, vid), lp = sm_p(S, left), rp = sm_p(S, right);
    const double dist = (pdist(vertex, lp) + pdist(vertex, rp)) / 2;
    // C:396-413 get_closet_points(front, vertex, [B[i-2], right, left, B[i+2]], dist): only whether it is empty matters
    const int e0 = sm_front(S, index - 2), e3 = sm_front(S, index + 2);
    bool any = false;
    for (int i = 0; i < n && !any; i++) {
        const int id = S.front[i];
        if (id == vid || id == e0 || id == right || id == left || id == e3) continue;
        if (pdist(vertex, sm_p(S, id)) <= dist) any = true;
    }
    // M:1129-1138 find_closest_segments(front, vertex, dist) with C:642-649 perpendicular_point
    for (int i = 0; i < n && !any; i++) {
        const int p1 = sm_front(S, i - 1), p2 = S.front[i];
        if (p1 == vid || p2 == vid) continue;
        const P2 a = sm_p(S, p1), b = sm_p(S, p2);
        const double A = b.x - a.x, B = b.y - a.y;
        const double s = (A * vertex.x + B * vertex.y - B * a.y - A * a.x) / (sm_sq(A) + sm_sq(B));
        if (0 <= s && s <= 1 && pdist(vertex, mk(a.x + s * A, a.y + s * B)) <= dist) any = true;
    }
    if (!any) return vertex;
    int times = 4;
    for (;;) {
        bool bad = false;
        P2 V1, V2;
        const double d = dist / times;
        const double W = d * pdist(vertex, lp) * cos(sm_rad((360 - v_angle) / 2));            // M:908-935 indention_vertex
        sm_circle_line(vertex.x, vertex.y, lp.x - vertex.x, lp.y - vertex.y, W, d, V1, V2, bad);
        if (bad) { S.failed = true; return vertex; }
        const P2 n_v = cw_angle(V1, lp, rp) < cw_angle(V2, lp, rp) ? V1 : V2;
        if (times >= 10) return vertex;                                                       // failed
        if (sm_ring_test(S, vid, n_v, left, right)) return n_v;
        times += 1;
    }
}

// M:895-906 inner_vertex(vertex, angle)
__device__ __noinline__ P2 sm_inner_vertex(const SmoothScratch &S, int vid, double angle) {
    const int index = sm_front_index(S, vid);
    const P2 left = sm_p(S, sm_front(S, index + 1)), right = sm_p(S, sm_front(S, index - 1)), vertex = sm_p(S, vid);
    const P2 m = mk((left.x + right.x) / 2, (left.y + right.y) / 2);
    const double d = pdist(m, right) * tan(sm_rad(angle));
    const double A = vertex.x - m.x, B = vertex.y - m.y;
    const double s = sqrt(sm_sq(d) / (sm_sq(A) + sm_sq(B)));
    return mk(m.x + s * A, m.y + s * B);
}

// C:475-481 Boundary2D.compute_boundary_angle of front position `index` (any integer)
__device__ __forceinline__ double sm_boundary_angle_deg(const SmoothScratch &S, int index) {
    index = ((index % S.n) + S.n) % S.n;
    return sm_deg(cw_angle(sm_p(S, S.front[index]), sm_p(S, sm_front(S, index + 1)), sm_p(S, sm_front(S, index - 1))));
}

// M:965-1060 smooth_current_boundary_3
__device__ __noinline__ void sm_smooth_front(SmoothScratch &S) {
    for (int i = 0; i < S.n && !S.failed; i++) {
        const int n = S.n, vid = S.front[i];
        if (vid < S.n0) continue;                                        // in self.original_vertices
        const int left = sm_front(S, i + 1), right = sm_front(S, i - 1);
        const double v_angle = sm_deg(cw_angle(sm_p(S, vid), sm_p(S, left), sm_p(S, right)));
        P2 n_v = sm_p(S, vid);
        if (v_angle <= 90) {
            double target_angle = v_angle >= 45 ? v_angle : 45;
            for (;;) {
                const P2 cand = sm_middle_vertex(sm_p(S, vid), sm_p(S, left), sm_p(S, right), target_angle);
                if (target_angle >= 135) break;                          // failed
                if (sm_ring_test(S, vid, cand, left, right)) { n_v = cand; break; }
                target_angle += 5;
            }
        } else if (v_angle <= 180) {
            const double left_angle = sm_boundary_angle_deg(S, i + 1), right_angle = sm_boundary_angle_deg(S, i - 1);
            if (right_angle < 45) n_v = sm_find_side_vertex(S, vid, left, right, sm_front(S, i - 2), right_angle);
            else if (left_angle < 45) n_v = sm_find_side_vertex(S, vid, right, left, sm_front(S, (i + 2) % n), left_angle);
            else n_v = sm_find_indention_vertex(S, vid, v_angle);
        } else if (v_angle <= 270) {
            n_v = sm_find_indention_vertex(S, vid, v_angle);
        } else {
            const P2 t = sm_inner_vertex(S, vid, 45);
            S.pool[vid] = make_double2(t.x, t.y);
            n_v = sm_find_indention_vertex(S, vid, v_angle);
        }
        S.pool[vid] = make_double2(n_v.x, n_v.y);
    }
}

// M:1284-1316 smooth_fixed_vertices(vertices not on the front, in id order, 400)
__device__ __noinline__ void sm_smooth_interior(SmoothScratch &S, int iteration) {
    double sum_coordinates = 0, diffs = 100;
    int it = 0;
    while (diffs > 0.001 && it < iteration) {
        it++;
        double new_sum = 0;
        for (int vid = S.n0; vid < S.nv; vid++) {
            if (S.onfront[vid]) continue;
            double x = 0, y = 0;
            const int count = S.deg[vid];
            const double2 self = S.pool[vid];
            for (int q = 0; q < count; q++) {
                const double2 c = S.pool[S.adj[vid * SM_MAXDEG + q]];
                x += c.x + self.x;
                y += c.y + self.y;
            }
            if (count == 0) continue;
            const double2 r = make_double2(x / (2 * count), y / (2 * count));
            S.pool[vid] = r;
            new_sum += r.x + r.y;
        }
        diffs = fabs(new_sum - sum_coordinates);
        sum_coordinates = new_sum;
    }
}

}  // namespace mg
