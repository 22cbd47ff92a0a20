/*
 * r3oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's rotated-geometry hot path
 * (SJTU-Thinklab-Det/r3det-pytorch).  It is the checker for the CUDA product path and the
 * "port" CPU baseline; the product never links, loads or calls it.  Every function cites the
 * reference file:line (relative to /root/reference) whose arithmetic it follows, operation by
 * operation, so that with `gcc -O2 -ffp-contract=off` it reproduces the reference's host build
 * bit-for-bit on non-degenerate input.  Pinned against the unmodified reference binaries in
 * oracle/_ref (tests/test_oracle_vs_ref.py) and against tests/golden/ fixtures.
 *
 * Build: python oracle/build.py   ->  oracle/_build/libr3oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct { float x, y; } pt;

static inline pt pt_add(pt a, pt b) { pt r = { a.x + b.x, a.y + b.y }; return r; }
static inline pt pt_sub(pt a, pt b) { pt r = { a.x - b.x, a.y - b.y }; return r; }
static inline pt pt_scale(float s, pt a) { pt r = { s * a.x, s * a.y }; return r; }
static inline float pt_dot(pt a, pt b) { return a.x * b.x + a.y * b.y; }
static inline float pt_cross(pt a, pt b) { return a.x * b.y - b.x * a.y; }

/* ================================================================================================
 * v1 geometry  (r3det/ops/rbbox_geo/src/rbbox_geo_kernel.cu:43-228; byte-identical twins in
 * r3det/ops/rnms/src/rcuda/rnms_kernel.cu:15-200 and r3det/ops/rnms/src/rcpu/rnms_cpu.cpp:11-221)
 * ============================================================================================== */

/* rbbox2points, rbbox_geo_kernel.cu:143-155 — absolute coordinates, rotation +a, cosf/sinf */
static void v1_corners(const float *rb, pt *vs)
{
    float x = rb[0], y = rb[1], w_2 = rb[2] / 2, h_2 = rb[3] / 2, a = rb[4];
    float cosa = cosf(a), sina = sinf(a);
    float wx = cosa * w_2, wy = sina * w_2;
    float hx = -sina * h_2, hy = cosa * h_2;
    vs[0].x = x + wx + hx; vs[0].y = y + wy + hy;
    vs[1].x = x - wx + hx; vs[1].y = y - wy + hy;
    vs[2].x = x - wx - hx; vs[2].y = y - wy - hy;
    vs[3].x = x + wx - hx; vs[3].y = y + wy - hy;
}

/* vertex_in_rbbox, rbbox_geo_kernel.cu:157-175 — strict inequalities */
static int v1_vertices_inside(const pt *v1, const pt *v2, pt *ps)
{
    pt center = pt_scale(0.5f, pt_add(v2[0], v2[2]));
    pt w_vec = pt_scale(0.5f, pt_sub(v2[1], v2[0]));
    pt h_vec = pt_scale(0.5f, pt_sub(v2[2], v2[1]));
    float h_vec_2 = pt_dot(h_vec, h_vec);
    float w_vec_2 = pt_dot(w_vec, w_vec);
    int n = 0;
    for (int i = 0; i < 4; i++) {
        pt pr = pt_sub(v1[i], center);
        if (fabsf(pt_dot(pr, h_vec)) < h_vec_2 && fabsf(pt_dot(pr, w_vec)) < w_vec_2)
            ps[n++] = v1[i];
    }
    return n;
}

/* LinSeg::InterSectWith, rbbox_geo_kernel.cu:94-140 */
static int v1_seg_intersect(pt a1, pt a2, pt b1, pt b2, pt *ps)
{
    pt A = pt_sub(a2, a1), B = pt_sub(b2, b1), Cv = pt_sub(a1, b1);
    if (Cv.x == 0 && Cv.y == 0) { ps[0] = a1; return 1; }
    float D = -pt_cross(A, B);
    if (D != 0) {
        float s = pt_cross(Cv, B) / D;
        float t = -pt_cross(A, Cv) / D;
        if (0 <= s && s < 1 && 0 <= t && t < 1) {       /* head vertex excluded */
            ps[0] = pt_add(a1, pt_scale(s, A));
            return 1;
        }
        return 0;
    }
    if (pt_cross(A, Cv) != 0) return 0;                    /* parallel, not colinear */
    int n = 0;
    float BdtC = pt_dot(B, Cv), BdtB = pt_dot(B, B);
    float AdtnC = -pt_dot(A, Cv), AdtA = pt_dot(A, A);
    if (BdtC >= 0 && BdtC < BdtB) ps[n++] = a1;
    if (AdtnC >= 0 && AdtnC < AdtA) ps[n++] = b1;
    return n;
}

/* rbbox_border_intsec, rbbox_geo_kernel.cu:177-191.  `cap` guards the reference's unchecked
 * u[16] buffer (generic position emits <= 8 points in total). */
static int v1_border_intersections(const pt *v1, const pt *v2, pt *ps, int cap)
{
    int n = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            pt tmp[2];
            int k = v1_seg_intersect(v1[i], v1[(i + 1) & 3], v2[j], v2[(j + 1) & 3], tmp);
            for (int q = 0; q < k && n < cap; q++) ps[n++] = tmp[q];
        }
    return n;
}

/* Point::operator<, rbbox_geo_kernel.cu:77-82 */
static inline int v1_less(pt p, pt q)
{
    if ((p.x == 0 && p.y == 0) && (q.x != 0 || q.y != 0)) return 1;
    return pt_cross(p, q) > 0;
}

/* area, rbbox_geo_kernel.cu:193-228 — 1e-2 de-dup, insertion sort about point 0, signed shoelace */
static float v1_area(pt *dirty, int n_dirty)
{
    const float numthres = (float)1e-2;
    pt vs[24];
    vs[0].x = 0; vs[0].y = 0;
    int n = 1;
    for (int i = 1; i < n_dirty; i++) {
        int clean = 1;
        dirty[i] = pt_sub(dirty[i], dirty[0]);
        for (int j = 0; j < n; j++) {
            pt d = pt_sub(dirty[i], vs[j]);
            if (fabsf(d.x) < numthres && fabsf(d.y) < numthres) { clean = 0; break; }
        }
        if (clean) vs[n++] = dirty[i];
    }
    for (int i = 1; i < n; i++) {
        vs[0] = vs[i];
        int j;
        for (j = i - 1; v1_less(vs[0], vs[j]); j--) vs[j + 1] = vs[j];
        vs[j + 1] = vs[0];
    }
    float a = 0;
    vs[0].x = 0; vs[0].y = 0;
    for (int i = 1; i < n; i++) a += pt_cross(vs[i], vs[(i + 1) % n]);
    return a / 2;
}

/* body of mat_iou_iof_kernel, rbbox_geo_kernel.cu:238-267 */
static float v1_pair(const float *rb1, const float *rb2, int iof)
{
    pt v1[4], v2[4], u[24];
    v1_corners(rb1, v1);
    v1_corners(rb2, v2);
    int n = 0;
    n += v1_vertices_inside(v1, v2, u + n);
    n += v1_vertices_inside(v2, v1, u + n);
    n += v1_border_intersections(v1, v2, u + n, 24 - n);
    if (n >= 3) {
        float s1 = rb1[2] * rb1[3];
        float s2 = rb2[2] * rb2[3];
        float su = v1_area(u, n);
        su = fminf(su, s1);
        su = fminf(su, s2);
        su = fmaxf(su, 0.0f);
        return iof ? su / s1 : su / (s1 + s2 - su);
    }
    return 0.0f;
}

/* ================================================================================================
 * v3 / v2 geometry (detectron2 lineage)
 *   v3: r3det/ops/box_iou_rotated/src/box_iou_rotated_utils.h:55-361 (rotation -a)
 *   v2: r3det/ops/ml_nms_rotated/src/box_iou_rotated_utils.h:55-347  (rotation +a, plain-cross pop test)
 * `sort_kind` 0 = the __CUDACC__ exchange sort (:203-216), 1 = host std::sort (:218-227; an
 * insertion sort for <= 16 points in libstdc++).
 * ============================================================================================== */
enum { ORC_V1 = 1, ORC_V2 = 2, ORC_V3 = 3 };

/* get_rotated_vertices, utils.h:55-74 (v3) / ml utils.h:55-77 (v2) */
static void d2_corners(float xc, float yc, float w, float h, float a, int variant, pt *p)
{
    double theta = a;
    float c2 = (float)cos(theta) * 0.5f;
    float s2 = (float)sin(theta) * 0.5f;
    if (variant == ORC_V3) {
        p[0].x = xc + s2 * h + c2 * w;
        p[0].y = yc + c2 * h - s2 * w;
        p[1].x = xc - s2 * h + c2 * w;
        p[1].y = yc - c2 * h - s2 * w;
    } else {
        p[0].x = xc - s2 * h - c2 * w;
        p[0].y = yc + c2 * h - s2 * w;
        p[1].x = xc + s2 * h - c2 * w;
        p[1].y = yc - c2 * h - s2 * w;
    }
    p[2].x = 2 * xc - p[0].x; p[2].y = 2 * yc - p[0].y;
    p[3].x = 2 * xc - p[1].x; p[3].y = 2 * yc - p[1].y;
}

/* get_intersection_points, utils.h:76-155 — inclusive tests, |det| <= 1e-14 skip */
static int d2_candidates(const pt *p1, const pt *p2, pt *out)
{
    pt e1[4], e2[4];
    for (int i = 0; i < 4; i++) {
        e1[i] = pt_sub(p1[(i + 1) % 4], p1[i]);
        e2[i] = pt_sub(p2[(i + 1) % 4], p2[i]);
    }
    int n = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float det = pt_cross(e2[j], e1[i]);
            if (fabs((double)det) <= 1e-14) continue;
            pt v12 = pt_sub(p2[j], p1[i]);
            float t1 = pt_cross(e2[j], v12) / det;
            float t2 = pt_cross(e1[i], v12) / det;
            if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f)
                out[n++] = pt_add(p1[i], pt_scale(t1, e1[i]));   /* vec1[i] * t1: same products */
        }
    {
        pt AB = e2[0], DA = e2[3];
        float ABdotAB = pt_dot(AB, AB), ADdotAD = pt_dot(DA, DA);
        for (int i = 0; i < 4; i++) {
            pt AP = pt_sub(p1[i], p2[0]);
            float APdotAB = pt_dot(AP, AB), APdotAD = -pt_dot(AP, DA);
            if (APdotAB >= 0 && APdotAD >= 0 && APdotAB <= ABdotAB && APdotAD <= ADdotAD) out[n++] = p1[i];
        }
    }
    {
        pt AB = e1[0], DA = e1[3];
        float ABdotAB = pt_dot(AB, AB), ADdotAD = pt_dot(DA, DA);
        for (int i = 0; i < 4; i++) {
            pt AP = pt_sub(p2[i], p1[0]);
            float APdotAB = pt_dot(AP, AB), APdotAD = -pt_dot(AP, DA);
            if (APdotAB >= 0 && APdotAD >= 0 && APdotAB <= ABdotAB && APdotAD <= ADdotAD) out[n++] = p2[i];
        }
    }
    return n;
}

static inline int d2_host_less(pt A, pt B)      /* utils.h:219-226 */
{
    float t = pt_cross(A, B);
    if (fabs((double)t) < 1e-6) return pt_dot(A, A) < pt_dot(B, B);
    return t > 0;
}

/* convex_hull_graham(shift_to_zero=true), utils.h:157-289 */
static int d2_hull(const pt *p, int num_in, pt *q, int variant, int sort_kind)
{
    int t = 0;
    for (int i = 1; i < num_in; i++)
        if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
    pt start = p[t];
    for (int i = 0; i < num_in; i++) q[i] = pt_sub(p[i], start);
    pt tmp = q[0]; q[0] = q[t]; q[t] = tmp;

    float dist[24];
    if (sort_kind == 0) {
        for (int i = 0; i < num_in; i++) dist[i] = pt_dot(q[i], q[i]);
        for (int i = 1; i < num_in - 1; i++)
            for (int j = i + 1; j < num_in; j++) {
                float cp = pt_cross(q[i], q[j]);
                if ((cp < -1e-6) || (fabs((double)cp) < 1e-6 && dist[i] > dist[j])) {
                    pt qt = q[i]; q[i] = q[j]; q[j] = qt;
                    float dt = dist[i]; dist[i] = dist[j]; dist[j] = dt;
                }
            }
    } else {
        for (int i = 2; i < num_in; i++) {             /* libstdc++ __insertion_sort on [1, num_in) */
            pt val = q[i];
            if (d2_host_less(val, q[1])) {
                memmove(&q[2], &q[1], sizeof(pt) * (size_t)(i - 1));
                q[1] = val;
            } else {
                int j = i;
                while (d2_host_less(val, q[j - 1])) { q[j] = q[j - 1]; j--; }
                q[j] = val;
            }
        }
        for (int i = 0; i < num_in; i++) dist[i] = pt_dot(q[i], q[i]);
    }

    int k;
    for (k = 1; k < num_in; k++)
        if (dist[k] > 1e-8) break;
    if (k == num_in) { q[0] = p[t]; return 1; }
    q[1] = q[k];
    int m = 2;
    for (int i = k + 1; i < num_in; i++) {
        while (m > 1) {
            pt q1 = pt_sub(q[i], q[m - 2]), q2 = pt_sub(q[m - 1], q[m - 2]);
            int pop = (variant == ORC_V3) ? (q1.x * q2.y >= q2.x * q1.y)       /* :258-267 */
                                          : (pt_cross(q1, q2) >= 0);            /* ml :253-254 */
            if (pop) m--; else break;
        }
        q[m++] = q[i];
    }
    return m;
}

/* polygon_area, utils.h:291-303 */
static float d2_area(const pt *q, int m)
{
    if (m <= 2) return 0;
    float area = 0;
    for (int i = 1; i < m - 1; i++)
        area += fabsf(pt_cross(pt_sub(q[i], q[0]), pt_sub(q[i + 1], q[0])));
    return (float)(area / 2.0);
}

/* single_box_iou_rotated, utils.h:331-361.  mode_iou: 1 = IoU, 0 = IoF (flag polarity of
 * box_iou_rotated_ext.overlaps, box_iou_rotated_wrapper.py:50-51). */
static float d2_pair(const float *b1, const float *b2, int mode_iou, int variant, int sort_kind)
{
    double sx = (b1[0] + b2[0]) / 2.0;
    double sy = (b1[1] + b2[1]) / 2.0;
    float x1 = (float)(b1[0] - sx), y1 = (float)(b1[1] - sy);
    float x2 = (float)(b2[0] - sx), y2 = (float)(b2[1] - sy);
    float area1 = b1[2] * b1[3];
    float area2 = b2[2] * b2[3];
    if (area1 < 1e-14 || area2 < 1e-14) return 0.f;

    pt p1[4], p2[4], cand[24], hull[24];
    d2_corners(x1, y1, b1[2], b1[3], b1[4], variant, p1);
    d2_corners(x2, y2, b2[2], b2[3], b2[4], variant, p2);
    int num = d2_candidates(p1, p2, cand);
    float inter = 0.0f;
    if (num > 2) {
        int m = d2_hull(cand, num, hull, variant, sort_kind);
        inter = d2_area(hull, m);
    }
    return mode_iou ? inter / (area1 + area2 - inter) : inter / area1;
}

/* ================================================================================================
 * Exact (float64, Sutherland–Hodgman) intersection — NOT a reference algorithm: the geometric
 * ground truth used to judge which of two disagreeing float32 implementations is closer.
 * ============================================================================================== */
static double exact_pair(const float *b1, const float *b2, int iof, double sign)
{
    double mx = 0.5 * ((double)b1[0] + b2[0]), my = 0.5 * ((double)b1[1] + b2[1]);
    double px[16], py[16], qx[16], qy[16];
    {   /* polygon = box 2 corners (CCW for w,h > 0) */
        double c = cos(sign * (double)b2[4]), s = sin(sign * (double)b2[4]);
        double cx = b2[0] - mx, cy = b2[1] - my, hw = 0.5 * b2[2], hh = 0.5 * b2[3];
        double ux = c * hw, uy = s * hw, vx = -s * hh, vy = c * hh;
        px[0] = cx - ux - vx; py[0] = cy - uy - vy;
        px[1] = cx + ux - vx; py[1] = cy + uy - vy;
        px[2] = cx + ux + vx; py[2] = cy + uy + vy;
        px[3] = cx - ux + vx; py[3] = cy - uy + vy;
    }
    int n = 4;
    double c = cos(sign * (double)b1[4]), s = sin(sign * (double)b1[4]);
    double cx = b1[0] - mx, cy = b1[1] - my, hw = 0.5 * fabs((double)b1[2]), hh = 0.5 * fabs((double)b1[3]);
    /* four half-planes of box 1: n·(p - c) <= h */
    double nx[4] = { c, -c, -s, s }, ny[4] = { s, -s, c, -c }, lim[4] = { hw, hw, hh, hh };
    for (int e = 0; e < 4 && n > 0; e++) {
        int m = 0;
        for (int i = 0; i < n; i++) {
            int j = (i + 1) % n;
            double di = nx[e] * (px[i] - cx) + ny[e] * (py[i] - cy) - lim[e];
            double dj = nx[e] * (px[j] - cx) + ny[e] * (py[j] - cy) - lim[e];
            if (di <= 0) { qx[m] = px[i]; qy[m] = py[i]; m++; }
            if ((di < 0 && dj > 0) || (di > 0 && dj < 0)) {
                double t = di / (di - dj);
                qx[m] = px[i] + t * (px[j] - px[i]); qy[m] = py[i] + t * (py[j] - py[i]); m++;
            }
        }
        n = m;
        memcpy(px, qx, sizeof(double) * (size_t)n); memcpy(py, qy, sizeof(double) * (size_t)n);
    }
    double a = 0;
    for (int i = 1; i + 1 < n; i++)
        a += (px[i] - px[0]) * (py[i + 1] - py[0]) - (px[i + 1] - px[0]) * (py[i] - py[0]);
    a = 0.5 * fabs(a);
    double s1 = fabs((double)b1[2] * b1[3]), s2 = fabs((double)b2[2] * b2[3]);
    if (s1 <= 0 || s2 <= 0) return 0;
    return iof ? a / s1 : a / (s1 + s2 - a);
}

/* ================================================================================================
 * Public IoU entry points
 * ============================================================================================== */

/* variant 1: rbbox_iou(rb1, rb2, vec=False, iof)  (r3det/ops/rbbox_geo/rbbox_geo.py:4-9)
 * variant 3: box_iou_rotated_ext.overlaps + the wrapper's too-small mask when `wrapper_mask`
 *            (r3det/ops/box_iou_rotated/box_iou_rotated_wrapper.py:50-60)
 * variant 2: mmcv.ops.box_iou_rotated restated from ml_nms_rotated's header (parity unpinned upstream)
 * mode: 0 = iou, 1 = iof. */
ORC_API void orc_iou_matrix_f32(const float *b1, int64_t m, const float *b2, int64_t n,
                                int variant, int mode, int wrapper_mask, int sort_kind, float *out)
{
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++) {
            const float *p = b1 + 5 * i, *q = b2 + 5 * j;
            float r;
            if (variant == ORC_V1) r = v1_pair(p, q, mode == 1);
            else r = d2_pair(p, q, mode == 0, variant, sort_kind);
            if (variant == ORC_V3 && wrapper_mask &&
                (fminf(p[2], p[3]) < 0.001f || fminf(q[2], q[3]) < 0.001f)) r = 0.f;
            out[i * n + j] = r;
        }
}

/* aligned: v1 -> vec_iou_iof_kernel modulo broadcast (rbbox_geo_kernel.cu:270-309), length max(n1,n2) */
ORC_API void orc_iou_aligned_f32(const float *b1, int64_t n1, const float *b2, int64_t n2,
                                 int variant, int mode, int sort_kind, float *out)
{
    int64_t n = n1 > n2 ? n1 : n2;
    for (int64_t i = 0; i < n; i++) {
        const float *p = b1 + 5 * (i % n1), *q = b2 + 5 * (i % n2);
        out[i] = (variant == ORC_V1) ? v1_pair(p, q, mode == 1) : d2_pair(p, q, mode == 0, variant, sort_kind);
    }
}

/* geometric ground truth in float64; angle_sign +1 (v1, v2) or -1 (v3) */
ORC_API void orc_iou_exact_f64(const float *b1, int64_t m, const float *b2, int64_t n,
                               int mode, int angle_sign, double *out)
{
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++)
            out[i * n + j] = exact_pair(b1 + 5 * i, b2 + 5 * j, mode == 1, (double)angle_sign);
}

/* ================================================================================================
 * NMS  (greedy in descending score)
 *   v1: rnms_ext.rnms            rnms_cpu.cpp:223-282 (>=) / rnms_kernel.cu:229-335 (>)
 *   v3: nms_rotated_ext          nms_rotated_cpu.cpp:8-61 (>=) / nms_rotated_cuda.cu:12-134 (>)
 *   v2: ml_nms_rotated           ml_nms_rotated/src/nms_rotated_cpu.cpp:36-57 (>=) / _cuda.cu:13-137 (>)
 * `inclusive` 1 = CPU rule (>=), 0 = GPU rule (>).  labels (float, may be NULL): IoU := 0 across
 * different labels (ml utils.h:317-322).  Output: kept original indices in descending-score order;
 * the v1 wrappers sort them ascending afterwards (rnms_kernel.cu:331-334).
 * ============================================================================================== */
typedef struct { float s; int64_t i; } sidx;
static int cmp_desc(const void *a, const void *b)
{
    const sidx *x = (const sidx *)a, *y = (const sidx *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);            /* stable on ties */
}

ORC_API int64_t orc_nms_f32(const float *boxes, int64_t box_stride, const float *scores, const float *labels,
                            int64_t K, float thr, int variant, int inclusive, int sort_kind, int64_t *keep)
{
    if (K <= 0) return 0;
    sidx *ord = (sidx *)malloc(sizeof(sidx) * (size_t)K);
    uint8_t *sup = (uint8_t *)calloc((size_t)K, 1);
    for (int64_t i = 0; i < K; i++) { ord[i].s = scores[i]; ord[i].i = i; }
    qsort(ord, (size_t)K, sizeof(sidx), cmp_desc);
    int64_t nk = 0;
    for (int64_t a = 0; a < K; a++) {
        int64_t i = ord[a].i;
        if (sup[i]) continue;
        keep[nk++] = i;
        for (int64_t b = a + 1; b < K; b++) {
            int64_t j = ord[b].i;
            if (sup[j]) continue;
            float iou;
            if (labels && labels[i] != labels[j]) iou = 0.0f;
            else if (variant == ORC_V1) iou = v1_pair(boxes + box_stride * i, boxes + box_stride * j, 0);
            else iou = d2_pair(boxes + box_stride * i, boxes + box_stride * j, 1, variant, sort_kind);
            if (inclusive ? (iou >= thr) : (iou > thr)) sup[j] = 1;
        }
    }
    free(ord); free(sup);
    return nk;
}

/* ================================================================================================
 * FRM  (r3det/ops/fr/src/feature_refine_kernel.cu)
 * ============================================================================================== */

/* bilinear_interpolate_gradient, feature_refine_kernel.cu:67-110 (weights/indices shared with
 * bilinear_interpolate :16-65).  Returns 0 when the sample is out of range. */
static int frm_taps(int H, int W, float y, float x, float w[4], int *xl, int *xh, int *yl, int *yh)
{
    if (y < -1.0 || y > H || x < -1.0 || x > W) return 0;
    if (y <= 0) y = 0;
    if (x <= 0) x = 0;
    *yl = (int)y; *xl = (int)x;
    if (*yl >= H - 1) { *yh = *yl = H - 1; y = (float)*yl; } else *yh = *yl + 1;
    if (*xl >= W - 1) { *xh = *xl = W - 1; x = (float)*xl; } else *xh = *xl + 1;
    float ly = y - *yl, lx = x - *xl;
    float hy = (float)(1. - ly), hx = (float)(1. - lx);
    w[0] = hy * hx; w[1] = hy * lx; w[2] = ly * hx; w[3] = ly * lx;
    return 1;
}

/* sample positions for one location, feature_refine_kernel.cu:127-151 (note box x -> row coordinate) */
static void frm_points(const float *bb, float scale, int points, float *px, float *py)
{
    float roi_y = bb[0] * scale;
    float roi_x = bb[1] * scale;
    px[0] = roi_x; py[0] = roi_y;
    if (points > 1) {
        float roi_w = bb[2] * scale, roi_h = bb[3] * scale, roi_a = bb[4];
        float w_2 = roi_w / 2, h_2 = roi_h / 2;
        float cosa = cosf(roi_a), sina = sinf(roi_a);
        float wx = cosa * w_2, wy = sina * w_2;
        float hx = -sina * h_2, hy = cosa * h_2;
        px[1] = roi_x + wx + hx; py[1] = roi_y + wy + hy;
        px[2] = roi_x - wx + hx; py[2] = roi_y - wy + hy;
        px[3] = roi_x - wx - hx; py[3] = roi_y - wy - hy;
        px[4] = roi_x + wx - hx; py[4] = roi_y + wy - hy;
    }
}

/* feature_refine_forward_kernel, feature_refine_kernel.cu:112-163.  out = in + sum_p bilinear(in, y_p, x_p) */
ORC_API void orc_frm_forward_f32(const float *feat, const float *boxes, int N, int Cc, int H, int W,
                                 float scale, int points, float *out)
{
    for (int n = 0; n < N; n++)
        for (int h = 0; h < H; h++)
            for (int w = 0; w < W; w++) {
                float px[5], py[5];
                frm_points(boxes + (((int64_t)n * H + h) * W + w) * 5, scale, points, px, py);
                float wt[5][4]; int xl[5], xh[5], yl[5], yh[5], ok[5];
                for (int p = 0; p < points; p++) ok[p] = frm_taps(H, W, py[p], px[p], wt[p], &xl[p], &xh[p], &yl[p], &yh[p]);
                for (int c = 0; c < Cc; c++) {
                    const float *pl = feat + ((int64_t)n * Cc + c) * H * W;
                    float v = pl[h * W + w];
                    for (int p = 0; p < points; p++) {
                        if (!ok[p]) { v += 0.0f; continue; }
                        float lt = pl[yl[p] * W + xl[p]], rt = pl[yl[p] * W + xh[p]];
                        float lb = pl[yh[p] * W + xl[p]], rb = pl[yh[p] * W + xh[p]];
                        v += (wt[p][0] * lt + wt[p][1] * rt + wt[p][2] * lb + wt[p][3] * rb);
                    }
                    out[((int64_t)n * Cc + c) * H * W + h * W + w] = v;
                }
            }
}

/* feature_refine_backward_kernel, feature_refine_kernel.cu:165-230.  The reference accumulates with
 * float atomics in a nondeterministic order; here the order is the element index order.  `acc64`
 * accumulates in double and rounds once (the tolerance anchor). */
ORC_API void orc_frm_backward_f32(const float *gout, const float *boxes, int N, int Cc, int H, int W,
                                  float scale, int points, int acc64, float *gin)
{
    int64_t total = (int64_t)N * Cc * H * W;
    double *acc = NULL;
    if (acc64) acc = (double *)calloc((size_t)total, sizeof(double));
    else memset(gin, 0, sizeof(float) * (size_t)total);
    for (int n = 0; n < N; n++)
        for (int c = 0; c < Cc; c++)
            for (int h = 0; h < H; h++)
                for (int w = 0; w < W; w++) {
                    int64_t base = ((int64_t)n * Cc + c) * H * W;
                    float g = gout[base + h * W + w];
                    float px[5], py[5];
                    frm_points(boxes + (((int64_t)n * H + h) * W + w) * 5, scale, points, px, py);
                    if (acc64) acc[base + h * W + w] += g; else gin[base + h * W + w] += g;
                    for (int p = 0; p < points; p++) {
                        float wt[4]; int xl, xh, yl, yh;
                        if (!frm_taps(H, W, py[p], px[p], wt, &xl, &xh, &yl, &yh)) continue;
                        float g1 = g * wt[0], g2 = g * wt[1], g3 = g * wt[2], g4 = g * wt[3];
                        if (acc64) {
                            acc[base + yl * W + xl] += g1; acc[base + yl * W + xh] += g2;
                            acc[base + yh * W + xl] += g3; acc[base + yh * W + xh] += g4;
                        } else {
                            gin[base + yl * W + xl] += g1; gin[base + yl * W + xh] += g2;
                            gin[base + yh * W + xl] += g3; gin[base + yh * W + xh] += g4;
                        }
                    }
                }
    if (acc64) {
        for (int64_t i = 0; i < total; i++) gin[i] = (float)acc[i];
        free(acc);
    }
}

/* ================================================================================================
 * polygon IoU / polygon NMS  (r3det/ops/nms_rotated/src/poly_nms_cuda.cu:22-194; the reference has no CPU
 * implementation of this op — nms_rotated_wrapper.py:70-71 raises for CPU tensors — so this restates the CUDA
 * device code; pinned against that code compiled unmodified in oracle/_ref/libref_cuda_polynms.so through
 * tests/golden/poly_refcuda.npz.  The reference's build contracts a*b-c*d into FMAs and this host build does not:
 * agreement is to FP32 cancellation noise of the origin-anchored triangle fans, not bit-for-bit.)
 * ================================================================================================ */
#define POLY_EPS 1E-8
static inline int poly_sig(float d) { return ((double)d > POLY_EPS) - ((double)d < -POLY_EPS); }          /* :25-27 */
static inline int poly_pt_eq(pt a, pt b) { return poly_sig(a.x - b.x) == 0 && poly_sig(a.y - b.y) == 0; } /* :29-31 */
static inline float poly_cross(pt o, pt a, pt b) { return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y); } /* :47-49 */

static float poly_area(pt *ps, int n)                                                                     /* :50-57 */
{
    ps[n] = ps[0];
    float res = 0;
    for (int i = 0; i < n; i++) res += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return (float)(res / 2.0);
}

static int poly_line_cross(pt a, pt b, pt c, pt d, pt *p)                                                 /* :58-67 */
{
    float s1 = poly_cross(a, b, c), s2 = poly_cross(a, b, d);
    if (poly_sig(s1) == 0 && poly_sig(s2) == 0) return 2;
    if (poly_sig(s2 - s1) == 0) return 0;
    p->x = (c.x * s2 - d.x * s1) / (s2 - s1);
    p->y = (c.y * s2 - d.y * s1) / (s2 - s1);
    return 1;
}

static void poly_cut(pt *p, int *n_io, pt a, pt b, pt *pp)                                                /* :69-83 */
{
    int n = *n_io, m = 0;
    p[n] = p[0];
    for (int i = 0; i < n; i++) {
        if (poly_sig(poly_cross(a, b, p[i])) > 0) pp[m++] = p[i];
        if (poly_sig(poly_cross(a, b, p[i])) != poly_sig(poly_cross(a, b, p[i + 1])))
            poly_line_cross(a, b, p[i], p[i + 1], &pp[m++]);
    }
    n = 0;
    for (int i = 0; i < m; i++)
        if (!i || !poly_pt_eq(pp[i], pp[i - 1])) p[n++] = pp[i];
    while (n > 1 && poly_pt_eq(p[n - 1], p[0])) n--;
    *n_io = n;
}

/* signed intersection area of the triangles (o,a,b) and (o,c,d), o = origin                                 :87-105 */
static float poly_tri_intersect(pt a, pt b, pt c, pt d)
{
    pt o = { 0.0f, 0.0f };
    int s1 = poly_sig(poly_cross(o, a, b)), s2 = poly_sig(poly_cross(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.0f;
    if (s1 == -1) { pt t = a; a = b; b = t; }
    if (s2 == -1) { pt t = c; c = d; d = t; }
    pt p[10] = { o, a, b }, pp[10];
    memset(pp, 0, sizeof pp);              /* the reference leaves pp uninitialised (read only if a crossing degenerates) */
    int n = 3;
    poly_cut(p, &n, o, c, pp);
    poly_cut(p, &n, c, d, pp);
    poly_cut(p, &n, d, o, pp);
    float res = fabsf(poly_area(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}

static void poly_reverse(pt *first, pt *last)                                                             /* :39-45 */
{
    while ((first != last) && (first != --last)) { pt t = *first; *first = *last; *last = t; ++first; }
}

static float poly_intersect_area(pt *ps1, int n1, pt *ps2, int n2)                                       /* :107-119 */
{
    if (poly_area(ps1, n1) < 0) poly_reverse(ps1, ps1 + n1);
    if (poly_area(ps2, n2) < 0) poly_reverse(ps2, ps2 + n2);
    ps1[n1] = ps1[0];
    ps2[n2] = ps2[0];
    float res = 0;
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++) res += poly_tri_intersect(ps1[i], ps1[i + 1], ps2[j], ps2[j + 1]);
    return res;
}

static float poly_iou(const float *p, const float *q)                                                     /* :122-143 */
{
    pt ps1[10], ps2[10];
    for (int i = 0; i < 4; i++) {
        ps1[i].x = p[i * 2]; ps1[i].y = p[i * 2 + 1];
        ps2[i].x = q[i * 2]; ps2[i].y = q[i * 2 + 1];
    }
    float inter = poly_intersect_area(ps1, 4, ps2, 4);
    float uni = fabsf(poly_area(ps1, 4)) + fabsf(poly_area(ps2, 4)) - inter;
    return (uni == 0) ? (inter + 1) / (uni + 1) : inter / uni;
}

/* aligned polygon IoU: out[i] = IoU(p[i], q[i]); rows of 8 floats with the given strides */
ORC_API void orc_poly_iou_aligned_f32(const float *p, int64_t ps, const float *q, int64_t qs, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = poly_iou(p + i * ps, q + i * qs);
}

/* greedy polygon NMS (poly_nms_cuda.cu:145-193 mask + :196-262 host scan): order by descending score (stable),
 * suppress j when IoU(kept i, j) > thr; keep holds original indices in descending-score order */
ORC_API int64_t orc_poly_nms_f32(const float *polys, int64_t stride, const float *scores, int64_t K, float thr, int64_t *keep)
{
    if (K <= 0) return 0;
    sidx *ord = (sidx *)malloc(sizeof(sidx) * (size_t)K);
    unsigned char *dead = (unsigned char *)calloc((size_t)K, 1);
    for (int64_t i = 0; i < K; i++) { ord[i].s = scores[i]; ord[i].i = i; }
    qsort(ord, (size_t)K, sizeof(sidx), cmp_desc);
    int64_t n = 0;
    for (int64_t a = 0; a < K; a++) {
        if (dead[a]) continue;
        const int64_t i = ord[a].i;
        keep[n++] = i;
        for (int64_t b = a + 1; b < K; b++)
            if (!dead[b] && poly_iou(polys + i * stride, polys + ord[b].i * stride) > thr) dead[b] = 1;
    }
    free(ord); free(dead);
    return n;
}
