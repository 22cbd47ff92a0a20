// poly.cuh — IoU of 4-point polygons for polygon NMS (the pair evaluator of the GEOM_QUAD instance of nms_rounds.cuh).
//
// Replaces (reference, relative to /root/reference):
//   devPolyIoU + poly_nms_kernel      r3det/ops/nms_rotated/src/poly_nms_cuda.cu:22-194
// The reference evaluates an origin-anchored triangle-fan clipping (16 triangle pairs, each clipped against three lines with
// vertex lists in local memory, ~3000 FP32 operations) for ALL K^2 / 2 pairs.  Here a pair first has to pass a bounding-box
// test, and the pairs that do are evaluated by OWN geometry when both quadrilaterals are convex — the DOTA case: obb2poly
// output — in registers and ~600 instructions:
//     area(P ∩ Q) = 1/2 * sum over the boundary segments of P ∩ Q of cross(start, end)              (Green's theorem)
// where every edge of P contributes its part inside Q and every edge of Q its part strictly inside P (coincident boundaries
// count once); an edge ∩ convex polygon is ONE parameter interval [t0, t1], found by clipping the edge against the four
// half-planes (Cyrus–Beck).  No polygon storage, no vertex counting, coordinates relative to P's first vertex.
// The reference's result carries FP32 cancellation noise from anchoring its fans at the image origin (~1e-5 * max|coord|^2 px^2
// of intersection area); that noise IS the specification wherever it can flip `IoU > thr`.  So pairs whose fast IoU lies within a
// margin of the threshold that covers this noise — and all pairs with a concave / degenerate quadrilateral — are decided by
// `quad_iou` below, a step-by-step restatement of the reference arithmetic (disclosed as such; it runs for a few per cent of
// the overlapping pairs instead of all of them).  Sorting, the greedy rounds and the keep-list emission are nms.cu's.
#pragma once
#include "common.cuh"

namespace r3g {
namespace poly {

struct P2 { float x, y; };

__device__ __forceinline__ int sgn(float d) { return ((double)d > 1E-8) - ((double)d < -1E-8); }               // :25-27
__device__ __forceinline__ bool same_pt(P2 a, P2 b) { return sgn(a.x - b.x) == 0 && sgn(a.y - b.y) == 0; }       // :29-31
__device__ __forceinline__ float crs(P2 o, P2 a, P2 b) { return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y); }

__device__ __forceinline__ float shoelace(P2* ps, int n) {                                                      // :50-57
    ps[n] = ps[0];
    float res = 0;
    for (int i = 0; i < n; i++) res += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return (float)(res / 2.0);
}

__device__ __forceinline__ void crossing(P2 a, P2 b, P2 c, P2 d, P2& p) {                                       // :58-67
    const float s1 = crs(a, b, c), s2 = crs(a, b, d);
    if (sgn(s1) == 0 && sgn(s2) == 0) return;
    if (sgn(s2 - s1) == 0) return;
    p.x = (c.x * s2 - d.x * s1) / (s2 - s1);
    p.y = (c.y * s2 - d.y * s1) / (s2 - s1);
}

// clip polygon p (n vertices) to the left of the directed line a -> b                                           :69-83
__device__ __noinline__ void clip_left(P2* p, int& n, P2 a, P2 b, P2* pp) {
    int m = 0;
    p[n] = p[0];
    for (int i = 0; i < n; i++) {
        const int si = sgn(crs(a, b, p[i]));
        if (si > 0) pp[m++] = p[i];
        if (si != sgn(crs(a, b, p[i + 1]))) crossing(a, b, p[i], p[i + 1], pp[m++]);
    }
    n = 0;
    for (int i = 0; i < m; i++)
        if (!i || !same_pt(pp[i], pp[i - 1])) p[n++] = pp[i];
    while (n > 1 && same_pt(p[n - 1], p[0])) n--;
}

// signed intersection area of the triangles (o, a, b) and (o, c, d), o = origin                                 :87-105
__device__ __forceinline__ float fan_pair(P2 a, P2 b, P2 c, P2 d) {
    const P2 o = { 0.0f, 0.0f };
    const int s1 = sgn(crs(o, a, b)), s2 = sgn(crs(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.0f;
    if (s1 == -1) { const P2 t = a; a = b; b = t; }
    if (s2 == -1) { const P2 t = c; c = d; d = t; }
    P2 p[10], pp[10];
#pragma unroll
    for (int k = 0; k < 10; k++) { p[k] = o; pp[k] = o; }
    p[1] = a; p[2] = b;
    int n = 3;
    clip_left(p, n, o, c, pp);
    clip_left(p, n, c, d, pp);
    clip_left(p, n, d, o, pp);
    float res = fabsf(shoelace(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}

__device__ __forceinline__ void orient_ccw(P2* ps, int n) {                                                     // :108-109, :39-45
    if (shoelace(ps, n) < 0) {
        int i = 0, j = n - 1;
        while (i < j) { const P2 t = ps[i]; ps[i] = ps[j]; ps[j] = t; i++; j--; }
    }
}

// IoU of two quadrilaterals given as 8 floats each                                                              :107-143
__device__ __noinline__ float quad_iou(const float* __restrict__ pa, const float* __restrict__ qa) {
    P2 ps1[6], ps2[6];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ps1[i].x = pa[2 * i]; ps1[i].y = pa[2 * i + 1];
        ps2[i].x = qa[2 * i]; ps2[i].y = qa[2 * i + 1];
    }
    orient_ccw(ps1, 4);
    orient_ccw(ps2, 4);
    ps1[4] = ps1[0]; ps2[4] = ps2[0];
    float inter = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inter += fan_pair(ps1[i], ps1[i + 1], ps2[j], ps2[j + 1]);
    const float uni = fabsf(shoelace(ps1, 4)) + fabsf(shoelace(ps2, 4)) - inter;
    return (uni == 0) ? (inter + 1) / (uni + 1) : inter / uni;
}

// ---- own geometry: convex quadrilaterals ---------------------------------------------------------------------------------
struct Quad { float x[4], y[4]; };             // counter-clockwise, coordinates relative to a common origin

__device__ __forceinline__ float fdiv_fast(float a, float b) { return __fdividef(a, b); }

// sum of cross(start, end) over the parts of P's edges inside Q (strict: parts on Q's boundary do not count)
__device__ __forceinline__ float clipped_boundary(const Quad& P, const Quad& Q, bool strict) {
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float ax = P.x[i], ay = P.y[i];
        const float dx = P.x[(i + 1) & 3] - ax, dy = P.y[(i + 1) & 3] - ay;
        float t0 = 0.0f, t1 = 1.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float ex = Q.x[(j + 1) & 3] - Q.x[j], ey = Q.y[(j + 1) & 3] - Q.y[j];
            const float f0 = ex * (ay - Q.y[j]) - ey * (ax - Q.x[j]);          // cross(e, a - q_j): >= 0 inside
            const float fd = ex * dy - ey * dx;                                // cross(e, d)
            const float t = fdiv_fast(-f0, fd);
            if (fd > 0.0f) t0 = fmaxf(t0, t);
            else if (fd < 0.0f) t1 = fminf(t1, t);
            else if (strict ? (f0 <= 0.0f) : (f0 < 0.0f)) t1 = -1.0f;          // parallel: wholly outside (or on the boundary)
        }
        if (t1 > t0) {
            const float sx = fmaf(t0, dx, ax), sy = fmaf(t0, dy, ay), ex2 = fmaf(t1, dx, ax), ey2 = fmaf(t1, dy, ay);
            sum += sx * ey2 - sy * ex2;
        }
    }
    return sum;
}

// per-polygon record written by the gather kernel: {signed area, 1 = strictly convex, max |coordinate|, -}
__device__ __forceinline__ float4 quad_meta(const float* v) {
    float a2 = 0.0f, m = 0.0f;
    int pos = 0, neg = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int j = (i + 1) & 3, k = (i + 2) & 3;
        a2 += (v[2 * i] - v[0]) * (v[2 * j + 1] - v[1]) - (v[2 * i + 1] - v[1]) * (v[2 * j] - v[0]);
        const float c = (v[2 * j] - v[2 * i]) * (v[2 * k + 1] - v[2 * j + 1]) - (v[2 * j + 1] - v[2 * i + 1]) * (v[2 * k] - v[2 * j]);
        pos += c > 0.0f; neg += c < 0.0f;
        m = fmaxf(m, fmaxf(fabsf(v[2 * i]), fabsf(v[2 * i + 1])));
    }
    const bool convex = (pos == 4 || neg == 4) && isfinite(a2) && a2 != 0.0f;
    return make_float4(0.5f * a2, convex ? 1.0f : 0.0f, m, 0.0f);
}

// Fast IoU of two convex quadrilaterals (8 floats each: first / second half in two float4).  Returns false when the pair must go
// to the reference restatement instead (concave input, or the result is not sane); `margin` = how far the reference's own
// FP32 noise can move its IoU away from this value.
__device__ __forceinline__ bool convex_quad_iou(float4 p0, float4 p1, float4 mp, float4 q0, float4 q1, float4 mq, float& iou, float& margin) {
    iou = 0.0f; margin = 1.0f;
    if (mp.y == 0.0f || mq.y == 0.0f) return false;
    const float ox = p0.x, oy = p0.y;
    Quad P, Q;
    const float px[4] = { p0.x, p0.z, p1.x, p1.z }, py[4] = { p0.y, p0.w, p1.y, p1.w };
    const float qx[4] = { q0.x, q0.z, q1.x, q1.z }, qy[4] = { q0.y, q0.w, q1.y, q1.w };
#pragma unroll
    for (int i = 0; i < 4; i++) {                       // counter-clockwise order, relative coordinates
        const int ip = (mp.x < 0.0f) ? ((4 - i) & 3) : i, iq = (mq.x < 0.0f) ? ((4 - i) & 3) : i;
        P.x[i] = px[ip] - ox; P.y[i] = py[ip] - oy;
        Q.x[i] = qx[iq] - ox; Q.y[i] = qy[iq] - oy;
    }
    const float a1 = fabsf(mp.x), a2 = fabsf(mq.x);
    float inter = 0.5f * (clipped_boundary(P, Q, false) + clipped_boundary(Q, P, true));
    if (!(inter >= -1e-3f * fminf(a1, a2)) || !(inter <= fminf(a1, a2) * 1.001f)) return false;      // also catches NaN
    inter = fminf(fmaxf(inter, 0.0f), fminf(a1, a2));
    const float uni = a1 + a2 - inter;
    if (!(uni > 0.0f)) return false;
    iou = inter / uni;
    const float M = fmaxf(mp.z, mq.z);
    margin = 1e-4f + 1.6e-5f * M * M / uni;             // origin-anchored fans: ~64 products of coordinates, each eps * M^2
    return true;
}

}  // namespace poly
}  // namespace r3g
