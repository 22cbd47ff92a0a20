// poly.cuh — polygon (4-point, arbitrary quadrilateral) IoU and the suppression-mask kernel of polygon NMS.
//
// Replaces (reference, relative to /root/reference):
//   devPolyIoU + poly_nms_kernel      r3det/ops/nms_rotated/src/poly_nms_cuda.cu:22-194
// The reference evaluates its origin-anchored triangle-fan clipping for ALL K^2/2 pairs.  Here a pair first has to
// pass an axis-aligned bounding-box test (disjoint boxes have IoU <= FP32 noise, far below any usable threshold);
// survivors are compacted with ballot/popc into a shared-memory queue and evaluated 32 at a time at full lane
// occupancy.  The IoU itself is NOT re-derived: quadrilaterals may be concave or self-intersecting and the reference's
// result (including its FP32 cancellation noise from anchoring the fans at the image origin) is the specification, so
// the arithmetic below restates it step by step.  Sorting, the greedy scan and the keep-list emission are the
// rotated-box NMS pipeline's (nms.cu).
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

}  // namespace poly
}  // namespace r3g
