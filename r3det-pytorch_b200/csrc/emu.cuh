// emu.cuh — operation-faithful FP32 restatements of the reference's two point-set IoU algorithms.
//
// The fast path (geom.cuh) computes the geometric intersection area.  The reference algorithms are not
// purely geometric: v1 drops polygon vertices closer than 1e-2 px and uses strict/half-open tests
// (r3det/ops/rbbox_geo/src/rbbox_geo_kernel.cu:94-228); the detectron2-lineage v2/v3 code orders points
// with 1e-6 / 1e-8 tolerances (r3det/ops/box_iou_rotated/src/box_iou_rotated_utils.h:157-289).  Where
// those quirks can change the answer beyond FP32 noise — a v1 pair flagged by v1_dedup_risk(), or an NMS
// pair whose IoU lies within a hair of the threshold — the kernels re-evaluate the pair with the
// functions below, which follow the reference arithmetic step by step on the RAW box tuples
// (absolute coordinates for v1, midpoint-shifted for v2/v3), so the decision is the reference's.
// These run for ~1e-3 of overlapping pairs; they use small local arrays and are compiled without
// fused-multiply-add contraction (explicit __fmul_rn/__fadd_rn) so that exact-zero and strict
// comparisons behave as in the reference's host build.
#pragma once
#include "geom.cuh"

namespace r3g {
namespace emu {

struct P2 { float x, y; };

#if defined(__CUDA_ARCH__)
R3G_HD float mulr(float a, float b) { return __fmul_rn(a, b); }
R3G_HD float addr(float a, float b) { return __fadd_rn(a, b); }
R3G_HD float subr(float a, float b) { return __fsub_rn(a, b); }
#else
R3G_HD float mulr(float a, float b) { volatile float r = a * b; return r; }
R3G_HD float addr(float a, float b) { volatile float r = a + b; return r; }
R3G_HD float subr(float a, float b) { volatile float r = a - b; return r; }
#endif
R3G_HD float dot2(P2 a, P2 b) { return addr(mulr(a.x, b.x), mulr(a.y, b.y)); }
R3G_HD float cross2(P2 a, P2 b) { return subr(mulr(a.x, b.y), mulr(b.x, a.y)); }
R3G_HD P2 sub2(P2 a, P2 b) { P2 r = { subr(a.x, b.x), subr(a.y, b.y) }; return r; }
R3G_HD P2 add2(P2 a, P2 b) { P2 r = { addr(a.x, b.x), addr(a.y, b.y) }; return r; }
R3G_HD P2 scale2(float s, P2 a) { P2 r = { mulr(s, a.x), mulr(s, a.y) }; return r; }

// ------------------------------------------------------------------------------------------------
// v1  (rbbox_geo_kernel.cu:143-267 == rnms_kernel.cu:115-227 == rnms_cpu.cpp:111-221)
// ------------------------------------------------------------------------------------------------
R3G_HD void v1_corners(const float* rb, P2* vs) {                       // :143-155
    float x = rb[0], y = rb[1], w_2 = rb[2] / 2, h_2 = rb[3] / 2, a = rb[4];
    // cosf/sinf as the reference calls them: CUDA's on the device (what rbbox_geo / rnms CUDA builds use), libm's on
    // the host.  The two differ in the last ulp for ~10 % of angles; on near-coincident boxes (< 1e-2 px apart) the v1
    // algorithm is chaotic in that bit, which is why the reference's own CPU and CUDA builds disagree there.
    float cosa = cosf(a), sina = sinf(a);
    float wx = mulr(cosa, w_2), wy = mulr(sina, w_2);
    float hx = mulr(-sina, h_2), hy = mulr(cosa, h_2);
    vs[0].x = addr(addr(x, wx), hx); vs[0].y = addr(addr(y, wy), hy);
    vs[1].x = addr(subr(x, wx), hx); vs[1].y = addr(subr(y, wy), hy);
    vs[2].x = subr(subr(x, wx), hx); vs[2].y = subr(subr(y, wy), hy);
    vs[3].x = subr(addr(x, wx), hx); vs[3].y = subr(addr(y, wy), hy);
}

R3G_HD int v1_inside(const P2* v1, const P2* v2, P2* ps) {              // :157-175, strict
    P2 center = scale2(0.5f, add2(v2[0], v2[2]));
    P2 wv = scale2(0.5f, sub2(v2[1], v2[0]));
    P2 hv = scale2(0.5f, sub2(v2[2], v2[1]));
    float h2 = dot2(hv, hv), w2 = dot2(wv, wv);
    int n = 0;
    for (int i = 0; i < 4; i++) {
        P2 pr = sub2(v1[i], center);
        if (fabsf(dot2(pr, hv)) < h2 && fabsf(dot2(pr, wv)) < w2) ps[n++] = v1[i];
    }
    return n;
}

R3G_HD int v1_seg(P2 a1, P2 a2, P2 b1, P2 b2, P2* ps) {                 // :94-140
    P2 A = sub2(a2, a1), B = sub2(b2, b1), C = sub2(a1, b1);
    if (C.x == 0 && C.y == 0) { ps[0] = a1; return 1; }
    float D = -cross2(A, B);
    if (D != 0) {
        float s = cross2(C, B) / D;
        float t = -cross2(A, C) / D;
        if (0 <= s && s < 1 && 0 <= t && t < 1) { ps[0] = add2(a1, scale2(s, A)); return 1; }
        return 0;
    }
    if (cross2(A, C) != 0) return 0;
    int n = 0;
    float BdtC = dot2(B, C), BdtB = dot2(B, B), AdtnC = -dot2(A, C), AdtA = dot2(A, A);
    if (BdtC >= 0 && BdtC < BdtB) ps[n++] = a1;
    if (AdtnC >= 0 && AdtnC < AdtA) ps[n++] = b1;
    return n;
}

R3G_HD bool v1_less(P2 p, P2 q) {                                        // :77-82
    if ((p.x == 0 && p.y == 0) && (q.x != 0 || q.y != 0)) return true;
    return cross2(p, q) > 0;
}

R3G_HD float v1_area(P2* dirty, int nd) {                                // :193-228
    P2 vs[24];
    vs[0].x = 0; vs[0].y = 0;
    int n = 1;
    for (int i = 1; i < nd; i++) {
        bool clean = true;
        dirty[i] = sub2(dirty[i], dirty[0]);
        for (int j = 0; j < n; j++) {
            P2 d = sub2(dirty[i], vs[j]);
            if (fabsf(d.x) < 1e-2f && fabsf(d.y) < 1e-2f) { clean = false; break; }
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
    for (int i = 1; i < n; i++) a = addr(a, cross2(vs[i], vs[(i + 1) % n]));
    return a / 2;
}

// raw 5-float boxes in the coordinates the reference sees (absolute, class offset already added)
R3G_HD float v1_pair(const float* rb1, const float* rb2, bool iof) {     // :238-267
    P2 v1[4], v2[4], u[24];
    v1_corners(rb1, v1);
    v1_corners(rb2, v2);
    int n = v1_inside(v1, v2, u);
    n += v1_inside(v2, v1, u + n);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            P2 tmp[2];
            int k = v1_seg(v1[i], v1[(i + 1) & 3], v2[j], v2[(j + 1) & 3], tmp);
            for (int q = 0; q < k && n < 24; q++) u[n++] = tmp[q];       // the reference's u[16] is unchecked
        }
    if (n < 3) return 0.0f;
    float s1 = mulr(rb1[2], rb1[3]), s2 = mulr(rb2[2], rb2[3]);
    float su = v1_area(u, n);
    su = fmaxf(fminf(fminf(su, s1), s2), 0.0f);
    return iof ? su / s1 : su / subr(addr(s1, s2), su);
}

// ------------------------------------------------------------------------------------------------
// v2 / v3  (box_iou_rotated_utils.h:55-361; the ml_nms_rotated copy for v2), device (__CUDACC__) sort
// ------------------------------------------------------------------------------------------------
R3G_HD void d2_corners(float xc, float yc, float w, float h, float a, int variant, P2* p) {   // :55-74
    double th = (double)a;
    float c2 = mulr((float)cos(th), 0.5f), s2 = mulr((float)sin(th), 0.5f);
    if (variant == V3) {
        p[0].x = addr(addr(xc, mulr(s2, h)), mulr(c2, w));
        p[0].y = subr(addr(yc, mulr(c2, h)), mulr(s2, w));
        p[1].x = addr(subr(xc, mulr(s2, h)), mulr(c2, w));
        p[1].y = subr(subr(yc, mulr(c2, h)), mulr(s2, w));
    } else {
        p[0].x = subr(subr(xc, mulr(s2, h)), mulr(c2, w));
        p[0].y = subr(addr(yc, mulr(c2, h)), mulr(s2, w));
        p[1].x = subr(addr(xc, mulr(s2, h)), mulr(c2, w));
        p[1].y = subr(subr(yc, mulr(c2, h)), mulr(s2, w));
    }
    p[2].x = subr(mulr(2, xc), p[0].x); p[2].y = subr(mulr(2, yc), p[0].y);
    p[3].x = subr(mulr(2, xc), p[1].x); p[3].y = subr(mulr(2, yc), p[1].y);
}

R3G_HD int d2_points(const P2* p1, const P2* p2, P2* out) {              // :76-155, inclusive
    P2 e1[4], e2[4];
    for (int i = 0; i < 4; i++) { e1[i] = sub2(p1[(i + 1) & 3], p1[i]); e2[i] = sub2(p2[(i + 1) & 3], p2[i]); }
    int n = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float det = cross2(e2[j], e1[i]);
            if (fabs((double)det) <= 1e-14) continue;
            P2 v12 = sub2(p2[j], p1[i]);
            float t1 = cross2(e2[j], v12) / det, t2 = cross2(e1[i], v12) / det;
            if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) out[n++] = add2(p1[i], scale2(t1, e1[i]));
        }
    for (int pass = 0; pass < 2; pass++) {
        const P2* a = pass ? p2 : p1;            // points tested
        const P2* r = pass ? p1 : p2;            // rectangle
        const P2* er = pass ? e1 : e2;
        P2 AB = er[0], DA = er[3];
        float ABAB = dot2(AB, AB), ADAD = dot2(DA, DA);
        for (int i = 0; i < 4; i++) {
            P2 AP = sub2(a[i], r[0]);
            float apab = dot2(AP, AB), apad = -dot2(AP, DA);
            if (apab >= 0 && apad >= 0 && apab <= ABAB && apad <= ADAD) out[n++] = a[i];
        }
    }
    return n;
}

R3G_HD float d2_hull_area(const P2* p, int num, int variant) {           // :157-303 (shift_to_zero=true)
    P2 q[24];
    float dist[24];
    int t = 0;
    for (int i = 1; i < num; i++)
        if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
    P2 start = p[t];
    for (int i = 0; i < num; i++) q[i] = sub2(p[i], start);
    P2 tmp = q[0]; q[0] = q[t]; q[t] = tmp;
    for (int i = 0; i < num; i++) dist[i] = dot2(q[i], q[i]);
    for (int i = 1; i < num - 1; i++)
        for (int j = i + 1; j < num; j++) {
            float cp = cross2(q[i], q[j]);
            if ((cp < -1e-6) || (fabs((double)cp) < 1e-6 && dist[i] > dist[j])) {
                P2 qt = q[i]; q[i] = q[j]; q[j] = qt;
                float dt = dist[i]; dist[i] = dist[j]; dist[j] = dt;
            }
        }
    int k;
    for (k = 1; k < num; k++)
        if (dist[k] > 1e-8) break;
    if (k == num) return 0.0f;
    q[1] = q[k];
    int m = 2;
    for (int i = k + 1; i < num; i++) {
        while (m > 1) {
            P2 q1 = sub2(q[i], q[m - 2]), q2 = sub2(q[m - 1], q[m - 2]);
            bool pop = (variant == V3) ? (mulr(q1.x, q2.y) >= mulr(q2.x, q1.y)) : (cross2(q1, q2) >= 0);
            if (pop) m--; else break;
        }
        q[m++] = q[i];
    }
    if (m <= 2) return 0.0f;
    float area = 0;
    for (int i = 1; i < m - 1; i++) area = addr(area, fabsf(cross2(sub2(q[i], q[0]), sub2(q[i + 1], q[0]))));
    return (float)(area / 2.0);
}

R3G_HD float d2_pair(const float* b1, const float* b2, bool mode_iou, int variant) {   // :331-361
    double sx = addr(b1[0], b2[0]) / 2.0, sy = addr(b1[1], b2[1]) / 2.0;
    float x1 = (float)(b1[0] - sx), y1 = (float)(b1[1] - sy);
    float x2 = (float)(b2[0] - sx), y2 = (float)(b2[1] - sy);
    float area1 = mulr(b1[2], b1[3]), area2 = mulr(b2[2], b2[3]);
    if (area1 < 1e-14 || area2 < 1e-14) return 0.0f;
    P2 p1[4], p2[4], cand[24];
    d2_corners(x1, y1, b1[2], b1[3], b1[4], variant, p1);
    d2_corners(x2, y2, b2[2], b2[3], b2[4], variant, p2);
    int num = d2_points(p1, p2, cand);
    float inter = (num > 2) ? d2_hull_area(cand, num, variant) : 0.0f;
    return mode_iou ? inter / subr(addr(area1, area2), inter) : inter / area1;
}

// Box preparation used by the kernels.  v2/v3: prep_box (their geometry is midpoint-shifted, so the
// per-box trig is all there is).  v1: the reference builds its corners in ABSOLUTE coordinates
// (rbbox_geo_kernel.cu:143-155), which quantises them to ulp(|coordinate|) — up to 1e-3 px once
// batched_rnms has added its class offsets.  To follow that, the prepared v1 box is re-fitted to the
// reference's own rounded corners: centre = (v0+v2)/2, axis = direction of the longer half-diagonal pair,
// half extents = |half edge vectors| (the same quantities vertex_in_rbbox derives, :160-164).
R3G_HD void prep_box_strict(const float* b, int variant, BoxP0& p0, BoxP1& p1) {
    prep_box(b, variant, p0, p1);
    if (variant != V1) return;
    P2 vs[4];
    v1_corners(b, vs);
    float wx = 0.5f * (vs[0].x - vs[1].x), wy = 0.5f * (vs[0].y - vs[1].y);   // +w half vector
    float hx = 0.5f * (vs[1].x - vs[2].x), hy = 0.5f * (vs[1].y - vs[2].y);   // +h half vector
    float hw = sqrtf(wx * wx + wy * wy), hh = sqrtf(hx * hx + hy * hy);
    if (!(hw > 0.0f) && !(hh > 0.0f)) return;                                    // degenerate: keep trig form
    if (hw >= hh) { p1.c = wx / hw; p1.s = wy / hw; }
    else          { p1.c = hy / hh; p1.s = -hx / hh; }
    float w = b[2], h = b[3];
    p1.hw = (w < 0.0f) ? -hw : hw;
    p1.hh = (h < 0.0f) ? -hh : hh;
    p0.cx = 0.5f * (vs[0].x + vs[2].x);
    p0.cy = 0.5f * (vs[0].y + vs[2].y);
    p0.r = sqrtf(hw * hw + hh * hh) * 1.00001f + 1e-6f;
}

R3G_HD float pair(const float* b1, const float* b2, int variant, int mode) {
    return (variant == V1) ? v1_pair(b1, b2, mode == MODE_IOF) : d2_pair(b1, b2, mode == MODE_IOU, variant);
}

}  // namespace emu
}  // namespace r3g
