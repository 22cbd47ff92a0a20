// geom.cuh — rotated-box pair geometry shared by the IoU and NMS kernels (sm_100a product code).
//
// One FP32, register-only, branch-light core serves RBboxOverlaps2D_v1/v2/v3 and nms v1/v2/v3:
//   * prepared boxes: per-box sincos, half extents, circumradius and area are computed ONCE per box
//     (O(M+N)) with the same trig the reference variant uses, not once per pair as the reference does
//     (rbbox_geo_kernel.cu:143-155, box_iou_rotated_utils.h:55-74);
//   * pair test: circumradius reject -> separating-axis reject -> intersection area;
//   * intersection area = closed line integral  ∮ X dY  of box B's boundary CLAMPED to box A
//     (A axis-aligned in its own frame).  Clamping to a convex set preserves winding numbers, so the
//     clamped boundary encloses exactly A∩B.  Each of B's 4 edges contributes
//         (clampY(q) - clampY(p)) * mean_{x in [x0,x1]} clampX(x)
//     — a fixed sequence of min/max/fma with no polygon storage, vertex counting or sorting.  It is
//     the Sutherland–Hodgman clip + shoelace sum evaluated edge-by-edge in closed form, and it is
//     continuous in its inputs (no topological decisions), unlike the reference's point-set
//     algorithms (rbbox_geo_kernel.cu:193-228 de-dup + angular sort; box_iou_rotated_utils.h:157-289
//     Graham scan).
// Variant-specific behaviour that is not pure geometry (v1's 1e-2 vertex de-dup, NMS threshold ties)
// is handled by the exact restatements in emu.cuh, invoked only for the rare flagged pairs.
//
// The header also compiles as plain C++ (R3G_HD empty) so tests can sweep 10^7 pairs on the CPU
// against the oracle without a GPU (tests/hostgeom).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define R3G_HD __host__ __device__ __forceinline__
#else
#define R3G_HD inline
#endif

namespace r3g {

enum Variant { V1 = 1, V2 = 2, V3 = 3 };
enum Mode { MODE_IOU = 0, MODE_IOF = 1 };

// Prepared box = two 16-byte records (kept as two planes so that lane-consecutive boxes are
// conflict-free 128-bit shared-memory / global accesses).
struct __attribute__((aligned(16))) BoxP0 { float cx, cy, r, area; };   // centre, safe circumradius, w*h
struct __attribute__((aligned(16))) BoxP1 { float c, s, hw, hh; };      // unit w-axis (c,s), half extents

// Per-box preparation with the variant's own trig:
//   v1: cosf/sinf, rotation +a           (rbbox_geo_kernel.cu:143-155)
//   v2: (float)cos/sin(double), +a       (ml_nms_rotated/src/box_iou_rotated_utils.h:60-77)
//   v3: (float)cos/sin(double), -a       (box_iou_rotated/src/box_iou_rotated_utils.h:60-73)
R3G_HD void prep_box(const float* b, int variant, BoxP0& p0, BoxP1& p1) {
    float w = b[2], h = b[3], a = b[4];
    float c, s;
    if (variant == V1) {
        c = cosf(a);
        s = sinf(a);
    } else {
        double th = (double)a;
        c = (float)cos(th);
        s = (float)sin(th);
        if (variant == V3) s = -s;
    }
    p1.c = c; p1.s = s;
    p1.hw = 0.5f * w; p1.hh = 0.5f * h;
    p0.cx = b[0]; p0.cy = b[1];
    // conservative radius: never rejects a pair whose boxes touch
    p0.r = sqrtf(p1.hw * p1.hw + p1.hh * p1.hh) * 1.00001f + 1e-6f;
    p0.area = w * h;
}

R3G_HD bool circle_reject(const BoxP0& A, const BoxP0& B) {
    float dx = B.cx - A.cx, dy = B.cy - A.cy, rr = A.r + B.r;
    return dx * dx + dy * dy > rr * rr;
}

R3G_HD float clampf(float x, float lim) { return fminf(fmaxf(x, -lim), lim); }

// 2-ulp division for the four per-edge weighted means (the result is a convex combination, so a relative
// error of 2^-22 moves the area by < 3e-7 of area(A)); IEEE division is kept for the two edge slopes.
#if defined(__CUDA_ARCH__)
R3G_HD float fast_div(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));      // 1 ulp; b is a sum of edge-length pieces, far from the ftz range
    return a * r;
}
// ~1-ulp quotient without the IEEE slow path: approximate reciprocal + one Newton step + residual correction
R3G_HD float nr_div(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    r = fmaf(fmaf(-b, r, 1.0f), r, r);
    float q = a * r;
    return fmaf(fmaf(-b, q, a), r, q);
}
#else
R3G_HD float fast_div(float a, float b) { return a / b; }
R3G_HD float nr_div(float a, float b) { return a / b; }
#endif

// Corners of B and the frame quantities of the pair, in A's frame (A = [-a,a]x[-b,b]).
struct PairFrame {
    float qx[4], qy[4];   // B's corners, CCW for positive w,h
    float su, sv;         // dx/dy of B's u-edges (q0->q1, q2->q3) and v-edges (q1->q2, q3->q0)
    float a, b;
    float cd, sd, dbx, dby;   // relative rotation and centre offset in B's frame (for the degeneracy test)
};

// Returns false when a separating axis exists (boxes disjoint -> intersection exactly 0).  `sat_passed`: the caller has
// already run pair_sat() on this pair (the kernels' queue stage), so the test is skipped — for a pair that pair_sat lets
// through by a rounding of the last bit the integral below returns the (continuous) ~0 area instead of an exact 0.
R3G_HD bool pair_frame(const BoxP0& A0, const BoxP1& A1, const BoxP0& B0, const BoxP1& B1, PairFrame& f, bool sat_passed = false) {
    float dx = B0.cx - A0.cx, dy = B0.cy - A0.cy;
    float a = A1.hw, b = A1.hh;
    // relative rotation (cos, sin of thetaB - thetaA)
    float cd = A1.c * B1.c + A1.s * B1.s;
    float sd = A1.c * B1.s - A1.s * B1.c;
    float acd = fabsf(cd), asd = fabsf(sd);
    // centre offset in A's frame and in B's frame
    float dlx = dx * A1.c + dy * A1.s, dly = dy * A1.c - dx * A1.s;
    float dbx = dx * B1.c + dy * B1.s, dby = dy * B1.c - dx * B1.s;
    // B's half-axis vectors in A's frame
    float ux = B1.hw * cd, uy = B1.hw * sd;
    float vx = -B1.hh * sd, vy = B1.hh * cd;
    // separating axes: A's two, then B's two
    if (!sat_passed) {
        bool sep = (fabsf(dlx) > a + fabsf(ux) + fabsf(vx)) | (fabsf(dly) > b + fabsf(uy) + fabsf(vy)) |
                   (fabsf(dbx) > B1.hw + a * acd + b * asd) | (fabsf(dby) > B1.hh + a * asd + b * acd);
        if (sep) return false;
    }
    float ex = dlx - vx, ey = dly - vy, gx = dlx + vx, gy = dly + vy;
    f.qx[0] = ex - ux; f.qy[0] = ey - uy;
    f.qx[1] = ex + ux; f.qy[1] = ey + uy;
    f.qx[2] = gx + ux; f.qy[2] = gy + uy;
    f.qx[3] = gx - ux; f.qy[3] = gy - uy;
    // an edge whose y-extent is below 1e-12*size contributes nothing: its slope is irrelevant
    f.su = (asd > 1e-12f) ? nr_div(cd, sd) : 0.0f;
    f.sv = (acd > 1e-12f) ? nr_div(-sd, cd) : 0.0f;
    f.a = a; f.b = b;
    f.cd = cd; f.sd = sd; f.dbx = dbx; f.dby = dby;
    return true;
}

// Separating-axis test alone (the cheap second-stage filter of the kernels): true = boxes overlap or touch.
R3G_HD bool pair_sat(const BoxP0& A0, const BoxP1& A1, const BoxP0& B0, const BoxP1& B1) {
    float dx = B0.cx - A0.cx, dy = B0.cy - A0.cy;
    float cd = A1.c * B1.c + A1.s * B1.s;
    float sd = A1.c * B1.s - A1.s * B1.c;
    float acd = fabsf(cd), asd = fabsf(sd);
    float dlx = dx * A1.c + dy * A1.s, dly = dy * A1.c - dx * A1.s;
    float dbx = dx * B1.c + dy * B1.s, dby = dy * B1.c - dx * B1.s;
    float aw = fabsf(A1.hw), ah = fabsf(A1.hh), bw = fabsf(B1.hw), bh = fabsf(B1.hh);
    bool sep = (fabsf(dlx) > aw + bw * acd + bh * asd) | (fabsf(dly) > ah + bw * asd + bh * acd) |
               (fabsf(dbx) > bw + aw * acd + ah * asd) | (fabsf(dby) > bh + aw * asd + ah * acd);
    return !sep;
}

// One edge p->q of B: (clampY(q)-clampY(p)) * mean of clampX over the part of the edge inside |y|<=b.
R3G_HD float edge_term(float px, float py, float Yp, float qx, float qy, float Yq, float slope, float a) {
    float dY = Yq - Yp;
    float x0 = fmaf(Yp - py, slope, px);      // x where the edge enters the slab (== px if p is inside)
    float x1 = fmaf(Yq - qy, slope, qx);      // x where it leaves           (== qx if q is inside)
    // integral of clamp(x,-a,a) from x0 to x1 via the three monotone pieces x = min(x,-a)+clamp(x)+max(x,a)
    float l0 = fminf(x0, -a), l1 = fminf(x1, -a);
    float m0 = clampf(x0, a), m1 = clampf(x1, a);
    float h0 = fmaxf(x0, a), h1 = fmaxf(x1, a);
    float d1 = l1 - l0, d2 = m1 - m0, d3 = h1 - h0;
    float num = fmaf(a, d3 - d1, 0.5f * (m0 + m1) * d2);
    float den = d1 + d2 + d3;
    float mean = (den != 0.0f) ? fast_div(num, den) : m0;
    return dY * mean;
}

R3G_HD float frame_area(const PairFrame& f) {
    float Y0 = clampf(f.qy[0], f.b), Y1 = clampf(f.qy[1], f.b);
    float Y2 = clampf(f.qy[2], f.b), Y3 = clampf(f.qy[3], f.b);
    float t0 = edge_term(f.qx[0], f.qy[0], Y0, f.qx[1], f.qy[1], Y1, f.su, f.a);
    float t1 = edge_term(f.qx[1], f.qy[1], Y1, f.qx[2], f.qy[2], Y2, f.sv, f.a);
    float t2 = edge_term(f.qx[2], f.qy[2], Y2, f.qx[3], f.qy[3], Y3, f.su, f.a);
    float t3 = edge_term(f.qx[3], f.qy[3], Y3, f.qx[0], f.qy[0], Y0, f.sv, f.a);
    return fabsf((t0 + t2) + (t1 + t3));
}

// Intersection area of two prepared boxes (0 when disjoint).
R3G_HD float pair_intersection(const BoxP0& A0, const BoxP1& A1, const BoxP0& B0, const BoxP1& B1) {
    PairFrame f;
    if (!pair_frame(A0, A1, B0, B1, f)) return 0.0f;
    return frame_area(f);
}

// v1 only: true when some corner of one box lies within `tau` of the other's boundary — the only
// configurations in which the reference's 1e-2 vertex de-dup / strict tests (rbbox_geo_kernel.cu:169-170,
// 195-213) can change the polygon, i.e. where v1 departs from the geometric area.  Conservative.
R3G_HD bool corner_near_boundary(const float* qx, const float* qy, float a, float b, float tau) {
    // "(|ex| < tau and ey < tau) or (|ey| < tau and ex < tau)" with ex = |x| - a, ey = |y| - b is the same set as
    // | max(ex, ey) | < tau (the signed Chebyshev distance to the rectangle's boundary): 3 instructions per corner,
    // and the four corners share one comparison through a running minimum.
    float d = 3.0e38f;
#pragma unroll
    for (int k = 0; k < 4; k++) d = fminf(d, fabsf(fmaxf(fabsf(qx[k]) - a, fabsf(qy[k]) - b)));
    return d < tau;
}

R3G_HD bool v1_dedup_risk(const BoxP0& A0, const BoxP1& A1, const BoxP0& B0, const BoxP1& B1,
                          const PairFrame& f, float tau) {
    (void)A0; (void)B0;
    if (corner_near_boundary(f.qx, f.qy, f.a, f.b, tau)) return true;
    // A's corners in B's frame: -db +- a*(cd,-sd) +- b*(sd,cd)
    const float P = f.a * f.cd, Q = f.b * f.sd, R = f.a * f.sd, S = f.b * f.cd;
    const float x0 = -f.dbx + P, x1 = -f.dbx - P, y0 = -f.dby - R, y1 = -f.dby + R;
    float px[4], py[4];
    px[0] = x0 + Q; py[0] = y0 + S;
    px[1] = x0 - Q; py[1] = y0 - S;
    px[2] = x1 + Q; py[2] = y1 + S;
    px[3] = x1 - Q; py[3] = y1 - S;
    return corner_near_boundary(px, py, fabsf(B1.hw), fabsf(B1.hh), tau);
}

// IoU / IoF from the intersection area under the variant's epilogue.
//   v1: clamp to [0, min(s1,s2)] (rbbox_geo_kernel.cu:254-256), no small-area guard
//   v2/v3: area < 1e-14 -> 0 (box_iou_rotated_utils.h:353-355)
R3G_HD float overlap_ratio(float inter, float s1, float s2, int variant, int mode) {
    if (variant == V1) {
        inter = fmaxf(fminf(fminf(inter, s1), s2), 0.0f);
    } else if (s1 < 1e-14f || s2 < 1e-14f) {
        return 0.0f;
    }
    float den = (mode == MODE_IOF) ? s1 : (s1 + s2 - inter);
    return (den > 0.0f) ? nr_div(inter, den) : 0.0f;
}

// Fast-path overlap of one prepared pair.  `risk` is set for pairs the caller must re-evaluate with the
// variant's restatement in emu.cuh (strict reference parity); tau = 0 disables the test.
R3G_HD float pair_overlap(const BoxP0& A0, const BoxP1& A1, const BoxP0& B0, const BoxP1& B1,
                          int variant, int mode, float tau, bool& risk, bool sat_passed = false) {
    risk = false;
    PairFrame f;
    if (!pair_frame(A0, A1, B0, B1, f, sat_passed)) return 0.0f;
    float inter = frame_area(f);
    if (tau > 0.0f) {
        // Thin boxes go to the restatement as well: below tau every vertex is "near" another one, and for v1
        // (absolute-coordinate arithmetic, corners quantised to ulp(|coordinate|) ~ 1e-4 px) a box thinner than
        // 4 px carries a relative area noise above the 1e-5 gate that only the same arithmetic reproduces.
        // (IoF divides by the first box alone: there the reference's own FP32 noise passes 1e-5 below 16 px, all variants.)
        const float thin = (mode == MODE_IOF) ? 8.0f : ((variant == V1) ? 2.0f : tau);
        const float mina = fminf(fabsf(A1.hw), fabsf(A1.hh)), minb = fminf(fabsf(B1.hw), fabsf(B1.hh));
        const float thin_b = (mode == MODE_IOF) ? ((variant == V1) ? 2.0f : tau) : thin;   // IoF divides by box A alone
        risk = (mina < thin) || (minb < thin_b) || v1_dedup_risk(A0, A1, B0, B1, f, tau);
    }
    return overlap_ratio(inter, A0.area, B0.area, variant, mode);
}

}  // namespace r3g
