// frm_tma.cuh — FRM forward / backward-apply with TMA-staged tiles (sm_100a: cp.async.bulk.tensor + mbarrier).
//
// Replaces (reference, relative to /root/reference):
//   feature_refine_forward_kernel / feature_refine_backward_kernel   r3det/ops/fr/src/feature_refine_kernel.cu:112-230
// The sample taps of a location depend on (n, h, w) only.  The reference samples box x as the ROW and box y as the COLUMN
// coordinate (feature_refine_kernel.cu:131-132), so the taps of the location tile (h0.., w0..) lie around the TRANSPOSED window
// (rows w0.., columns h0..) of the same plane.  The gather kernel of frm.cu reads those taps straight from global memory:
// every request touches 3-5 cache lines and the kernel is bound by L1 wavefronts (86 % L1TEX, 30 % of HBM).  Here, per
// channel plane, the TMA unit copies two boxes into shared memory — the CTA's own 32 x 16 tile and the 22 x 44 (points = 1) or
// 32 x 52 (points = 5) window its taps fall into (its first column on a 16-byte boundary, as the TMA unit requires) — through a 4-stage mbarrier ring, and the threads only do conflict-free
// shared-memory reads, FMAs and one store.  Window pitches (44, 52 floats) are = 12, 20 mod 32, so that the 8 (w) x 4 (h)
// lanes of a warp hit 32 different banks.  Taps that fall outside the window (boxes far from their location) are read from
// global memory instead: results do not depend on the window.  Out-of-range box coordinates are zero-filled by the TMA unit and
// never read (their weights are zero).  The same kernel serves the backward apply: gather of grad_out through the per-target
// CSR of frm.cu's plan (rows of up to 8 entries live in registers, longer rows finish from global memory).
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace r3g {
namespace ftma {

constexpr int TH = 32, TW = 16;                  // location tile: 32 rows (h) x 16 columns (w); a thread owns (h, w) and (h, w + 8)
constexpr int THREADS = 256;
#ifndef R3G_FRM_STAGES
#define R3G_FRM_STAGES 4                           // ring depth (measured: 3 / 4 / 5 / 6)
#endif
constexpr int STAGES = R3G_FRM_STAGES;
constexpr int CC = 32;                           // channels per CTA
constexpr int MAXL = 8;
#ifndef R3G_FRM_EMPTY_BARRIER
#define R3G_FRM_EMPTY_BARRIER 0                  // 1: slot release through per-slot mbarriers (one arrival per warp, delayed refill by the
                                                 // producer thread) instead of __syncthreads.  MEASURED SLOWER (forward 0.102 vs 0.091 ms, backward
                                                 // apply 0.210 vs 0.162 ms): the producer's warp stalls on the slowest warp and the ring runs one
                                                 // slot shallower; kept as a build variant for the record
#endif

template <int P, bool BWD = false> struct Win {  // window of the taps: rows [w0 - HR, w0 + TW + HR), columns [h0 - HC, h0 - HC + RC)
    // the TMA unit wants the first column of a box on a 16-byte boundary: the left halo HC is a multiple of 4 floats.
    // (forward taps of a location spread over offsets -3 .. +4, the sources of a backward target over -4 .. +3)
    static constexpr int HR = (P == 1) ? (BWD ? 4 : 3) : 8, HC = (P == 1) ? 8 : 12;
    static constexpr int RR = TW + 2 * HR, RC = (P == 1) ? 44 : 52;  // 22 x 44 (columns -8 .. +4) / 32 x 52 (columns -12 .. +8)
    static_assert(HC % 4 == 0, "window start must be 16-byte aligned");
    static constexpr int OWN_BYTES = TH * TW * 4;
    static constexpr int WIN_BYTES = RR * RC * 4;
    static constexpr int WIN_PAD = (WIN_BYTES + 16 + 127) / 128 * 128;   // + a zero word the TMA unit never writes (padding entries read it)
    static constexpr int ZERO_OFF = OWN_BYTES + WIN_BYTES;
    static constexpr int STAGE_BYTES = OWN_BYTES + WIN_PAD;
    static_assert(RC % 4 == 0 && (RC % 32 == 12 || RC % 32 == 20), "window pitch: 16-byte rows, 8 rows on 8 different bank groups");
};

struct Level {
    CUtensorMap own, win;                        // 3-D maps over (N * C, H, W): boxes {TW, TH, 1} and {RC, RR, 1}
    const float* feat; const float* boxes; const float* residual; float* out;
    int H, W; float scale;
    int tiles_x, tiles_y, cchunks;
    unsigned block0, loc0;
};
struct Levels { Level lv[MAXL]; int L, N, C; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// one bilinear sample: weights and the four taps as BYTE offsets into the staged window (>= 0), or -(plane offset + 1) for
// taps outside the window (read from global memory instead)
struct Samp { float w[4]; int o[4]; };

__device__ __forceinline__ bool make_samp(int H, int W, float y, float x, int r0, int c0, int RR, int RC, Samp& s) {
    // feature_refine_kernel.cu:16-65 (same arithmetic as frm_taps in frm.cu); returns true when a tap lies outside the window
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
        s.w[0] = s.w[1] = s.w[2] = s.w[3] = 0.0f; s.o[0] = s.o[1] = s.o[2] = s.o[3] = 0;
        return false;
    }
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
    const float ly = y - (float)yl, lx = x - (float)xl;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    s.w[0] = hy * hx; s.w[1] = hy * lx; s.w[2] = ly * hx; s.w[3] = ly * lx;
    const bool far = (yl < r0 || yh >= r0 + RR || xl < c0 || xh >= c0 + RC);
    if (!far) {
        s.o[0] = ((yl - r0) * RC + (xl - c0)) * 4; s.o[1] = ((yl - r0) * RC + (xh - c0)) * 4;
        s.o[2] = ((yh - r0) * RC + (xl - c0)) * 4; s.o[3] = ((yh - r0) * RC + (xh - c0)) * 4;
    } else {
        s.o[0] = -(yl * W + xl + 1); s.o[1] = -(yl * W + xh + 1); s.o[2] = -(yh * W + xl + 1); s.o[3] = -(yh * W + xh + 1);
    }
    return far;
}

template <int P>
__global__ void __launch_bounds__(THREADS, (P == 1) ? 5 : 2) frm_forward_tma_kernel(const __grid_constant__ Levels S) {
    typedef Win<P> Wn;
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ __align__(8) unsigned long long full[STAGES];
    __shared__ __align__(8) unsigned long long empty[STAGES];     // one arrival per warp that has finished reading the slot
    int li = 0;
#pragma unroll
    for (int i = 1; i < MAXL; i++) if (i < S.L && blockIdx.x >= S.lv[i].block0) li = i;
    const Level& lv = S.lv[li];
    const int C = S.C, H = lv.H, W = lv.W;
    const int tiles = lv.tiles_x * lv.tiles_y;
    int bid = blockIdx.x - lv.block0;
    const int tl = bid % tiles; bid /= tiles;
    const int chunk = bid % lv.cchunks;
    const int n = bid / lv.cchunks;
    const int h0 = (tl / lv.tiles_x) * TH, w0 = (tl % lv.tiles_x) * TW;
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int hl = wi * 4 + (lane >> 3), wl = lane & 7;           // lanes: 8 along w, 4 along h
    const int h = h0 + hl;
    const int HW = H * W;
    const int r0 = w0 - Wn::HR, c0w = h0 - Wn::HC;               // window origin (row, column) in the plane
    const int ca = chunk * CC, cb = min(C, ca + CC);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int slot, int c) {
        unsigned char* base = fsm + slot * Wn::STAGE_BYTES;
        const int plane = n * C + c;
        mbar_expect_tx(&full[slot], Wn::OWN_BYTES + Wn::WIN_BYTES);
        tma_load_3d(base, &lv.own, w0, h0, plane, &full[slot]);
        tma_load_3d(base + Wn::OWN_BYTES, &lv.win, c0w, r0, plane, &full[slot]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) if (ca + s < cb) issue(s, ca + s);
    }

    // this thread's two locations and their samples (feature_refine_kernel.cu:127-151: px = column, py = row coordinate)
    Samp sp[2][P];
    bool valid[2];
    bool anyfar = false;
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int w = w0 + wl + 8 * q;
        valid[q] = (h < H) && (w < W);
        if (valid[q]) {
            const float* bb = lv.boxes + ((size_t)n * HW + (size_t)h * W + w) * 5;
            const float roi_y = __ldg(bb) * lv.scale, roi_x = __ldg(bb + 1) * lv.scale;
            float px[5], py[5];
            px[0] = roi_x; py[0] = roi_y;
            if (P > 1) {
                const float roi_w = __ldg(bb + 2) * lv.scale, roi_h = __ldg(bb + 3) * lv.scale, roi_a = __ldg(bb + 4);
                const float w_2 = roi_w / 2, h_2 = roi_h / 2;
                const float cosa = cosf(roi_a), sina = sinf(roi_a);
                const float wx = cosa * w_2, wy = sina * w_2;
                const float hx = -sina * h_2, hy = cosa * h_2;
                px[1] = roi_x + wx + hx; py[1] = roi_y + wy + hy;
                px[2] = roi_x - wx + hx; py[2] = roi_y - wy + hy;
                px[3] = roi_x - wx - hx; py[3] = roi_y - wy - hy;
                px[4] = roi_x + wx - hx; py[4] = roi_y + wy - hy;
            }
#pragma unroll
            for (int p = 0; p < P; p++) anyfar = make_samp(H, W, py[p], px[p], r0, c0w, Wn::RR, Wn::RC, sp[q][p]) || anyfar;
        } else {
#pragma unroll
            for (int p = 0; p < P; p++) make_samp(H, W, -4.0f, -4.0f, r0, c0w, Wn::RR, Wn::RC, sp[q][p]);
        }
    }
    anyfar = __any_sync(0xffffffffu, anyfar);                    // warp-uniform: the shared-memory-only loop, or the general one

    // running pointers (advance by one plane per channel); the second location is 8 floats further
    const size_t o0 = ((size_t)n * C + ca) * HW + (size_t)h * W + w0 + wl;
    const float* plane = lv.feat + ((size_t)n * C + ca) * HW;
    float* optr = lv.out + o0;
    const float* rptr = lv.residual ? lv.residual + o0 : nullptr;
    const unsigned char* ownb = fsm + (hl * TW + wl) * 4;        // this thread's own element inside a stage
    for (int cbase = ca, round = 0; cbase < cb; cbase += STAGES, round++) {
#pragma unroll
        for (int slot = 0; slot < STAGES; slot++) {               // the slot index is a compile-time constant: offsets fold into the loads
            if (cbase + slot < cb) {
                mbar_wait(&full[slot], (unsigned)(round & 1));
                const unsigned char* st = fsm + slot * Wn::STAGE_BYTES;
                const unsigned char* win = st + Wn::OWN_BYTES;
                float v[2];
                v[0] = *reinterpret_cast<const float*>(ownb + slot * Wn::STAGE_BYTES);
                v[1] = *reinterpret_cast<const float*>(ownb + slot * Wn::STAGE_BYTES + 32);
#pragma unroll
                for (int q = 0; q < 2; q++) {
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        const Samp& s = sp[q][p];
                        float t[4];
                        if (!anyfar) {
#pragma unroll
                            for (int k = 0; k < 4; k++) t[k] = *reinterpret_cast<const float*>(win + s.o[k]);
                        } else {                                   // some tap of the warp lies outside the staged window
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                t[k] = (s.o[k] >= 0) ? *reinterpret_cast<const float*>(win + s.o[k]) : __ldg(plane + (-s.o[k] - 1));
                        }
                        float sacc = s.w[0] * t[0];
                        sacc = fmaf(s.w[1], t[1], sacc);
                        sacc = fmaf(s.w[2], t[2], sacc);
                        sacc = fmaf(s.w[3], t[3], sacc);
                        v[q] += sacc;
                    }
                }
                if (rptr) {
                    if (valid[0]) v[0] = __ldg(rptr) + v[0];
                    if (valid[1]) v[1] = __ldg(rptr + 8) + v[1];
                    rptr += HW;
                }
                if (valid[0]) __stcs(optr, v[0]);
                if (valid[1]) __stcs(optr + 8, v[1]);
                optr += HW; plane += HW;
#if R3G_FRM_EMPTY_BARRIER
                // consumer release: no block-wide barrier — a warp signals that it is done with the slot and moves on; the
                // producer thread refills the PREVIOUS slot (whose readers have normally all passed by now) once its eight
                // arrivals are in, so the warps of a CTA drift apart and hide each other's shared-memory latency
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                if (tid == 0) {
                    const int pslot = (slot + STAGES - 1) % STAGES;
                    const int pc = cbase + slot - 1;                               // channel the previous slot held
                    if (pc >= ca && pc + STAGES < cb) {
                        mbar_wait(&empty[pslot], (unsigned)(((pc - ca) / STAGES) & 1));
                        issue(pslot, pc + STAGES);
                    }
                }
#else
                __syncthreads();                                   // every thread is done with this slot: the TMA unit may refill it
                if (tid == 0 && cbase + slot + STAGES < cb) issue(slot, cbase + slot + STAGES);
#endif
            }
        }
    }
}

// ---- backward apply: grad_in[t] = grad_out[t] + sum over the CSR row of t of w_e * grad_out[src_e] ------------------------
// The sources of target (y, x) are the locations whose sample landed on it: locations (h ~ x, w ~ y) — the transposed window
// again.  A thread owns two targets and keeps the first EMAX entries of each CSR row in registers (window offset + weight);
// longer rows (rare for points = 1, where the mean row has 4 entries) finish from global memory in CSR order, so the
// summation order — and the result, bit for bit — is that of the gather kernel in frm.cu.
#ifndef R3G_FRM_BWD_GROUP
#define R3G_FRM_BWD_GROUP 4                      // backward apply: entries of a row per warp-uniform test (measured: 1 -> 0.161, 2-4 -> 0.151, 6-12 -> 0.161 ms)
#endif
#ifndef R3G_FRM_EMAX
#define R3G_FRM_EMAX 12
#endif
#ifndef R3G_FRM_BWD_TMA_MINB
#define R3G_FRM_BWD_TMA_MINB 3
#endif
constexpr int EMAX = R3G_FRM_EMAX;      // P(row longer than 12) ~ 3e-4 at points = 1

template <int P>
__global__ void __launch_bounds__(THREADS, R3G_FRM_BWD_TMA_MINB) frm_backward_tma_kernel(const __grid_constant__ Levels S, const unsigned* __restrict__ row_start,
                                                                      const unsigned* __restrict__ src, const float* __restrict__ wsorted) {
    typedef Win<P, true> Wn;
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ __align__(8) unsigned long long full[STAGES];
    __shared__ __align__(8) unsigned long long empty[STAGES];
    int li = 0;
#pragma unroll
    for (int i = 1; i < MAXL; i++) if (i < S.L && blockIdx.x >= S.lv[i].block0) li = i;
    const Level& lv = S.lv[li];
    const int C = S.C, H = lv.H, W = lv.W;
    const int tiles = lv.tiles_x * lv.tiles_y;
    int bid = blockIdx.x - lv.block0;
    const int tl = bid % tiles; bid /= tiles;
    const int chunk = bid % lv.cchunks;
    const int n = bid / lv.cchunks;
    const int h0 = (tl / lv.tiles_x) * TH, w0 = (tl % lv.tiles_x) * TW;
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int hl = wi * 4 + (lane >> 3), wl = lane & 7;
    const int h = h0 + hl;
    const int HW = H * W;
    const int r0 = w0 - Wn::HR, c0w = h0 - Wn::HC;
    const int ca = chunk * CC, cb = min(C, ca + CC);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int slot, int c) {
        unsigned char* base = fsm + slot * Wn::STAGE_BYTES;
        const int plane = n * C + c;
        mbar_expect_tx(&full[slot], Wn::OWN_BYTES + Wn::WIN_BYTES);
        tma_load_3d(base, &lv.own, w0, h0, plane, &full[slot]);
        tma_load_3d(base + Wn::OWN_BYTES, &lv.win, c0w, r0, plane, &full[slot]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) if (ca + s < cb) issue(s, ca + s);
    }

    // CSR rows of this thread's two targets: entry e -> (byte offset of the source inside a stage, weight).  Entries a row does
    // not have read the stage's zero word with weight 0 (no predicates in the loop); a source outside the window is kept as
    // -(plane offset + 1).  Warps whose rows all fit (<= EMAX entries, all inside the window — almost all warps) run the
    // shared-memory-only loop: one LDS + FFMA per entry; the others take the general loop.
    float ew[2][EMAX];
    int eo[2][EMAX];
    unsigned e0[2], cnt[2];
    bool valid[2];
    bool slow = false;
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int w = w0 + wl + 8 * q;
        valid[q] = (h < H) && (w < W);
        e0[q] = 0; cnt[q] = 0;
        if (valid[q]) {
            const size_t t = (size_t)lv.loc0 + (size_t)n * HW + (size_t)h * W + w;
            e0[q] = __ldg(row_start + t);
            cnt[q] = __ldg(row_start + t + 1) - e0[q];
        }
        slow = slow || cnt[q] > (unsigned)EMAX;
#pragma unroll
        for (int e = 0; e < EMAX; e++) {
            ew[q][e] = 0.0f; eo[q][e] = Wn::ZERO_OFF;
            if ((unsigned)e < cnt[q]) {
                const int l = (int)__ldg(src + e0[q] + e);
                const int sr = l / W, sc = l - sr * W;
                ew[q][e] = __ldg(wsorted + e0[q] + e);
                const bool in = sr >= r0 && sr < r0 + Wn::RR && sc >= c0w && sc < c0w + Wn::RC;
                eo[q][e] = in ? (Wn::OWN_BYTES + ((sr - r0) * Wn::RC + (sc - c0w)) * 4) : -(l + 1);
                slow = slow || !in;
            }
        }
    }
    // the longest row of the warp bounds the unrolled loop (warp-uniform: no issue slots for entries nobody has)
    unsigned wmax = max(min(cnt[0], (unsigned)EMAX), min(cnt[1], (unsigned)EMAX));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    slow = __any_sync(0xffffffffu, slow);
    if (tid < STAGES) *reinterpret_cast<float*>(fsm + tid * Wn::STAGE_BYTES + Wn::ZERO_OFF) = 0.0f;
    __syncthreads();

    const float* plane = lv.feat + ((size_t)n * C + ca) * HW;
    float* optr = lv.out + ((size_t)n * C + ca) * HW + (size_t)h * W + w0 + wl;
    const int ownoff = (hl * TW + wl) * 4;
    for (int cbase = ca, round = 0; cbase < cb; cbase += STAGES, round++) {
#pragma unroll
        for (int slot = 0; slot < STAGES; slot++) {               // compile-time slot: stage offsets fold into the loads
            if (cbase + slot < cb) {
                mbar_wait(&full[slot], (unsigned)(round & 1));
                const unsigned char* st = fsm + slot * Wn::STAGE_BYTES;
                float v0 = *reinterpret_cast<const float*>(st + ownoff), v1 = *reinterpret_cast<const float*>(st + ownoff + 32);
                if (!slow) {
                    // entries in groups of R3G_FRM_BWD_GROUP: one warp-uniform test per group, the group's loads issued together
                    // (slots past a row's end point at the zero word with weight 0: reading them is harmless); the summation order
                    // of a row is unchanged
#pragma unroll
                    for (int e0g = 0; e0g < EMAX; e0g += R3G_FRM_BWD_GROUP) {
                        if ((unsigned)e0g >= wmax) break;
                        float g0[R3G_FRM_BWD_GROUP], g1[R3G_FRM_BWD_GROUP];
#pragma unroll
                        for (int k = 0; k < R3G_FRM_BWD_GROUP; k++) {
                            if (e0g + k < EMAX) {
                                g0[k] = *reinterpret_cast<const float*>(st + eo[0][e0g + k]);
                                g1[k] = *reinterpret_cast<const float*>(st + eo[1][e0g + k]);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < R3G_FRM_BWD_GROUP; k++) {
                            if (e0g + k < EMAX) { v0 = fmaf(ew[0][e0g + k], g0[k], v0); v1 = fmaf(ew[1][e0g + k], g1[k], v1); }
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        float v = q ? v1 : v0;
#pragma unroll
                        for (int e = 0; e < EMAX; e++) {
                            if ((unsigned)e < cnt[q]) {
                                const int o = eo[q][e];
                                const float g = (o >= 0) ? *reinterpret_cast<const float*>(st + o) : __ldg(plane + (-o - 1));
                                v = fmaf(ew[q][e], g, v);
                            }
                        }
                        for (unsigned e = e0[q] + EMAX; e < e0[q] + cnt[q]; e++)         // rows longer than EMAX: the rest, in CSR order
                            v = fmaf(__ldg(wsorted + e), __ldg(plane + __ldg(src + e)), v);
                        if (q) v1 = v; else v0 = v;
                    }
                }
                if (valid[0]) __stcs(optr, v0);
                if (valid[1]) __stcs(optr + 8, v1);
                optr += HW; plane += HW;
#if R3G_FRM_EMPTY_BARRIER
                __syncwarp();                                      // consumer release + delayed refill, as in the forward kernel
                if (lane == 0) mbar_arrive(&empty[slot]);
                if (tid == 0) {
                    const int pslot = (slot + STAGES - 1) % STAGES;
                    const int pc = cbase + slot - 1;
                    if (pc >= ca && pc + STAGES < cb) {
                        mbar_wait(&empty[pslot], (unsigned)(((pc - ca) / STAGES) & 1));
                        issue(pslot, pc + STAGES);
                    }
                }
#else
                __syncthreads();
                if (tid == 0 && cbase + slot + STAGES < cb) issue(slot, cbase + slot + STAGES);
#endif
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3-D map over (planes, H, W) of a contiguous NCHW tensor with a {bw, bh, 1} box; false when TMA cannot address it
static bool make_map(CUtensorMap* m, const float* base, int planes, int H, int W, int bw, int bh) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || (W % 4) != 0 || (((uintptr_t)base) & 15u) != 0 || bw > 256 || bh > 256) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes };
    const cuuint64_t strides[2] = { (cuuint64_t)W * 4, (cuuint64_t)H * W * 4 };
    const cuuint32_t box[3] = { (cuuint32_t)bw, (cuuint32_t)bh, 1 };
    const cuuint32_t es[3] = { 1, 1, 1 };
    static int promo = -1;
    if (promo < 0) { const char* e = getenv("R3G_TMA_L2PROMO"); promo = e ? atoi(e) : 2; }
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (getenv("R3G_TMA_DEBUG")) fprintf(stderr, "[r3g] tensor map %p planes %d H %d W %d box %d x %d -> %d\n", (const void*)base, planes, H, W, bw, bh, (int)r);
    return r == CUDA_SUCCESS;
}

}  // namespace ftma
}  // namespace r3g
