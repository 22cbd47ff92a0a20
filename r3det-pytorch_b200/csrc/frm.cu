// frm.cu — FRM feature refinement (forward + atomic-free backward) on sm_100a.
//
// Replaces (reference, relative to /root/reference):
//   feature_refine_forward_kernel / feature_refine_backward_kernel   r3det/ops/fr/src/feature_refine_kernel.cu:112-230
// Reference: one thread per NCHW element; every thread re-reads its location's box, re-derives the sample
// points (sinf/cosf for points=5) and the bilinear weights — identical work repeated for all C=256
// channels — and the backward pass issues 1+4*points float atomics per element into a pre-zeroed buffer.
// Here the sample taps depend only on (n, h, w), so:
//   forward : a thread owns one location, derives its taps ONCE into registers and streams the channels.  Lanes run
//             along h: with the reference's x/y swap the taps of h-adjacent locations are adjacent in memory, so
//             the gathers coalesce; the location's own element / the output go through a padded shared-memory
//             tile with a row-oriented mapping, software-pipelined one round (8 channels) ahead in registers.
//             Blocks are ordered (tile, channel-chunk, image) so concurrently running CTAs gather from the same
//             few planes (L2-resident).  HBM traffic: read feat once + write out once = 8 B/element; no zero-fill
//             pass (the reference's Python zero-fills `output` first).
//   backward: the taps of an image are inverted once per call into a per-target CSR by a radix sort on the
//             target index ("sorted scatter"); a thread owns one target pixel and GATHERS
//             grad_in[t] = grad_out[t] + sum_e w_e * grad_out[src_e] for 32 channels per pass.  No atomics,
//             no pre-zeroed output, bit-reproducible (stable sort => fixed summation order).
//   All FPN levels of a batch go through ONE launch sequence (r3g_frm_*_multi_f32): the three small levels are pure
//   launch latency when run alone (a few hundred CTAs each), and the backward's sort / CSR build is shared.
//   The module's residual add  x + FR(...)  (feature_refine_module.py:126) can ride along in the forward epilogue.
// Sample-point math follows feature_refine_kernel.cu:16-65 (interpolation) and :127-151 (points), including
// the reference's swap of box x -> row and box y -> column.
#include <cub/cub.cuh>
#include <stdlib.h>
#include "common.cuh"
#include "frm_tma.cuh"

namespace r3g {

constexpr int FRM_THREADS = 256;
constexpr int FRM_CCHUNK_FWD = 32;   // channels per forward CTA
#ifndef R3G_FRM_BWD_CC
#define R3G_FRM_BWD_CC 32
#endif
#ifndef R3G_FRM_BWD_MINB
#define R3G_FRM_BWD_MINB 2
#endif
constexpr int FRM_CC_BWD = R3G_FRM_BWD_CC;       // channels accumulated together per backward thread

struct Taps4 { float w[4]; int o[4]; };

// bilinear taps of one sample; invalid samples get zero weights and offset 0 (feature_refine_kernel.cu:16-110)
__device__ __forceinline__ bool frm_taps(int H, int W, float y, float x, Taps4& t) {
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
#pragma unroll
        for (int k = 0; k < 4; k++) { t.w[k] = 0.0f; t.o[k] = 0; }
        return false;
    }
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
    const float ly = y - (float)yl, lx = x - (float)xl;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    t.w[0] = hy * hx; t.w[1] = hy * lx; t.w[2] = ly * hx; t.w[3] = ly * lx;
    t.o[0] = yl * W + xl; t.o[1] = yl * W + xh; t.o[2] = yh * W + xl; t.o[3] = yh * W + xh;
    return true;
}

// sample positions of one location (feature_refine_kernel.cu:127-151): px = column coordinate, py = row coordinate
template <int P>
__device__ __forceinline__ void frm_points(const float* __restrict__ bb, float scale, float* px, float* py) {
    const float roi_y = bb[0] * scale;       // box x -> ROW coordinate   (:131)
    const float roi_x = bb[1] * scale;       // box y -> COLUMN coordinate (:132)
    px[0] = roi_x; py[0] = roi_y;
    if (P > 1) {
        const float roi_w = bb[2] * scale, roi_h = bb[3] * scale, roi_a = bb[4];
        const float w_2 = roi_w / 2, h_2 = roi_h / 2;
        const float cosa = cosf(roi_a), sina = sinf(roi_a);
        const float wx = cosa * w_2, wy = sina * w_2;
        const float hx = -sina * h_2, hy = cosa * h_2;
        px[1] = roi_x + wx + hx; py[1] = roi_y + wy + hy;
        px[2] = roi_x - wx + hx; py[2] = roi_y - wy + hy;
        px[3] = roi_x - wx - hx; py[3] = roi_y - wy - hy;
        px[4] = roi_x + wx - hx; py[4] = roi_y + wy - hy;
    }
}

// Location tile of a CTA: tile_h consecutive rows (h) x tile_w consecutive columns (w), tile_h * tile_w = 256.
// Compute mapping: thread t owns location (h0 + t % tile_h, w0 + t / tile_h) — LANES RUN ALONG h.  The reference
// samples box x as the ROW coordinate and box y as the COLUMN coordinate (feature_refine_kernel.cu:131-132), so for
// refined boxes near their own location the taps of lane-adjacent locations (h, h+1, ...) are adjacent in memory and
// every gather is a handful of coalesced sectors instead of 32 scattered ones.  The location's own element (and the
// output) would be strided by W in that mapping, so they are moved through a padded shared-memory tile with a
// row-oriented mapping (t / tile_w, t % tile_w): global traffic stays fully coalesced in both directions.
struct TileGeom { int tile_w, tile_h, tiles_x, tiles_y; };

static TileGeom tile_geom(int H, int W) {
    TileGeom g;
    g.tile_h = 32;
    while (g.tile_h > 4 && g.tile_h / 2 >= H) g.tile_h /= 2;
    g.tile_w = FRM_THREADS / g.tile_h;
    (void)W;
    g.tiles_x = (W + g.tile_w - 1) / g.tile_w;
    g.tiles_y = (H + g.tile_h - 1) / g.tile_h;
    return g;
}

constexpr int FRM_MAX_LEVELS = 8;
struct FrmLevel {
    const float* feat;       // forward: features; backward: grad_out        (N, C, H, W)
    const float* boxes;      // (N*H*W, 5)
    const float* residual;   // forward only, optional: added to the output   (N, C, H, W)
    float* out;              // forward: out; backward: grad_in
    int H, W;
    float scale;
    TileGeom g;
    int cchunks;
    unsigned block0;         // first CTA of this level
    unsigned loc0;           // first global location (n, h, w) of this level (backward CSR indexing)
};
struct FrmLevels { FrmLevel lv[FRM_MAX_LEVELS]; int L, N, C; unsigned nl; };

__device__ __forceinline__ int frm_level_of_block(const FrmLevels& S, unsigned b) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < FRM_MAX_LEVELS; i++) if (i < S.L && b >= S.lv[i].block0) l = i;
    return l;
}
__device__ __forceinline__ int frm_level_of_loc(const FrmLevels& S, unsigned i) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < FRM_MAX_LEVELS; k++) if (k < S.L && i >= S.lv[k].loc0) l = k;
    return l;
}

constexpr int FRM_TILE_PITCH = 296;      // >= tile_h * (tile_w + 1) for every geometry (32x9, 16x17, 8x33, 4x65)
constexpr int FRM_CH_FWD = 8;            // channels per shared-memory round (forward)

template <int P>
__global__ void __launch_bounds__(FRM_THREADS, (P == 1) ? 4 : 2) frm_forward_kernel(const __grid_constant__ FrmLevels S) {
    __shared__ float tile[FRM_CH_FWD][FRM_TILE_PITCH];
    const FrmLevel& lv = S.lv[frm_level_of_block(S, blockIdx.x)];
    const float* __restrict__ feat = lv.feat;
    const float* __restrict__ boxes = lv.boxes;
    const float* __restrict__ resid = lv.residual;
    float* __restrict__ out = lv.out;
    const int C = S.C, H = lv.H, W = lv.W, cchunks = lv.cchunks;
    const float scale = lv.scale;
    const TileGeom g = lv.g;
    // block = tile + tiles * (chunk + cchunks * n): tiles fastest so that co-resident CTAs share planes
    const int tiles = g.tiles_x * g.tiles_y;
    int bid = blockIdx.x - lv.block0;
    const int tl = bid % tiles; bid /= tiles;
    const int chunk = bid % cchunks;
    const int n = bid / cchunks;
    const int h0 = (tl / g.tiles_x) * g.tile_h, w0 = (tl % g.tiles_x) * g.tile_w;
    const int pitch = g.tile_w + 1;
    // compute mapping (lanes along h)
    const int ch = threadIdx.x % g.tile_h, cw = threadIdx.x / g.tile_h;
    const int h = h0 + ch, w = w0 + cw;
    const bool cvalid = (h < H) && (w < W);
    // row-oriented mapping for the coalesced global load / store of the tile
    const int rh = threadIdx.x / g.tile_w, rw = threadIdx.x % g.tile_w;
    const bool rvalid = (h0 + rh < H) && (w0 + rw < W);
    const int HW = H * W;
    const int rloc = (h0 + rh) * W + (w0 + rw);
    const int cslot = ch * pitch + cw, rslot = rh * pitch + rw;

    Taps4 t[P];
    if (cvalid) {
        float px[5], py[5];
        const float* bb = boxes + ((size_t)n * HW + (size_t)h * W + w) * 5;
        float b5[5] = { __ldg(bb), __ldg(bb + 1), __ldg(bb + 2), __ldg(bb + 3), __ldg(bb + 4) };
        frm_points<P>(b5, scale, px, py);
#pragma unroll
        for (int p = 0; p < P; p++) frm_taps(H, W, py[p], px[p], t[p]);
    } else {
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int k = 0; k < 4; k++) { t[p].w[k] = 0.0f; t[p].o[k] = 0; }
    }

    const int c0 = chunk * FRM_CCHUNK_FWD, c1 = min(C, c0 + FRM_CCHUNK_FWD);
    // software pipeline: the next round's own-elements are in flight (registers) while this round gathers
    float nxt[FRM_CH_FWD];
    {
        const float* plane = feat + ((size_t)n * C + c0) * HW + rloc;
#pragma unroll
        for (int cc = 0; cc < FRM_CH_FWD; cc++) nxt[cc] = (rvalid && c0 + cc < c1) ? __ldg(plane + (size_t)cc * HW) : 0.0f;
    }
    for (int cb = c0; cb < c1; cb += FRM_CH_FWD) {
        const float* plane = feat + ((size_t)n * C + cb) * HW;
        float* oplane = out + ((size_t)n * C + cb) * HW + rloc;
#pragma unroll
        for (int cc = 0; cc < FRM_CH_FWD; cc++) tile[cc][rslot] = nxt[cc];
        __syncthreads();
        if (cb + FRM_CH_FWD < c1) {
            const float* np = plane + (size_t)FRM_CH_FWD * HW + rloc;
#pragma unroll
            for (int cc = 0; cc < FRM_CH_FWD; cc++)
                nxt[cc] = (rvalid && cb + FRM_CH_FWD + cc < c1) ? __ldg(np + (size_t)cc * HW) : 0.0f;
        }
        if (cvalid) {
#pragma unroll
            for (int cc = 0; cc < FRM_CH_FWD; cc++) {
                if (cb + cc < c1) {
                    const float* pl = plane + (size_t)cc * HW;
                    float v = tile[cc][cslot];
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        float sacc = t[p].w[0] * __ldg(pl + t[p].o[0]);
                        sacc = fmaf(t[p].w[1], __ldg(pl + t[p].o[1]), sacc);
                        sacc = fmaf(t[p].w[2], __ldg(pl + t[p].o[2]), sacc);
                        sacc = fmaf(t[p].w[3], __ldg(pl + t[p].o[3]), sacc);
                        v += sacc;
                    }
                    tile[cc][cslot] = v;
                }
            }
        }
        __syncthreads();
        // each thread re-reads only its own row-oriented slot here and overwrites only that slot at the top of the
        // next round, and every compute-mapped read of this round happened before the barrier above: no third barrier
        if (rvalid) {
            if (resid != nullptr) {
                const float* rp = resid + ((size_t)n * C + cb) * HW + rloc;
                float rv[FRM_CH_FWD];
#pragma unroll
                for (int cc = 0; cc < FRM_CH_FWD; cc++) rv[cc] = (cb + cc < c1) ? __ldg(rp + (size_t)cc * HW) : 0.0f;
#pragma unroll
                for (int cc = 0; cc < FRM_CH_FWD; cc++)
                    if (cb + cc < c1) __stcs(oplane + (size_t)cc * HW, rv[cc] + tile[cc][rslot]);
            } else {
#pragma unroll
                for (int cc = 0; cc < FRM_CH_FWD; cc++)
                    if (cb + cc < c1) __stcs(oplane + (size_t)cc * HW, tile[cc][rslot]);
            }
        }
    }
}

// ---- backward: tap generation -> sort by target -> CSR -> gather ------------------------------------------

// Plan = the taps of the call inverted into a per-target CSR.  The targets of the taps are counted (one integer atomic per tap),
// an exclusive scan gives the row starts, the taps are dropped into their rows in arrival order and every row (4 entries on
// average for points = 1) is then sorted by (source location, point, corner) — the order a stable sort of the taps by target
// would produce — so the summation order of the gather is fixed: bit-reproducible, and no radix sort over all taps.
template <int P>
__global__ void frm_bwd_count_kernel(const __grid_constant__ FrmLevels S, unsigned* __restrict__ keys, float* __restrict__ wts,
                                     unsigned* __restrict__ cnt) {
    const size_t nl = S.nl;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nl) return;
    const FrmLevel& lv = S.lv[frm_level_of_loc(S, (unsigned)i)];
    const int H = lv.H, W = lv.W, HW = H * W;
    const float scale = lv.scale;
    const unsigned li = (unsigned)i - lv.loc0;               // location inside the level
    const int n = (int)(li / HW);
    const float* bb = lv.boxes + (size_t)li * 5;
    float b5[5] = { __ldg(bb), __ldg(bb + 1), __ldg(bb + 2), __ldg(bb + 3), __ldg(bb + 4) };
    float px[5], py[5];
    frm_points<P>(b5, scale, px, py);
    const unsigned none = 0xffffffffu;                       // invalid samples take no part
#pragma unroll
    for (int p = 0; p < P; p++) {
        Taps4 t;
        const bool ok = frm_taps(H, W, py[p], px[p], t);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const size_t e = (i * P + p) * 4 + k;
            const unsigned key = ok ? (lv.loc0 + (unsigned)((size_t)n * HW + t.o[k])) : none;
            keys[e] = key;
            wts[e] = t.w[k];
            if (ok) atomicAdd(cnt + key, 1u);
        }
    }
}

__global__ void frm_bwd_fill_kernel(const __grid_constant__ FrmLevels S, const unsigned* __restrict__ keys, const float* __restrict__ wts,
                                    size_t E, int P, const unsigned* __restrict__ row_start, unsigned* __restrict__ cnt,
                                    unsigned* __restrict__ src, float* __restrict__ wsorted) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const unsigned key = keys[e];
    if (key == 0xffffffffu) return;
    const unsigned slot = atomicSub(cnt + key, 1u) - 1u;     // the counters run back down to zero
    const unsigned pos = row_start[key] + slot;
    const unsigned loc = (unsigned)(e / (size_t)(4 * P));
    const unsigned tie = (unsigned)(e - (size_t)loc * (4 * P));                     // point * 4 + corner, < 20
    const FrmLevel& lv = S.lv[frm_level_of_loc(S, loc)];
    src[pos] = (((loc - lv.loc0) % (unsigned)(lv.H * lv.W)) << 5) | tie;           // source location inside its image (< 2^24), tie-break
    wsorted[pos] = wts[e];
}

// one thread per target: its row sorted by (source, tie), then the tie bits are dropped.  Rows of up to CAP entries are ranked
// in registers (keys are unique: rank = number of smaller keys; every entry is written straight to its final slot); longer
// rows — boxes piling onto one pixel — fall back to an insertion sort in global memory.
constexpr unsigned FRM_SORT_MAX_ROW = 512;
template <int CAP>
__global__ void frm_bwd_sort_rows_kernel(const unsigned* __restrict__ row_start, size_t nl, unsigned* __restrict__ src, float* __restrict__ wsorted) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nl) return;
    const unsigned a = row_start[t], b = row_start[t + 1];
    const unsigned n = b - a;
    if (n <= (unsigned)CAP) {
        unsigned k[CAP];
        float w[CAP];
#pragma unroll
        for (int i = 0; i < CAP; i++) {
            k[i] = ((unsigned)i < n) ? src[a + i] : 0xffffffffu;
            w[i] = ((unsigned)i < n) ? wsorted[a + i] : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < CAP; i++) {
            if ((unsigned)i < n) {
                unsigned r = 0;
#pragma unroll
                for (int j = 0; j < CAP; j++) r += (k[j] < k[i]) ? 1u : 0u;
                src[a + r] = k[i] >> 5;
                wsorted[a + r] = w[i];
            }
        }
        return;
    }
    // pathological pile-ups (hundreds of boxes sampling one pixel) stay in arrival order: still correct — the reference sums with
    // float atomics in arbitrary order — but not bit-reproducible, instead of a quadratic serial sort that could run for seconds
    if (n > FRM_SORT_MAX_ROW) { for (unsigned i = a; i < b; i++) src[i] >>= 5; return; }
    for (unsigned i = a + 1; i < b; i++) {
        const unsigned kk = src[i];
        const float ww = wsorted[i];
        unsigned j = i;
        while (j > a && src[j - 1] > kk) { src[j] = src[j - 1]; wsorted[j] = wsorted[j - 1]; j--; }
        src[j] = kk; wsorted[j] = ww;
    }
    for (unsigned i = a; i < b; i++) src[i] >>= 5;
}

__global__ void __launch_bounds__(FRM_THREADS, R3G_FRM_BWD_MINB) frm_backward_kernel(
    const __grid_constant__ FrmLevels S, const unsigned* __restrict__ row_start, const unsigned* __restrict__ src,
    const float* __restrict__ wsorted) {
    __shared__ float tile[FRM_CC_BWD][FRM_TILE_PITCH];
    const FrmLevel& lv = S.lv[frm_level_of_block(S, blockIdx.x)];
    const float* __restrict__ gout = lv.feat;
    float* __restrict__ gin = lv.out;
    const int C = S.C, H = lv.H, W = lv.W, cchunks = lv.cchunks;
    const TileGeom g = lv.g;
    const int tiles = g.tiles_x * g.tiles_y;
    int bid = blockIdx.x - lv.block0;
    const int tl = bid % tiles; bid /= tiles;
    const int chunk = bid % cchunks;
    const int n = bid / cchunks;
    const int h0 = (tl / g.tiles_x) * g.tile_h, w0 = (tl % g.tiles_x) * g.tile_w;
    const int pitch = g.tile_w + 1;
    const int ch = threadIdx.x % g.tile_h, cw = threadIdx.x / g.tile_h;     // lanes along h: the sources of
    const int h = h0 + ch, w = w0 + cw;                                     // lane-adjacent targets are adjacent
    const bool cvalid = (h < H) && (w < W);
    const int rh = threadIdx.x / g.tile_w, rw = threadIdx.x % g.tile_w;
    const bool rvalid = (h0 + rh < H) && (w0 + rw < W);
    const int HW = H * W;
    const int rloc = (h0 + rh) * W + (w0 + rw);
    const int cslot = ch * pitch + cw, rslot = rh * pitch + rw;
    const int c0 = chunk * FRM_CC_BWD;
    const float* base = gout + ((size_t)n * C + c0) * HW;
    float* obase = gin + ((size_t)n * C + c0) * HW;

#pragma unroll
    for (int k = 0; k < FRM_CC_BWD; k++)
        if (rvalid && c0 + k < C) tile[k][rslot] = __ldg(base + (size_t)k * HW + rloc);
    __syncthreads();
    if (cvalid) {
        const size_t t = (size_t)lv.loc0 + (size_t)n * HW + (size_t)h * W + w;
        const unsigned e0 = __ldg(row_start + t), e1 = __ldg(row_start + t + 1);
        float acc[FRM_CC_BWD];
#pragma unroll
        for (int k = 0; k < FRM_CC_BWD; k++) acc[k] = tile[k][cslot];
        if (c0 + FRM_CC_BWD <= C) {
            for (unsigned e = e0; e < e1; e++) {
                const unsigned l = __ldg(src + e);
                const float wt = __ldg(wsorted + e);
#pragma unroll
                for (int k = 0; k < FRM_CC_BWD; k++) acc[k] = fmaf(wt, __ldg(base + (size_t)k * HW + l), acc[k]);
            }
        } else {
            for (unsigned e = e0; e < e1; e++) {
                const unsigned l = __ldg(src + e);
                const float wt = __ldg(wsorted + e);
#pragma unroll
                for (int k = 0; k < FRM_CC_BWD; k++)
                    if (c0 + k < C) acc[k] = fmaf(wt, __ldg(base + (size_t)k * HW + l), acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < FRM_CC_BWD; k++) tile[k][cslot] = acc[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FRM_CC_BWD; k++)
        if (rvalid && c0 + k < C) __stcs(obase + (size_t)k * HW + rloc, tile[k][rslot]);
}

struct FrmBwdWs {
    unsigned *keys, *cnt, *row_start, *src;
    float *wts, *wsorted;
    void* cub_tmp; size_t cub_bytes;
    size_t bytes;
};

static FrmBwdWs carve_frm(void* ws, size_t nl, int P) {
    FrmBwdWs w;
    const size_t E = nl * P * 4;
    char* p = (char*)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes, 256); return (void*)r; };
    w.keys = (unsigned*)take(4 * E);
    w.wts = (float*)take(4 * E); w.wsorted = (float*)take(4 * E);
    w.src = (unsigned*)take(4 * E);
    w.cnt = (unsigned*)take(4 * (nl + 1));
    w.row_start = (unsigned*)take(4 * (nl + 1));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (unsigned*)nullptr, (unsigned*)nullptr, (int)(nl + 1));
    w.cub_bytes = align_up(tb, 256);
    w.cub_tmp = take(w.cub_bytes);
    w.bytes = off;
    return w;
}

// Validate a level list and lay the levels out (CTA ranges, location ranges).  Returns 1 when there is nothing to do.
static int frm_plan(const char* who, int L, const float* const* in, const float* const* boxes, const float* const* residuals,
                    float* const* out, int N, int C, const int* hw, const float* scales, int points, int cc, FrmLevels* S,
                    size_t* blocks_out) {
    R3G_REQUIRE(L >= 1 && L <= FRM_MAX_LEVELS, "%s: 1..%d levels per call", who, FRM_MAX_LEVELS);
    R3G_REQUIRE(N >= 0 && C >= 0 && hw != nullptr, "%s: bad arguments", who);
    R3G_REQUIRE(points == 1 || points == 5, "%s: points must be 1 or 5 (got %d)", who, points);   // feature_refine_module.py:19
    size_t blocks = 0, nl = 0;
    S->L = 0; S->N = N; S->C = C;
    for (int l = 0; l < L; l++) {
        const int H = hw[2 * l], W = hw[2 * l + 1];
        R3G_REQUIRE(H >= 0 && W >= 0, "%s: negative dimension", who);
        if ((size_t)N * C * H * W == 0) continue;                      // empty level: nothing to compute
        R3G_REQUIRE((size_t)H * W < (1ull << 24), "%s: feature map too large", who);
        R3G_REQUIRE(in && boxes && out && in[l] && boxes[l] && out[l] && scales, "%s: null pointer", who);
        FrmLevel& lv = S->lv[S->L++];
        lv.feat = in[l]; lv.boxes = boxes[l]; lv.residual = residuals ? residuals[l] : nullptr; lv.out = out[l];
        lv.H = H; lv.W = W; lv.scale = scales[l];
        lv.g = tile_geom(H, W);
        lv.cchunks = (C + cc - 1) / cc;
        lv.block0 = (unsigned)blocks; lv.loc0 = (unsigned)nl;
        blocks += (size_t)lv.g.tiles_x * lv.g.tiles_y * lv.cchunks * N;
        nl += (size_t)N * H * W;
    }
    R3G_REQUIRE(blocks < (1ull << 31), "%s: grid too large", who);
    R3G_REQUIRE(nl * (size_t)points * 4 < (1ull << 31), "%s: too many sample taps for 32-bit indexing", who);
    S->nl = (unsigned)nl;
    *blocks_out = blocks;
    return S->L == 0 ? 1 : 0;
}

}  // namespace r3g

using namespace r3g;

// TMA path of the forward (frm_tma.cuh): every level must be addressable by a tensor map (W % 4 == 0, 16-byte aligned base)
static bool frm_tma_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("R3G_FRM_TMA"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

// tensor maps + tile layout of every level; false when a level cannot be addressed by the TMA unit (the gather kernels serve it)
template <int P, bool BWD>
static bool frm_tma_levels(const FrmLevels& S, ftma::Levels& T, size_t* blocks_out) {
    T.L = S.L; T.N = S.N; T.C = S.C;
    size_t blocks = 0;
    for (int l = 0; l < S.L; l++) {
        const FrmLevel& a = S.lv[l];
        ftma::Level& b = T.lv[l];
        if (a.H > 32767 || a.W > 32767) return false;
        if (!ftma::make_map(&b.own, a.feat, S.N * S.C, a.H, a.W, ftma::TW, ftma::TH)) return false;
        if (!ftma::make_map(&b.win, a.feat, S.N * S.C, a.H, a.W, ftma::Win<P, BWD>::RC, ftma::Win<P, BWD>::RR)) return false;
        b.feat = a.feat; b.boxes = a.boxes; b.residual = a.residual; b.out = a.out;
        b.H = a.H; b.W = a.W; b.scale = a.scale;
        b.tiles_x = (a.W + ftma::TW - 1) / ftma::TW; b.tiles_y = (a.H + ftma::TH - 1) / ftma::TH;
        b.cchunks = (S.C + ftma::CC - 1) / ftma::CC;
        b.block0 = (unsigned)blocks; b.loc0 = a.loc0;
        blocks += (size_t)b.tiles_x * b.tiles_y * b.cchunks * S.N;
    }
    *blocks_out = blocks;
    return blocks < (1ull << 31);
}

template <int P>
static int frm_forward_tma(const FrmLevels& S, cudaStream_t st, bool* done) {
    *done = false;
    ftma::Levels T;
    size_t blocks = 0;
    if (!frm_tma_levels<P, false>(S, T, &blocks)) return R3G_OK;
    constexpr int smem = ftma::STAGES * ftma::Win<P>::STAGE_BYTES;
    if (smem > 48 * 1024) R3G_CUDA_OK(cudaFuncSetAttribute(ftma::frm_forward_tma_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ftma::frm_forward_tma_kernel<P><<<(unsigned)blocks, ftma::THREADS, smem, st>>>(T);
    R3G_LAUNCH_OK("frm_forward_tma_kernel");
    *done = true;
    return R3G_OK;
}

R3G_API int r3g_frm_forward_multi_f32(int L, const float* const* feats, const float* const* boxes, const float* const* residuals,
                                      int N, int C, const int* level_hw, const float* spatial_scales, int points,
                                      float* const* outs, void* stream) {
    FrmLevels S;
    size_t blocks = 0;
    int rc = frm_plan("r3g_frm_forward_multi_f32", L, feats, boxes, residuals, outs, N, C, level_hw, spatial_scales, points,
                      FRM_CCHUNK_FWD, &S, &blocks);
    if (rc < 0) return rc;
    if (rc == 1) return R3G_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (frm_tma_enabled()) {
        bool done = false;
        rc = (points == 1) ? frm_forward_tma<1>(S, st, &done) : frm_forward_tma<5>(S, st, &done);
        if (rc != R3G_OK) return rc;
        if (done) return R3G_OK;
    }
    if (points == 1) frm_forward_kernel<1><<<(unsigned)blocks, FRM_THREADS, 0, st>>>(S);
    else frm_forward_kernel<5><<<(unsigned)blocks, FRM_THREADS, 0, st>>>(S);
    R3G_LAUNCH_OK("frm_forward_kernel");
    return R3G_OK;
}

R3G_API int r3g_frm_forward_f32(const float* feat, const float* boxes, int N, int C, int H, int W,
                                float spatial_scale, int points, float* out, void* stream) {
    const int hw[2] = { H, W };
    return r3g_frm_forward_multi_f32(1, &feat, &boxes, nullptr, N, C, hw, &spatial_scale, points, &out, stream);
}

R3G_API int r3g_frm_backward_multi_workspace_bytes(int L, int N, const int* level_hw, int points, size_t* bytes) {
    R3G_REQUIRE(bytes != nullptr && L >= 1 && L <= FRM_MAX_LEVELS && N >= 0 && level_hw != nullptr && (points == 1 || points == 5),
                "r3g_frm_backward_multi_workspace_bytes: bad arguments");
    size_t nl = 0;
    for (int l = 0; l < L; l++) {
        R3G_REQUIRE(level_hw[2 * l] >= 0 && level_hw[2 * l + 1] >= 0, "r3g_frm_backward_multi_workspace_bytes: negative dimension");
        nl += (size_t)N * level_hw[2 * l] * level_hw[2 * l + 1];
    }
    R3G_REQUIRE(nl * (size_t)points * 4 < (1ull << 31), "r3g_frm_backward_multi_workspace_bytes: too many sample taps");
    *bytes = carve_frm(nullptr, nl > 0 ? nl : 1, points).bytes;
    return R3G_OK;
}

R3G_API int r3g_frm_backward_workspace_bytes(int N, int H, int W, int points, size_t* bytes) {
    const int hw[2] = { H, W };
    return r3g_frm_backward_multi_workspace_bytes(1, N, hw, points, bytes);
}

// The backward in two halves: the PLAN (tap sort + CSR) depends on the boxes only, so a training step can build it on a
// side stream while the forward pass is still running; APPLY is the gather over that CSR.
static int frm_bwd_plan(const FrmLevels& S, int points, void* workspace, size_t workspace_bytes, cudaStream_t st, const char* who) {
    R3G_REQUIRE(workspace != nullptr, "%s: null workspace", who);
    const size_t nl = S.nl, E = nl * points * 4;
    FrmBwdWs w = carve_frm(workspace, nl, points);
    if (workspace_bytes < w.bytes) {
        set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    const int tpb = 256;
    R3G_CUDA_OK(cudaMemsetAsync(w.cnt, 0, 4 * (nl + 1), st));
    if (points == 1) frm_bwd_count_kernel<1><<<(unsigned)((nl + tpb - 1) / tpb), tpb, 0, st>>>(S, w.keys, w.wts, w.cnt);
    else frm_bwd_count_kernel<5><<<(unsigned)((nl + tpb - 1) / tpb), tpb, 0, st>>>(S, w.keys, w.wts, w.cnt);
    size_t tb = w.cub_bytes;
    R3G_CUDA_OK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.cnt, w.row_start, (int)(nl + 1), st));
    frm_bwd_fill_kernel<<<(unsigned)((E + tpb - 1) / tpb), tpb, 0, st>>>(S, w.keys, w.wts, E, points, w.row_start, w.cnt, w.src, w.wsorted);
    if (points == 1) frm_bwd_sort_rows_kernel<8><<<(unsigned)((nl + 127) / 128), 128, 0, st>>>(w.row_start, nl, w.src, w.wsorted);
    else frm_bwd_sort_rows_kernel<32><<<(unsigned)((nl + 127) / 128), 128, 0, st>>>(w.row_start, nl, w.src, w.wsorted);
    R3G_LAUNCH_OK("frm backward plan kernels");
    return R3G_OK;
}

static int frm_bwd_apply(const FrmLevels& S, size_t blocks, int points, void* workspace, size_t workspace_bytes, cudaStream_t st,
                         const char* who) {
    R3G_REQUIRE(workspace != nullptr, "%s: null workspace", who);
    FrmBwdWs w = carve_frm(workspace, S.nl, points);
    if (workspace_bytes < w.bytes) {
        set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    if (points == 1 && frm_tma_enabled()) {           // points = 5: rows of ~20 entries — the register-resident rows do not pay there
        ftma::Levels T;
        size_t tb = 0;
        if (frm_tma_levels<1, true>(S, T, &tb)) {
            ftma::frm_backward_tma_kernel<1><<<(unsigned)tb, ftma::THREADS, ftma::STAGES * ftma::Win<1, true>::STAGE_BYTES, st>>>(
                T, w.row_start, w.src, w.wsorted);
            R3G_LAUNCH_OK("frm_backward_tma_kernel");
            return R3G_OK;
        }
    }
    frm_backward_kernel<<<(unsigned)blocks, FRM_THREADS, 0, st>>>(S, w.row_start, w.src, w.wsorted);
    R3G_LAUNCH_OK("frm_backward_kernel");
    return R3G_OK;
}

R3G_API int r3g_frm_backward_plan_multi_f32(int L, const float* const* boxes, int N, const int* level_hw, const float* spatial_scales,
                                            int points, void* workspace, size_t workspace_bytes, void* stream) {
    // the plan ignores the feature pointers and the channel count: a one-channel table over the boxes pointers
    FrmLevels S;
    size_t blocks = 0;
    int rc = frm_plan("r3g_frm_backward_plan_multi_f32", L, boxes, boxes, nullptr, (float* const*)boxes, N, 1, level_hw, spatial_scales,
                      points, FRM_CC_BWD, &S, &blocks);
    if (rc < 0) return rc;
    if (rc == 1) return R3G_OK;
    return frm_bwd_plan(S, points, workspace, workspace_bytes, (cudaStream_t)stream, "r3g_frm_backward_plan_multi_f32");
}

R3G_API int r3g_frm_backward_apply_multi_f32(int L, const float* const* grad_outs, const float* const* boxes, int N, int C,
                                             const int* level_hw, const float* spatial_scales, int points, float* const* grad_ins,
                                             const void* workspace, size_t workspace_bytes, void* stream) {
    FrmLevels S;
    size_t blocks = 0;
    int rc = frm_plan("r3g_frm_backward_apply_multi_f32", L, grad_outs, boxes, nullptr, grad_ins, N, C, level_hw, spatial_scales, points,
                      FRM_CC_BWD, &S, &blocks);
    if (rc < 0) return rc;
    if (rc == 1) return R3G_OK;
    return frm_bwd_apply(S, blocks, points, const_cast<void*>(workspace), workspace_bytes, (cudaStream_t)stream,
                         "r3g_frm_backward_apply_multi_f32");
}

R3G_API int r3g_frm_backward_multi_f32(int L, const float* const* grad_outs, const float* const* boxes, int N, int C,
                                       const int* level_hw, const float* spatial_scales, int points, float* const* grad_ins,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    FrmLevels S;
    size_t blocks = 0;
    int rc = frm_plan("r3g_frm_backward_multi_f32", L, grad_outs, boxes, nullptr, grad_ins, N, C, level_hw, spatial_scales, points,
                      FRM_CC_BWD, &S, &blocks);
    if (rc < 0) return rc;
    if (rc == 1) return R3G_OK;
    rc = frm_bwd_plan(S, points, workspace, workspace_bytes, (cudaStream_t)stream, "r3g_frm_backward_multi_f32");
    if (rc != R3G_OK) return rc;
    return frm_bwd_apply(S, blocks, points, workspace, workspace_bytes, (cudaStream_t)stream, "r3g_frm_backward_multi_f32");
}

R3G_API int r3g_frm_backward_f32(const float* grad_out, const float* boxes, int N, int C, int H, int W,
                                 float spatial_scale, int points, float* grad_in,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    const int hw[2] = { H, W };
    return r3g_frm_backward_multi_f32(1, &grad_out, &boxes, N, C, hw, &spatial_scale, points, &grad_in, workspace, workspace_bytes, stream);
}
