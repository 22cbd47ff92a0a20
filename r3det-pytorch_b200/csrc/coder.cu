// coder.cu — the rotated box coder and the steps on either side of multiclass NMS / FRM that use it, on sm_100a.
//
// Replaces (reference, relative to /root/reference):
//   DeltaXYWHAOBBoxCoder.encode / .decode               r3det/core/bbox/coder/delta_xywha_rbbox_coder.py:46-101
//     bbox2delta_v1 :104-139  delta2bbox_v1 :142-211  bbox2delta_v2 :214-252  delta2bbox_v2 :255-311
//     bbox2delta_v3 :314-360  delta2bbox_v3 :363-423
//   RAnchorHead._get_bboxes_single up to the NMS call    r3det/models/dense_heads/rotate_anchor_head.py:626-662
//     (permute/reshape, sigmoid, max over classes, per-level top-k(nms_pre), three gathers, decode, rescale, zero
//      background column: ~15 torch launches per level per image -> 5 launches + one segmented sort per BATCH)
//   RRetinaHead.filter_bboxes                            r3det/models/dense_heads/rotate_retina_head.py:117-179
//   RRetinaRefineHead.refine_bboxes                      r3det/models/dense_heads/rotate_retina_refine_head.py:56-97
//
// Arithmetic follows the torch expressions operation by operation in FP32 with explicitly rounded intrinsics (no FMA
// contraction), so results equal eager torch on the same device up to libm differences in exp/log/sin/cos.
// All network outputs are read in their native NCHW layout (no permute/reshape copies).  HBM-bound gather work.
#include <cub/block/block_scan.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <math.h>
#include "common.cuh"

namespace r3g {

constexpr float PI_F = 3.14159265358979323846f;
constexpr float HALF_PI_F = 1.57079632679489661923f;
constexpr float QUARTER_PI_F = 0.78539816339744830962f;
constexpr int SEL_MAX_LEVELS = 8;
constexpr int SEL_MAX_IMAGES = 64;

struct CoderP { float mean[5], stdv[5]; float max_ratio, ctr_clamp; int add_ctr_clamp, variant; };

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// torch.remainder: result takes the sign of the divisor
__device__ __forceinline__ float py_rem(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.0f && ((b < 0.0f) != (m < 0.0f))) m += b;
    return m;
}
__device__ __forceinline__ float wrap(float a, float off) { return sub(py_rem(add(a, off), PI_F), off); }   // (a + off) % pi - off

// delta2bbox_v1/v2/v3 for one (roi, delta) pair
__device__ __forceinline__ void decode_one(const CoderP& P, const float* roi, const float* d, bool clamp, float max_x, float max_y,
                                           float* o) {
    float dn[5];
#pragma unroll
    for (int k = 0; k < 5; k++) dn[k] = add(mul(d[k], P.stdv[k]), P.mean[k]);
    const float px = roi[0], py = roi[1], pw = roi[2], ph = roi[3], pa = roi[4];
    float dw = dn[2], dh = dn[3];
    if (P.variant == 1) {                                                          // :171-211
        float dxw = mul(pw, dn[0]), dyh = mul(ph, dn[1]);
        if (P.add_ctr_clamp) {
            dxw = clampf(dxw, -P.ctr_clamp, P.ctr_clamp); dyh = clampf(dyh, -P.ctr_clamp, P.ctr_clamp);
            dw = fminf(dw, P.max_ratio); dh = fminf(dh, P.max_ratio);
        } else {
            dw = clampf(dw, -P.max_ratio, P.max_ratio); dh = clampf(dh, -P.max_ratio, P.max_ratio);
        }
        float gx = add(px, dxw), gy = add(py, dyh);
        if (clamp) { gx = clampf(gx, 0.0f, max_x); gy = clampf(gy, 0.0f, max_y); }
        o[0] = gx; o[1] = gy; o[2] = mul(pw, expf(dw)); o[3] = mul(ph, expf(dh)); o[4] = add(pa, dn[4]);
        return;
    }
    dw = clampf(dw, -P.max_ratio, P.max_ratio); dh = clampf(dh, -P.max_ratio, P.max_ratio);
    const float ang = (P.variant == 2) ? pa : -pa;                                 // v2 :298-299, v3 :407-408
    const float c = cosf(ang), s = sinf(ang);
    const float ax = mul(dn[0], pw), ay = mul(dn[1], ph);
    o[0] = add(sub(mul(ax, c), mul(ay, s)), px);
    o[1] = add(add(mul(ax, s), mul(ay, c)), py);
    const float gw = mul(pw, expf(dw)), gh = mul(ph, expf(dh));
    if (P.variant == 2) {
        o[2] = gw; o[3] = gh;
        o[4] = wrap(add(mul(dn[4], PI_F), pa), QUARTER_PI_F);                      // :289, :302-303
    } else {
        const float gt = add(dn[4], pa);
        const bool big = gw > gh;                                                  // :413-417
        o[2] = big ? gw : gh; o[3] = big ? gh : gw;
        o[4] = wrap(big ? gt : add(gt, HALF_PI_F), HALF_PI_F);
    }
}

// bbox2delta_v1/v2/v3 for one (proposal, gt) pair
__device__ __forceinline__ void encode_one(const CoderP& P, const float* p, const float* g, float* o) {
    const float px = p[0], py = p[1], pw = p[2], ph = p[3], pa = p[4];
    const float gx = g[0], gy = g[1], gw = g[2], gh = g[3], ga = g[4];
    float d[5];
    if (P.variant == 1) {                                                          // :127-133
        d[0] = dvd(sub(gx, px), pw); d[1] = dvd(sub(gy, py), ph);
        d[2] = logf(dvd(gw, pw)); d[3] = logf(dvd(gh, ph)); d[4] = sub(ga, pa);
    } else {
        const float ang = (P.variant == 2) ? pa : -pa;
        const float c = cosf(ang), s = sinf(ang);
        const float ex = sub(gx, px), ey = sub(gy, py);
        d[0] = dvd(add(mul(c, ex), mul(s, ey)), pw);
        d[1] = dvd(add(mul(-s, ex), mul(c, ey)), ph);
        if (P.variant == 2) {                                                      // :237-243
            d[2] = logf(dvd(gw, pw)); d[3] = logf(dvd(gh, ph));
            d[4] = dvd(wrap(sub(ga, pa), QUARTER_PI_F), PI_F);
        } else {                                                                   // :338-352
            const float d1 = wrap(sub(ga, pa), HALF_PI_F);
            const float d2 = wrap(add(sub(ga, pa), HALF_PI_F), HALF_PI_F);
            const bool first = fabsf(d1) < fabsf(d2);
            d[2] = logf(dvd(first ? gw : gh, pw)); d[3] = logf(dvd(first ? gh : gw, ph));
            d[4] = first ? d1 : d2;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) o[k] = dvd(sub(d[k], P.mean[k]), P.stdv[k]);
}

__device__ __forceinline__ float sigmoidf(float x) { return dvd(1.0f, add(1.0f, expf(-x))); }

// ------------------------------------------------------------------------------------------------ plain coder calls
__global__ void delta2bbox_kernel(const float* __restrict__ rois, int64_t roi_stride, const float* __restrict__ deltas,
                                  int64_t n, int groups, CoderP P, int clamp, float max_x, float max_y, float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * groups) return;
    const int64_t i = t / groups;
    float roi[5], d[5], o[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { roi[k] = rois[i * roi_stride + k]; d[k] = deltas[t * 5 + k]; }
    decode_one(P, roi, d, clamp != 0, max_x, max_y, o);
#pragma unroll
    for (int k = 0; k < 5; k++) out[t * 5 + k] = o[k];
}

__global__ void bbox2delta_kernel(const float* __restrict__ prop, int64_t ps, const float* __restrict__ gt, int64_t gs, int64_t n,
                                  CoderP P, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p[5], g[5], o[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { p[k] = prop[i * ps + k]; g[k] = gt[i * gs + k]; }
    encode_one(P, p, g, o);
#pragma unroll
    for (int k = 0; k < 5; k++) out[i * 5 + k] = o[k];
}

// ------------------------------------------------------------------------------------------------ FRM prologue
// filter_bboxes: per location keep the anchor whose best class logit is largest (first maximum), decode it.
// A CTA covers 32 consecutive locations x all anchors: thread (x = location, y = anchor) reduces its C class planes
// (coalesced along H*W), the anchors are then compared through shared memory and thread y == 0 decodes.  One thread per
// location alone would leave ~175k threads with 135 serial loads each: latency-bound at 7 % of HBM.
constexpr int FILTER_LOCS = 32;
__global__ void filter_bboxes_kernel(const float* __restrict__ cls, const float* __restrict__ reg, const float* __restrict__ anchors,
                                     int B, int A, int C, int HW, CoderP P, float* __restrict__ out) {
    extern __shared__ float best_of[];                          // [A][FILTER_LOCS]
    const int b = blockIdx.y, hw = blockIdx.x * FILTER_LOCS + threadIdx.x, a = threadIdx.y;
    const bool valid = hw < HW;
    float m = 0.0f;
    if (valid) {
        const float* cb = cls + ((int64_t)b * A * C + (int64_t)a * C) * HW + hw;
        m = __ldg(cb);
        for (int c = 1; c < C; c++) {
            const float v = __ldg(cb + (int64_t)c * HW);
            m = (v > m || v != v) ? v : m;                      // torch.max propagates NaN
        }
    }
    best_of[a * FILTER_LOCS + threadIdx.x] = m;
    __syncthreads();
    if (a != 0 || !valid) return;
    int best = 0; float bestv = best_of[threadIdx.x];
    for (int k = 1; k < A; k++) {
        const float v = best_of[k * FILTER_LOCS + threadIdx.x];
        if ((v > bestv && bestv == bestv) || (v != v && bestv == bestv)) { best = k; bestv = v; }   // first maximum, NaN wins
    }
    float roi[5], d[5], o[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        roi[k] = __ldg(anchors + ((int64_t)hw * A + best) * 5 + k);
        d[k] = __ldg(reg + ((int64_t)b * A * 5 + best * 5 + k) * HW + hw);
    }
    decode_one(P, roi, d, false, 0.0f, 0.0f, o);
    const int64_t t = (int64_t)b * HW + hw;
#pragma unroll
    for (int k = 0; k < 5; k++) out[t * 5 + k] = o[k];
}

// refine_bboxes: decode the (B,5,H,W) deltas against per-image rois (B, H*W, 5)
__global__ void refine_bboxes_kernel(const float* __restrict__ reg, const float* __restrict__ rois, int B, int HW, CoderP P,
                                     float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * HW) return;
    const int b = (int)(t / HW), hw = (int)(t - (int64_t)b * HW);
    float roi[5], d[5], o[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { roi[k] = __ldg(rois + t * 5 + k); d[k] = __ldg(reg + ((int64_t)b * 5 + k) * HW + hw); }
    decode_one(P, roi, d, false, 0.0f, 0.0f, o);
#pragma unroll
    for (int k = 0; k < 5; k++) out[t * 5 + k] = o[k];
}

// ------------------------------------------------------------------------------------------------ get_bboxes tail
struct LevelDesc {
    const float* cls;            // (B, A*C, H, W)
    const float* reg;            // (B, A*5, H, W)
    const float* anchors;        // (n, 5) shared by the batch, or (B, n, 5) when anchor_batch_stride != 0
    int64_t anchor_batch_stride; // in floats
    int HW, n, k;                // locations, rows (= HW * A), rows kept (= min(nms_pre, n) or n)
    int row0, out0;              // prefix sums of n and k over the levels
    int blk0;                    // first CTA (of 256 rows) of this level inside an image's grid row
};
struct SelectArgs {
    LevelDesc lv[SEL_MAX_LEVELS];
    int L, B, A, C, n_total, k_total, blocks_per_image;
    int row_bits;                // bits of a row index inside its level; sort key = (u - 0xC0000000) << row_bits | row
    int clamp, rescale;
    float max_x[SEL_MAX_IMAGES], max_y[SEL_MAX_IMAGES];
    float sf[SEL_MAX_IMAGES][4];
    CoderP P;
};

constexpr int SEL_BIN_BITS = 12;                 // histogram resolution of the top-k threshold search
constexpr int SEL_BINS = 1 << SEL_BIN_BITS;

__device__ __forceinline__ int level_of_out(const SelectArgs& S, int r) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < SEL_MAX_LEVELS; i++) if (i < S.L && r >= S.lv[i].out0) l = i;
    return l;
}
__device__ __forceinline__ int level_of_block(const SelectArgs& S, int b) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < SEL_MAX_LEVELS; i++) if (i < S.L && b >= S.lv[i].blk0) l = i;
    return l;
}

// Per-level top-k(nms_pre) by max-class sigmoid score, in three steps that never sort the ~200k rows of an image:
//   1. one key per row, u = ~bits(sigmoid(max logit)) (ascending u = descending score; sigmoid is monotone, so this is
//      the reference's `scores.max(1)`), plus a per-(image, level) histogram of the top 12 bits of u;
//   2. per segment the smallest bin T whose cumulative count reaches k;
//   3. rows with bin <= T (k plus the tail of one bin) are compacted as 64-bit keys (u << row_bits | row) and ONLY those are
//      sorted (segmented radix sort): ties resolve to the lower row index, deterministically.
// Levels with n <= nms_pre are kept whole in row order (as the reference leaves them) and skip all of this.
// grid = (blocks_per_image, B); a CTA works inside one (image, level) segment, threads run along H*W (coalesced planes).
__global__ void __launch_bounds__(256) select_hist_kernel(const __grid_constant__ SelectArgs S, unsigned* __restrict__ ukey,
                                                          int* __restrict__ hist) {
    __shared__ int h[SEL_BINS];
    const int l = level_of_block(S, blockIdx.x);
    const LevelDesc& lv = S.lv[l];
    if (lv.k >= lv.n) return;                               // level kept whole
    const int b = blockIdx.y;
    for (int t = threadIdx.x; t < SEL_BINS; t += 256) h[t] = 0;
    __syncthreads();
    const int local = (blockIdx.x - lv.blk0) * 256 + threadIdx.x;        // anchor-major, location-minor
    int bin = -1;
    if (local < lv.n) {
        const int a = local / lv.HW, hw = local - a * lv.HW;
        const float* cb = lv.cls + ((int64_t)b * S.A * S.C + (int64_t)a * S.C) * lv.HW + hw;
        float m = __ldg(cb);
        int c = 1;
        for (; c + 4 <= S.C; c += 4) {                      // four class planes in flight per step
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = __ldg(cb + (int64_t)(c + k) * lv.HW);
#pragma unroll
            for (int k = 0; k < 4; k++) m = (v[k] > m || v[k] != v[k]) ? v[k] : m;     // torch.max propagates NaN
        }
        for (; c < S.C; c++) {
            const float v = __ldg(cb + (int64_t)c * lv.HW);
            m = (v > m || v != v) ? v : m;
        }
        unsigned u = ~__float_as_uint(sigmoidf(m));
        u = u < 0xC0000000u ? 0xC0000000u : u;             // NaN scores rank first, as torch.topk ranks them
        ukey[(int64_t)b * S.n_total + lv.row0 + hw * S.A + a] = u;       // at the row's own slot
        bin = (int)(u >> (32 - SEL_BIN_BITS));
    }
    // background rows crowd into a handful of bins: one shared-memory atomic per distinct bin and warp instead of one per row
    {
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin >= 0 && (int)(threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&h[bin], __popc(peers));
    }
    __syncthreads();
    int* g = hist + (int64_t)(b * S.L + l) * SEL_BINS;
    for (int t = threadIdx.x; t < SEL_BINS; t += 256)
        if (h[t]) atomicAdd(g + t, h[t]);
}

// one CTA per (image, level): threshold bin + the segment's compaction cursor
__global__ void __launch_bounds__(256) select_threshold_kernel(const __grid_constant__ SelectArgs S, const int* __restrict__ hist,
                                                               int* __restrict__ thr, int* __restrict__ seg_begin,
                                                               int* __restrict__ seg_end) {
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int best;
    const int seg = blockIdx.x, b = seg / S.L, l = seg - b * S.L;
    const LevelDesc& lv = S.lv[l];
    const int base = b * S.n_total + lv.row0;
    if (threadIdx.x == 0) { seg_begin[seg] = base; seg_end[seg] = base; best = SEL_BINS - 1; }
    if (lv.k >= lv.n) { if (threadIdx.x == 0) thr[seg] = -1; return; }
    __syncthreads();
    constexpr int PER = SEL_BINS / 256;
    const int* g = hist + (int64_t)seg * SEL_BINS + threadIdx.x * PER;
    int c[PER], sum = 0;
#pragma unroll
    for (int t = 0; t < PER; t++) { c[t] = g[t]; sum += c[t]; }
    int before;
    Scan(tmp).ExclusiveSum(sum, before);
    if (before < lv.k && before + sum >= lv.k) {            // the k-th best row lies in one of this thread's bins
        int run = before, t = 0;
        for (; t < PER; t++) { run += c[t]; if (run >= lv.k) break; }
        best = threadIdx.x * PER + t;
    }
    __syncthreads();
    if (threadIdx.x == 0) thr[seg] = best;
}

__global__ void __launch_bounds__(256) select_compact_kernel(const __grid_constant__ SelectArgs S, const unsigned* __restrict__ ukey,
                                                             const int* __restrict__ thr, int* __restrict__ seg_end,
                                                             unsigned long long* __restrict__ keys) {
    const int l = level_of_block(S, blockIdx.x);
    const LevelDesc& lv = S.lv[l];
    if (lv.k >= lv.n) return;
    const int b = blockIdx.y, seg = b * S.L + l;
    const int n = (blockIdx.x - lv.blk0) * 256 + threadIdx.x;            // row order here: plain coalesced key reads
    bool take = false;
    unsigned u = 0;
    if (n < lv.n) {
        u = ukey[(int64_t)b * S.n_total + lv.row0 + n];
        take = (int)(u >> (32 - SEL_BIN_BITS)) <= thr[seg];
    }
    // warp-aggregated append to the segment's candidate list
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (bal) {
        const unsigned lane = threadIdx.x & 31u;
        int base = 0;
        if (lane == (unsigned)(__ffs(bal) - 1)) base = atomicAdd(seg_end + seg, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        // sigmoid lies in [0, 1], so u >= 0xC07FFFFF: the two top bits carry nothing and are dropped from the sort key
        if (take) keys[base + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)(u - 0xC0000000u) << S.row_bits) | (unsigned)n;
    }
}

// One thread per kept row: gather anchor + deltas, decode, clamp / rescale; then the C scores and the zero column.
__global__ void select_decode_kernel(const __grid_constant__ SelectArgs S, const unsigned long long* __restrict__ sorted,
                                     float* __restrict__ boxes, float* __restrict__ scores) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)S.B * S.k_total) return;
    const int b = (int)(t / S.k_total), r = (int)(t - (int64_t)b * S.k_total);
    const int l = level_of_out(S, r);
    const LevelDesc& lv = S.lv[l];
    const int rr = r - lv.out0;
    const int n = (lv.k < lv.n) ? (int)(sorted[(int64_t)b * S.n_total + lv.row0 + rr] & ((1ull << S.row_bits) - 1ull)) : rr;
    const int hw = n / S.A, a = n - hw * S.A;
    float roi[5], d[5], o[5];
    const float* an = lv.anchors + (int64_t)b * lv.anchor_batch_stride + (int64_t)n * 5;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        roi[k] = __ldg(an + k);
        d[k] = __ldg(lv.reg + ((int64_t)b * S.A * 5 + a * 5 + k) * lv.HW + hw);
    }
    decode_one(S.P, roi, d, S.clamp != 0, S.max_x[b], S.max_y[b], o);
    if (S.rescale) {
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = dvd(o[k], S.sf[b][k]);
    }
#pragma unroll
    for (int k = 0; k < 5; k++) boxes[t * 5 + k] = o[k];
    const float* cb = lv.cls + ((int64_t)b * S.A * S.C + (int64_t)a * S.C) * lv.HW + hw;
    float* so = scores + t * (S.C + 1);
    for (int c = 0; c < S.C; c++) so[c] = sigmoidf(__ldg(cb + (int64_t)c * lv.HW));
    so[S.C] = 0.0f;                                      // the dummy background column (rotate_anchor_head.py:659-664)
}

static int make_coder(const char* who, const float* means, const float* stds, int variant, double wh_ratio_clip, int add_ctr_clamp,
                      float ctr_clamp, CoderP* P) {
    R3G_REQUIRE(variant >= 1 && variant <= 3, "%s: variant must be 1, 2 or 3", who);
    R3G_REQUIRE(means != nullptr && stds != nullptr, "%s: null means / stds", who);
    R3G_REQUIRE(wh_ratio_clip > 0.0, "%s: wh_ratio_clip must be positive", who);
    for (int k = 0; k < 5; k++) { P->mean[k] = means[k]; P->stdv[k] = stds[k]; }
    P->max_ratio = (float)fabs(log(wh_ratio_clip));
    P->ctr_clamp = ctr_clamp; P->add_ctr_clamp = add_ctr_clamp; P->variant = variant;
    return R3G_OK;
}

struct SelectWs {
    unsigned long long *keys_in, *keys_out;
    unsigned* ukey;
    int *hist, *thr, *seg_begin, *seg_end;
    void* cub; size_t cub_bytes, bytes;
};

static int carve_select(void* ws, int64_t items, int64_t segments, SelectWs* w) {
    char* p = (char*)ws;
    size_t off = 0;
    const size_t n = (size_t)(items > 0 ? items : 1), sg = (size_t)(segments > 0 ? segments : 1);
    w->keys_in = (unsigned long long*)(p + off); off += align_up(8 * n, 256);
    w->keys_out = (unsigned long long*)(p + off); off += align_up(8 * n, 256);
    w->ukey = (unsigned*)(p + off); off += align_up(4 * n, 256);
    w->hist = (int*)(p + off); off += align_up(4 * sg * SEL_BINS, 256);
    w->thr = (int*)(p + off); off += align_up(4 * sg, 256);
    w->seg_begin = (int*)(p + off); off += align_up(4 * sg, 256);
    w->seg_end = (int*)(p + off); off += align_up(4 * sg, 256);
    // the sort ping-pongs between keys_in and keys_out (cub::DoubleBuffer), so cub itself needs no scratch worth
    // mentioning; a fixed kilobyte is handed over (and no device query is needed to size the workspace)
    w->cub = p + off; w->cub_bytes = 1024; off += 1024;
    w->bytes = off;
    return R3G_OK;
}

}  // namespace r3g

using namespace r3g;

R3G_API int r3g_delta2bbox_f32(const float* rois, int64_t n, int64_t roi_stride, const float* deltas, int64_t groups,
                               const float* means, const float* stds, int variant, const int* max_shape_hw,
                               double wh_ratio_clip, int add_ctr_clamp, float ctr_clamp, float* out, void* stream) {
    CoderP P;
    int rc = make_coder("r3g_delta2bbox_f32", means, stds, variant, wh_ratio_clip, add_ctr_clamp, ctr_clamp, &P);
    if (rc != R3G_OK) return rc;
    R3G_REQUIRE(n >= 0 && groups >= 1 && roi_stride >= 5, "r3g_delta2bbox_f32: bad sizes");
    if (n == 0) return R3G_OK;
    R3G_REQUIRE(rois && deltas && out, "r3g_delta2bbox_f32: null pointer");
    const bool clamp = (variant == 1) && max_shape_hw != nullptr;      // only delta2bbox_v1 takes max_shape (:97-99)
    const float mx = clamp ? (float)(max_shape_hw[1] - 1) : 0.0f, my = clamp ? (float)(max_shape_hw[0] - 1) : 0.0f;
    const int64_t total = n * groups;
    delta2bbox_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rois, roi_stride, deltas, n, (int)groups, P,
                                                                                      clamp ? 1 : 0, mx, my, out);
    R3G_LAUNCH_OK("delta2bbox_kernel");
    return R3G_OK;
}

R3G_API int r3g_bbox2delta_f32(const float* proposals, int64_t proposal_stride, const float* gt, int64_t gt_stride, int64_t n,
                               const float* means, const float* stds, int variant, float* out, void* stream) {
    CoderP P;
    int rc = make_coder("r3g_bbox2delta_f32", means, stds, variant, 16.0 / 1000.0, 0, 0.0f, &P);
    if (rc != R3G_OK) return rc;
    R3G_REQUIRE(n >= 0 && proposal_stride >= 5 && gt_stride >= 5, "r3g_bbox2delta_f32: bad sizes");
    if (n == 0) return R3G_OK;
    R3G_REQUIRE(proposals && gt && out, "r3g_bbox2delta_f32: null pointer");
    bbox2delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(proposals, proposal_stride, gt, gt_stride, n, P, out);
    R3G_LAUNCH_OK("bbox2delta_kernel");
    return R3G_OK;
}

R3G_API int r3g_filter_bboxes_f32(const float* cls_score, const float* bbox_pred, const float* anchors, int64_t B, int64_t A,
                                  int64_t C, int64_t H, int64_t W, const float* means, const float* stds, int variant,
                                  double wh_ratio_clip, int add_ctr_clamp, float ctr_clamp, float* out, void* stream) {
    CoderP P;
    int rc = make_coder("r3g_filter_bboxes_f32", means, stds, variant, wh_ratio_clip, add_ctr_clamp, ctr_clamp, &P);
    if (rc != R3G_OK) return rc;
    R3G_REQUIRE(B >= 0 && A >= 1 && C >= 1 && H >= 0 && W >= 0 && H * W * A < (1ll << 31), "r3g_filter_bboxes_f32: bad sizes");
    R3G_REQUIRE(A <= 32 && B < 65536, "r3g_filter_bboxes_f32: at most 32 anchors per location and 65535 images per call");
    const int64_t total = B * H * W;
    if (total == 0) return R3G_OK;
    R3G_REQUIRE(cls_score && bbox_pred && anchors && out, "r3g_filter_bboxes_f32: null pointer");
    const dim3 grid((unsigned)((H * W + FILTER_LOCS - 1) / FILTER_LOCS), (unsigned)B), block(FILTER_LOCS, (unsigned)A);
    filter_bboxes_kernel<<<grid, block, sizeof(float) * FILTER_LOCS * (size_t)A, (cudaStream_t)stream>>>(cls_score, bbox_pred, anchors, (int)B,
                                                                                                      (int)A, (int)C, (int)(H * W), P, out);
    R3G_LAUNCH_OK("filter_bboxes_kernel");
    return R3G_OK;
}

R3G_API int r3g_refine_bboxes_f32(const float* bbox_pred, const float* rois, int64_t B, int64_t H, int64_t W, const float* means,
                                  const float* stds, int variant, double wh_ratio_clip, int add_ctr_clamp, float ctr_clamp,
                                  float* out, void* stream) {
    CoderP P;
    int rc = make_coder("r3g_refine_bboxes_f32", means, stds, variant, wh_ratio_clip, add_ctr_clamp, ctr_clamp, &P);
    if (rc != R3G_OK) return rc;
    R3G_REQUIRE(B >= 0 && H >= 0 && W >= 0 && H * W < (1ll << 31), "r3g_refine_bboxes_f32: bad sizes");
    const int64_t total = B * H * W;
    if (total == 0) return R3G_OK;
    R3G_REQUIRE(bbox_pred && rois && out, "r3g_refine_bboxes_f32: null pointer");
    refine_bboxes_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(bbox_pred, rois, (int)B, (int)(H * W), P, out);
    R3G_LAUNCH_OK("refine_bboxes_kernel");
    return R3G_OK;
}

static int select_sizes(const char* who, int64_t L, int64_t B, int64_t A, const int64_t* hw, int64_t nms_pre, int64_t* n_total,
                        int64_t* k_total) {
    R3G_REQUIRE(L >= 1 && L <= SEL_MAX_LEVELS, "%s: 1..%d levels supported", who, SEL_MAX_LEVELS);
    R3G_REQUIRE(B >= 0 && B <= SEL_MAX_IMAGES, "%s: at most %d images per call", who, SEL_MAX_IMAGES);
    R3G_REQUIRE(A >= 1 && hw != nullptr, "%s: bad arguments", who);
    int64_t nt = 0, kt = 0;
    for (int l = 0; l < L; l++) {
        R3G_REQUIRE(hw[2 * l] >= 0 && hw[2 * l + 1] >= 0, "%s: negative feature-map size", who);
        const int64_t n = hw[2 * l] * hw[2 * l + 1] * A;
        nt += n; kt += (nms_pre > 0 && n > nms_pre) ? nms_pre : n;
    }
    R3G_REQUIRE(nt < (1ll << 31) && B * nt < (1ll << 31), "%s: too many rows", who);
    *n_total = nt; *k_total = kt;
    return R3G_OK;
}

R3G_API int r3g_select_decode_sizes(int64_t L, int64_t B, int64_t A, const int64_t* level_hw, int64_t nms_pre,
                                    int64_t* rows_per_image, size_t* workspace_bytes) {
    int64_t nt = 0, kt = 0;
    int rc = select_sizes("r3g_select_decode_sizes", L, B, A, level_hw, nms_pre, &nt, &kt);
    if (rc != R3G_OK) return rc;
    R3G_REQUIRE(rows_per_image && workspace_bytes, "r3g_select_decode_sizes: null output");
    SelectWs w;
    rc = carve_select(nullptr, B * nt, B * L, &w);
    if (rc != R3G_OK) return rc;
    *rows_per_image = kt; *workspace_bytes = w.bytes;
    return R3G_OK;
}

R3G_API int r3g_select_decode_f32(int64_t L, const float* const* cls_scores, const float* const* bbox_preds,
                                  const float* const* anchors, const int64_t* anchor_batch_strides, const int64_t* level_hw,
                                  int64_t B, int64_t A, int64_t C, int64_t nms_pre,
                                  const float* means, const float* stds, int variant, double wh_ratio_clip, int add_ctr_clamp,
                                  float ctr_clamp, const int* max_shapes_hw, const float* scale_factors,
                                  float* boxes_out, float* scores_out, void* workspace, size_t workspace_bytes, void* stream) {
    int64_t nt = 0, kt = 0;
    int rc = select_sizes("r3g_select_decode_f32", L, B, A, level_hw, nms_pre, &nt, &kt);
    if (rc != R3G_OK) return rc;
    SelectArgs S;
    rc = make_coder("r3g_select_decode_f32", means, stds, variant, wh_ratio_clip, add_ctr_clamp, ctr_clamp, &S.P);
    if (rc != R3G_OK) return rc;
    R3G_REQUIRE(C >= 1, "r3g_select_decode_f32: need at least one class");
    if (B == 0 || kt == 0) return R3G_OK;
    R3G_REQUIRE(cls_scores && bbox_preds && anchors && boxes_out && scores_out && workspace, "r3g_select_decode_f32: null pointer");
    SelectWs w;
    rc = carve_select(workspace, B * nt, B * L, &w);
    if (rc != R3G_OK) return rc;
    if (workspace_bytes < w.bytes) {
        set_error("r3g_select_decode_f32: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    S.L = (int)L; S.B = (int)B; S.A = (int)A; S.C = (int)C; S.n_total = (int)nt; S.k_total = (int)kt;
    int row0 = 0, out0 = 0, blk0 = 0;
    bool any_topk = false;
    for (int l = 0; l < SEL_MAX_LEVELS; l++) {
        LevelDesc& lv = S.lv[l];
        if (l >= L) { lv = S.lv[0]; lv.n = lv.k = 0; lv.row0 = row0; lv.out0 = out0; lv.blk0 = blk0; continue; }
        const int64_t HW = level_hw[2 * l] * level_hw[2 * l + 1];
        R3G_REQUIRE(HW == 0 || (cls_scores[l] && bbox_preds[l] && anchors[l]), "r3g_select_decode_f32: null level pointer");
        lv.cls = cls_scores[l]; lv.reg = bbox_preds[l]; lv.anchors = anchors[l];
        lv.anchor_batch_stride = anchor_batch_strides ? anchor_batch_strides[l] : 0;
        lv.HW = (int)HW; lv.n = (int)(HW * A); lv.k = (nms_pre > 0 && lv.n > nms_pre) ? (int)nms_pre : lv.n;
        lv.row0 = row0; lv.out0 = out0; lv.blk0 = blk0;
        row0 += lv.n; out0 += lv.k; blk0 += (lv.n + 255) / 256;
        any_topk |= lv.k < lv.n;
    }
    S.blocks_per_image = blk0;
    int max_n = 1;
    for (int l = 0; l < L; l++) max_n = S.lv[l].n > max_n ? S.lv[l].n : max_n;
    S.row_bits = 1;
    while ((1ll << S.row_bits) < max_n) S.row_bits++;
    S.clamp = (variant == 1 && max_shapes_hw != nullptr) ? 1 : 0;
    S.rescale = scale_factors != nullptr ? 1 : 0;
    for (int b = 0; b < B; b++) {
        S.max_x[b] = S.clamp ? (float)(max_shapes_hw[2 * b + 1] - 1) : 0.0f;
        S.max_y[b] = S.clamp ? (float)(max_shapes_hw[2 * b] - 1) : 0.0f;
        for (int k = 0; k < 4; k++) S.sf[b][k] = S.rescale ? scale_factors[4 * b + k] : 1.0f;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long* sorted_keys = w.keys_out;
    if (any_topk) {
        const int segs = (int)(B * L);
        R3G_CUDA_OK(cudaMemsetAsync(w.hist, 0, sizeof(int) * (size_t)segs * SEL_BINS, st));
        const dim3 grid((unsigned)S.blocks_per_image, (unsigned)B);
        select_hist_kernel<<<grid, 256, 0, st>>>(S, w.ukey, w.hist);
        select_threshold_kernel<<<segs, 256, 0, st>>>(S, w.hist, w.thr, w.seg_begin, w.seg_end);
        select_compact_kernel<<<grid, 256, 0, st>>>(S, w.ukey, w.thr, w.seg_end, w.keys_in);
        R3G_LAUNCH_OK("select kernels");
        size_t tb = w.cub_bytes;
        cub::DoubleBuffer<unsigned long long> db(w.keys_in, w.keys_out);
        R3G_CUDA_OK(cub::DeviceSegmentedRadixSort::SortKeys(w.cub, tb, db, B * nt, segs, (const int*)w.seg_begin,
                                                            (const int*)w.seg_end, 0, 30 + S.row_bits, st));
        sorted_keys = db.Current();
    }
    const int64_t rows = B * kt;
    select_decode_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(S, sorted_keys, boxes_out, scores_out);
    R3G_LAUNCH_OK("select_decode_kernel");
    return R3G_OK;
}
