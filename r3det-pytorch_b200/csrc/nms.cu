// nms.cu — rotated NMS (single class and per-class segmented) and polygon NMS entirely on the device, sm_100a.
//
// Replaces (reference, relative to /root/reference):
//   nmsr_kernel + nmsr_cuda host scan            r3det/ops/rnms/src/rcuda/rnms_kernel.cu:229-335        (v1)
//   nms_rotated_cuda_kernel + host scan          r3det/ops/nms_rotated/src/nms_rotated_cuda.cu:12-134    (v3)
//   ml_nms_rotated                               r3det/ops/ml_nms_rotated/src/nms_rotated_cuda.cu:13-137 (v2)
//   poly_nms_kernel + host scan                  r3det/ops/nms_rotated/src/poly_nms_cuda.cu:122-262
// The reference computes the FULL K x K bitmask (cross-class pairs included, via a coordinate offset
// trick), copies it to the host and scans it serially on the CPU.  Here:
//   1. position space p = (segment asc, score desc), segment = label or (image, label): two stable radix sorts, or —
//      for K <= 16384, where the sorts are pure launch latency — ranks COUNTED in one O(K^2) kernel;
//   2. boxes are gathered/prepared once in p-order (class offsets applied in FP32 exactly as the
//      reference's batched wrappers do, so the geometry sees the same rounded coordinates);
//   3. ONE persistent kernel (nms_rounds.cuh) runs the greedy selection in rounds — chunk of the next <= 2048 alive
//      candidates per segment, chunk-local bitmask, scan, kept rows applied to the rest of the segment — so that only
//      pairs (kept row, later column) and pairs inside a chunk are ever evaluated, the workspace is linear in K and no
//      host round trip exists;
//   4. keep bits -> flags -> exclusive scan -> keep list in score order or index order.
// Pairs whose IoU lies within `margin` of the threshold (or that trip the degeneracy test) are decided
// by the reference's own algorithm (emu.cuh), which makes the keep set the reference's.
#include <cub/cub.cuh>
#include <stdlib.h>
#include "common.cuh"
#include "emu.cuh"
#include "geom.cuh"
#include "nms_rounds.cuh"
#include "poly.cuh"

namespace r3g {

constexpr float NMS_SLACK = 1.0f / 262144.0f;  // 2^-18 relative slack of the expanded circumradius test
constexpr int NMS_SMALL_K = 16384;             // up to here ranks are counted instead of sorted

// rows of a mask item are split over this many warps when K is small (the kernel is then bound by one warp's serial work)
static int nms_split_of(int64_t K) { return K <= 4096 ? 8 : (K <= 8192 ? 4 : (K <= 16384 ? 2 : 1)); }

struct NmsWs {
    unsigned char* ctrl;                            // 1 KB: rn::Ctrl[2] at 0, barrier counter at 128, counters / time stamps at 512
    unsigned *keyA, *keyA2, *keyB, *keyB2;
    int *ord_rank, *ord_tmp, *pos_rank, *pos_tmp;   // ord_rank[r] = original index of rank r; pos_rank[p] = rank at position p
    unsigned* pos_label;
    float4 *p0, *p1, *p2r, *p2c; float* raw;
    unsigned long long* alive; size_t alive_bytes, zero_bytes, zero_bytes_small;
    int *seg_cur, *seg_pe, *act;
    rn::Entry* ent;
    int *spos, *klist, *ownerB, *ownerD;
    int *flag, *pref, *keep_p;
    void* cub_tmp; size_t cub_bytes;
    unsigned long long* mask;
    size_t bytes;
};

static size_t cub_temp_bytes(int64_t K) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr, (int)K);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, (int)K);
    return align_up(a > b ? a : b, 256);
}

static NmsWs carve_nms(void* ws, int64_t K) {
    NmsWs w;
    char* p = (char*)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes, 256); return (void*)r; };
    // zeroed by one memset per call: control block, alive bits, keep flags (and the two rank counters of the small-K path)
    w.ctrl = (unsigned char*)take(1024);
    w.alive_bytes = 16 * (size_t)((K + 127) / 128);                     // two 64-bit words per 128-column group
    w.alive = (unsigned long long*)take(w.alive_bytes);
    w.keep_p = (int*)take(4 * K);
    w.zero_bytes = off;
    w.keyA = (unsigned*)take(4 * K); w.keyA2 = (unsigned*)take(4 * K);
    w.zero_bytes_small = off;
    w.keyB = (unsigned*)take(4 * K); w.keyB2 = (unsigned*)take(4 * K);
    w.ord_rank = (int*)take(4 * K); w.ord_tmp = (int*)take(4 * K);
    w.pos_rank = (int*)take(4 * K); w.pos_tmp = (int*)take(4 * K);
    w.pos_label = (unsigned*)take(4 * K);
    w.p0 = (float4*)take(16 * K); w.p1 = (float4*)take(16 * K);
    w.p2r = (float4*)take(32 * K); w.p2c = (float4*)take(16 * K);       // row plane: every constant twice (f32x2 operands)
    w.raw = (float*)take(20 * K);
    w.seg_cur = (int*)take(4 * K); w.seg_pe = (int*)take(4 * K); w.act = (int*)take(8 * K);
    w.ent = (rn::Entry*)take(sizeof(rn::Entry) * (size_t)(K / 2 + 2));  // chunks of >= 2 rows
    w.spos = (int*)take(4 * K); w.klist = (int*)take(4 * K);
    // mask items: <= K/2 single-block chunks + 17 n / 64 for the others; apply items: <= K B / 8192 + K / 32
    w.ownerB = (int*)take(4 * (size_t)(K + 64) * (size_t)nms_split_of(K));
    w.ownerD = (int*)take(4 * (size_t)(K / 2 + 64));
    w.flag = (int*)take(4 * K); w.pref = (int*)take(4 * K);
    w.cub_bytes = cub_temp_bytes(K);
    w.cub_tmp = take(w.cub_bytes);
    w.mask = (unsigned long long*)take((size_t)8 * (size_t)K * (rn::B_MAX / 64 + 2));   // sum over chunks of n * (ceil(n / 64) rounded up to even)
    w.bytes = off;
    return w;
}

__device__ __forceinline__ unsigned score_key_desc(float s) {
    unsigned b = __float_as_uint(s);
    unsigned asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~asc;
}

__global__ void nms_keys_kernel(const float* __restrict__ scores, int K, unsigned* keyA, int* idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) { keyA[i] = score_key_desc(scores[i]); idx[i] = i; }
}

// segment key of candidate i: the label, or (image << 16 | label) for multi-image batches (image-major: an image's
// segments are contiguous in position space).  In a batch both fields are 16 bits wide.  A candidate that cannot take part —
// image id outside [0, n_batches), label outside [0, 65536); the padding of r3g_mc_candidates_batched_f32 is of this kind —
// gets the ONE reserved key 0xffffffff instead of aliasing into another image's or class's segment.  One key for all of them
// matters: the rounds kernel (phase 0) relies on every key occupying exactly ONE run of the position order.
constexpr unsigned NMS_DEAD_KEY = 0xffffffffu;
__device__ __forceinline__ unsigned seg_key_of(const int64_t* labels, const int64_t* batch_ids, int i, int n_batches) {
    const int64_t lab = labels ? labels[i] : 0;
    if (!batch_ids) return (unsigned)lab;
    const int64_t img = batch_ids[i];
    const bool bad = img < 0 || img >= n_batches || lab < 0 || lab > 0xffff;
    return bad ? NMS_DEAD_KEY : (((unsigned)img << 16) | (unsigned)lab);
}

__global__ void nms_label_keys_kernel(const int64_t* __restrict__ labels, const int64_t* __restrict__ batch_ids,
                                      const int* __restrict__ ord_rank, int K, int n_batches, unsigned* keyB, int* rank_iota) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < K) {
        keyB[r] = seg_key_of(labels, batch_ids, ord_rank[r], n_batches);
        rank_iota[r] = r;
    }
}

// image key of rank r for the stable image-major pass: ids that take no part (reserved 0xffff) sort behind the last image, so
// the key needs only bit_length(n_batches) bits — one radix digit pass for up to 255 images; the rank order is copied alongside
__global__ void nms_batch_keys_kernel(const int64_t* __restrict__ labels, const int64_t* __restrict__ batch_ids, const int* __restrict__ order,
                                      int K, unsigned n_batches, unsigned* key, int* order_copy) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < K) {
        const int o = order[r];
        key[r] = min(seg_key_of(labels, batch_ids, o, (int)n_batches) >> 16, n_batches);
        order_copy[r] = o;
    }
}

// Gather + prepare boxes in position order; class offsets in FP32 as the reference wrappers compute them:
//   offsets = label.to(float) * scale ; box[:, :2] += offsets      (rnms_wrapper.py:61-64, nms_rotated_wrapper.py:84-90)
// Whole warps run to the ballot that builds the `alive` bit vector (blockDim is a multiple of 32).
__global__ void nms_gather_kernel(const float* __restrict__ boxes, int64_t stride, const int* __restrict__ ord_rank,
                                  const int* __restrict__ pos_rank, const unsigned* __restrict__ pos_label, int K,
                                  int variant, int drop_small, const float* __restrict__ class_offset, int has_labels,
                                  int batched, int n_batches, const float* __restrict__ scores, const long long* __restrict__ k_valid,
                                  float4* p0, float4* p1, float4* p2r, float4* p2c, float* raw, unsigned* alive32) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    // counted call: positions behind the device-side candidate count hold padding (it sorts last) — nothing reads their planes
    // and their alive words are already zero (reset stage); whole warps leave
    if (k_valid != nullptr && (long long)(p & ~31) >= __ldg(k_valid)) return;
    bool ok = false;
    if (p < K) {
        const int idx = ord_rank[pos_rank[p]];
        const float* b = boxes + (int64_t)idx * stride;
        float off = 0.0f;
        ok = scores[idx] > -INFINITY;                                             // padding candidates (score -inf / NaN) take no part
        if (batched && (pos_label[p] >> 16) >= (unsigned)n_batches) ok = false;    // image id out of range: takes no part
        float x[5] = { 0.0f, 0.0f, 1.0f, 1.0f, 0.0f };                            // (their boxes may be unwritten memory: not read)
        if (ok) {
#pragma unroll
            for (int k = 0; k < 5; k++) x[k] = b[k];
        }
        if (class_offset != nullptr && has_labels) {
            const unsigned key = pos_label[p];
            const float scale = batched ? class_offset[min(key >> 16, (unsigned)n_batches - 1u)] : class_offset[0];   // per-image offset scale
            off = __fmul_rn((float)(int)(batched ? (key & 0xffffu) : key), scale);
            x[0] = __fadd_rn(x[0], off);
            x[1] = __fadd_rn(x[1], off);
        }
        BoxP0 a; BoxP1 c;
        emu::prep_box_strict(x, variant, a, c);
        p0[p] = make_float4(a.cx, a.cy, a.r, a.area);
        p1[p] = make_float4(c.c, c.s, c.hw, c.hh);
#pragma unroll
        for (int k = 0; k < 5; k++) raw[(int64_t)p * 5 + k] = x[k];
        if (drop_small && fminf(x[2], x[3]) < 0.001f) ok = false;           // nms_rotated_wrapper.py:40-46
        // expanded circumradius test planes, in class-local coordinates (the class offset is removed again so that the
        // squares stay small): row form {-2X, -2Y, -2r, |X|^2 - r^2 - slack}, column form {X, Y, r, |X|^2 - r^2 - slack};
        // boxes that take no part are pushed to +inf (never pass)
        const float X = a.cx - off, Y = a.cy - off;
        const float q = X * X + Y * Y, rr = a.r * a.r;
        const float kk = ok ? (q - rr) - NMS_SLACK * (q + rr) - 1e-6f : 3.0e38f;
        // row constants stored twice: two 16-byte halves that load straight into aligned register pairs (packed f32x2 stage 1)
        p2r[2 * (size_t)p] = make_float4(-2.0f * X, -2.0f * X, -2.0f * Y, -2.0f * Y);
        p2r[2 * (size_t)p + 1] = make_float4(-2.0f * a.r, -2.0f * a.r, kk, kk);
        p2c[p] = make_float4(X, Y, a.r, kk);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && p < K) alive32[p >> 5] = bal;
}

// polygons: corners into the two 16-byte planes, bounding box into the column plane
__global__ void poly_gather_kernel(const float* __restrict__ polys, int64_t stride, const int* __restrict__ ord_rank,
                                   const int* __restrict__ pos_rank, int K, float4* p0, float4* p1, float4* p2r, float4* p2c,
                                   unsigned* alive32) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < K) {
        const float* s = polys + (int64_t)ord_rank[pos_rank[p]] * stride;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = s[k];
        p0[p] = make_float4(v[0], v[1], v[2], v[3]);
        p1[p] = make_float4(v[4], v[5], v[6], v[7]);
        const float x0 = fminf(fminf(v[0], v[2]), fminf(v[4], v[6])), x1 = fmaxf(fmaxf(v[0], v[2]), fmaxf(v[4], v[6]));
        const float y0 = fminf(fminf(v[1], v[3]), fminf(v[5], v[7])), y1 = fmaxf(fmaxf(v[1], v[3]), fmaxf(v[5], v[7]));
        p2c[p] = make_float4(x0, y0, x1, y1);
        p2r[2 * (size_t)p] = poly::quad_meta(v);                          // signed area, convexity, coordinate magnitude
    }
    const unsigned bal = __ballot_sync(0xffffffffu, p < K);
    if ((threadIdx.x & 31) == 0 && p < K) alive32[p >> 5] = bal;
}

// keep bits by position -> flags in output-slot order (rank order, or original index for R3G_NMS_ORDER_INDEX)
__global__ void nms_flags_kernel(const int* __restrict__ keep_p, const int* __restrict__ pos_rank,
                                 const int* __restrict__ ord_rank, int K, int order_index, int* __restrict__ flag) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= K) return;
    const int rank = pos_rank[p];
    flag[order_index ? ord_rank[rank] : rank] = keep_p[p];
}

__global__ void nms_emit_kernel(const int* __restrict__ flag, const int* __restrict__ pref, const int* __restrict__ ord_rank,
                                const int64_t* __restrict__ batch_ids, int n_batches, int K, int order_index, int64_t* keep_out,
                                unsigned long long* num_keep) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    int img = -1;                                                          // image of a kept candidate (batched calls)
    if (s < K && flag[s]) {
        const int idx = order_index ? s : ord_rank[s];
        keep_out[pref[s]] = (int64_t)idx;
        if (batch_ids) {
            const int64_t b = batch_ids[idx];
            if (b >= 0 && b < n_batches) img = (int)b;
        }
    }
    if (batch_ids) {
        // per-image counts (zeroed by the launcher): one atomic per image and warp — neighbours in the keep order are mostly
        // of one image, and thousands of single increments on a handful of addresses serialise in L2
        unsigned todo = __ballot_sync(0xffffffffu, img >= 0);
        while (todo) {
            const int lead = __ffs((int)todo) - 1;
            const int li = __shfl_sync(0xffffffffu, img, lead);
            const unsigned peers = __ballot_sync(0xffffffffu, img == li);
            if ((int)(threadIdx.x & 31) == lead) atomicAdd(num_keep + li, (unsigned long long)__popc(peers));
            todo &= ~peers;
        }
    } else if (s == K - 1) {
        *num_keep = (unsigned long long)(pref[s] + flag[s]);
    }
}

// ---- small-K path (K <= NMS_SMALL_K): the sorts, the structure scans and the compaction are launch latency there ----
// (two cub radix sorts = 12-18 launches of ~10 us each for a few thousand keys).  Ranks are COUNTED instead:
//   rank(i) = #{ j : (image_j, score_key_j, j) < (image_i, score_key_i, i) }           (the stable score order)
//   pos(i)  = #{ j : (segment_j, score_key_j, j) < (segment_i, score_key_i, i) }       (segment-major position order)
// O(K^2) compares spread over (K/256) x slices CTAs — a few microseconds — and exactly the permutations the stable
// radix sorts produce.  The final compaction becomes one single-CTA kernel.

template <bool BATCHED>
__global__ void __launch_bounds__(256) nms_rank_count_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels,
                                                             const int64_t* __restrict__ batch_ids, int K, int n_batches, int slice_len,
                                                             int* __restrict__ rcnt, int* __restrict__ pcnt) {
    __shared__ uint4 tile[256];                              // {key low (index), key high (score key), segment key, -}
    __shared__ unsigned img_lo, img_hi;                      // BATCHED: image range of the tile
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool iv = i < K;
    const unsigned long long ai = iv ? (((unsigned long long)score_key_desc(scores[i]) << 32) | (unsigned)i) : 0ull;
    const unsigned si = iv ? seg_key_of(labels, batch_ids, i, n_batches) : 0u;
    const unsigned mi = si >> 16;                            // BATCHED: the image (segment keys are image-major)
    const int j_begin = blockIdx.y * slice_len, j_end = min(K, j_begin + slice_len);
    int r = 0, p = 0, r1 = 0, p1 = 0, r2 = 0, p2 = 0, r3 = 0, p3 = 0;
    for (int j0 = j_begin; j0 < j_end; j0 += 256) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        if (BATCHED && threadIdx.x == 0) { img_lo = 0xffffffffu; img_hi = 0u; }
        __syncthreads();
        unsigned lo = 0xffffffffu, hi = 0u;
        if (j < j_end) {
            const unsigned sj = seg_key_of(labels, batch_ids, j, n_batches);
            tile[threadIdx.x] = make_uint4((unsigned)j, score_key_desc(scores[j]), sj, 0u);
            lo = hi = sj >> 16;
        }
        if (BATCHED) {                                       // one shared-memory atomic pair per warp
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(&img_lo, lo); atomicMax(&img_hi, hi); }
        }
        __syncthreads();
        const int n = min(256, j_end - j0);
        if (BATCHED) {
            // candidates usually arrive image by image: a tile wholly of earlier images counts in full, one wholly of later
            // images not at all — both without looking at its entries
            if (img_hi < mi) { r += n; p += n; continue; }
            if (img_lo > mi) continue;
        }
        auto one = [&](int t, int& rr, int& pp) {
            const uint4 e = tile[t];                         // one broadcast read per candidate
            const bool lt = (((unsigned long long)e.y << 32) | e.x) < ai;
            if (BATCHED) {
                const unsigned mj = e.z >> 16;
                rr += (mj < mi) | ((mj == mi) & lt);
            } else {
                rr += lt;
            }
            pp += (e.z < si) | ((e.z == si) & lt);
        };
        int t = 0;
#pragma unroll 2
        for (; t + 4 <= n; t += 4) {                         // four independent counter chains
            one(t, r, p); one(t + 1, r1, p1); one(t + 2, r2, p2); one(t + 3, r3, p3);
        }
        for (; t < n; t++) one(t, r, p);
    }
    r += r1 + r2 + r3; p += p1 + p2 + p3;
    if (iv) { atomicAdd(rcnt + i, r); atomicAdd(pcnt + i, p); }
}

__global__ void nms_rank_scatter_kernel(const int64_t* __restrict__ labels, const int64_t* __restrict__ batch_ids, int K, int n_batches,
                                        const int* __restrict__ rcnt, const int* __restrict__ pcnt,
                                        int* __restrict__ ord_rank, int* __restrict__ pos_rank, unsigned* __restrict__ pos_label) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const int r = rcnt[i], p = pcnt[i];
    ord_rank[r] = i;
    pos_rank[p] = r;
    pos_label[p] = seg_key_of(labels, batch_ids, i, n_batches);
}

// nms_flags_kernel + exclusive sum + nms_emit_kernel in one CTA
__global__ void __launch_bounds__(1024) nms_finish_small_kernel(const int* __restrict__ keep_p, const int* __restrict__ pos_rank,
                                                                const int* __restrict__ ord_rank, const int64_t* __restrict__ batch_ids,
                                                                int n_batches, int K, int order_index, int* __restrict__ flag,
                                                                int64_t* __restrict__ keep_out, unsigned long long* num_keep) {
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int carry;
    for (int p = threadIdx.x; p < K; p += 1024) {
        const int rank = pos_rank[p];
        flag[order_index ? ord_rank[rank] : rank] = keep_p[p];
    }
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int s0 = 0; s0 < K; s0 += 1024) {
        const int s = s0 + threadIdx.x;
        const int f = (s < K) ? flag[s] : 0;
        int pre, total;
        Scan(tmp).ExclusiveSum(f, pre, total);
        const int base = carry;
        if (f) {
            const int idx = order_index ? s : ord_rank[s];
            keep_out[base + pre] = (int64_t)idx;
            if (batch_ids) {
                const int64_t b = batch_ids[idx];
                if (b >= 0 && b < n_batches) atomicAdd(num_keep + b, 1ull);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) carry = base + total;
        __syncthreads();
    }
    if (!batch_ids && threadIdx.x == 0) *num_keep = (unsigned long long)carry;
}

}  // namespace r3g

using namespace r3g;

// ---- host stages shared by the rotated-box and the polygon entry points ------------------------------------------
static int nms_order_stage(NmsWs& w, const float* scores, const int64_t* labels, const int64_t* batch_ids, int Ki, cudaStream_t st,
                           bool small, int label_bits, int n_batches = 1) {
    const int tpb = 256, gK = (Ki + tpb - 1) / tpb;
    if (small) {
        // counted ranks: keyA / keyA2 double as the two count arrays
        int* rcnt = (int*)w.keyA; int* pcnt = (int*)w.keyA2;           // zeroed by nms_reset_stage
        // ~4 CTAs per SM in total; batches get single-tile slices: most (i-block, tile) pairs are of different images and
        // return at once, so the remaining work needs the finer split to spread over the SMs
        int slices = batch_ids ? (Ki + 255) / 256 : (device_sm_count() * 4 + gK - 1) / gK;
        const int max_slices = (Ki + 255) / 256;
        if (slices > max_slices) slices = max_slices;
        if (slices < 1) slices = 1;
        const int slice_len = ((Ki + slices - 1) / slices + 255) / 256 * 256;
        slices = (Ki + slice_len - 1) / slice_len;
        if (batch_ids) nms_rank_count_kernel<true><<<dim3(gK, slices), 256, 0, st>>>(scores, labels, batch_ids, Ki, n_batches, slice_len, rcnt, pcnt);
        else nms_rank_count_kernel<false><<<dim3(gK, slices), 256, 0, st>>>(scores, labels, batch_ids, Ki, n_batches, slice_len, rcnt, pcnt);
        nms_rank_scatter_kernel<<<gK, tpb, 0, st>>>(labels, batch_ids, Ki, n_batches, rcnt, pcnt, w.ord_rank, w.pos_rank, w.pos_label);
        R3G_LAUNCH_OK("nms rank kernels");
        return R3G_OK;
    }
    // 1. rank order by descending score (stable: ties keep ascending index)
    nms_keys_kernel<<<gK, tpb, 0, st>>>(scores, Ki, w.keyA, w.ord_tmp);
    size_t tb = w.cub_bytes;
    R3G_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keyA, w.keyA2, w.ord_tmp, w.ord_rank, Ki, 0, 32, st));
    if (batch_ids) {
        // multi-image batch: rank order becomes (image asc, score desc) by a stable pass over the image id
        nms_batch_keys_kernel<<<gK, tpb, 0, st>>>(labels, batch_ids, w.ord_rank, Ki, (unsigned)n_batches, w.keyA, w.ord_tmp);
        int img_bits = 1;
        while ((1 << img_bits) <= n_batches) img_bits++;
        tb = w.cub_bytes;
        R3G_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keyA, w.keyA2, w.ord_tmp, w.ord_rank, Ki, 0, img_bits, st));
    }
    // 2. position order: stable by segment key (label, or label and image) on top of the rank order
    nms_label_keys_kernel<<<gK, tpb, 0, st>>>(labels, batch_ids, w.ord_rank, Ki, n_batches, w.keyB, w.pos_tmp);
    if (labels || batch_ids) {
        tb = w.cub_bytes;
        // in a batch the rank order is already image-major: sorting on the 16 label bits alone keeps every (image, label)
        // segment contiguous (label-major), the full keys travel with the permutation
        // (one 8-bit digit pass per 8 label bits: a caller that knows its class count saves three of the four passes)
        R3G_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keyB, w.pos_label, w.pos_tmp, w.pos_rank, Ki, 0,
                                                    batch_ids ? (label_bits < 16 ? label_bits : 16) : label_bits, st));
    } else {
        R3G_CUDA_OK(cudaMemcpyAsync(w.pos_label, w.keyB, 4 * (size_t)Ki, cudaMemcpyDeviceToDevice, st));
        R3G_CUDA_OK(cudaMemcpyAsync(w.pos_rank, w.pos_tmp, 4 * (size_t)Ki, cudaMemcpyDeviceToDevice, st));
    }
    return R3G_OK;
}

// state the rounds kernel starts from: control block, keep flags and the tail of the alive vector zero (the gather kernels
// write the alive words that hold candidates)
static int nms_reset_stage(NmsWs& w, bool small, cudaStream_t st) {
    R3G_CUDA_OK(cudaMemsetAsync(w.ctrl, 0, small ? w.zero_bytes_small : w.zero_bytes, st));
    return R3G_OK;
}

static int nms_env_int(const char* name, int dflt, int lo, int hi) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    const int v = atoi(s);
    return v < lo ? lo : (v > hi ? hi : v);
}

// greedy selection in rounds: one cooperative launch (the grid barrier needs every CTA resident)
template <int GEOM>
static int nms_rounds_stage(NmsWs& w, int Ki, int variant, int inclusive, float thr, float tau, float margin, int prefilter,
                            const int64_t* k_valid, cudaStream_t st) {
    static int occ_of[64] = {0};
    int& occ = occ_of[current_device_slot()];
    if (occ == 0) {
        R3G_CUDA_OK(cudaFuncSetAttribute(rn::nms_rounds_kernel<GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rn::SMEM_BYTES));
        int o = 0;
        R3G_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, rn::nms_rounds_kernel<GEOM>, rn::THREADS, rn::SMEM_BYTES));
        if (o < 1) { set_error("nms_rounds_kernel does not fit on this device"); return R3G_ERR_CUDA; }
        occ = o;
    }
    static const int chunk_env = nms_env_int("R3G_NMS_CHUNK", rn::B_MAX, 64, rn::B_MAX) / 64 * 64;     // tuning knobs
    static const int grid_env = nms_env_int("R3G_NMS_GRID", 0, 0, 1 << 20);
    static const int div_env = nms_env_int("R3G_NMS_CHUNK_DIV", 0, -1, 64);           // 0: decided on the device from the segment lengths
    static const int work_env = nms_env_int("R3G_NMS_WORK_THR_M", 24, 0, 1 << 20);     // ... threshold, in millions of pair tests
    static const int min_env = nms_env_int("R3G_NMS_CHUNK_MIN", 512, 64, rn::B_MAX) / 64 * 64;
    rn::Args a;
    a.p0 = w.p0; a.p1 = w.p1; a.p2r = w.p2r; a.p2c = w.p2c; a.raw = w.raw; a.label = w.pos_label; a.K = Ki;
    a.k_valid = reinterpret_cast<const long long*>(k_valid);
    a.alive = w.alive; a.seg_cur = w.seg_cur; a.seg_pe = w.seg_pe; a.act = w.act; a.ent = w.ent; a.spos = w.spos; a.klist = w.klist;
    a.ownerB = w.ownerB; a.ownerD = w.ownerD; a.mask = w.mask; a.keep_p = w.keep_p;
    a.ctrl = reinterpret_cast<rn::Ctrl*>(w.ctrl); a.bar = reinterpret_cast<unsigned*>(w.ctrl + 128);
    a.dbg = reinterpret_cast<unsigned long long*>(w.ctrl + 512);
    a.B = chunk_env; a.split = nms_split_of(Ki); a.chunk_div = div_env; a.chunk_min = min_env; a.work_thr = (unsigned long long)work_env * 1000000ull;
    a.variant = variant; a.inclusive = inclusive; a.prefilter = prefilter; a.thr = thr; a.tau = tau; a.margin = margin;
    const long long cap = (long long)device_sm_count() * occ;
    long long grid = Ki / 32;
    if (grid < 64) grid = 64;
    if (grid_env > 0) grid = grid_env;
    if (grid > cap) grid = cap;
    void* params[] = { (void*)&a };
    R3G_CUDA_OK(cudaLaunchCooperativeKernel((const void*)rn::nms_rounds_kernel<GEOM>, dim3((unsigned)grid), dim3(rn::THREADS), params,
                                            rn::SMEM_BYTES, st));
    return R3G_OK;
}

static int nms_finish_stage(NmsWs& w, int Ki, const int64_t* batch_ids, int n_batches, int flags, bool small,
                            int64_t* keep_out, int64_t* num_keep_out, cudaStream_t st) {
    const int tpb = 256, gK = (Ki + tpb - 1) / tpb;
    const int order_index = (flags & R3G_NMS_ORDER_INDEX) ? 1 : 0;
    if (small && Ki <= 4096) {            // one CTA walks K / 1024 chunks: past a few thousand the three-kernel path is faster
        nms_finish_small_kernel<<<1, 1024, 0, st>>>(w.keep_p, w.pos_rank, w.ord_rank, batch_ids, n_batches, Ki, order_index, w.flag,
                                                    keep_out, (unsigned long long*)num_keep_out);
        R3G_LAUNCH_OK("nms_finish_small_kernel");
        return R3G_OK;
    }
    nms_flags_kernel<<<gK, tpb, 0, st>>>(w.keep_p, w.pos_rank, w.ord_rank, Ki, order_index, w.flag);
    // compaction in the requested order
    size_t tb = w.cub_bytes;
    R3G_CUDA_OK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.flag, w.pref, Ki, st));
    nms_emit_kernel<<<gK, tpb, 0, st>>>(w.flag, w.pref, w.ord_rank, batch_ids, n_batches, Ki, order_index, keep_out,
                                        (unsigned long long*)num_keep_out);
    R3G_LAUNCH_OK("nms_emit_kernel");
    return R3G_OK;
}

R3G_API int r3g_nms_workspace_bytes(int64_t K, size_t* bytes) {
    R3G_REQUIRE(bytes != nullptr && K >= 0, "r3g_nms_workspace_bytes: bad arguments");
    R3G_REQUIRE(K <= (1ll << 26), "r3g_nms_workspace_bytes: K too large (limit 2^26 candidates per call)");
    *bytes = carve_nms(nullptr, K > 0 ? K : 1).bytes;
    return R3G_OK;
}

R3G_API int r3g_nms_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                        int64_t K, float thr, int variant, int flags, const float* class_offset,
                        int64_t* keep_out, int64_t* num_keep_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
    return r3g_nms_batched_f32(boxes, stride, scores, labels, nullptr, 1, K, thr, variant, flags, class_offset,
                               keep_out, num_keep_out, workspace, workspace_bytes, stream);
}

R3G_API int r3g_nms_batched_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                                const int64_t* batch_ids, int n_batches,
                                int64_t K, float thr, int variant, int flags, const float* class_offset,
                                int64_t* keep_out, int64_t* num_keep_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
    return r3g_nms_batched_counted_f32(boxes, stride, scores, labels, batch_ids, n_batches, K, nullptr, thr, variant, flags, class_offset,
                                       keep_out, num_keep_out, workspace, workspace_bytes, stream);
}

R3G_API int r3g_nms_batched_counted_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                                        const int64_t* batch_ids, int n_batches,
                                        int64_t K, const int64_t* count_dev, float thr, int variant, int flags, const float* class_offset,
                                        int64_t* keep_out, int64_t* num_keep_out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    R3G_REQUIRE(K >= 0 && K <= (1ll << 26), "r3g_nms_f32: bad K (limit 2^26 candidates per call)");
    R3G_REQUIRE(n_batches >= 1 && n_batches <= 65535, "r3g_nms_batched_f32: n_batches must be in [1, 65535]");
    R3G_REQUIRE(batch_ids != nullptr || n_batches == 1, "r3g_nms_batched_f32: n_batches > 1 needs batch_ids");
    R3G_REQUIRE(count_dev == nullptr || batch_ids != nullptr, "r3g_nms_batched_counted_f32: a device-side count needs batch_ids (padding carries image id 65535)");
    R3G_REQUIRE(variant >= 1 && variant <= 3, "r3g_nms_f32: variant must be 1, 2 or 3 (got %d)", variant);
    R3G_REQUIRE(num_keep_out != nullptr, "r3g_nms_f32: null num_keep_out");
    cudaStream_t st = (cudaStream_t)stream;
    R3G_CUDA_OK(cudaMemsetAsync(num_keep_out, 0, sizeof(int64_t) * (size_t)n_batches, st));
    if (K == 0) return R3G_OK;
    R3G_REQUIRE(boxes && scores && keep_out && workspace, "r3g_nms_f32: null pointer");
    R3G_REQUIRE(stride >= 5, "r3g_nms_f32: box stride must be >= 5 floats");
    NmsWs w = carve_nms(workspace, K);
    if (workspace_bytes < w.bytes) {
        set_error("r3g_nms_f32: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    const int Ki = (int)K;
    const int tpb = 256, gK = (Ki + tpb - 1) / tpb;
    const bool small = Ki <= NMS_SMALL_K && !(flags & R3G_NMS_SORT_PATH);
    int rc = nms_reset_stage(w, small, st);
    if (rc != R3G_OK) return rc;
    int label_bits = (flags >> 8) & 63;                 // R3G_NMS_LABEL_BITS(n): labels < 2^n; 0 = unknown
    if (label_bits == 0 || label_bits > 32) label_bits = 32;
    rc = nms_order_stage(w, scores, labels, batch_ids, Ki, st, small, label_bits, n_batches);
    if (rc != R3G_OK) return rc;
    nms_gather_kernel<<<gK, tpb, 0, st>>>(boxes, stride, w.ord_rank, w.pos_rank, w.pos_label, Ki, variant,
                                          (flags & R3G_NMS_DROP_SMALL) ? 1 : 0, class_offset, labels ? 1 : 0,
                                          batch_ids ? 1 : 0, n_batches, scores, (const long long*)count_dev, w.p0, w.p1, w.p2r, w.p2c, w.raw,
                                          (unsigned*)w.alive);
    R3G_LAUNCH_OK("nms_gather_kernel");
    rc = nms_rounds_stage<rn::GEOM_BOX>(w, Ki, variant, (flags & R3G_NMS_INCLUSIVE) ? 1 : 0, thr,
                                        (flags & R3G_NMS_STRICT) ? 2e-2f : 0.0f, (variant == R3G_V1) ? 1e-3f : 5e-5f, 0, count_dev, st);
    if (rc != R3G_OK) return rc;
    return nms_finish_stage(w, Ki, batch_ids, n_batches, flags, small, keep_out, num_keep_out, st);
}

// ---- polygon NMS ------------------------------------------------------------------------------------------------------
R3G_API int r3g_poly_nms_workspace_bytes(int64_t K, size_t* bytes) {
    return r3g_nms_workspace_bytes(K, bytes);
}

R3G_API int r3g_poly_nms_f32(const float* polys, int64_t stride, const float* scores, const int64_t* labels, int64_t K, float thr,
                             int64_t* keep_out, int64_t* num_keep_out, void* workspace, size_t workspace_bytes, void* stream) {
    R3G_REQUIRE(K >= 0 && K <= (1ll << 26), "r3g_poly_nms_f32: bad K (limit 2^26 candidates per call)");
    R3G_REQUIRE(num_keep_out != nullptr, "r3g_poly_nms_f32: null num_keep_out");
    cudaStream_t st = (cudaStream_t)stream;
    R3G_CUDA_OK(cudaMemsetAsync(num_keep_out, 0, sizeof(int64_t), st));
    if (K == 0) return R3G_OK;
    R3G_REQUIRE(polys && scores && keep_out && workspace, "r3g_poly_nms_f32: null pointer");
    R3G_REQUIRE(stride >= 8, "r3g_poly_nms_f32: polygon stride must be >= 8 floats");
    NmsWs w = carve_nms(workspace, K);
    if (workspace_bytes < w.bytes) {
        set_error("r3g_poly_nms_f32: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    const int Ki = (int)K, tpb = 256, gK = (Ki + tpb - 1) / tpb;
    const bool small = Ki <= NMS_SMALL_K;
    int rc = nms_reset_stage(w, small, st);
    if (rc != R3G_OK) return rc;
    rc = nms_order_stage(w, scores, labels, nullptr, Ki, st, small, 32);
    if (rc != R3G_OK) return rc;
    poly_gather_kernel<<<gK, tpb, 0, st>>>(polys, stride, w.ord_rank, w.pos_rank, Ki, w.p0, w.p1, w.p2r, w.p2c, (unsigned*)w.alive);
    R3G_LAUNCH_OK("poly_gather_kernel");
    // below 1e-3 the FP32 noise of disjoint pairs could exceed the threshold: test every pair (no bounding-box filter)
    rc = nms_rounds_stage<rn::GEOM_QUAD>(w, Ki, 0, 0, thr, 0.0f, 0.0f, thr >= 1e-3f ? 1 : 0, nullptr, st);      // poly_nms_cuda.cu:183: IoU > thr
    if (rc != R3G_OK) return rc;
    return nms_finish_stage(w, Ki, nullptr, 1, 0, small, keep_out, num_keep_out, st);
}

// ---- padded keep records -------------------------------------------------------------------------------------------
// What the callers of NMS do next (rotate_anchor_head.py:626-673, bbox_nms_rotated.py:127-131: `dets[keep][:max_num]`, one
// Python slice per image after a host read of the keep count) as ONE kernel with fixed-size outputs: image b's first
// min(num_keep[b], max_per_img) kept candidates -> out_dets[b, o, :] = <x, y, w, h, a, score>, out_labels[b, o]; the rest of
// the (max_per_img) rows is zero; out_counts[b] = rows filled.  No host synchronisation is needed anywhere downstream.
namespace r3g {

__global__ void __launch_bounds__(256) nms_pack_kernel(const float* __restrict__ boxes, int64_t stride, const float* __restrict__ scores,
                                                       const int64_t* __restrict__ labels, const int64_t* __restrict__ keep,
                                                       const int64_t* __restrict__ num_keep, const int64_t* __restrict__ batch_ids,
                                                       int n_batches, int64_t K, int max_per_img, int drop_last,
                                                       float* __restrict__ out_dets, int64_t* __restrict__ out_labels,
                                                       int64_t* __restrict__ out_counts) {
    extern __shared__ long long starts[];                        // [n_batches + 1] exclusive prefix of num_keep
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_batches; b0 += 256) {               // blocked scan, 256 images at a time
        const int b = b0 + threadIdx.x;
        long long v = (b < n_batches) ? num_keep[b] : 0;
        typedef cub::BlockScan<long long, 256> Scan;
        __shared__ typename Scan::TempStorage tmp;
        long long pre, tot;
        Scan(tmp).ExclusiveSum(v, pre, tot);
        if (b < n_batches) starts[b] = carry + pre;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) starts[n_batches] = carry;
    __syncthreads();
    const long long total = starts[n_batches];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int64_t)n_batches * max_per_img) {                  // padding rows and counts
        const int b = (int)(t / max_per_img), o = (int)(t - (int64_t)b * max_per_img);
        long long cnt = num_keep[b] - (drop_last ? 1 : 0);
        cnt = cnt < 0 ? 0 : (cnt > max_per_img ? max_per_img : cnt);
        if (o == 0) out_counts[b] = cnt;
        if (o >= cnt) {
#pragma unroll
            for (int q = 0; q < 6; q++) out_dets[t * 6 + q] = 0.0f;
            out_labels[t] = 0;
        }
    }
    if (t < K && t < total) {                                    // kept entry t of the concatenated keep list
        const int64_t idx = keep[t];
        int b = 0;
        if (batch_ids) {
            const int64_t bb = batch_ids[idx];
            b = (bb < 0 || bb >= n_batches) ? -1 : (int)bb;
        }
        if (b >= 0) {
            const long long o = t - starts[b];
            long long cnt = num_keep[b] - (drop_last ? 1 : 0);
            cnt = cnt > max_per_img ? max_per_img : cnt;
            if (o >= 0 && o < cnt) {
                const int64_t r = (int64_t)b * max_per_img + o;
                const float* bx = boxes + idx * stride;
#pragma unroll
                for (int q = 0; q < 5; q++) out_dets[r * 6 + q] = bx[q];
                out_dets[r * 6 + 5] = scores[idx];
                out_labels[r] = labels ? labels[idx] : 0;
            }
        }
    }
}

}  // namespace r3g

R3G_API int r3g_nms_pack_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                             const int64_t* keep, const int64_t* num_keep, const int64_t* batch_ids, int n_batches, int64_t K,
                             int max_per_img, int drop_last, float* out_dets, int64_t* out_labels, int64_t* out_counts,
                             void* stream) {
    R3G_REQUIRE(K >= 0 && n_batches >= 1 && n_batches <= 8192 && max_per_img >= 1, "r3g_nms_pack_f32: bad sizes (1..8192 images)");
    R3G_REQUIRE(batch_ids != nullptr || n_batches == 1, "r3g_nms_pack_f32: n_batches > 1 needs batch_ids");
    R3G_REQUIRE(num_keep && out_dets && out_labels && out_counts, "r3g_nms_pack_f32: null pointer");
    R3G_REQUIRE(K == 0 || (boxes && scores && keep), "r3g_nms_pack_f32: null pointer");
    R3G_REQUIRE(stride >= 5, "r3g_nms_pack_f32: box stride must be >= 5 floats");
    const int64_t slots = (int64_t)n_batches * max_per_img;
    const int64_t work = slots > K ? slots : K;
    R3G_REQUIRE(work < (1ll << 31) * 256, "r3g_nms_pack_f32: problem too large");
    nms_pack_kernel<<<(unsigned)((work + 255) / 256), 256, sizeof(long long) * (size_t)(n_batches + 1), (cudaStream_t)stream>>>(
        boxes, stride, scores, labels, keep, num_keep, batch_ids, n_batches, K, max_per_img, drop_last ? 1 : 0, out_dets, out_labels,
        out_counts);
    R3G_LAUNCH_OK("nms_pack_kernel");
    return R3G_OK;
}

// ---- multiclass candidate extraction --------------------------------------------------------------------------------
// replaces the torch prologue of multiclass_nms_rotated (r3det/core/post_processing/bbox_nms_rotated.py:34-41, 98-103:
// expand / boolean-mask / nonzero, five kernels and a sync) with ONE pass: candidate k enumerates the (box, class)
// pairs with score > score_thr in row-major order — exactly the order `bboxes[valid_mask]` / `nonzero()[:, 1]` produce.
namespace r3g {

__global__ void mc_flags_kernel(const float* __restrict__ scores, int64_t n, int C, int64_t score_stride, float thr,
                                int* __restrict__ flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * C) return;
    const int64_t i = t / C;
    const int c = (int)(t - i * C);
    flag[t] = (scores[i * score_stride + c] > thr) ? 1 : 0;
}

__global__ void mc_emit_kernel(const float* __restrict__ boxes, int box_cols, const float* __restrict__ scores,
                               const float* __restrict__ factors, int64_t n, int C, int64_t score_stride,
                               const int* __restrict__ flag, const int* __restrict__ pref,
                               float* __restrict__ out_boxes, float* __restrict__ out_scores,
                               int64_t* __restrict__ out_labels, int64_t* __restrict__ out_src, int64_t* __restrict__ count) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * C) return;
    if (t == n * C - 1) *count = (int64_t)pref[t] + flag[t];
    if (!flag[t]) return;
    const int64_t i = t / C;
    const int c = (int)(t - i * C);
    const int64_t k = pref[t];
    const float* b = boxes + i * box_cols + (box_cols > 5 ? (int64_t)c * 5 : 0);
#pragma unroll
    for (int q = 0; q < 5; q++) out_boxes[k * 5 + q] = b[q];
    float s = scores[i * score_stride + c];
    if (factors) s = s * factors[i];
    out_scores[k] = s;
    out_labels[k] = c;
    out_src[k] = t;
}

}  // namespace r3g

// ---- batched, synchronisation-free candidate extraction ----------------------------------------------------------
// The same enumeration for a whole batch, with everything the batched NMS call needs produced on the device: candidates are
// compacted to the front of capacity-sized arrays (capacity = B * n * C), the tail is PADDING (score -inf: r3g_nms_batched_f32
// ignores such candidates), every candidate carries its image id, and the per-image class-offset scale of the reference's
// batched wrappers is reduced in the same pass (rule 1: max over the candidate boxes' five columns + 1, rnms_wrapper.py:61-64;
// rule 2: span of their horizontal bounding boxes + 1, nms_rotated_wrapper.py:84-90; images without candidates get 1).
namespace r3g {

__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
constexpr int MCB_MAX_IMAGES = 64;
constexpr int MCB_NONE_HI = (int)0x80808080, MCB_NONE_LO = 0x7f7f7f7f;     // memset patterns: "no candidate seen"

__global__ void __launch_bounds__(256) mcb_emit_kernel(const float* __restrict__ boxes, int box_cols, const float* __restrict__ scores,
                                                       int64_t n_img, int B, int C, int64_t score_stride, const int* __restrict__ flag,
                                                       const int* __restrict__ pref, int rule, float* __restrict__ out_boxes,
                                                       float* __restrict__ out_scores, int64_t* __restrict__ out_labels,
                                                       int64_t* __restrict__ out_batch, int64_t* __restrict__ out_src,
                                                       int64_t* __restrict__ count, int* __restrict__ ghi, int* __restrict__ glo) {
    __shared__ int shi[MCB_MAX_IMAGES], slo[MCB_MAX_IMAGES];
    const int64_t T = (int64_t)B * n_img * C;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x < MCB_MAX_IMAGES) { shi[threadIdx.x] = MCB_NONE_HI; slo[threadIdx.x] = MCB_NONE_LO; }
    __syncthreads();
    const int64_t total = (int64_t)pref[T - 1] + flag[T - 1];
    if (t == 0) *count = total;
    if (t < T) {
        if (flag[t]) {
            const int64_t i = t / C;
            const int c = (int)(t - i * C);
            const int b = (int)(i / n_img);
            const int64_t k = pref[t];
            const float* bx = boxes + i * box_cols + (box_cols > 5 ? (int64_t)c * 5 : 0);
            float v[5];
#pragma unroll
            for (int q = 0; q < 5; q++) { v[q] = bx[q]; out_boxes[k * 5 + q] = v[q]; }
            out_scores[k] = scores[i * score_stride + c];
            out_labels[k] = c;
            out_batch[k] = b;
            out_src[k] = t;
            if (rule == 1) {
                atomicMax(&shi[b], f2ord(fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), v[4])));
            } else if (rule == 2) {                       // obb2hbb of nms_rotated_wrapper.py:7-20, FP32 like the torch ops
                const float cs = cosf(v[4]), sn = sinf(v[4]);
                const float xb = fabsf(v[2] / 2 * cs) + fabsf(v[3] / 2 * sn), yb = fabsf(v[2] / 2 * sn) + fabsf(v[3] / 2 * cs);
                atomicMax(&shi[b], f2ord(fmaxf(v[0] + xb, v[1] + yb)));
                atomicMin(&slo[b], f2ord(fminf(v[0] - xb, v[1] - yb)));
            }
        }
        // padding sorts behind every candidate in both position orders (image-major when ranks are counted, label-major on the
        // radix path, whose class pass looks at the low label bits only: all ones)
        if (t >= total) { out_scores[t] = -INFINITY; out_labels[t] = 65535; out_batch[t] = 65535; out_src[t] = -1; }
    }
    __syncthreads();
    if (threadIdx.x < B) {
        if (shi[threadIdx.x] != MCB_NONE_HI) atomicMax(ghi + threadIdx.x, shi[threadIdx.x]);
        if (slo[threadIdx.x] != MCB_NONE_LO) atomicMin(glo + threadIdx.x, slo[threadIdx.x]);
    }
}

__global__ void mcb_scale_kernel(const int* __restrict__ ghi, const int* __restrict__ glo, int B, int rule, float* __restrict__ scale) {
    const int b = threadIdx.x;
    if (b >= B) return;
    float s = 1.0f;
    if (rule == 1 && ghi[b] != MCB_NONE_HI) s = ord2f(ghi[b]) + 1.0f;
    if (rule == 2 && ghi[b] != MCB_NONE_HI) s = (ord2f(ghi[b]) - ord2f(glo[b])) + 1.0f;
    scale[b] = s;
}

}  // namespace r3g

R3G_API int r3g_mc_candidates_batched_f32(const float* multi_bboxes, int box_cols, const float* multi_scores, int64_t score_stride,
                                          int64_t n_per_image, int B, int C, float score_thr, int offset_rule,
                                          float* out_boxes, float* out_scores, int64_t* out_labels, int64_t* out_batch,
                                          int64_t* out_src, int64_t* count_out, float* scale_out,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    R3G_REQUIRE(n_per_image >= 0 && C >= 0 && B >= 1 && B <= MCB_MAX_IMAGES, "r3g_mc_candidates_batched_f32: bad sizes (1..64 images)");
    const int64_t n = (int64_t)B * n_per_image, T = n * C;
    R3G_REQUIRE(T < (1ll << 31), "r3g_mc_candidates_batched_f32: too many (box, class) pairs");
    R3G_REQUIRE(count_out && scale_out && (offset_rule >= 0 && offset_rule <= 2), "r3g_mc_candidates_batched_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) {
        R3G_CUDA_OK(cudaMemsetAsync(count_out, 0, sizeof(int64_t), st));
        mcb_scale_kernel<<<1, MCB_MAX_IMAGES, 0, st>>>(nullptr, nullptr, B, 0, scale_out);
        return R3G_OK;
    }
    R3G_REQUIRE(box_cols == 5 || box_cols == 5 * C, "r3g_mc_candidates_batched_f32: boxes must be (n, 5) or (n, 5*C)");
    R3G_REQUIRE(multi_bboxes && multi_scores && out_boxes && out_scores && out_labels && out_batch && out_src && workspace,
                "r3g_mc_candidates_batched_f32: null pointer");
    size_t need = 0;
    r3g_mc_candidates_workspace_bytes(n, C, &need);
    need += 512;
    if (workspace_bytes < need) {
        set_error("r3g_mc_candidates_batched_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
        return R3G_ERR_WORKSPACE;
    }
    char* p = (char*)workspace;
    int* ghi = (int*)p; int* glo = (int*)(p + 256); p += 512;
    int* flag = (int*)p; p += align_up(4 * (size_t)T, 256);
    int* pref = (int*)p; p += align_up(4 * (size_t)T, 256);
    size_t tb = need - 512 - 2 * align_up(4 * (size_t)T, 256);
    R3G_CUDA_OK(cudaMemsetAsync(ghi, 0x80, 256, st));
    R3G_CUDA_OK(cudaMemsetAsync(glo, 0x7f, 256, st));
    const unsigned grid = (unsigned)((T + 255) / 256);
    mc_flags_kernel<<<grid, 256, 0, st>>>(multi_scores, n, C, score_stride, score_thr, flag);
    R3G_CUDA_OK(cub::DeviceScan::ExclusiveSum(p, tb, flag, pref, (int)T, st));
    mcb_emit_kernel<<<grid, 256, 0, st>>>(multi_bboxes, box_cols, multi_scores, n_per_image, B, C, score_stride, flag, pref, offset_rule,
                                          out_boxes, out_scores, out_labels, out_batch, out_src, count_out, ghi, glo);
    mcb_scale_kernel<<<1, MCB_MAX_IMAGES, 0, st>>>(ghi, glo, B, offset_rule, scale_out);
    R3G_LAUNCH_OK("batched candidate kernels");
    return R3G_OK;
}

R3G_API int r3g_mc_candidates_workspace_bytes(int64_t n, int C, size_t* bytes) {
    R3G_REQUIRE(bytes && n >= 0 && C >= 0 && n * (int64_t)C < (1ll << 31), "r3g_mc_candidates_workspace_bytes: bad arguments");
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int*)nullptr, (int*)nullptr, (int)(n * C > 0 ? n * C : 1));
    *bytes = align_up(tb, 256) + 2 * align_up(4 * (size_t)(n * C > 0 ? n * C : 1), 256);
    return R3G_OK;
}

R3G_API int r3g_mc_candidates_f32(const float* multi_bboxes, int box_cols, const float* multi_scores, int64_t score_stride,
                                  const float* score_factors, int64_t n, int C, float score_thr,
                                  float* out_boxes, float* out_scores, int64_t* out_labels, int64_t* out_src,
                                  int64_t* count_out, void* workspace, size_t workspace_bytes, void* stream) {
    R3G_REQUIRE(n >= 0 && C >= 0 && n * (int64_t)C < (1ll << 31), "r3g_mc_candidates_f32: bad sizes");
    R3G_REQUIRE(count_out != nullptr, "r3g_mc_candidates_f32: null count_out");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t T = n * C;
    if (T == 0) {
        R3G_CUDA_OK(cudaMemsetAsync(count_out, 0, sizeof(int64_t), st));
        return R3G_OK;
    }
    R3G_REQUIRE(box_cols == 5 || box_cols == 5 * C, "r3g_mc_candidates_f32: boxes must be (n, 5) or (n, 5*C)");
    R3G_REQUIRE(multi_bboxes && multi_scores && out_boxes && out_scores && out_labels && out_src && workspace,
                "r3g_mc_candidates_f32: null pointer");
    size_t need = 0;
    r3g_mc_candidates_workspace_bytes(n, C, &need);
    if (workspace_bytes < need) {
        set_error("r3g_mc_candidates_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
        return R3G_ERR_WORKSPACE;
    }
    char* p = (char*)workspace;
    int* flag = (int*)p; p += align_up(4 * (size_t)T, 256);
    int* pref = (int*)p; p += align_up(4 * (size_t)T, 256);
    size_t tb = need - 2 * align_up(4 * (size_t)T, 256);
    const unsigned grid = (unsigned)((T + 255) / 256);
    mc_flags_kernel<<<grid, 256, 0, st>>>(multi_scores, n, C, score_stride, score_thr, flag);
    R3G_CUDA_OK(cub::DeviceScan::ExclusiveSum(p, tb, flag, pref, (int)T, st));
    mc_emit_kernel<<<grid, 256, 0, st>>>(multi_bboxes, box_cols, multi_scores, score_factors, n, C, score_stride, flag, pref,
                                         out_boxes, out_scores, out_labels, out_src, count_out);
    R3G_LAUNCH_OK("mc candidate kernels");
    return R3G_OK;
}
