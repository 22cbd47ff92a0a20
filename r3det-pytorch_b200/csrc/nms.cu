// nms.cu — rotated NMS (single class and per-class segmented) entirely on the device, sm_100a.
//
// Replaces (reference, relative to /root/reference):
//   nmsr_kernel + nmsr_cuda host scan            r3det/ops/rnms/src/rcuda/rnms_kernel.cu:229-335        (v1)
//   nms_rotated_cuda_kernel + host scan          r3det/ops/nms_rotated/src/nms_rotated_cuda.cu:12-134    (v3)
//   ml_nms_rotated                               r3det/ops/ml_nms_rotated/src/nms_rotated_cuda.cu:13-137 (v2)
// The reference computes the FULL K x K bitmask (cross-class pairs included, via a coordinate offset
// trick), copies it to the host and scans it serially on the CPU.  Here:
//   1. position space p = (segment asc, score desc), segment = label or (label, image): two stable radix sorts, or —
//      for K <= 16384, where the sorts are pure launch latency — ranks COUNTED in one O(K^2) kernel;
//   2. boxes are gathered/prepared once in p-order (class offsets applied in FP32 exactly as the
//      reference's batched wrappers do, so the geometry sees the same rounded coordinates);
//   3. mask kernel: only upper-triangular 64x64 tiles INSIDE a segment are visited; persistent warps pull
//      (64 rows x 128 columns) items from a ticket (rows split over up to 8 warps when the grid has spare warps);
//      a lane owns 4 adjacent columns, the circumradius test shifts its sign bit into a 32-bit mask per 8 rows,
//      survivors are compacted into shared-memory queues and evaluated 32 at a time (separating axes, then the
//      clamped-boundary integral of geom.cuh); suppression bits are OR-ed into shared-memory words, written once;
//   4. scan kernel: one 512-thread CTA per segment, superblocks of 4 x 64 rows: warp 0 resolves the chains from a
//      shared-memory window with ffs-jumps, the other warps prefetch the next window and OR the kept rows' words
//      into the shared-memory `removed` bit-vector (no host round trip);
//   5. keep bits -> flags -> exclusive scan -> keep list in score order or index order.
// Pairs whose IoU lies within `margin` of the threshold (or that trip the degeneracy test) are decided
// by the reference's own algorithm (emu.cuh), which makes the keep set the reference's.
#include <cub/cub.cuh>
#include "common.cuh"
#include "emu.cuh"
#include "geom.cuh"
#include "poly.cuh"

namespace r3g {

constexpr int NMS_THREADS = 256;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int NMS_G = 2;                       // column blocks (of 64) per item
constexpr int NMS_TN = NMS_G * 64;             // 128 columns per item
constexpr int NMS_CPL = NMS_TN / 32;           // 4 adjacent columns per lane
constexpr int NMS_RG = 8;                      // rows per mask group (8 rows x 4 columns = 32 mask bits per lane)
constexpr int NMS_Q1CAP = 32 + NMS_RG * NMS_TN;
constexpr float NMS_SLACK = 1.0f / 262144.0f;  // 2^-18 relative slack of the expanded circumradius test

struct NmsWs {
    unsigned *keyA, *keyA2, *keyB, *keyB2;
    int *ord_rank, *ord_tmp, *pos_rank, *pos_tmp;   // ord_rank[r] = original index of rank r; pos_rank[p] = rank at position p
    unsigned* pos_label;
    BoxP0* p0; BoxP1* p1; float4* p2r; float4* p2c; float* raw; unsigned char* valid;
    int* blk_end; long long *nw, *row_base, *ng, *item_base;
    int* seg_list; int* counters;                   // counters[0] = nseg
    int *flag, *pref, *keep_p;
    void* cub_tmp; size_t cub_bytes;
    unsigned long long* mask;
    size_t bytes;
};

static size_t cub_temp_bytes(int64_t K, int64_t nblk) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr, (int)K);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, (int)K);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (long long*)nullptr, (long long*)nullptr, (int)nblk + 1);
    size_t m = a > b ? a : b;
    return align_up(m > c ? m : c, 256);
}

static NmsWs carve_nms(void* ws, int64_t K) {
    NmsWs w;
    const int64_t nblk = (K + 63) / 64;
    char* p = (char*)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes, 256); return (void*)r; };
    w.counters = (int*)take(256);
    w.keyA = (unsigned*)take(4 * K); w.keyA2 = (unsigned*)take(4 * K);
    w.keyB = (unsigned*)take(4 * K); w.keyB2 = (unsigned*)take(4 * K);
    w.ord_rank = (int*)take(4 * K); w.ord_tmp = (int*)take(4 * K);
    w.pos_rank = (int*)take(4 * K); w.pos_tmp = (int*)take(4 * K);
    w.pos_label = (unsigned*)take(4 * K);
    w.p0 = (BoxP0*)take(16 * K); w.p1 = (BoxP1*)take(16 * K);
    w.p2r = (float4*)take(32 * K); w.p2c = (float4*)take(16 * K);       // row plane: every constant twice (f32x2 operands)
    w.raw = (float*)take(20 * K); w.valid = (unsigned char*)take(K);
    w.blk_end = (int*)take(4 * nblk);
    w.nw = (long long*)take(8 * (nblk + 1)); w.row_base = (long long*)take(8 * (nblk + 1));
    w.ng = (long long*)take(8 * (nblk + 1)); w.item_base = (long long*)take(8 * (nblk + 1));
    w.seg_list = (int*)take(4 * K);
    w.flag = (int*)take(4 * K); w.pref = (int*)take(4 * K); w.keep_p = (int*)take(4 * K);
    w.cub_bytes = cub_temp_bytes(K, nblk);
    w.cub_tmp = take(w.cub_bytes);
    w.mask = (unsigned long long*)take((size_t)8 * 64 * (size_t)(nblk * (nblk + 1) / 2));
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

// segment key of rank r: the label, or (image << 16 | label) for multi-image batches (image-major: an image's
// segments are contiguous in position space)
__global__ void nms_label_keys_kernel(const int64_t* __restrict__ labels, const int64_t* __restrict__ batch_ids,
                                      const int* __restrict__ ord_rank, int K, unsigned* keyB, int* rank_iota) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < K) {
        const int idx = ord_rank[r];
        unsigned key = labels ? (unsigned)labels[idx] : 0u;
        if (batch_ids) key = (((unsigned)batch_ids[idx] & 0xffffu) << 16) | (key & 0xffffu);
        keyB[r] = key;
        rank_iota[r] = r;
    }
}

__global__ void nms_batch_keys_kernel(const int64_t* __restrict__ batch_ids, const int* __restrict__ order, int K, unsigned* key) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < K) key[r] = (unsigned)batch_ids[order[r]] & 0xffffu;
}

// Gather + prepare boxes in position order; class offsets in FP32 as the reference wrappers compute them:
//   offsets = label.to(float) * scale ; box[:, :2] += offsets      (rnms_wrapper.py:61-64, nms_rotated_wrapper.py:84-90)
__global__ void nms_gather_kernel(const float* __restrict__ boxes, int64_t stride, const int* __restrict__ ord_rank,
                                  const int* __restrict__ pos_rank, const unsigned* __restrict__ pos_label, int K,
                                  int variant, int drop_small, const float* __restrict__ class_offset, int has_labels,
                                  int batched,
                                  BoxP0* p0, BoxP1* p1, float4* p2r, float4* p2c, float* raw, unsigned char* valid) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= K) return;
    const int idx = ord_rank[pos_rank[p]];
    const float* b = boxes + (int64_t)idx * stride;
    float x[5] = { b[0], b[1], b[2], b[3], b[4] };
    float off = 0.0f;
    if (class_offset != nullptr && has_labels) {
        const unsigned key = pos_label[p];
        const float scale = batched ? class_offset[key >> 16] : class_offset[0];           // per-image offset scale
        off = __fmul_rn((float)(int)(batched ? (key & 0xffffu) : key), scale);
        x[0] = __fadd_rn(x[0], off);
        x[1] = __fadd_rn(x[1], off);
    }
    BoxP0 a; BoxP1 c;
    emu::prep_box_strict(x, variant, a, c);
    p0[p] = a; p1[p] = c;
#pragma unroll
    for (int k = 0; k < 5; k++) raw[(int64_t)p * 5 + k] = x[k];
    const bool ok = !(drop_small && fminf(x[2], x[3]) < 0.001f);      // nms_rotated_wrapper.py:40-46
    valid[p] = ok ? 1 : 0;
    // expanded circumradius test planes, in class-local coordinates (the class offset is removed again so that the
    // squares stay small): row form {-2X, -2Y, -2r, |X|^2 - r^2 - slack}, column form {X, Y, r, |X|^2 - r^2 - slack};
    // boxes that take no part are pushed to +inf (never pass)
    const float X = a.cx - off, Y = a.cy - off;
    const float q = X * X + Y * Y, rr = a.r * a.r;
    const float kk = ok ? (q - rr) - NMS_SLACK * (q + rr) - 1e-6f : 3.0e38f;
    // row constants stored twice: two 16-byte halves that load straight into aligned register pairs (packed f32x2 stage 1)
    p2r[2 * p] = make_float4(-2.0f * X, -2.0f * X, -2.0f * Y, -2.0f * Y);
    p2r[2 * p + 1] = make_float4(-2.0f * a.r, -2.0f * a.r, kk, kk);
    p2c[p] = make_float4(X, Y, a.r, kk);
}

// per 64-row block: last column block its rows can interact with (end of the class segment of its last row)
__global__ void nms_blocks_kernel(const unsigned* __restrict__ pos_label, int K, int nblk,
                                  int* blk_end, long long* nw, long long* ng) {
    int rb = blockIdx.x * blockDim.x + threadIdx.x;
    if (rb > nblk) return;
    if (rb == nblk) { nw[rb] = 0; ng[rb] = 0; return; }
    int last = min(K, rb * 64 + 64) - 1;
    unsigned L = pos_label[last];
    int lo = last, hi = K;                          // first p > last with label != L (labels sorted ascending)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (pos_label[mid] == L) lo = mid; else hi = mid;
    }
    int be = lo >> 6;
    blk_end[rb] = be;
    long long w = be - rb + 1;
    nw[rb] = w * 64;                                // words of this row block (64 rows x w words)
    ng[rb] = (w + NMS_G - 1) / NMS_G;               // items of this row block
}

__global__ void nms_segments_kernel(const unsigned* __restrict__ pos_label, int K, int* seg_list, int* counters) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= K) return;
    if (p == 0 || pos_label[p] != pos_label[p - 1]) seg_list[atomicAdd(&counters[0], 1)] = p;
}

__device__ __noinline__ float nms_emu_call(const float* b1, const float* b2, int variant) {
    float x[5] = { b1[0], b1[1], b1[2], b1[3], b1[4] };
    float y[5] = { b2[0], b2[1], b2[2], b2[3], b2[4] };
    return emu::pair(x, y, variant, MODE_IOU);
}

struct MaskArgs {
    const BoxP0* p0; const BoxP1* p1; const float4* p2r; const float4* p2c; const float* raw; const unsigned char* valid;
    const unsigned* label; unsigned long long* ticket;
    const int* blk_end; const long long* row_base; const long long* item_base;
    unsigned long long* mask;
    int K, nblk, variant, inclusive;
    float thr, tau, margin;
};

__device__ __forceinline__ float4 nldg4(const void* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// packed f32x2 arithmetic (FFMA2 / FADD2 on sm_100): low half = element 0
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__global__ void __launch_bounds__(NMS_THREADS, 3) nms_mask_kernel(const MaskArgs A) {
    // per-warp: q1 circumradius survivors -> q2 separating-axis survivors -> (q3 near-threshold / degenerate pairs);
    // q1/q2 entries are item-relative (row << 7 | col); all queues are flushed at item end (bits land in `sm`)
    __shared__ unsigned short q1_all[NMS_WARPS][NMS_Q1CAP];
    __shared__ unsigned short q2_all[NMS_WARPS][64];
    __shared__ uint2 q3_all[NMS_WARPS][96];          // (row, col) positions of pairs left to the reference restatement; persists across items
    __shared__ unsigned long long sm_all[NMS_WARPS][64 * NMS_G];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id(), lt = lanemask_lt();
    unsigned short* q1 = q1_all[warp];
    unsigned short* q2 = q2_all[warp];
    uint2* q3 = q3_all[warp];
    unsigned long long* sm = sm_all[warp];
    // When there are fewer items than warps in the grid (small K: a detection head's per-image call), an item's 64 rows
    // are split over up to 8 warps: the kernel is then bound by ONE warp's serial work (its SAT / area / restatement
    // batches), not by throughput.
    const long long base_items = A.item_base[A.nblk];
    int split = 1;
    while (split < 8 && base_items * (split * 2) <= (long long)gridDim.x * NMS_WARPS) split *= 2;
    const long long total = base_items * split;
    const int rows_per_sub = 64 / split;
    int c1 = 0, c2 = 0, c3 = 0;
    int i0 = 0, j0 = 0;

    int rb = -1, cb0 = 0, ncb = 0, r_lo = 0, r_hi = 64;      // r_lo .. r_hi: this warp's rows of the item in flight
    auto decide = [&](int il, int jl, float r) {
        const bool sup = A.inclusive ? (r >= A.thr) : (r > A.thr);
        if (sup) atomicOr(&sm[il * NMS_G + (jl >> 6)], 1ull << (jl & 63));
    };
    // Pairs decided by the reference's own algorithm.  Entries of the item in flight go to its shared-memory words;
    // entries of finished items (their words are already in global memory) are OR-ed into the global mask.
    auto drain_emu = [&](int nb) {
        __syncwarp();
        if ((int)lane < nb) {
            const uint2 e = q3[c3 - nb + lane];
            const int i = (int)e.x, j = (int)e.y;
            const float r = nms_emu_call(A.raw + (int64_t)i * 5, A.raw + (int64_t)j * 5, A.variant);
            const bool sup = A.inclusive ? (r >= A.thr) : (r > A.thr);
            if (sup) {
                const int rbi = i >> 6, cbj = j >> 6;
                if (rbi == rb && cbj >= cb0 && cbj < cb0 + ncb && (i & 63) >= r_lo && (i & 63) < r_hi) {
                    atomicOr(&sm[(i & 63) * NMS_G + (cbj - cb0)], 1ull << (j & 63));
                } else {
                    const long long base = A.row_base[rbi];
                    const int nwr = A.blk_end[rbi] - rbi + 1;
                    atomicOr(A.mask + base + (long long)(i & 63) * nwr + (cbj - rbi), 1ull << (j & 63));
                }
            }
        }
        __syncwarp();
        c3 -= nb;
    };
    auto drain_area = [&](int nb) {
        __syncwarp();
        bool emu = false;
        unsigned e = 0;
        if ((int)lane < nb) {
            e = q2[c2 - nb + lane];
            const int il = (int)(e >> 7), jl = (int)(e & 127u);
            const int i = i0 + il, j = j0 + jl;
            float4 a0 = nldg4(A.p0 + i), a1 = nldg4(A.p1 + i), b0 = nldg4(A.p0 + j), b1 = nldg4(A.p1 + j);
            BoxP0 A0 = { a0.x, a0.y, a0.z, a0.w }; BoxP1 A1 = { a1.x, a1.y, a1.z, a1.w };
            BoxP0 B0 = { b0.x, b0.y, b0.z, b0.w }; BoxP1 B1 = { b1.x, b1.y, b1.z, b1.w };
            bool risk;
            const float r = pair_overlap(A0, A1, B0, B1, A.variant, MODE_IOU, A.tau, risk);
            emu = A.tau > 0.0f && (risk || fabsf(r - A.thr) < A.margin);
            if (!emu) decide(il, jl, r);
        }
        __syncwarp();
        c2 -= nb;
        const unsigned bal = __ballot_sync(0xffffffffu, emu);
        if (bal) {
            if (emu) q3[c3 + __popc(bal & lt)] = make_uint2((unsigned)(i0 + (int)(e >> 7)), (unsigned)(j0 + (int)(e & 127u)));
            c3 += __popc(bal);
            if (c3 >= 64) drain_emu(32);
        }
    };
    auto drain_sat = [&](int nb) {
        __syncwarp();
        bool ok = false;
        unsigned e = 0;
        if ((int)lane < nb) {
            e = q1[c1 - nb + lane];
            const int i = i0 + (int)(e >> 7), j = j0 + (int)(e & 127u);
            float4 a0 = nldg4(A.p0 + i), a1 = nldg4(A.p1 + i), b0 = nldg4(A.p0 + j), b1 = nldg4(A.p1 + j);
            BoxP0 A0 = { a0.x, a0.y, a0.z, a0.w }; BoxP1 A1 = { a1.x, a1.y, a1.z, a1.w };
            BoxP0 B0 = { b0.x, b0.y, b0.z, b0.w }; BoxP1 B1 = { b1.x, b1.y, b1.z, b1.w };
            ok = pair_sat(A0, A1, B0, B1);
        }
        __syncwarp();
        c1 -= nb;
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (ok) q2[c2 + __popc(bal & lt)] = (unsigned short)e;
        c2 += __popc(bal);
        if (c2 >= 32) drain_area(32);
    };

    while (true) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(A.ticket, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const int sub = (int)(item % split);
        item /= split;
        // row block of this item: last rb with item_base[rb] <= item
        int lo = 0, hi = A.nblk;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (A.item_base[mid] <= item) lo = mid; else hi = mid;
        }
        rb = lo;
        r_lo = sub * rows_per_sub; r_hi = r_lo + rows_per_sub;
        const int g = (int)(item - A.item_base[rb]);
        const int be = A.blk_end[rb];
        cb0 = rb + g * NMS_G;                                 // first column block of the item
        ncb = min(NMS_G, be - cb0 + 1);                       // valid column blocks
        i0 = rb * 64;
        j0 = cb0 * 64;
        const int i1 = min(A.K, i0 + r_hi);
        const int j1 = min(A.K, j0 + ncb * 64);
        const int jb = j0 + (int)lane * NMS_CPL;

        for (int k = lane; k < 64 * NMS_G; k += 32) sm[k] = 0ull;
        float cx[NMS_CPL], cy[NMS_CPL], cr[NMS_CPL], ck[NMS_CPL];
        unsigned cl[NMS_CPL];
#pragma unroll
        for (int k = 0; k < NMS_CPL; k++) {
            if (jb + k < j1) {
                const float4 b = nldg4(A.p2c + jb + k);
                cx[k] = b.x; cy[k] = b.y; cr[k] = b.z; ck[k] = b.w;
                cl[k] = A.label[jb + k];
            } else {
                cx[k] = 0.0f; cy[k] = 0.0f; cr[k] = 0.0f; ck[k] = 3.0e38f; cl[k] = 0xffffffffu;
            }
        }
        unsigned long long CX[NMS_CPL / 2], CY[NMS_CPL / 2], CR[NMS_CPL / 2], CK[NMS_CPL / 2];     // two columns per register pair
#pragma unroll
        for (int h = 0; h < NMS_CPL / 2; h++) {
            CX[h] = pack2(cx[2 * h], cx[2 * h + 1]); CY[h] = pack2(cy[2 * h], cy[2 * h + 1]);
            CR[h] = pack2(cr[2 * h], cr[2 * h + 1]); CK[h] = pack2(ck[2 * h], ck[2 * h + 1]);
        }
        // interior items (one label throughout, all columns after all rows) need no per-pair label / order test
        const bool interior = (cb0 > rb) && (A.label[i0] == A.label[j1 - 1]);
        __syncwarp();

        for (int ig = i0 + r_lo; ig < i1; ig += NMS_RG) {
            const int nr = min(NMS_RG, i1 - ig);
            unsigned m = 0;
#pragma unroll
            for (int r = 0; r < NMS_RG; r++) {
                if (r < nr) {
                    // two columns per instruction (FFMA2 / FADD2): the kernel is issue-bound, and the packed form does the same
                    // operations in the same order as the scalar one
                    const ulonglong2 ra = __ldg(reinterpret_cast<const ulonglong2*>(A.p2r + 2 * (ig + r)));       // {-2X, -2X}, {-2Y, -2Y}
                    const ulonglong2 rb = __ldg(reinterpret_cast<const ulonglong2*>(A.p2r + 2 * (ig + r) + 1));   // {-2r, -2r}, {k, k}
#pragma unroll
                    for (int h = 0; h < NMS_CPL / 2; h++) {
                        unsigned long long t = add2(CK[h], rb.y);
                        t = fma2(ra.x, CX[h], t);
                        t = fma2(ra.y, CY[h], t);
                        t = fma2(rb.x, CR[h], t);
                        m = __funnelshift_l((unsigned)t, m, 1);
                        m = __funnelshift_l((unsigned)(t >> 32), m, 1);
                    }
                }
            }
            if (!interior && m != 0) {      // boundary items only (diagonal block / class boundary): label and order per pair
                unsigned allow = 0;
                for (int r = 0; r < nr; r++) {
                    const unsigned la = A.label[ig + r];
#pragma unroll
                    for (int k = 0; k < NMS_CPL; k++)
                        allow = (allow << 1) | ((cl[k] == la && jb + k > ig + r) ? 1u : 0u);
                }
                m &= allow;
            }
            const int cnt = __popc(m);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            if (tot) {
                int pos = c1 + incl - cnt;
                const int nbits = nr * NMS_CPL;
                const unsigned rowbase = (unsigned)(ig - i0);
                while (m) {
                    const int b = 31 - __clz(m);
                    m ^= 1u << b;
                    const unsigned idx = (unsigned)(nbits - 1 - b);
                    q1[pos++] = (unsigned short)(((rowbase + (idx >> 2)) << 7) | (lane * NMS_CPL + (idx & 3u)));
                }
                c1 += tot;
                while (c1 >= 32) drain_sat(32);
            }
        }
        if (c1 > 0) drain_sat(c1);
        if (c2 > 0) drain_area(c2);
        __syncwarp();
        // write the item's words: row r of block rb holds (be - rb + 1) words, word index = cb - rb
        const long long base = A.row_base[rb];
        const int nwr = be - rb + 1;
        for (int k = r_lo * NMS_G + lane; k < r_hi * NMS_G; k += 32) {
            const int il = k / NMS_G, c = k % NMS_G;
            if (c < ncb) A.mask[base + (long long)il * nwr + (cb0 - rb + c)] = sm[k];
        }
        __syncwarp();
        rb = -1;                               // the item's words are in global memory now
        while (c3 >= 32) drain_emu(32);
    }
    rb = -1;
    if (c3 > 0) drain_emu(c3);
}

struct ScanArgs {
    const unsigned long long* mask; const unsigned char* valid; const unsigned* label;
    const int* blk_end; const long long* row_base; const int* seg_list; const int* counters;
    const int* pos_rank; const int* ord_rank;
    int* keep_p;                                       // per position: 1 = kept
    int K, remv_cap;
};

// Greedy scan, one CTA per class segment, software-pipelined in SUPERBLOCKS of 4 x 64 rows so that one barrier and one
// global-memory round trip are paid per 256 rows and none of it sits on the serial chain.  For superblock S (column
// window = its own 4 blocks + the next 4):
//   warp 0      : for each of the 4 blocks: chain over the 64 diagonal words (shared memory) with ffs jumps -> kept;
//                 lanes 0..7 OR the kept rows' window words into removed[] so the next block / superblock sees them;
//   warps 1..7  : prefetch the 256 x 8 window words + valid bits of superblock S+1 into the other buffer, apply the
//                 kept rows of superblock S-1 to removed[] beyond its window (global mask words, 4 loads in flight),
//                 and write the keep flags of S-1.
// removed[] lives in shared memory and is updated with atomicOr.
constexpr int NMS_SB = 4;                       // blocks per superblock
constexpr int NMS_WIN = 2 * NMS_SB;             // window words per row
constexpr int SCAN_THREADS = 512;               // 1 resolver warp + 15 warps of prefetch / apply work

__global__ void __launch_bounds__(SCAN_THREADS) nms_scan_kernel(const ScanArgs A) {
    extern __shared__ unsigned long long remv[];          // remv_cap words
    __shared__ unsigned long long Wd[2][64 * NMS_SB][NMS_WIN];
    __shared__ unsigned long long vbits[2][NMS_SB], kept_s[2][NMS_SB];
    __shared__ long long mbase[4][NMS_SB];                // row_base / words-per-row, ring over superblocks S-1 .. S+2
    __shared__ int mnwr[4][NMS_SB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nseg = A.counters[0];
    for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
        const int ps = A.seg_list[s];
        const unsigned L = A.label[ps];
        int lo = ps, hi = A.K;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (A.label[mid] == L) lo = mid; else hi = mid;
        }
        const int pe = lo + 1;
        const int bf = ps >> 6, bl = (pe - 1) >> 6;
        const int nb = bl - bf + 1;
        const int nsb = (nb + NMS_SB - 1) / NMS_SB;
        __syncthreads();                                   // previous segment fully done with shared memory
        for (int w = tid; w < nb; w += SCAN_THREADS) remv[w] = 0ull;
        auto load_meta = [&](int S) {                      // one thread per block of superblock S
            const int q = tid & (NMS_SB - 1), b = bf + S * NMS_SB + q;
            if (b <= bl) { mbase[S & 3][q] = A.row_base[b]; mnwr[S & 3][q] = A.blk_end[b] - b + 1; }
        };
        if (tid < NMS_SB) { load_meta(0); kept_s[0][tid] = 0ull; kept_s[1][tid] = 0ull; }
        else if (tid < 2 * NMS_SB && nsb > 1) load_meta(1);
        __syncthreads();

        // window words + valid bits of superblock S into buffer `buf`.  `nt` threads (rank `u`) share the 256 x 8 words as
        // flattened (row, word) pairs so that every thread has ~9 independent loads in flight; the valid bits are ballots
        // over whole warps (callers pass warp-aligned thread ranges).
        auto prefetch = [&](int S, int buf, int u, int nt) {
            const int b0 = bf + S * NMS_SB;
            constexpr int PAIRS = 64 * NMS_SB * NMS_WIN;
            constexpr int PER = (PAIRS + (SCAN_THREADS - 32) - 1) / (SCAN_THREADS - 32);
            unsigned long long v[PER];
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const int idx = u + j * nt;
                unsigned long long x = 0ull;
                if (idx < PAIRS) {
                    const int r = idx >> 3, k = idx & (NMS_WIN - 1), q = r >> 6, t = r & 63, b = b0 + q, col = b0 + k;
                    const int p = b * 64 + t;
                    if (b <= bl && col >= b && col <= bl && p >= ps && p < pe)
                        x = A.mask[mbase[S & 3][q] + (long long)t * mnwr[S & 3][q] + (col - b)];
                }
                v[j] = x;
            }
            for (int r0 = 0; r0 < 64 * NMS_SB; r0 += nt) {             // valid bits (rows of other segments / dropped boxes = 0)
                const int r = r0 + u;
                bool ok = false;
                if (r < 64 * NMS_SB) {
                    const int p = (b0 + (r >> 6)) * 64 + (r & 63);
                    ok = (b0 + (r >> 6) <= bl) && p >= ps && p < pe && A.valid[p];
                }
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (r < 64 * NMS_SB && (r & 31) == 0) reinterpret_cast<unsigned*>(&vbits[buf][r >> 6])[(r >> 5) & 1] = bal;
            }
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const int idx = u + j * nt;
                if (idx < PAIRS) Wd[buf][idx >> 3][idx & (NMS_WIN - 1)] = v[j];
            }
        };
        prefetch(0, 0, tid, SCAN_THREADS);

        for (int S = 0; S < nsb; S++) {
            const int buf = S & 1;
            const int b0 = bf + S * NMS_SB;
            __syncthreads();
            if (warp == 0) {
                for (int q = 0; q < NMS_SB; q++) {
                    const int b = b0 + q;
                    if (b > bl) break;
                    const unsigned long long vb = vbits[buf][q];
                    unsigned long long cur = remv[b - bf], kept = 0ull, acc = 0ull;
                    unsigned long long avail = vb & ~cur;
                    while (avail) {                           // warp-uniform; one kept row per trip
                        const int t = __ffsll((long long)avail) - 1;
                        kept |= 1ull << t;
                        cur |= Wd[buf][q * 64 + t][q];
                        if (lane < NMS_WIN) acc |= Wd[buf][q * 64 + t][lane];
                        const unsigned long long above = (t == 63) ? 0ull : (~0ull << (t + 1));
                        avail = vb & ~cur & above;
                    }
                    if (lane < NMS_WIN && lane > q && b0 + lane <= bl && acc) atomicOr(&remv[b0 + lane - bf], acc);
                    if (lane == 0) kept_s[buf][q] = kept;
                    __syncwarp();
                }
            } else {
                const int u = tid - 32;                        // 0 .. SCAN_THREADS - 33
                const int nw_threads = SCAN_THREADS - 32;
                if (S + 1 < nsb) prefetch(S + 1, buf ^ 1, u, nw_threads);
                if (u < NMS_SB && S + 2 < nsb) load_meta(S + 2);
                if (S > 0) {
                    const int pb0 = b0 - NMS_SB;                 // previous superblock
                    for (int r = u; r < 64 * NMS_SB; r += nw_threads) {      // its keep bits, by position (store only)
                        const int q = r >> 6, t = r & 63, p = (pb0 + q) * 64 + t;
                        if (pb0 + q <= bl && p >= ps && p < pe) A.keep_p[p] = (int)((kept_s[buf ^ 1][q] >> t) & 1ull);
                    }
                    // its kept rows -> removed[] beyond its window (columns pb0 + NMS_WIN .. bl): the first four kept rows
                    // of each of the four blocks are fetched together (16 independent loads), the rest in a tail loop
                    const int first = pb0 + NMS_WIN;
                    unsigned long long kq[NMS_SB];
#pragma unroll
                    for (int q = 0; q < NMS_SB; q++) kq[q] = (pb0 + q <= bl) ? kept_s[buf ^ 1][q] : 0ull;
                    if (kq[0] | kq[1] | kq[2] | kq[3]) {
                        for (int col = first + u; col <= bl; col += nw_threads) {
                            unsigned long long kk[NMS_SB], acc = 0ull;
#pragma unroll
                            for (int q = 0; q < NMS_SB; q++) kk[q] = kq[q];
                            while (kk[0] | kk[1] | kk[2] | kk[3]) {                 // 16 independent loads per trip
                                unsigned long long v[NMS_SB * 4];
#pragma unroll
                                for (int q = 0; q < NMS_SB; q++) {
                                    const unsigned long long* cp = A.mask + mbase[(S - 1) & 3][q] + (col - (pb0 + q));
                                    const int nwr = mnwr[(S - 1) & 3][q];
#pragma unroll
                                    for (int j = 0; j < 4; j++) {
                                        const int t = kk[q] ? __ffsll((long long)kk[q]) - 1 : -1;
                                        kk[q] &= kk[q] - 1;
                                        v[q * 4 + j] = (t >= 0) ? cp[(long long)t * nwr] : 0ull;
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < NMS_SB * 4; j++) acc |= v[j];
                            }
                            if (acc) atomicOr(&remv[col - bf], acc);
                        }
                    }
                }
            }
        }
        __syncthreads();
        {   // keep bits of the segment's last superblock
            const int S = nsb - 1, pb0 = bf + S * NMS_SB;
            for (int r = tid; r < 64 * NMS_SB; r += SCAN_THREADS) {
                const int q = r >> 6, t = r & 63, p = (pb0 + q) * 64 + t;
                if (pb0 + q <= bl && p >= ps && p < pe) A.keep_p[p] = (int)((kept_s[S & 1][q] >> t) & 1ull);
            }
        }
    }
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
                                const int64_t* __restrict__ batch_ids, int K, int order_index, int64_t* keep_out,
                                unsigned long long* num_keep) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= K) return;
    if (flag[s]) {
        const int idx = order_index ? s : ord_rank[s];
        keep_out[pref[s]] = (int64_t)idx;
        if (batch_ids) atomicAdd(num_keep + batch_ids[idx], 1ull);       // per-image counts (zeroed by the launcher)
    }
    if (!batch_ids && s == K - 1) *num_keep = (unsigned long long)(pref[s] + flag[s]);
}

// ---- small-K path (K <= NMS_SMALL_K): the sorts, the structure scans and the compaction are launch latency there ----
// (two cub radix sorts = 12-18 launches of ~10 us each for a few thousand keys).  Ranks are COUNTED instead:
//   rank(i) = #{ j : (image_j, score_key_j, j) < (image_i, score_key_i, i) }           (the stable score order)
//   pos(i)  = #{ j : (segment_j, score_key_j, j) < (segment_i, score_key_i, i) }       (segment-major position order)
// O(K^2) compares spread over (K/256) x slices CTAs — a few microseconds — and exactly the permutations the stable
// radix sorts produce.  The block structure and the final compaction each become one single-CTA kernel.
constexpr int NMS_SMALL_K = 16384;

__device__ __forceinline__ unsigned seg_key_of(const int64_t* labels, const int64_t* batch_ids, int i) {
    unsigned key = labels ? (unsigned)labels[i] : 0u;
    if (batch_ids) key = (((unsigned)batch_ids[i] & 0xffffu) << 16) | (key & 0xffffu);
    return key;
}

template <bool BATCHED>
__global__ void __launch_bounds__(256) nms_rank_count_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels,
                                                             const int64_t* __restrict__ batch_ids, int K, int slice_len,
                                                             int* __restrict__ rcnt, int* __restrict__ pcnt) {
    __shared__ uint4 tile[256];                              // {key low (index), key high (score key), segment key, -}
    __shared__ unsigned img_lo, img_hi;                      // BATCHED: image range of the tile
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool iv = i < K;
    const unsigned long long ai = iv ? (((unsigned long long)score_key_desc(scores[i]) << 32) | (unsigned)i) : 0ull;
    const unsigned si = iv ? seg_key_of(labels, batch_ids, i) : 0u;
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
            const unsigned sj = seg_key_of(labels, batch_ids, j);
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

__global__ void nms_rank_scatter_kernel(const int64_t* __restrict__ labels, const int64_t* __restrict__ batch_ids, int K,
                                        const int* __restrict__ rcnt, const int* __restrict__ pcnt,
                                        int* __restrict__ ord_rank, int* __restrict__ pos_rank, unsigned* __restrict__ pos_label) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const int r = rcnt[i], p = pcnt[i];
    ord_rank[r] = i;
    pos_rank[p] = r;
    pos_label[p] = seg_key_of(labels, batch_ids, i);
}

// nms_blocks_kernel + the two exclusive sums + nms_segments_kernel in one CTA (nblk + 1 <= 1024)
__global__ void __launch_bounds__(1024) nms_structure_small_kernel(const unsigned* __restrict__ pos_label, int K, int nblk,
                                                                   int* blk_end, long long* row_base, long long* item_base,
                                                                   int* seg_list, int* counters) {
    typedef cub::BlockScan<long long, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int rb = threadIdx.x;
    long long nw = 0, ng = 0;
    if (rb < nblk) {
        const int last = min(K, rb * 64 + 64) - 1;
        const unsigned L = pos_label[last];
        int lo = last, hi = K;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pos_label[mid] == L) lo = mid; else hi = mid;
        }
        const int be = lo >> 6;
        blk_end[rb] = be;
        const long long w = be - rb + 1;
        nw = w * 64;
        ng = (w + NMS_G - 1) / NMS_G;
    }
    long long a, b;
    Scan(tmp).ExclusiveSum(nw, a);
    __syncthreads();
    Scan(tmp).ExclusiveSum(ng, b);
    if (rb <= nblk) { row_base[rb] = a; item_base[rb] = b; }
    for (int p = threadIdx.x; p < K; p += 1024)
        if (p == 0 || pos_label[p] != pos_label[p - 1]) seg_list[atomicAdd(&counters[0], 1)] = p;
}

// nms_flags_kernel + exclusive sum + nms_emit_kernel in one CTA
__global__ void __launch_bounds__(1024) nms_finish_small_kernel(const int* __restrict__ keep_p, const int* __restrict__ pos_rank,
                                                                const int* __restrict__ ord_rank, const int64_t* __restrict__ batch_ids,
                                                                int K, int order_index, int* __restrict__ flag,
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
            if (batch_ids) atomicAdd(num_keep + batch_ids[idx], 1ull);
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
                           bool small) {
    const int tpb = 256, gK = (Ki + tpb - 1) / tpb;
    if (small) {
        // counted ranks: keyA / keyA2 double as the two count arrays
        int* rcnt = (int*)w.keyA; int* pcnt = (int*)w.keyA2;
        R3G_CUDA_OK(cudaMemsetAsync(rcnt, 0, 4 * (size_t)Ki, st));
        R3G_CUDA_OK(cudaMemsetAsync(pcnt, 0, 4 * (size_t)Ki, st));
        // ~4 CTAs per SM in total; batches get single-tile slices: most (i-block, tile) pairs are of different images and
        // return at once, so the remaining work needs the finer split to spread over the SMs
        int slices = batch_ids ? (Ki + 255) / 256 : (device_sm_count() * 4 + gK - 1) / gK;
        const int max_slices = (Ki + 255) / 256;
        if (slices > max_slices) slices = max_slices;
        if (slices < 1) slices = 1;
        const int slice_len = ((Ki + slices - 1) / slices + 255) / 256 * 256;
        slices = (Ki + slice_len - 1) / slice_len;
        if (batch_ids) nms_rank_count_kernel<true><<<dim3(gK, slices), 256, 0, st>>>(scores, labels, batch_ids, Ki, slice_len, rcnt, pcnt);
        else nms_rank_count_kernel<false><<<dim3(gK, slices), 256, 0, st>>>(scores, labels, batch_ids, Ki, slice_len, rcnt, pcnt);
        nms_rank_scatter_kernel<<<gK, tpb, 0, st>>>(labels, batch_ids, Ki, rcnt, pcnt, w.ord_rank, w.pos_rank, w.pos_label);
        R3G_LAUNCH_OK("nms rank kernels");
        return R3G_OK;
    }
    // 1. rank order by descending score (stable: ties keep ascending index)
    nms_keys_kernel<<<gK, tpb, 0, st>>>(scores, Ki, w.keyA, w.ord_tmp);
    size_t tb = w.cub_bytes;
    R3G_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keyA, w.keyA2, w.ord_tmp, w.ord_rank, Ki, 0, 32, st));
    if (batch_ids) {
        // multi-image batch: rank order becomes (image asc, score desc) by a stable 16-bit pass over the image id
        nms_batch_keys_kernel<<<gK, tpb, 0, st>>>(batch_ids, w.ord_rank, Ki, w.keyA);
        R3G_CUDA_OK(cudaMemcpyAsync(w.ord_tmp, w.ord_rank, 4 * (size_t)Ki, cudaMemcpyDeviceToDevice, st));
        tb = w.cub_bytes;
        R3G_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keyA, w.keyA2, w.ord_tmp, w.ord_rank, Ki, 0, 16, st));
    }
    // 2. position order: stable by segment key (label, or label and image) on top of the rank order
    nms_label_keys_kernel<<<gK, tpb, 0, st>>>(labels, batch_ids, w.ord_rank, Ki, w.keyB, w.pos_tmp);
    if (labels || batch_ids) {
        tb = w.cub_bytes;
        // in a batch the rank order is already image-major: sorting on the 16 label bits alone keeps every (image, label)
        // segment contiguous (label-major), the full keys travel with the permutation
        R3G_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keyB, w.pos_label, w.pos_tmp, w.pos_rank, Ki, 0,
                                                    batch_ids ? 16 : 32, st));
    } else {
        R3G_CUDA_OK(cudaMemcpyAsync(w.pos_label, w.keyB, 4 * (size_t)Ki, cudaMemcpyDeviceToDevice, st));
        R3G_CUDA_OK(cudaMemcpyAsync(w.pos_rank, w.pos_tmp, 4 * (size_t)Ki, cudaMemcpyDeviceToDevice, st));
    }
    return R3G_OK;
}

static int nms_structure_stage(NmsWs& w, int Ki, int nblk, cudaStream_t st, bool small) {
    const int tpb = 256, gK = (Ki + tpb - 1) / tpb;
    size_t tb = 0;
    if (small) {
        nms_structure_small_kernel<<<1, 1024, 0, st>>>(w.pos_label, Ki, nblk, w.blk_end, w.row_base, w.item_base, w.seg_list, w.counters);
        R3G_LAUNCH_OK("nms_structure_small_kernel");
        return R3G_OK;
    }
    // 4. block / segment structure
    nms_blocks_kernel<<<(nblk + 1 + tpb - 1) / tpb, tpb, 0, st>>>(w.pos_label, Ki, nblk, w.blk_end, w.nw, w.ng);
    tb = w.cub_bytes;
    R3G_CUDA_OK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.nw, w.row_base, nblk + 1, st));
    tb = w.cub_bytes;
    R3G_CUDA_OK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.ng, w.item_base, nblk + 1, st));
    nms_segments_kernel<<<gK, tpb, 0, st>>>(w.pos_label, Ki, w.seg_list, w.counters);
    R3G_LAUNCH_OK("nms structure kernels");

    return R3G_OK;
}

static int nms_finish_stage(NmsWs& w, int Ki, int nblk, const int64_t* labels, const int64_t* batch_ids, int flags, int64_t K, bool small,
                            int64_t* keep_out, int64_t* num_keep_out, cudaStream_t st) {
    const int tpb = 256, gK = (Ki + tpb - 1) / tpb;
    size_t tb = 0;
    // 6. greedy scan per class segment
    ScanArgs sa;
    sa.mask = w.mask; sa.valid = w.valid; sa.label = w.pos_label; sa.blk_end = w.blk_end; sa.row_base = w.row_base;
    sa.seg_list = w.seg_list; sa.counters = w.counters; sa.pos_rank = w.pos_rank; sa.ord_rank = w.ord_rank;
    sa.keep_p = w.keep_p; sa.K = Ki;
    const int order_index = (flags & R3G_NMS_ORDER_INDEX) ? 1 : 0;
    sa.remv_cap = nblk;
    size_t smem = (size_t)nblk * 8;
    if (smem > 190 * 1024) {
        set_error("r3g_nms_f32: K=%lld exceeds the single-image limit of this build (%d boxes)", (long long)K, 190 * 1024 / 8 * 64);
        return R3G_ERR_ARG;
    }
    static bool smem_set_of[64] = {false};                  // per device
    bool& smem_set = smem_set_of[current_device_slot()];
    if (!smem_set) {      // static (window buffers) + dynamic (removed[]) shared memory may exceed the 48 KB default
        R3G_CUDA_OK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024));
        smem_set = true;
    }
    int sgrid = device_sm_count() * 2;
    if (!labels && !batch_ids) sgrid = 1;
    nms_scan_kernel<<<sgrid, SCAN_THREADS, smem, st>>>(sa);
    R3G_LAUNCH_OK("nms_scan_kernel");
    if (small && Ki <= 4096) {            // one CTA walks K / 1024 chunks: past a few thousand the three-kernel path is faster
        nms_finish_small_kernel<<<1, 1024, 0, st>>>(w.keep_p, w.pos_rank, w.ord_rank, batch_ids, Ki, order_index, w.flag, keep_out,
                                                    (unsigned long long*)num_keep_out);
        R3G_LAUNCH_OK("nms_finish_small_kernel");
        return R3G_OK;
    }
    nms_flags_kernel<<<gK, tpb, 0, st>>>(w.keep_p, w.pos_rank, w.ord_rank, Ki, order_index, w.flag);

    // 7. compaction in the requested order
    tb = w.cub_bytes;
    R3G_CUDA_OK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.flag, w.pref, Ki, st));
    nms_emit_kernel<<<gK, tpb, 0, st>>>(w.flag, w.pref, w.ord_rank, batch_ids, Ki, order_index, keep_out,
                                        (unsigned long long*)num_keep_out);
    R3G_LAUNCH_OK("nms_emit_kernel");
    return R3G_OK;
}

R3G_API int r3g_nms_workspace_bytes(int64_t K, size_t* bytes) {
    R3G_REQUIRE(bytes != nullptr && K >= 0, "r3g_nms_workspace_bytes: bad arguments");
    R3G_REQUIRE(K < (1ll << 30), "r3g_nms_workspace_bytes: K too large");
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
    R3G_REQUIRE(K >= 0 && K < (1ll << 30), "r3g_nms_f32: bad K");
    R3G_REQUIRE(n_batches >= 1 && n_batches <= 65535, "r3g_nms_batched_f32: n_batches must be in [1, 65535]");
    R3G_REQUIRE(batch_ids != nullptr || n_batches == 1, "r3g_nms_batched_f32: n_batches > 1 needs batch_ids");
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
    const int nblk = (Ki + 63) / 64;
    const int tpb = 256, gK = (Ki + tpb - 1) / tpb;
    R3G_CUDA_OK(cudaMemsetAsync(w.counters, 0, 256, st));

    const bool small = Ki <= NMS_SMALL_K && !(flags & R3G_NMS_SORT_PATH);
    int rc = nms_order_stage(w, scores, labels, batch_ids, Ki, st, small);
    if (rc != R3G_OK) return rc;
    // 3. gather + prepare
    nms_gather_kernel<<<gK, tpb, 0, st>>>(boxes, stride, w.ord_rank, w.pos_rank, w.pos_label, Ki, variant,
                                          (flags & R3G_NMS_DROP_SMALL) ? 1 : 0, class_offset, labels ? 1 : 0,
                                          batch_ids ? 1 : 0, w.p0, w.p1, w.p2r, w.p2c, w.raw, w.valid);
    rc = nms_structure_stage(w, Ki, nblk, st, small);
    if (rc != R3G_OK) return rc;

    // 5. suppression bitmask
    MaskArgs ma;
    ma.p0 = w.p0; ma.p1 = w.p1; ma.p2r = w.p2r; ma.p2c = w.p2c; ma.raw = w.raw; ma.valid = w.valid; ma.label = w.pos_label;
    ma.ticket = (unsigned long long*)(w.counters + 8);
    ma.blk_end = w.blk_end; ma.row_base = w.row_base; ma.item_base = w.item_base; ma.mask = w.mask;
    ma.K = Ki; ma.nblk = nblk; ma.variant = variant; ma.inclusive = (flags & R3G_NMS_INCLUSIVE) ? 1 : 0;
    ma.thr = thr;
    ma.tau = (flags & R3G_NMS_STRICT) ? 2e-2f : 0.0f;
    ma.margin = (variant == R3G_V1) ? 1e-3f : 5e-5f;
    static int occ_of[64] = {0};
    int& occ = occ_of[current_device_slot()];
    if (occ == 0) {
        R3G_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nms_mask_kernel, NMS_THREADS, 0));
        if (occ < 1) occ = 1;
    }
    // persistent warps pulling items from a ticket counter (zeroed with the other counters above);
    // upper bound on items: triangular number of (row block, group) pairs
    long long max_items = (long long)nblk * ((nblk + NMS_G - 1) / NMS_G + 1);
    long long grid = (max_items + NMS_WARPS - 1) / NMS_WARPS;
    const long long cap = (long long)device_sm_count() * occ;
    if (grid > cap) grid = cap;
    nms_mask_kernel<<<(unsigned)grid, NMS_THREADS, 0, st>>>(ma);
    R3G_LAUNCH_OK("nms_mask_kernel");

    return nms_finish_stage(w, Ki, nblk, labels, batch_ids, flags, K, small, keep_out, num_keep_out, st);
}

// ---- polygon NMS ------------------------------------------------------------------------------------------------------
struct PolyWs { float* quad; float4* aabb; size_t bytes; };
static PolyWs carve_poly(void* ws, size_t nms_bytes, int64_t K) {
    PolyWs w;
    char* p = (char*)ws;
    size_t off = align_up(nms_bytes, 256);
    w.quad = (float*)(p + off); off += align_up(32 * (size_t)K, 256);
    w.aabb = (float4*)(p + off); off += align_up(16 * (size_t)K, 256);
    w.bytes = off;
    return w;
}

R3G_API int r3g_poly_nms_workspace_bytes(int64_t K, size_t* bytes) {
    R3G_REQUIRE(bytes != nullptr && K >= 0 && K < (1ll << 30), "r3g_poly_nms_workspace_bytes: bad arguments");
    const int64_t k = K > 0 ? K : 1;
    *bytes = carve_poly(nullptr, carve_nms(nullptr, k).bytes, k).bytes;
    return R3G_OK;
}

R3G_API int r3g_poly_nms_f32(const float* polys, int64_t stride, const float* scores, const int64_t* labels, int64_t K, float thr,
                             int64_t* keep_out, int64_t* num_keep_out, void* workspace, size_t workspace_bytes, void* stream) {
    R3G_REQUIRE(K >= 0 && K < (1ll << 30), "r3g_poly_nms_f32: bad K");
    R3G_REQUIRE(num_keep_out != nullptr, "r3g_poly_nms_f32: null num_keep_out");
    cudaStream_t st = (cudaStream_t)stream;
    R3G_CUDA_OK(cudaMemsetAsync(num_keep_out, 0, sizeof(int64_t), st));
    if (K == 0) return R3G_OK;
    R3G_REQUIRE(polys && scores && keep_out && workspace, "r3g_poly_nms_f32: null pointer");
    R3G_REQUIRE(stride >= 8, "r3g_poly_nms_f32: polygon stride must be >= 8 floats");
    NmsWs w = carve_nms(workspace, K);
    PolyWs pw = carve_poly(workspace, w.bytes, K);
    if (workspace_bytes < pw.bytes) {
        set_error("r3g_poly_nms_f32: workspace too small (%zu < %zu)", workspace_bytes, pw.bytes);
        return R3G_ERR_WORKSPACE;
    }
    const int Ki = (int)K, nblk = (Ki + 63) / 64, tpb = 256, gK = (Ki + tpb - 1) / tpb;
    R3G_REQUIRE((size_t)nblk * 8 <= 190 * 1024, "r3g_poly_nms_f32: K exceeds the single-call limit of this build");
    R3G_CUDA_OK(cudaMemsetAsync(w.counters, 0, 256, st));
    const bool small = Ki <= NMS_SMALL_K;
    int rc = nms_order_stage(w, scores, labels, nullptr, Ki, st, small);
    if (rc != R3G_OK) return rc;
    poly::gather_kernel<<<gK, tpb, 0, st>>>(polys, stride, w.ord_rank, w.pos_rank, Ki, pw.quad, pw.aabb, w.valid);
    rc = nms_structure_stage(w, Ki, nblk, st, small);
    if (rc != R3G_OK) return rc;
    poly::PolyMaskArgs ma;
    ma.quad = pw.quad; ma.aabb = pw.aabb; ma.label = w.pos_label; ma.blk_end = w.blk_end; ma.row_base = w.row_base; ma.mask = w.mask;
    ma.ticket = (unsigned long long*)(w.counters + 8);
    ma.K = Ki; ma.nblk = nblk; ma.thr = thr;
    ma.prefilter = thr >= 1e-3f ? 1 : 0;      // below that, FP32 noise of disjoint pairs could exceed the threshold: test every pair
    const long long tiles = (long long)nblk * (nblk + 1) / 2;
    long long grid = (tiles + 7) / 8;
    const long long cap = (long long)device_sm_count() * 4;
    if (grid > cap) grid = cap;
    poly::mask_kernel<<<(unsigned)grid, 256, 0, st>>>(ma);
    R3G_LAUNCH_OK("poly mask kernel");
    return nms_finish_stage(w, Ki, nblk, labels, nullptr, 0, K, small, keep_out, num_keep_out, st);
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
