// iou.cu — pairwise rotated IoU / IoF for RBboxOverlaps2D_v1/v2/v3 on sm_100a.
//
// Replaces (reference, relative to /root/reference):
//   mat_iou_iof_kernel / vec_iou_iof_kernel       r3det/ops/rbbox_geo/src/rbbox_geo_kernel.cu:230-309
//   box_iou_rotated_cuda_kernel                   r3det/ops/box_iou_rotated/src/box_iou_rotated_cuda.cu:13-63
//
// Design (B200-first):
//   1. prep kernels: per-box trig/extent/radius ONCE (O(M+N)) into float4 planes in the workspace.
//   2. matrix kernel: persistent warps pull (64 rows x 128 cols) items from an atomic counter.
//      stage 1 — a lane owns 4 adjacent columns (one 16-byte store per row).  The circumradius test is evaluated in
//                its expanded form  |c|^2 - r^2 + |a|^2 - ra^2 - 2(a.c + ra r) - slack < 0 : three FFMA + one FADD per
//                pair on per-row / per-column constants, and the SIGN BIT is shifted straight into a 32-bit mask
//                (8 rows x 4 columns per lane) — no predicates, no branches.  Each row's zeros leave as one coalesced
//                512-byte streaming store per warp.  Every 8 rows the masks are compacted (warp scan of popc) into a
//                shared-memory queue of 16-bit (row, col) entries.
//      stage 2 — whenever >= 32 entries are queued the warp runs the separating-axis test at FULL lane occupancy;
//      stage 3 — the truly overlapping pairs (5 % on detection workloads) are queued again and evaluated 32 at a time
//                with the clamped-boundary area integral (geom.cuh) + variant epilogue, 4-byte store over the zero
//                already in L2;
//      stage 4 — strict mode: flagged degenerate pairs are queued a third time and re-evaluated with the reference's
//                own point-set algorithm (emu.cuh), again 32 at a time.
//      The reference runs its whole point-set algorithm in every thread of a warp as soon as one lane overlaps
//      (~80 % of warps at 5 % overlap) and recomputes sinf/cosf per pair.
//   Bound: HBM store (4 B/pair) on detection-like inputs, FP32 issue on dense-overlap inputs.
#include "common.cuh"
#include "emu.cuh"
#include "geom.cuh"

namespace r3g {

#ifndef R3G_IOU_THREADS
#define R3G_IOU_THREADS 256
#endif
#ifndef R3G_IOU_MINB
#define R3G_IOU_MINB 3
#endif
#ifndef R3G_IOU_TM
#define R3G_IOU_TM 64
#endif
#ifndef R3G_IOU_BATCHLOAD
#define R3G_IOU_BATCHLOAD 1        // stage 1: the 8 row records of a group are loaded together, ahead of the first use
#endif
#ifndef R3G_IOU_FASTGROUP
#define R3G_IOU_FASTGROUP 1        // stage 1, matrix mode: straight-line path for full 8-row groups of full tiles
#endif
#ifndef R3G_IOU_PREFETCH_BEST
#define R3G_IOU_PREFETCH_BEST 1    // assigner sweep: read the two running maxima of a pair before its area is computed
#endif
#ifndef R3G_IOU_TABCOMPACT
#define R3G_IOU_TABCOMPACT 1       // compaction loop: leading-zero count + offset table instead of per-entry index arithmetic
#endif
#ifndef R3G_IOU_UNROLLCOMPACT
#define R3G_IOU_UNROLLCOMPACT 0    // compaction variant: 32 predicated stores per lane instead of the find-leading-one loop (measured: +0.7 %)
#endif
#ifndef R3G_IOU_ZERO_CS
#define R3G_IOU_ZERO_CS 1          // zero rows leave as streaming (evict-first) stores
#endif
constexpr int IOU_THREADS = R3G_IOU_THREADS;
constexpr int IOU_WARPS = IOU_THREADS / 32;
constexpr int IOU_CPL = 4;                  // columns per lane
constexpr int IOU_TN = 32 * IOU_CPL;        // 128 columns per warp item
constexpr int IOU_TM = R3G_IOU_TM;          // rows per warp item
constexpr int IOU_RG = 8;                   // rows per mask group (8 rows x 4 columns = 32 mask bits per lane)
constexpr int IOU_Q1CAP = 32 + IOU_RG * IOU_TN;   // worst case: 31 queued + one full group of survivors
constexpr float IOU_SLACK = 1.0f / 262144.0f;     // 2^-18 relative slack of the expanded circle test (conservative)

// Row plane for the expanded circumradius test, relative to the origin (ox, oy):
//   {-2(ax-ox), -2(ay-oy), -2 ra, |a-o|^2 - ra^2 - slack_a}
// RowP2D holds every constant TWICE: its two 16-byte halves load straight into aligned register pairs for the packed f32x2
// form of stage 1 (FFMA2 / FADD2: two columns per issue slot at the same FMA-pipe time).  Measured: the packed form wins
// in the assigner sweeps (0.43 -> 0.38 ms), the scalar form in the matrix sweep (whose stage 1 also issues the row's
// streaming store: 0.354 vs 0.370 ms), so each mode keeps its own.
struct __attribute__((aligned(16))) RowP2 { float mx, my, mr, k; };
struct __attribute__((aligned(32))) RowP2D { float mx, mx2, my, my2, mr, mr2, k, k2; };

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

__global__ void prep_boxes_kernel(const float* __restrict__ boxes, int64_t n, int64_t stride, int variant,
                                  const float* __restrict__ origin_box,
                                  BoxP0* __restrict__ p0, BoxP1* __restrict__ p1, RowP2* __restrict__ p2, RowP2D* __restrict__ p2d) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + i * stride;
    float raw[5] = { b[0], b[1], b[2], b[3], b[4] };
    BoxP0 a; BoxP1 c;
    emu::prep_box_strict(raw, variant, a, c);
    p0[i] = a;
    p1[i] = c;
    if (p2 != nullptr) {
        const float ox = isfinite(origin_box[0]) ? origin_box[0] : 0.0f, oy = isfinite(origin_box[1]) ? origin_box[1] : 0.0f;
        const float x = a.cx - ox, y = a.cy - oy;
        const float q = x * x + y * y, rr = a.r * a.r;
        RowP2 r;
        r.mx = -2.0f * x; r.my = -2.0f * y; r.mr = -2.0f * a.r;
        r.k = (q - rr) - IOU_SLACK * (q + rr) - 1e-6f;
        p2[i] = r;
        if (p2d != nullptr) { RowP2D d = { r.mx, r.mx, r.my, r.my, r.mr, r.mr, r.k, r.k }; p2d[i] = d; }
    }
}

// rows and columns of one problem in ONE launch (blocks [0, row_blocks) prepare the rows, the rest the columns); block 0 also
// zeroes the sweep's counters, so that the matrix op is two stream operations: this kernel and the pair kernel
__global__ void prep_pair_kernel(const float* __restrict__ boxes1, int64_t m, int64_t stride1, const float* __restrict__ boxes2, int64_t n,
                                 int64_t stride2, int variant, unsigned row_blocks, BoxP0* __restrict__ r0, BoxP1* __restrict__ r1,
                                 RowP2* __restrict__ r2, RowP2D* __restrict__ r2d, BoxP0* __restrict__ c0, BoxP1* __restrict__ c1,
                                 unsigned long long* __restrict__ stats) {
    if (blockIdx.x == 0 && threadIdx.x < 32) stats[threadIdx.x] = 0ull;
    const bool rows = blockIdx.x < row_blocks;
    const int64_t i = (int64_t)(rows ? blockIdx.x : blockIdx.x - row_blocks) * blockDim.x + threadIdx.x;
    if (i >= (rows ? m : n)) return;
    const float* b = rows ? boxes1 + i * stride1 : boxes2 + i * stride2;
    float raw[5] = { b[0], b[1], b[2], b[3], b[4] };
    BoxP0 a; BoxP1 c;
    emu::prep_box_strict(raw, variant, a, c);
    if (!rows) { c0[i] = a; c1[i] = c; return; }
    r0[i] = a; r1[i] = c;
    const float ox = isfinite(boxes1[0]) ? boxes1[0] : 0.0f, oy = isfinite(boxes1[1]) ? boxes1[1] : 0.0f;
    const float x = a.cx - ox, y = a.cy - oy;
    const float q = x * x + y * y, rr = a.r * a.r;
    RowP2 r;
    r.mx = -2.0f * x; r.my = -2.0f * y; r.mr = -2.0f * a.r;
    r.k = (q - rr) - IOU_SLACK * (q + rr) - 1e-6f;
    r2[i] = r;
    RowP2D d = { r.mx, r.mx, r.my, r.my, r.mr, r.mr, r.k, r.k };
    r2d[i] = d;
}

__device__ __noinline__ float emu_pair_call(const float* b1, const float* b2, int variant, int mode) {
    float x[5] = { b1[0], b1[1], b1[2], b1[3], b1[4] };
    float y[5] = { b2[0], b2[1], b2[2], b2[3], b2[4] };
    return emu::pair(x, y, variant, mode);
}

struct __attribute__((aligned(16))) TieRec { int row, slot; float iou; int val; };   // global row, lowq slot, overlap, row-in-image + 1

// Batched assignment: image p owns rows [row0, row0 + m) of the row planes and columns [col0, col0 + n) of the column
// planes (col0 = 0 for every image when the anchors are shared); its per-anchor results live at slot = column + cb_shift.
constexpr int ASSIGN_MAX_IMAGES = 64;
struct AssignProb { int row0, m, col0, n, cb_shift, pad; long long item0; };
struct AssignTable { AssignProb p[ASSIGN_MAX_IMAGES]; int nprob, pad; long long total_items; };

struct IouArgs {
    const BoxP0* r0; const BoxP1* r1; const RowP2* r2; const RowP2D* r2d; int m;
    const BoxP0* c0; const BoxP1* c1; int n;
    const float* raw1; int64_t s1;
    const float* raw2; int64_t s2;
    const float* origin_box;
    int variant, mode, small_mask;
    float tau;
    float* out;
    unsigned long long* stats;      // [0..3] counters, [4] item ticket
    // fused assigner (OUT = 1, 2): packed (iou bits << 32 | ~index) running maxima, and the low-quality match table
    unsigned long long* col_best;   // per column (anchor): best row (GT)
    unsigned long long* row_best;   // per row (GT): best column (anchor)
    int* lowq;                      // per column: largest (row + 1) whose overlap equals that row's maximum
    float min_pos_iou;
    // tie candidates of pass 1 (gt_max_assign_all): pairs whose overlap reached their row's running maximum when they
    // were evaluated.  The running maximum only grows, so every pair that equals the FINAL maximum is in the list.
    TieRec* ties; unsigned tie_cap;  // stats[5] counts the records (may exceed tie_cap: then pass 2 sweeps instead)
    const AssignTable* table;        // OUT = 1, 2: the images of the batch (device memory)
};

struct ItemCtx { int row0, col0, cb_shift, prob; };      // the image an item (or a queued pair) belongs to

// OUT selects what the kernel does with the overlaps it computes:
//   0  store the (M, N) matrix                                         (RBboxOverlaps2D_v*)
//   1  fused assigner pass 1: per-column and per-row max / argmax, no matrix (MaxIoUAssigner's two reductions)
//   2  fused assigner pass 2: columns whose overlap with a row EQUALS that row's maximum (gt_max_assign_all)
enum { OUT_MATRIX = 0, OUT_ASSIGN_MAX = 1, OUT_ASSIGN_TIES = 2 };

__device__ __forceinline__ unsigned long long pack_best(float v, int idx) {
    return ((unsigned long long)__float_as_uint(fmaxf(v, 0.0f)) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ void update_best(unsigned long long* slot, unsigned long long cand) {
    // the maxima settle after a few updates: read first (L2), issue the atomic only when it would change the slot
    if (cand > __ldcg(slot)) atomicMax(slot, cand);
}

// `pre` (assigner pass 1): the two running maxima of the pair, read BEFORE its overlap was computed (the L2 latency hides under
// the area stage).  They can only be older, i.e. smaller, than the slots are now: an atomic may be issued that no longer
// changes its slot, a tie record may be appended that the check kernel drops again — never the other way round.
struct BestPre { unsigned long long col, row; bool have; };

template <int OUT>
__device__ __forceinline__ void emit_overlap(const IouArgs& A, int i, int j, float r, const ItemCtx& c, BestPre pre = BestPre{ 0ull, 0ull, false }) {
    if (OUT == OUT_MATRIX) {
        A.out[(int64_t)i * A.n + j] = r;
    } else if (OUT == OUT_ASSIGN_MAX) {
        if (r > 0.0f) {
            const unsigned long long ccand = pack_best(r, i - c.row0);
            if (ccand > (pre.have ? pre.col : __ldcg(A.col_best + j + c.cb_shift))) atomicMax(A.col_best + j + c.cb_shift, ccand);
            const unsigned long long cand = pack_best(r, j - c.col0), cur = pre.have ? pre.row : __ldcg(A.row_best + i);
            if (cand > cur) atomicMax(A.row_best + i, cand);
            if (A.ties != nullptr && r >= __uint_as_float((unsigned)(cur >> 32)) && r >= A.min_pos_iou) {
                const unsigned long long t = atomicAdd(A.stats + 5, 1ull);
                if (t < A.tie_cap) { TieRec e = { i, j + c.cb_shift, r, i - c.row0 + 1 }; A.ties[t] = e; }
            }
        }
    } else {
        const float gm = __uint_as_float((unsigned)(__ldcg(A.row_best + i) >> 32));
        if (r == gm && gm >= A.min_pos_iou) atomicMax(A.lowq + j + c.cb_shift, i - c.row0 + 1);
    }
}

__device__ __forceinline__ float4 ldg4(const void* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct PairBoxes { BoxP0 A0, B0; BoxP1 A1, B1; };
__device__ __forceinline__ PairBoxes load_pair(const IouArgs& A, int i, int j) {
    float4 a0 = ldg4(A.r0 + i), a1 = ldg4(A.r1 + i), b0 = ldg4(A.c0 + j), b1 = ldg4(A.c1 + j);
    PairBoxes P;
    P.A0 = { a0.x, a0.y, a0.z, a0.w }; P.A1 = { a1.x, a1.y, a1.z, a1.w };
    P.B0 = { b0.x, b0.y, b0.z, b0.w }; P.B1 = { b1.x, b1.y, b1.z, b1.w };
    return P;
}

// Per-warp shared-memory working set: the item's prepared boxes (staged with cp.async) and the queues.
struct __attribute__((aligned(16))) WarpSmem {
    float4 r0[IOU_TM], r1[IOU_TM];                 // rows of the item: BoxP0, BoxP1 (RowP2 is read by broadcast LDG)
    float4 c0[IOU_TN], c1[IOU_TN];                 // columns of the item: BoxP0, BoxP1
    uint2 q3[64];                                  // flagged pairs (absolute row, col); persists across items
    unsigned short q1[IOU_Q1CAP];                  // circumradius survivors, item-relative (row << 7 | col)
    unsigned short q2[64];                         // separating-axis survivors
    int ctx[4];                                    // assigner modes: row0, col0, cb_shift, image of the item in flight
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ BoxP0 as_p0(float4 v) { BoxP0 b = { v.x, v.y, v.z, v.w }; return b; }
__device__ __forceinline__ BoxP1 as_p1(float4 v) { BoxP1 b = { v.x, v.y, v.z, v.w }; return b; }

template <bool VEC, int OUT, bool BATCH>
__global__ void __launch_bounds__(IOU_THREADS, R3G_IOU_MINB) iou_matrix_kernel(const IouArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    WarpSmem& W = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
#if R3G_IOU_TABCOMPACT
    __shared__ unsigned q1_offset[32];          // queue entry of pair t = 4 row + column of a group, relative to (row base << 7 | 4 lane)
    if (threadIdx.x < 32) { const unsigned t = 31u - threadIdx.x; q1_offset[threadIdx.x] = ((t & ~3u) << 5) + (t & 3u); }   // bit b holds pair t = 31 - b
    __syncthreads();
#endif
    int c1 = 0, c2 = 0, c3 = 0;
    const unsigned lt = lanemask_lt();
    // matrix mode: one (m, n) problem.  Assigner modes: the images of the batch, items enumerated image by image.
    const int tiles_n_all = (A.n + IOU_TN - 1) / IOU_TN;
    const int tiles_m = (A.m + IOU_TM - 1) / IOU_TM;
    const long long total = BATCH ? A.table->total_items : (long long)tiles_m * tiles_n_all;
    unsigned n_circle = 0, n_sat = 0, n_emu = 0;
    if (OUT == OUT_ASSIGN_TIES && A.ties != nullptr && __ldcg(A.stats + 5) <= A.tie_cap) return;   // the list was complete
    float ox = __ldg(A.origin_box), oy = __ldg(A.origin_box + 1);     // common origin of the expanded circle test
    if (!isfinite(ox)) ox = 0.0f;
    if (!isfinite(oy)) oy = 0.0f;

    int i0 = 0, j0 = 0, i1 = 0, jb = 0, cshift = 0;
    float cx[IOU_CPL], cy[IOU_CPL], cr[IOU_CPL], ck[IOU_CPL];
    unsigned long long CX[IOU_CPL / 2], CY[IOU_CPL / 2], CR[IOU_CPL / 2], CK[IOU_CPL / 2];   // the same, two columns per register pair
    float cm[IOU_CPL] = { 0.f, 0.f, 0.f, 0.f };     // OUT_ASSIGN_TIES: the columns' best overlap (from pass 1)
    float gm_lo = 0.f, gm_hi = 0.f, cmax_item = 0.f;  // OUT_ASSIGN_TIES: row maxima of the item, best column maximum
    bool full4 = false;
    float* orow = A.out;

    // ---- stage 4: reference restatement for a batch of flagged pairs (a __noinline__ call: its registers / stack do not
    //      tax the fast path) ----
    auto stage4 = [&](int nb) {
        __syncwarp();
        if ((int)lane < nb) {
            uint2 e = W.q3[c3 - nb + lane];
            ItemCtx ec = { 0, 0, 0, 0 };
            if (BATCH) {                              // the pair may belong to an earlier image: its id rides in the top bits
                const AssignProb& pr = A.table->p[e.x >> 26];
                ec.row0 = pr.row0; ec.col0 = pr.col0; ec.cb_shift = pr.cb_shift;
                e.x &= 0x3ffffffu;
            }
            float r = emu_pair_call(A.raw1 + (int64_t)e.x * A.s1, A.raw2 + (int64_t)e.y * A.s2, A.variant, A.mode);
            if (A.small_mask) {
                const float4 a1 = ldg4(A.r1 + e.x), b1 = ldg4(A.c1 + e.y);
                if (fminf(a1.z, a1.w) * 2.0f < 0.001f || fminf(b1.z, b1.w) * 2.0f < 0.001f) r = 0.0f;
            }
            emit_overlap<OUT>(A, (int)e.x, (int)e.y, r, ec);
        }
        __syncwarp();
        c3 -= nb;
        n_emu += nb;
    };

    // Items (64 rows x 128 columns) come from a dynamic ticket (the counter is zeroed by the launcher).  Inside an item the
    // groups of 8 rows run stage 1 and feed the queues; after every group (and once more after the last one, with `fl` set, to
    // flush what is left) the queue stages run deepest first.  Every stage body exists exactly once in the instruction stream.
    while (true) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(A.stats + 4, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        int tiles_n = tiles_n_all, i_end = A.m, j_end = A.n;      // row / column limits of the item's problem
        if (BATCH) {
            // image of this item: last p with item0[p] <= item
            int lo = 0, hi = A.table->nprob;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (A.table->p[mid].item0 <= item) lo = mid; else hi = mid;
            }
            const AssignProb pr = A.table->p[lo];
            item -= pr.item0;
            tiles_n = (pr.n + IOU_TN - 1) / IOU_TN;
            __syncwarp();
            if (lane == 0) { W.ctx[0] = pr.row0; W.ctx[1] = pr.col0; W.ctx[2] = pr.cb_shift; W.ctx[3] = lo; }
            i_end = pr.row0 + pr.m; j_end = pr.col0 + pr.n;
            cshift = pr.cb_shift;
            i0 = pr.row0; j0 = pr.col0;
        } else {
            i0 = 0; j0 = 0;
        }
        const int tm = (int)(item / tiles_n), tn = (int)(item - (long long)tm * tiles_n);
        i0 += tm * IOU_TM; j0 += tn * IOU_TN;
        i1 = min(i_end, i0 + IOU_TM);
        jb = j0 + (int)lane * IOU_CPL;
        __syncwarp();
        // stage the item's prepared boxes: 2 x 64 row records + 2 x 128 column records, 16 bytes each
        for (int t = lane; t < IOU_TM; t += 32) {
            const int i = min(i0 + t, i_end - 1);
            cp_async16(&W.r0[t], A.r0 + i); cp_async16(&W.r1[t], A.r1 + i);
        }
        for (int t = lane; t < IOU_TN; t += 32) {
            const int j = min(j0 + t, j_end - 1);
            cp_async16(&W.c0[t], A.c0 + j); cp_async16(&W.c1[t], A.c1 + j);
        }
        cp_async_wait_all();
        __syncwarp();
        // this lane's 4 columns relative to the origin; invalid columns are pushed to +inf (always rejected)
#pragma unroll
        for (int k = 0; k < IOU_CPL; k++) {
            if (jb + k < j_end) {
                const float4 b = W.c0[lane * IOU_CPL + k];
                cx[k] = b.x - ox; cy[k] = b.y - oy; cr[k] = b.z;
                const float q = cx[k] * cx[k] + cy[k] * cy[k], rr = b.z * b.z;
                ck[k] = (q - rr) - IOU_SLACK * (q + rr);
            } else {
                cx[k] = 0.0f; cy[k] = 0.0f; cr[k] = 0.0f; ck[k] = 3.0e38f;
            }
        }
        if (OUT == OUT_ASSIGN_MAX) {
#pragma unroll
            for (int h = 0; h < IOU_CPL / 2; h++) {
                CX[h] = pack2(cx[2 * h], cx[2 * h + 1]); CY[h] = pack2(cy[2 * h], cy[2 * h + 1]);
                CR[h] = pack2(cr[2 * h], cr[2 * h + 1]); CK[h] = pack2(ck[2 * h], ck[2 * h + 1]);
            }
        }
        full4 = VEC && (jb + IOU_CPL <= j_end);
        const bool full_tile = VEC && (j0 + IOU_TN <= j_end);        // warp-uniform: every lane owns four valid columns
        if (OUT == OUT_MATRIX) orow = A.out + (int64_t)i0 * A.n + jb;
        if (OUT == OUT_ASSIGN_TIES) {
#pragma unroll
            for (int k = 0; k < IOU_CPL; k++)
                cm[k] = (jb + k < j_end) ? __uint_as_float((unsigned)(__ldcg(A.col_best + jb + k + cshift) >> 32)) : -1.0f;
            // row maxima of the item (lane t holds rows t and t+32) and the best column maximum of the item:
            // a row whose maximum exceeds every column maximum of the item cannot tie here and is skipped whole
            gm_lo = (i0 + (int)lane < i1) ? __uint_as_float((unsigned)(__ldcg(A.row_best + i0 + lane) >> 32)) : 0.0f;
            gm_hi = (i0 + 32 + (int)lane < i1) ? __uint_as_float((unsigned)(__ldcg(A.row_best + i0 + 32 + lane) >> 32)) : 0.0f;
            cmax_item = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cmax_item = fmaxf(cmax_item, __shfl_xor_sync(0xffffffffu, cmax_item, o));
        }

        for (int ig = i0; ; ig += IOU_RG) {
            const bool fl = ig >= i1;                  // the flush pass behind the item's last group
            if (!fl) {
                // ---- stage 1: one group of 8 rows x (4 columns per lane) ----
                const int nr = min(IOU_RG, i1 - ig);
                unsigned m = 0;
#if R3G_IOU_FASTGROUP
                // matrix mode, a full group inside a full tile (all but the edge items): straight-line code — no per-row bounds
                // or tail tests, one base address for the row records, rows loaded four at a time
                if (OUT == OUT_MATRIX && nr == IOU_RG && full_tile) {
                    const RowP2* __restrict__ rp = A.r2 + ig;
#pragma unroll
                    for (int h = 0; h < IOU_RG; h += 4) {
                        float4 a[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) a[r] = ldg4(rp + h + r);
#pragma unroll
                        for (int r = 0; r < 4; r++) {
#pragma unroll
                            for (int k = 0; k < IOU_CPL; k++) {
                                float s = fmaf(a[r].x, cx[k], ck[k] + a[r].w);
                                s = fmaf(a[r].y, cy[k], s);
                                s = fmaf(a[r].z, cr[k], s);
                                m = __funnelshift_l(__float_as_uint(s), m, 1);
                            }
                            st_cs_f4(orow, make_float4(0.f, 0.f, 0.f, 0.f));
                            orow += A.n;
                        }
                    }
                } else if (OUT == OUT_ASSIGN_MAX && nr == IOU_RG && full_tile) {
                    // the assigner sweep's full groups: the same straight-line form in packed f32x2 arithmetic (two columns
                    // per instruction; add, then three fused multiply-adds, exactly as in the general path below)
                    const ulonglong2* __restrict__ rp = reinterpret_cast<const ulonglong2*>(A.r2d + ig);
#pragma unroll
                    for (int h = 0; h < IOU_RG; h += 4) {
                        ulonglong2 ra[4], rb[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) { ra[r] = __ldg(rp + 2 * (h + r)); rb[r] = __ldg(rp + 2 * (h + r) + 1); }
#pragma unroll
                        for (int r = 0; r < 4; r++) {
#pragma unroll
                            for (int q = 0; q < IOU_CPL / 2; q++) {
                                unsigned long long t = add2(CK[q], rb[r].y);
                                t = fma2(ra[r].x, CX[q], t);
                                t = fma2(ra[r].y, CY[q], t);
                                t = fma2(rb[r].x, CR[q], t);
                                m = __funnelshift_l((unsigned)t, m, 1);
                                m = __funnelshift_l((unsigned)(t >> 32), m, 1);
                            }
                        }
                    }
                } else {
#endif
#if R3G_IOU_BATCHLOAD
                float4 arow[IOU_RG];                   // matrix mode: all row records in flight before the first use
                if (OUT == OUT_MATRIX) {
#pragma unroll
                    for (int r = 0; r < IOU_RG; r++) arow[r] = ldg4(A.r2 + ig + min(r, nr - 1));
                }
#endif
#pragma unroll
                for (int r = 0; r < IOU_RG; r++) {
                    if (r < nr) {
                        if (OUT == OUT_ASSIGN_TIES) {
                            static_assert(IOU_TM <= 64, "row maxima are cached two per lane");
                            const int rl = ig - i0 + r;
                            const float gm = __shfl_sync(0xffffffffu, rl < 32 ? gm_lo : gm_hi, rl & 31);
                            if (!(gm > 0.0f) || gm > cmax_item) { m <<= IOU_CPL; continue; }      // warp-uniform skip
                            const float4 a = ldg4(A.r2 + ig + r);
#pragma unroll
                            for (int k = 0; k < IOU_CPL; k++) {
                                float s = fmaf(a.x, cx[k], ck[k] + a.w);
                                s = fmaf(a.y, cy[k], s);
                                s = fmaf(a.z, cr[k], s);
                                m = (m << 1) | ((s < 0.0f && cm[k] >= gm) ? 1u : 0u);
                            }
                            continue;
                        }
                        // s < 0  <=>  centres closer than the sum of the (conservative) circumradii
                        if (OUT == OUT_MATRIX) {
#if R3G_IOU_BATCHLOAD
                            const float4 a = arow[r];
#else
                            const float4 a = ldg4(A.r2 + ig + r);
#endif
#pragma unroll
                            for (int k = 0; k < IOU_CPL; k++) {
                                float s = fmaf(a.x, cx[k], ck[k] + a.w);
                                s = fmaf(a.y, cy[k], s);
                                s = fmaf(a.z, cr[k], s);
                                m = __funnelshift_l(__float_as_uint(s), m, 1);
                            }
                        } else {
                            // two columns per instruction, the same operations in the same order (add, then three fused multiply-adds)
                            const ulonglong2 ra = __ldg(reinterpret_cast<const ulonglong2*>(A.r2d + ig + r));        // {mx, mx}, {my, my}
                            const ulonglong2 rb = __ldg(reinterpret_cast<const ulonglong2*>(A.r2d + ig + r) + 1);    // {mr, mr}, {k, k}
#pragma unroll
                            for (int h = 0; h < IOU_CPL / 2; h++) {
                                unsigned long long t = add2(CK[h], rb.y);
                                t = fma2(ra.x, CX[h], t);
                                t = fma2(ra.y, CY[h], t);
                                t = fma2(rb.x, CR[h], t);
                                m = __funnelshift_l((unsigned)t, m, 1);
                                m = __funnelshift_l((unsigned)(t >> 32), m, 1);
                            }
                        }
                        if (OUT == OUT_MATRIX) {
                            if (full4) {
#if R3G_IOU_ZERO_CS
                                st_cs_f4(orow, make_float4(0.f, 0.f, 0.f, 0.f));
#else
                                *reinterpret_cast<float4*>(orow) = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
                            } else {
#pragma unroll
                                for (int k = 0; k < IOU_CPL; k++)
                                    if (jb + k < A.n) st_cs_f1(orow + k, 0.f);      // matrix mode only: one problem
                            }
                            orow += A.n;
                        }
                    }
                }
#if R3G_IOU_FASTGROUP
                }
#endif
                // compact the group's survivors: warp scan of popc, then each lane emits its own bits (row-major order)
                const int cnt = __popc(m);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    // the shuffle's own predicate (source lane in range) guards the add: two instructions per step
                    asm volatile("{ .reg .pred p; .reg .s32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.s32 %0, %0, t; }"
                                 : "+r"(incl) : "r"(d));
                }
                const int tot = __shfl_sync(0xffffffffu, incl, 31);
                if (tot) {
                    int pos = c1 + incl - cnt;
                    const int nbits = nr * IOU_CPL;
                    // entry = (row << 7 | column), row = rowbase + idx / 4, column = 4 lane + idx % 4, idx = nbits - 1 - bit
                    const unsigned ebase = ((unsigned)(ig - i0) << 7) | (lane * IOU_CPL);
#if R3G_IOU_UNROLLCOMPACT
                    // 32 predicated stores: no find-leading-one (XU pipe, long dependent chain), no divergent trip counts
                    m <<= 32 - nbits;                                  // pair t = 4 row + column now sits at bit 31 - t
                    unsigned short* qp = W.q1 + pos;
#pragma unroll
                    for (int t = 0; t < 32; t++) {
                        if (m & (0x80000000u >> t)) { *qp = (unsigned short)(ebase + (unsigned)(((t & ~3) << 5) + (t & 3))); qp++; }
                    }
#else
#if R3G_IOU_TABCOMPACT
                    // pair t = 4 row + column moved to bit 31 - t; the entry offset of a bit comes from a 32-entry table
                    m <<= 32 - nbits;
                    unsigned short* qp = W.q1 + pos;
                    while (m) {
                        unsigned b;
                        asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(m));   // one find-leading-one; the table is indexed by the bit
                        m ^= 1u << b;
                        *qp++ = (unsigned short)(ebase + q1_offset[b]);
                    }
#else
                    while (m) {
                        const int b = 31 - __clz(m);
                        m ^= 1u << b;
                        const unsigned idx = (unsigned)(nbits - 1 - b);
                        W.q1[pos++] = (unsigned short)(ebase + ((idx & ~3u) << 5) + (idx & 3u));
                    }
#endif
#endif
                    c1 += tot;
                }
            }
            // ---- queue stages, deepest first ----
            while (true) {
                if (c3 >= 32) { stage4(32); continue; }
                if (c2 >= 32 || (fl && c1 == 0 && c2 > 0)) {
                    // ---- stage 3: area integral + epilogue for truly overlapping pairs ----
                    const int nb = min(32, c2);
                    __syncwarp();
                    bool risk = false;
                    int i = 0, j = 0;
                    if ((int)lane < nb) {
                        const unsigned e = W.q2[c2 - nb + lane];
                        const int il = (int)(e >> 7), jl = (int)(e & 127u);
                        i = i0 + il; j = j0 + jl;
                        ItemCtx ctx = { 0, 0, 0, 0 };
                        if (BATCH) { ctx.row0 = W.ctx[0]; ctx.col0 = W.ctx[1]; ctx.cb_shift = W.ctx[2]; }
                        BestPre pre = { 0ull, 0ull, false };
#if R3G_IOU_PREFETCH_BEST
                        if (OUT == OUT_ASSIGN_MAX) { pre.col = __ldcg(A.col_best + j + ctx.cb_shift); pre.row = __ldcg(A.row_best + i); pre.have = true; }
#endif
                        const BoxP0 A0 = as_p0(W.r0[il]), B0 = as_p0(W.c0[jl]);
                        const BoxP1 A1 = as_p1(W.r1[il]), B1 = as_p1(W.c1[jl]);
                        float r = pair_overlap(A0, A1, B0, B1, A.variant, A.mode, A.tau, risk, true);      // survivors of stage 2
                        if (A.small_mask && (fminf(A1.hw, A1.hh) * 2.0f < 0.001f || fminf(B1.hw, B1.hh) * 2.0f < 0.001f)) r = 0.0f;
                        if (!risk && r != 0.0f) emit_overlap<OUT>(A, i, j, r, ctx, pre);
                    }
                    __syncwarp();
                    c2 -= nb;
                    n_sat += nb;
                    const unsigned bal = __ballot_sync(0xffffffffu, risk);
                    if (risk) W.q3[c3 + __popc(bal & lt)] = make_uint2((unsigned)i | (BATCH ? ((unsigned)W.ctx[3] << 26) : 0u), (unsigned)j);
                    c3 += __popc(bal);
                    continue;
                }
                if (c1 >= 32 || (fl && c1 > 0)) {
                    // ---- stage 2: separating-axis test ----
                    const int nb = min(32, c1);
                    __syncwarp();
                    bool ok = false;
                    unsigned e = 0;
                    if ((int)lane < nb) {
                        e = W.q1[c1 - nb + lane];
                        const int il = (int)(e >> 7), jl = (int)(e & 127u);
                        ok = pair_sat(as_p0(W.r0[il]), as_p1(W.r1[il]), as_p0(W.c0[jl]), as_p1(W.c1[jl]));
                    }
                    __syncwarp();
                    c1 -= nb;
                    n_circle += nb;
                    const unsigned bal = __ballot_sync(0xffffffffu, ok);
                    if (ok) W.q2[c2 + __popc(bal & lt)] = (unsigned short)e;
                    c2 += __popc(bal);
                    continue;
                }
                break;
            }
            if (fl) break;
        }
    }
    while (c3 > 0) stage4(min(32, c3));     // flagged pairs persist across items (absolute indices): the last ones

    if (A.stats) {
        if (blockIdx.x == 0 && threadIdx.x == 0) A.stats[3] = (unsigned long long)A.m * (unsigned long long)A.n;
        if (lane == 0) {
            atomicAdd(A.stats + 0, (unsigned long long)n_circle);
            atomicAdd(A.stats + 1, (unsigned long long)n_sat);
            atomicAdd(A.stats + 2, (unsigned long long)n_emu);
        }
    }
}

__global__ void iou_aligned_kernel(const float* __restrict__ b1, int64_t n1, int64_t s1,
                                   const float* __restrict__ b2, int64_t n2, int64_t s2,
                                   int variant, int mode, float tau, int small_mask, float* __restrict__ out) {
    const int64_t n = n1 > n2 ? n1 : n2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float* p = b1 + (i % n1) * s1;
        const float* q = b2 + (i % n2) * s2;
        float x[5] = { p[0], p[1], p[2], p[3], p[4] };
        float y[5] = { q[0], q[1], q[2], q[3], q[4] };
        BoxP0 A0, B0; BoxP1 A1, B1;
        emu::prep_box_strict(x, variant, A0, A1);
        emu::prep_box_strict(y, variant, B0, B1);
        float r = 0.0f;
        if (!circle_reject(A0, B0)) {
            bool risk;
            r = pair_overlap(A0, A1, B0, B1, variant, mode, tau, risk);
            if (risk) r = emu_pair_call(p, q, variant, mode);
        }
        if (small_mask && (fminf(x[2], x[3]) < 0.001f || fminf(y[2], y[3]) < 0.001f)) r = 0.0f;
        out[i] = r;
    }
}

struct IouWorkspace {
    unsigned long long* stats;
    BoxP0 *r0, *c0;
    BoxP1 *r1, *c1;
    RowP2* r2;
    RowP2D* r2d;
    size_t bytes;
};

static IouWorkspace carve(void* ws, int64_t m, int64_t n) {
    IouWorkspace w;
    char* p = (char*)ws;
    size_t off = 0;
    w.stats = (unsigned long long*)(p + off); off += 256;
    w.r0 = (BoxP0*)(p + off); off += align_up(sizeof(BoxP0) * (size_t)m, 256);
    w.r1 = (BoxP1*)(p + off); off += align_up(sizeof(BoxP1) * (size_t)m, 256);
    w.c0 = (BoxP0*)(p + off); off += align_up(sizeof(BoxP0) * (size_t)n, 256);
    w.c1 = (BoxP1*)(p + off); off += align_up(sizeof(BoxP1) * (size_t)n, 256);
    w.r2 = (RowP2*)(p + off); off += align_up(sizeof(RowP2) * (size_t)m, 256);
    w.r2d = (RowP2D*)(p + off); off += align_up(sizeof(RowP2D) * (size_t)m, 256);
    w.bytes = off;
    return w;
}

}  // namespace r3g

using namespace r3g;

R3G_API int r3g_iou_workspace_bytes(int64_t m, int64_t n, size_t* bytes) {
    R3G_REQUIRE(bytes != nullptr && m >= 0 && n >= 0, "r3g_iou_workspace_bytes: bad arguments");
    *bytes = carve(nullptr, m, n).bytes;
    return R3G_OK;
}

// Launch the pair sweep in one of its output modes (persistent grid = SMs x occupancy, dynamic item tickets).
template <bool VEC, int OUT, bool BATCH = false>
static int launch_sweep(const IouArgs& a, cudaStream_t st, int64_t item_count = -1) {
    const size_t smem = sizeof(WarpSmem) * IOU_WARPS;
    static int occ_of[64] = {0};          // per device: function attributes live in the device's context
    int& occ = occ_of[current_device_slot()];
    if (occ == 0) {
        R3G_CUDA_OK(cudaFuncSetAttribute(iou_matrix_kernel<VEC, OUT, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        R3G_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, iou_matrix_kernel<VEC, OUT, BATCH>, IOU_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    const int64_t items = item_count >= 0 ? item_count
                                           : (((int64_t)a.m + IOU_TM - 1) / IOU_TM) * (((int64_t)a.n + IOU_TN - 1) / IOU_TN);
    int64_t grid = (items + IOU_WARPS - 1) / IOU_WARPS;
    if (grid < 1) grid = 1;
    const int64_t cap = (int64_t)device_sm_count() * occ;
    if (grid > cap) grid = cap;
    iou_matrix_kernel<VEC, OUT, BATCH><<<(unsigned)grid, IOU_THREADS, smem, st>>>(a);
    R3G_LAUNCH_OK("iou_matrix_kernel");
    return R3G_OK;
}

static int iou_check_common(const char* who, int64_t m, int64_t n, int variant, int mode) {
    R3G_REQUIRE(m >= 0 && n >= 0, "%s: negative size", who);
    R3G_REQUIRE(variant >= 1 && variant <= 3, "%s: variant must be 1, 2 or 3 (got %d)", who, variant);
    R3G_REQUIRE(mode == R3G_MODE_IOU || mode == R3G_MODE_IOF, "%s: mode must be iou(0) or iof(1)", who);
    R3G_REQUIRE(m < (1ll << 30) && n < (1ll << 30), "%s: more than 2^30 boxes", who);
    return R3G_OK;
}

R3G_API int r3g_iou_prepare_f32(const float* boxes1, int64_t m, int64_t stride1,
                                const float* boxes2, int64_t n, int64_t stride2, int variant,
                                void* workspace, size_t workspace_bytes, void* stream) {
    int rc = iou_check_common("r3g_iou_prepare_f32", m, n, variant, R3G_MODE_IOU);
    if (rc != R3G_OK) return rc;
    if (m == 0 || n == 0) return R3G_OK;
    R3G_REQUIRE(boxes1 && boxes2 && workspace, "r3g_iou_prepare_f32: null pointer");
    R3G_REQUIRE(stride1 >= 5 && stride2 >= 5, "r3g_iou_prepare_f32: box stride must be >= 5 floats");
    IouWorkspace w = carve(workspace, m, n);
    if (workspace_bytes < w.bytes) {
        set_error("r3g_iou_prepare_f32: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned rb = (unsigned)((m + 255) / 256), cbk = (unsigned)((n + 255) / 256);
    prep_pair_kernel<<<rb + cbk, 256, 0, st>>>(boxes1, m, stride1, boxes2, n, stride2, variant, rb, w.r0, w.r1, w.r2, w.r2d, w.c0, w.c1, w.stats);
    R3G_LAUNCH_OK("prep_pair_kernel");
    return R3G_OK;
}

static int iou_matrix_prepared(const float* boxes1, int64_t m, int64_t stride1, const float* boxes2, int64_t n, int64_t stride2,
                               int variant, int mode, int flags, float* out, void* workspace, size_t workspace_bytes, void* stream,
                               bool counters_zeroed) {
    int rc = iou_check_common("r3g_iou_matrix_prepared_f32", m, n, variant, mode);
    if (rc != R3G_OK) return rc;
    if (m == 0 || n == 0) return R3G_OK;
    R3G_REQUIRE(boxes1 && boxes2 && out && workspace, "r3g_iou_matrix_prepared_f32: null pointer");
    IouWorkspace w = carve(workspace, m, n);
    if (workspace_bytes < w.bytes) {
        set_error("r3g_iou_matrix_prepared_f32: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (!counters_zeroed) R3G_CUDA_OK(cudaMemsetAsync(w.stats, 0, 256, st));       // (the prepare kernel of the same call has done it)
    IouArgs a = {};
    a.r0 = w.r0; a.r1 = w.r1; a.r2 = w.r2; a.r2d = w.r2d; a.m = (int)m; a.c0 = w.c0; a.c1 = w.c1; a.n = (int)n;
    a.raw1 = boxes1; a.s1 = stride1; a.raw2 = boxes2; a.s2 = stride2; a.origin_box = boxes1;
    a.variant = variant; a.mode = mode;
    a.small_mask = (variant == R3G_V3 && (flags & R3G_FLAG_SMALL_MASK)) ? 1 : 0;
    a.tau = (flags & R3G_FLAG_EMULATE_ALL) ? 1e30f : ((flags & R3G_FLAG_STRICT) ? 2e-2f : 0.0f);
    a.out = out; a.stats = w.stats;

    a.col_best = nullptr; a.row_best = nullptr; a.lowq = nullptr; a.min_pos_iou = 0.0f;
    const bool vec = (n % 4 == 0) && (((uintptr_t)out & 15u) == 0);
    rc = vec ? launch_sweep<true, OUT_MATRIX>(a, st) : launch_sweep<false, OUT_MATRIX>(a, st);
    if (rc != R3G_OK) return rc;
    return R3G_OK;
}

R3G_API int r3g_iou_matrix_prepared_f32(const float* boxes1, int64_t m, int64_t stride1,
                                        const float* boxes2, int64_t n, int64_t stride2,
                                        int variant, int mode, int flags, float* out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    return iou_matrix_prepared(boxes1, m, stride1, boxes2, n, stride2, variant, mode, flags, out, workspace, workspace_bytes, stream, false);
}

R3G_API int r3g_iou_matrix_f32(const float* boxes1, int64_t m, int64_t stride1,
                               const float* boxes2, int64_t n, int64_t stride2,
                               int variant, int mode, int flags, float* out,
                               void* workspace, size_t workspace_bytes, void* stream) {
    int rc = iou_check_common("r3g_iou_matrix_f32", m, n, variant, mode);
    if (rc != R3G_OK) return rc;
    rc = r3g_iou_prepare_f32(boxes1, m, stride1, boxes2, n, stride2, variant, workspace, workspace_bytes, stream);
    if (rc != R3G_OK) return rc;
    return iou_matrix_prepared(boxes1, m, stride1, boxes2, n, stride2, variant, mode, flags, out, workspace, workspace_bytes, stream,
                               m > 0 && n > 0);
}

R3G_API int r3g_iou_aligned_f32(const float* boxes1, int64_t n1, int64_t stride1,
                                const float* boxes2, int64_t n2, int64_t stride2,
                                int variant, int mode, int flags, float* out, void* stream) {
    R3G_REQUIRE(n1 >= 0 && n2 >= 0, "r3g_iou_aligned_f32: negative size");
    R3G_REQUIRE(variant >= 1 && variant <= 3, "r3g_iou_aligned_f32: variant must be 1, 2 or 3 (got %d)", variant);
    R3G_REQUIRE(mode == R3G_MODE_IOU || mode == R3G_MODE_IOF, "r3g_iou_aligned_f32: mode must be iou(0) or iof(1)");
    if (n1 == 0 || n2 == 0) return R3G_OK;
    R3G_REQUIRE(boxes1 && boxes2 && out, "r3g_iou_aligned_f32: null pointer");
    R3G_REQUIRE(stride1 >= 5 && stride2 >= 5, "r3g_iou_aligned_f32: box stride must be >= 5 floats");
    const int64_t n = n1 > n2 ? n1 : n2;
    int64_t grid = (n + 255) / 256;
    const int64_t cap = (int64_t)device_sm_count() * 8;
    if (grid > cap) grid = cap;
    iou_aligned_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        boxes1, n1, stride1, boxes2, n2, stride2, variant, mode, (flags & R3G_FLAG_EMULATE_ALL) ? 1e30f : ((flags & R3G_FLAG_STRICT) ? 2e-2f : 0.0f),
        (variant == R3G_V3 && (flags & R3G_FLAG_SMALL_MASK)) ? 1 : 0, out);
    R3G_LAUNCH_OK("iou_aligned_kernel");
    return R3G_OK;
}

// ---- fused assigner --------------------------------------------------------------------------------------------------
// mmdet-2.19 MaxIoUAssigner.assign_wrt_overlaps around the reference's calculator call (caller:
// r3det/models/dense_heads/rotate_anchor_head.py:220-228; semantics recalled in SURVEY.md A6) without materialising the
// (G, A) overlap matrix: pass 1 keeps the per-anchor and per-GT max/argmax as packed 64-bit atomics, pass 2 (only for
// gt_max_assign_all) re-visits the few pairs whose column maximum reaches their row maximum to find exact ties, and a
// finalize kernel applies the thresholds.  The IoU values are the matrix kernel's, bit for bit (same code path).
namespace r3g {

// the image table travels in kernel-parameter space (graph-capturable: no host buffer outlives the call)
__global__ void assign_table_kernel(const __grid_constant__ AssignTable T, AssignTable* out) {
    if (threadIdx.x < ASSIGN_MAX_IMAGES) out->p[threadIdx.x] = T.p[threadIdx.x];
    if (threadIdx.x == 0) { out->nprob = T.nprob; out->pad = 0; out->total_items = T.total_items; }
}

__global__ void assign_init_kernel(unsigned long long* col_best, int64_t NC, unsigned long long* row_best, int64_t NR, int* lowq,
                                   int* zero_gt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long zero_first = 0xFFFFFFFFull;            // overlap 0, index 0 (torch.max returns the first maximum)
    if (t < NC) { col_best[t] = zero_first; lowq[t] = 0; }
    if (t < NR) row_best[t] = zero_first;
    if (t < ASSIGN_MAX_IMAGES) zero_gt[t] = 0;
}

__device__ __forceinline__ int image_of_row(const AssignTable* T, int i) {
    int lo = 0, hi = T->nprob;                                       // last p with row0[p] <= i and m[p] > 0 covering i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (T->p[mid].row0 <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// per GT: low-quality matching that does not need the tie pass
//   gt_max_assign_all == 0: assigned[gt_argmax[i]] = i + 1 for gt_max[i] >= min_pos_iou        (later GTs win)
//   gt_max[i] == 0 with min_pos_iou <= 0: `overlaps[i, :] == gt_max[i]` holds for EVERY anchor (mmdet quirk)
__global__ void assign_gt_kernel(const unsigned long long* __restrict__ row_best, int64_t NR, const AssignTable* __restrict__ T,
                                 float min_pos_iou, int assign_all, int* lowq, int* zero_gt, float* gt_max, int64_t* gt_argmax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NR) return;
    const unsigned long long b = row_best[i];
    const float v = __uint_as_float((unsigned)(b >> 32));
    const int arg = (int)(0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFull));
    if (gt_max) gt_max[i] = v;
    if (gt_argmax) gt_argmax[i] = arg;
    if (v >= min_pos_iou) {
        const int p = image_of_row(T, (int)i);
        const AssignProb& pr = T->p[p];
        const int rel = (int)i - pr.row0 + 1;
        if (!assign_all) atomicMax(lowq + pr.col0 + arg + pr.cb_shift, rel);
        else if (v == 0.0f) atomicMax(zero_gt + p, rel);
    }
}

// gt_max_assign_all from the tie list of pass 1: a recorded pair matches when its overlap equals the row's final maximum
__global__ void assign_ties_kernel(const TieRec* __restrict__ ties, const unsigned long long* __restrict__ count, unsigned cap,
                                   const unsigned long long* __restrict__ row_best, float min_pos_iou, int* lowq) {
    const unsigned long long n = *count;
    if (n > cap) return;                                             // overflow: the sweep (OUT_ASSIGN_TIES) does it
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (unsigned long long)gridDim.x * blockDim.x) {
        const TieRec e = ties[t];
        const float gm = __uint_as_float((unsigned)(row_best[e.row] >> 32));
        if (e.iou == gm && gm >= min_pos_iou) atomicMax(lowq + e.slot, e.val);
    }
}

// one thread per (image, anchor) slot
__global__ void assign_finalize_kernel(const unsigned long long* __restrict__ col_best, const int* __restrict__ lowq,
                                       const int* __restrict__ zero_gt, const AssignTable* __restrict__ T, int64_t A,
                                       float pos_thr, float neg_thr, int match_low_quality, int64_t* assigned,
                                       float* max_overlaps, int64_t* argmax) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (j >= A) return;
    const AssignProb& pr = T->p[p];
    const int64_t o = (int64_t)p * A + j;                            // output slot; results live at column + cb_shift
    const int64_t slot = pr.col0 + j + pr.cb_shift;
    if (pr.m == 0) {                                                 // no GT in this image: everything is background
        assigned[o] = 0; if (max_overlaps) max_overlaps[o] = 0.0f; if (argmax) argmax[o] = 0;
        return;
    }
    const unsigned long long b = col_best[slot];
    const float v = __uint_as_float((unsigned)(b >> 32));
    const int arg = (int)(0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFull));
    int64_t a = -1;
    if (v >= 0.0f && v < neg_thr) a = 0;
    if (v >= pos_thr) a = arg + 1;
    if (match_low_quality) {
        const int lq = max(lowq[slot], zero_gt[p]);
        if (lq > 0) a = lq;
    }
    assigned[o] = a;
    if (max_overlaps) max_overlaps[o] = v;
    if (argmax) argmax[o] = arg;
}

struct AssignWs {
    IouWorkspace iou; unsigned long long *col_best, *row_best; int *lowq, *zero_gt; TieRec* ties; unsigned tie_cap;
    AssignTable* table; size_t bytes;
};

// NR rows (all GT of the batch), NCP column records (anchors: A shared, or B * A), NC result slots (B * A)
static AssignWs carve_assign(void* ws, int64_t NR, int64_t NCP, int64_t NC) {
    AssignWs w;
    w.iou = carve(ws, NR, NCP);
    char* p = (char*)ws;
    size_t off = w.iou.bytes;
    w.col_best = (unsigned long long*)(p + off); off += align_up(8 * (size_t)(NC > 0 ? NC : 1), 256);
    w.row_best = (unsigned long long*)(p + off); off += align_up(8 * (size_t)(NR > 0 ? NR : 1), 256);
    w.lowq = (int*)(p + off); off += align_up(4 * (size_t)(NC > 0 ? NC : 1), 256);
    w.zero_gt = (int*)(p + off); off += 256;
    w.table = (AssignTable*)(p + off); off += align_up(sizeof(AssignTable), 256);
    // expected records: a few (~ln of the overlapping columns) per row; far above that the list gives way to the sweep
    w.tie_cap = (unsigned)(65536 + 64 * (size_t)(NR > 0 ? NR : 0));
    w.ties = (TieRec*)(p + off); off += align_up(sizeof(TieRec) * (size_t)w.tie_cap, 256);
    w.bytes = off;
    return w;
}

}  // namespace r3g

using namespace r3g;

static int assign_sizes(const char* who, int64_t B, const int64_t* gt_counts, int64_t A, int shared, int64_t* NR, int64_t* NCP) {
    R3G_REQUIRE(B >= 1 && B <= ASSIGN_MAX_IMAGES, "%s: 1..%d images per call", who, ASSIGN_MAX_IMAGES);
    R3G_REQUIRE(gt_counts != nullptr && A >= 0, "%s: bad arguments", who);
    int64_t nr = 0;
    for (int64_t b = 0; b < B; b++) { R3G_REQUIRE(gt_counts[b] >= 0, "%s: negative GT count", who); nr += gt_counts[b]; }
    R3G_REQUIRE(nr < (1ll << 26) && B * A < (1ll << 31), "%s: problem too large", who);
    *NR = nr; *NCP = shared ? A : B * A;
    return R3G_OK;
}

R3G_API int r3g_assign_batched_workspace_bytes(int64_t B, const int64_t* gt_counts, int64_t A, int anchors_shared, size_t* bytes) {
    int64_t NR = 0, NCP = 0;
    int rc = assign_sizes("r3g_assign_batched_workspace_bytes", B, gt_counts, A, anchors_shared, &NR, &NCP);
    if (rc != R3G_OK) return rc;
    R3G_REQUIRE(bytes != nullptr, "r3g_assign_batched_workspace_bytes: null output");
    *bytes = carve_assign(nullptr, NR, NCP, B * A).bytes;
    return R3G_OK;
}

R3G_API int r3g_assign_workspace_bytes(int64_t G, int64_t A, size_t* bytes) {
    R3G_REQUIRE(bytes != nullptr && G >= 0 && A >= 0, "r3g_assign_workspace_bytes: bad arguments");
    return r3g_assign_batched_workspace_bytes(1, &G, A, 1, bytes);
}

R3G_API int r3g_max_iou_assign_batched_f32(int64_t B, const float* gt, const int64_t* gt_counts, int64_t gt_stride,
                                           const float* anchors, int64_t A, int64_t anchor_stride, int anchors_shared,
                                           int variant, int flags, float pos_iou_thr, float neg_iou_thr, float min_pos_iou,
                                           int match_low_quality, int gt_max_assign_all,
                                           int64_t* assigned_gt_inds, float* max_overlaps, int64_t* argmax_overlaps,
                                           float* gt_max_overlaps, int64_t* gt_argmax_overlaps,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "r3g_max_iou_assign_batched_f32";
    int64_t NR = 0, NCP = 0;
    int rc = assign_sizes(who, B, gt_counts, A, anchors_shared, &NR, &NCP);
    if (rc != R3G_OK) return rc;
    rc = iou_check_common(who, NR, NCP, variant, R3G_MODE_IOU);
    if (rc != R3G_OK) return rc;
    if (A == 0) return R3G_OK;
    R3G_REQUIRE(anchors && assigned_gt_inds && workspace, "%s: null pointer", who);
    R3G_REQUIRE(NR == 0 || gt != nullptr, "%s: null gt", who);
    const int64_t NC = B * A;
    AssignWs w = carve_assign(workspace, NR, NCP, NC);
    if (workspace_bytes < w.bytes) {
        set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, w.bytes);
        return R3G_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int tpb = 256;
    AssignTable T;
    int64_t row0 = 0, item0 = 0;
    for (int b = 0; b < ASSIGN_MAX_IMAGES; b++) {
        AssignProb& pr = T.p[b];
        const int64_t m = b < B ? gt_counts[b] : 0;
        pr.row0 = (int)row0; pr.m = (int)m;
        pr.col0 = (b < B && !anchors_shared) ? (int)(b * A) : 0;
        pr.n = b < B ? (int)A : 0;
        pr.cb_shift = (b < B && anchors_shared) ? (int)(b * A) : 0;        // slot = column + cb_shift = b * A + anchor
        pr.pad = 0; pr.item0 = item0;
        if (b < B) { row0 += m; item0 += ((m + IOU_TM - 1) / IOU_TM) * ((A + IOU_TN - 1) / IOU_TN); }
    }
    T.nprob = (int)B; T.pad = 0; T.total_items = item0;
    assign_table_kernel<<<1, ASSIGN_MAX_IMAGES, 0, st>>>(T, w.table);
    const int64_t mx = NC > NR ? NC : NR;
    assign_init_kernel<<<(unsigned)((mx + tpb - 1) / tpb), tpb, 0, st>>>(w.col_best, NC, w.row_best, NR, w.lowq, w.zero_gt);
    if (NR > 0) {
        rc = r3g_iou_prepare_f32(gt, NR, gt_stride, anchors, NCP, anchor_stride, variant, workspace, w.iou.bytes, stream);
        if (rc != R3G_OK) return rc;
        IouArgs a = {};
        a.r0 = w.iou.r0; a.r1 = w.iou.r1; a.r2 = w.iou.r2; a.r2d = w.iou.r2d; a.m = (int)NR; a.c0 = w.iou.c0; a.c1 = w.iou.c1; a.n = (int)NCP;
        a.raw1 = gt; a.s1 = gt_stride; a.raw2 = anchors; a.s2 = anchor_stride; a.origin_box = gt;
        a.variant = variant; a.mode = R3G_MODE_IOU;
        a.small_mask = (variant == R3G_V3 && (flags & R3G_FLAG_SMALL_MASK)) ? 1 : 0;
        a.tau = (flags & R3G_FLAG_EMULATE_ALL) ? 1e30f : ((flags & R3G_FLAG_STRICT) ? 2e-2f : 0.0f);
        a.out = nullptr; a.stats = w.iou.stats;
        a.col_best = w.col_best; a.row_best = w.row_best; a.lowq = w.lowq; a.min_pos_iou = min_pos_iou;
        a.table = w.table;
        const bool want_ties = match_low_quality && gt_max_assign_all;
        a.ties = want_ties ? w.ties : nullptr; a.tie_cap = w.tie_cap;
        R3G_CUDA_OK(cudaMemsetAsync(w.iou.stats, 0, 256, st));
        const bool batch = !(B == 1 && anchors_shared);      // one image over shared anchors: the plain single-problem kernels
        rc = batch ? launch_sweep<false, OUT_ASSIGN_MAX, true>(a, st, item0) : launch_sweep<false, OUT_ASSIGN_MAX, false>(a, st);
        if (rc != R3G_OK) return rc;
        assign_gt_kernel<<<(unsigned)((NR + tpb - 1) / tpb), tpb, 0, st>>>(w.row_best, NR, w.table, min_pos_iou, gt_max_assign_all,
                                                                         w.lowq, w.zero_gt, gt_max_overlaps, gt_argmax_overlaps);
        if (want_ties) {
            assign_ties_kernel<<<(unsigned)((w.tie_cap + 4 * tpb - 1) / (4 * tpb)), tpb, 0, st>>>(w.ties, w.iou.stats + 5, w.tie_cap,
                                                                                            w.row_best, min_pos_iou, w.lowq);
            // list overflow (massive ties, e.g. many identical anchors): exact sweep; its warps return at once otherwise.
            // stats[5] must survive the reset of the item ticket, so only the ticket is cleared.
            R3G_CUDA_OK(cudaMemsetAsync(w.iou.stats + 4, 0, 8, st));
            rc = batch ? launch_sweep<false, OUT_ASSIGN_TIES, true>(a, st, item0) : launch_sweep<false, OUT_ASSIGN_TIES, false>(a, st);
            if (rc != R3G_OK) return rc;
        }
    }
    const dim3 fgrid((unsigned)((A + tpb - 1) / tpb), (unsigned)B);
    assign_finalize_kernel<<<fgrid, tpb, 0, st>>>(w.col_best, w.lowq, w.zero_gt, w.table, A, pos_iou_thr, neg_iou_thr,
                                                  match_low_quality, assigned_gt_inds, max_overlaps, argmax_overlaps);
    R3G_LAUNCH_OK("assign kernels");
    return R3G_OK;
}

R3G_API int r3g_max_iou_assign_f32(const float* gt, int64_t G, int64_t gt_stride,
                                   const float* anchors, int64_t A, int64_t anchor_stride,
                                   int variant, int flags, float pos_iou_thr, float neg_iou_thr, float min_pos_iou,
                                   int match_low_quality, int gt_max_assign_all,
                                   int64_t* assigned_gt_inds, float* max_overlaps, int64_t* argmax_overlaps,
                                   float* gt_max_overlaps, int64_t* gt_argmax_overlaps,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    R3G_REQUIRE(G >= 0, "r3g_max_iou_assign_f32: negative size");
    return r3g_max_iou_assign_batched_f32(1, gt, &G, gt_stride, anchors, A, anchor_stride, 1, variant, flags, pos_iou_thr, neg_iou_thr,
                                          min_pos_iou, match_low_quality, gt_max_assign_all, assigned_gt_inds, max_overlaps,
                                          argmax_overlaps, gt_max_overlaps, gt_argmax_overlaps, workspace, workspace_bytes, stream);
}
