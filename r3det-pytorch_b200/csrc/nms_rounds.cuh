// nms_rounds.cuh — greedy rotated / polygon NMS as ROUNDS of (chunk, resolve, apply) inside one persistent kernel.
//
// Replaces the K x K suppression bitmask of the reference (r3det/ops/rnms/src/rcuda/rnms_kernel.cu:229-335,
// r3det/ops/nms_rotated/src/nms_rotated_cuda.cu:71-134, r3det/ops/nms_rotated/src/poly_nms_cuda.cu:122-262: mask kernel over
// ALL pairs, D2H copy, serial host scan) and this library's own round-1 design (per-class triangular mask + scan), both of
// which evaluate every pair of a class although the rows of ~95 % of the candidates of a detector are suppressed and their
// mask rows never read.  Greedy NMS only ever needs the pairs (kept row, later column).  Per class segment, in score order:
//
//   A  select   the next <= B candidates that are still alive form the segment's CHUNK          (one warp per segment)
//   B  mask     pairs INSIDE the chunk -> chunk-local bitmask (upper triangle, 64 x 128 items)     (all warps, ticket)
//   C  resolve  greedy scan of the chunk-local mask -> the chunk's kept rows                       (one CTA per segment)
//   D  apply    kept rows of the chunk x all LATER alive candidates of the segment -> clear `alive` (all warps, ticket)
//
// and again with the survivors until every segment is exhausted.  All segments (classes, or (image, class) pairs of a
// batch) advance together; phases are separated by grid-wide barriers (the grid is sized to be co-resident).  On detector-
// like input a class of 13k candidates is done in 2-3 rounds and ~20x fewer pair evaluations than the triangular mask; when
// nothing is suppressed the rounds degenerate into the blocked triangular sweep (same pair count as before).  The workspace
// is linear in K: chunk-local masks need at most K * (B / 64 + 1) words.
//
// Pair evaluation is the three-stage pipeline of the IoU kernel (iou.cu): stage 1 = a lane owns 4 adjacent columns and
// shifts the sign bit of the expanded circumradius test (boxes; FFMA2 / FADD2) or the bounding-box overlap predicate
// (polygons) into a 32-bit mask per 8 rows; survivors are compacted into shared-memory queues and evaluated 32 at a time —
// separating axes, then the clamped-boundary area integral (geom.cuh) or the polygon clip (poly.cuh); pairs that are near the
// threshold or degenerate are decided by the reference's own arithmetic (emu.cuh).  Rows and columns of an item are staged
// in shared memory with cp.async (gathered through the chunk's position list, so no staging copies exist in global memory).
#pragma once
#include "common.cuh"
#include "emu.cuh"
#include "geom.cuh"
#include "poly.cuh"

#ifndef R3G_NMS_INITCHECK_CLEAN
#define R3G_NMS_INITCHECK_CLEAN 0
#endif

namespace r3g {
namespace rn {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int B_MAX = 2048;                 // chunk rows per segment and round: one `removed` word per lane in the scan
constexpr int TN = 128, CPL = 4, RG = 8;    // item = 64 rows x 128 columns; a lane owns 4 adjacent columns; 8 rows per mask group
constexpr int Q1CAP = 32 + RG * TN;
constexpr int Q3CAP = 64;
constexpr float SLACK = 1.0f / 262144.0f;

enum { GEOM_BOX = 0, GEOM_QUAD = 1 };

struct __attribute__((aligned(64))) Ctrl {      // one per round parity; zero at launch
    unsigned n_act;        // segments listed for this round
    unsigned n_ent;        // segments of this round with a chunk of >= 2 rows
    unsigned rows_used;    // chunk position-list slots handed out
    unsigned kept_used;    // kept-list slots handed out
    unsigned itemsB, itemsD, ticketB, ticketD;
    unsigned long long words_used;
    unsigned long long work;   // ctrl[0] only: sum over segments of min(length, B)^2 / 2 — the pair tests of a first round of full chunks
    unsigned long long pad[2];
};

struct __attribute__((aligned(16))) Entry {     // one segment's chunk in one round (64 bytes, read as four 16-byte words)
    int seg, row_base, n, nb;                   // chunk rows: spos[row_base .. row_base + n), nb = ceil(n / 64)
    unsigned wbase; int itemB0, cur_new, pe;    // mask words [wbase, wbase + n * pitch), pitch = nb rounded up to even; columns left to apply to: [cur_new, pe)
    int nk, kbase, itemD0, ncg;                 // kept rows: klist[kbase .. kbase + nk); apply items
    int cg0, pad0, pad1, pad2;                  // first 128-column group of the apply range
};

struct Args {
    // position-space candidates (read-only here; written by the gather kernel)
    const float4* p0; const float4* p1;         // boxes: BoxP0 / BoxP1; polygons: corners 0-1 / corners 2-3
    const float4* p2r;                          // boxes: row plane, two 16-byte halves per candidate; polygons: unused
    const float4* p2c;                          // boxes: column plane {X, Y, r, k}; polygons: bounding box {x0, y0, x1, y1}
    const float* raw;                           // boxes: (K, 5) tuples with class offsets applied (restatement input)
                                                // polygons: p2r[2 p] = {signed area, convex, max |coordinate|, -}
    const unsigned* label;                      // segment key per position (sorted ascending)
    int K;
    const long long* k_valid;                   // optional (device): only positions < *k_valid hold candidates (padding sorts last)
    // state
    unsigned long long* alive;                  // bit p: candidate p is valid and not suppressed so far
    int* seg_cur; int* seg_pe; int* act;        // per segment: cursor, end; act[2][K]: segment lists of this / the next round
    Entry* ent; int* spos; int* klist; int* ownerB; int* ownerD;
    unsigned long long* mask;
    int* keep_p;                                // out: 1 per kept position (zeroed by the launcher)
    Ctrl* ctrl;                                 // [2]
    unsigned* bar;                              // grid barrier counter (zero at launch)
    unsigned long long* dbg;                    // [0] stage-1 pairs [1] separating-axis tests [2] area evaluations [3] restatements
                                                // [4] rounds [5] mask items [6] apply items [8] n stamps [9..] phase time stamps (ns)
    int B, split;                               // chunk rows (<= B_MAX); row split of a mask item (1, 2, 4, 8)
    int chunk_div, chunk_min;                   // chunk = clamp(remaining / chunk_div, chunk_min, B) rows; chunk_div = 0: the kernel decides
                                                // (4 when the first round of full chunks would cost more than work_thr pair tests, else
                                                // full chunks), < 0: always full chunks
    unsigned long long work_thr;
    int variant, inclusive, prefilter;
    float thr, tau, margin;
};

struct __attribute__((aligned(16))) WarpSmem {
    float4 r2[128];                 // boxes: row plane of the item's 64 rows (2 halves each); polygons: [t] = bounding box
    float4 r0[64], r1[64];          // rows
    float4 c0[128], c1[128];        // columns
    int rpos[64], cpos[128];        // positions of the item's rows / columns
    unsigned long long sm[128];     // mask phase: 64 rows x 2 words; apply phase: 2 words
    uint4 q3[Q3CAP];                // pairs left to the restatement: {row position, column position, word index, meta}
    unsigned short q1[Q1CAP];       // stage-1 survivors (row << 7 | col), item-relative
    unsigned short q2[64];          // separating-axis survivors
};
constexpr size_t SMEM_BYTES = sizeof(WarpSmem) * WARPS;
static_assert(2 * 256 * 8 * 8 + 64 * 8 + 64 + 128 + 2 * 256 * 2 <= SMEM_BYTES, "scan buffers alias the per-warp working sets");

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __noinline__ float emu_call(const float* b1, const float* b2, int variant) {
    float x[5] = { b1[0], b1[1], b1[2], b1[3], b1[4] };
    float y[5] = { b2[0], b2[1], b2[2], b2[3], b2[4] };
    return emu::pair(x, y, variant, MODE_IOU);
}
__device__ __noinline__ float quad_call(float4 a0, float4 a1, float4 b0, float4 b1) {
    const float p[8] = { a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w };
    const float q[8] = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w };
    return poly::quad_iou(p, q);
}

// number of (row block, 128-column group) items of a chunk of nb 64-row blocks: sum over rb of ceil((nb - rb) / 2)
__host__ __device__ __forceinline__ int items_of(int nb) { const int h = nb >> 1; return (nb & 1) ? (h + 1) * (h + 1) : h * (h + 1); }

// meta word of a q3 entry: bits 0-5 bit index, 6-12 shared-memory word of the item in flight, 30 apply phase, 31 item in flight
constexpr unsigned META_INFLIGHT = 0x80000000u, META_APPLY = 0x40000000u;

template <int GEOM>
__global__ void __launch_bounds__(THREADS, 2) nms_rounds_kernel(const Args A) {
    extern __shared__ __align__(16) unsigned char rn_smem[];
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = lane_id(), lt = lanemask_lt();
    WarpSmem& W = reinterpret_cast<WarpSmem*>(rn_smem)[warp];
    const unsigned gwarp = blockIdx.x * WARPS + warp, nwarps = gridDim.x * WARPS;
    const unsigned FULL = 0xffffffffu;

    unsigned long long n_s1 = 0, n_sat = 0, n_area = 0, n_emu = 0;      // work counters of this warp (lane 0 publishes them)
    int n_stamp = 0;
    auto stamp = [&]() {                                                // CTA 0, thread 0: a time stamp per phase boundary
        if (blockIdx.x == 0 && tid == 0 && n_stamp < 54) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.dbg[9 + n_stamp++] = t;
            A.dbg[8] = (unsigned long long)n_stamp;
        }
    };
    stamp();
    // Grid barrier (the launch is cooperative: every CTA is resident).  Arrivals count on one cache line; the last arriver
    // publishes the generation on another line, which is the only one the waiting CTAs poll (read-shared in L2, so the
    // pollers do not queue behind the arrival atomics).
    unsigned gen = 0;
    auto grid_barrier = [&]() {
        __syncthreads();
        gen++;
        if (tid == 0) {
            __threadfence();
            const unsigned old = atomicAdd(A.bar, 1u);
            if (old + 1u == gen * gridDim.x) {
                asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(A.bar + 32), "r"(gen) : "memory");
            } else {
                while (ld_acquire_u32(A.bar + 32) < gen) __nanosleep(20);
            }
            __threadfence();
        }
        __syncthreads();
        stamp();
    };

    // ---- phase 0: class segments of the position space (a run start finds its end by bisection: every key occupies one run) ----
    const int KV = A.k_valid ? (int)min((long long)A.K, __ldg(A.k_valid)) : A.K;      // padding behind the candidates is never visited
    for (int p = blockIdx.x * THREADS + (int)tid; p < KV; p += gridDim.x * THREADS) {
        const unsigned L = __ldg(A.label + p);
        if (p == 0 || __ldg(A.label + p - 1) != L) {
            int lo = p, hi = KV;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(A.label + mid) == L) lo = mid; else hi = mid;
            }
            const unsigned s = atomicAdd(&A.ctrl[0].n_act, 1u);
            A.seg_cur[s] = p; A.seg_pe[s] = lo + 1; A.act[s] = (int)s;
            const unsigned long long len = (unsigned long long)min(lo + 1 - p, A.B);
            if (len >= 128) atomicAdd(&A.ctrl[0].work, len * len / 2ull);
        }
    }
    grid_barrier();

    int c1 = 0, c2 = 0, c3 = 0;             // queue fill counts (warp-uniform)
    bool pd = false;                        // apply phase (decisions clear `alive`) or mask phase (decisions set chunk-mask bits)
    unsigned wb_item = 0; int nbw = 0;      // mask phase: word index of (row 0, column block 0) of the item, words per mask row
    int cg_item = 0;                        // apply phase: 128-column group of the item

    auto drain_emu = [&](int nb) {
        __syncwarp();
        if ((int)lane < nb) {
            const uint4 e = W.q3[c3 - nb + lane];
            float r;
            if (GEOM == GEOM_BOX) r = emu_call(A.raw + (int64_t)e.x * 5, A.raw + (int64_t)e.y * 5, A.variant);
            else r = quad_call(__ldg(A.p0 + e.x), __ldg(A.p1 + e.x), __ldg(A.p0 + e.y), __ldg(A.p1 + e.y));   // the reference's own arithmetic
            const bool sup = A.inclusive ? (r >= A.thr) : (r > A.thr);
            if (sup) {
                const unsigned long long bit = 1ull << (e.w & 63u);
                if (e.w & META_INFLIGHT) atomicOr(&W.sm[(e.w >> 6) & 127u], bit);
                else if (e.w & META_APPLY) atomicAnd(A.alive + e.z, ~bit);
                else atomicOr(A.mask + e.z, bit);
            }
        }
        __syncwarp();
        c3 -= nb;
        n_emu += nb;
    };
    // stage 3: the pairs queued in `q` (item-relative row << 7 | col) -> decision
    auto drain_area = [&](unsigned short* q, int& cq, int nb) {
        __syncwarp();
        bool emu = false;
        unsigned e = 0;
        if ((int)lane < nb) {
            e = q[cq - nb + lane];
            const int il = (int)(e >> 7), jl = (int)(e & 127u);
            float r;
            if (GEOM == GEOM_BOX) {
                const float4 a0 = W.r0[il], a1 = W.r1[il], b0 = W.c0[jl], b1 = W.c1[jl];
                const BoxP0 A0 = { a0.x, a0.y, a0.z, a0.w }; const BoxP1 A1 = { a1.x, a1.y, a1.z, a1.w };
                const BoxP0 B0 = { b0.x, b0.y, b0.z, b0.w }; const BoxP1 B1 = { b1.x, b1.y, b1.z, b1.w };
                bool risk;
                r = pair_overlap(A0, A1, B0, B1, A.variant, MODE_IOU, A.tau, risk, true);       // survivors of the separating-axis stage
                emu = A.tau > 0.0f && (risk || fabsf(r - A.thr) < A.margin);
            } else {
                // own convex-quadrilateral geometry; concave input and pairs within the reference's noise of the threshold are
                // left to the restatement
                float margin;
                const bool ok = poly::convex_quad_iou(W.r0[il], W.r1[il], __ldg(A.p2r + 2 * (size_t)W.rpos[il]),
                                                      W.c0[jl], W.c1[jl], __ldg(A.p2r + 2 * (size_t)W.cpos[jl]), r, margin);
                emu = !ok || fabsf(r - A.thr) < margin;
            }
            if (!emu) {
                const bool sup = A.inclusive ? (r >= A.thr) : (r > A.thr);
                if (sup) atomicOr(&W.sm[pd ? (jl >> 6) : (il * 2 + (jl >> 6))], 1ull << (jl & 63));
            }
        }
        __syncwarp();
        cq -= nb;
        n_area += nb;
        {
            const unsigned bal = __ballot_sync(FULL, emu);
            if (bal) {
                if (emu) {
                    const int il = (int)(e >> 7), jl = (int)(e & 127u);
                    uint4 t;
                    t.x = (unsigned)W.rpos[il]; t.y = (unsigned)W.cpos[jl];
                    if (pd) { t.z = (unsigned)(cg_item * 2 + (jl >> 6)); t.w = (unsigned)(jl & 63) | META_APPLY; }
                    else {
                        t.z = wb_item + (unsigned)(il * nbw + (jl >> 6));
                        t.w = (unsigned)(jl & 63) | ((unsigned)(il * 2 + (jl >> 6)) << 6) | META_INFLIGHT;
                    }
                    W.q3[c3 + __popc(bal & lt)] = t;
                }
                c3 += __popc(bal);
                if (c3 > Q3CAP - 32) drain_emu(32);
            }
        }
    };
    auto drain_sat = [&](int nb) {
        __syncwarp();
        bool ok = false;
        unsigned e = 0;
        if ((int)lane < nb) {
            e = W.q1[c1 - nb + lane];
            const int il = (int)(e >> 7), jl = (int)(e & 127u);
            const float4 a0 = W.r0[il], a1 = W.r1[il], b0 = W.c0[jl], b1 = W.c1[jl];
            const BoxP0 A0 = { a0.x, a0.y, a0.z, a0.w }; const BoxP1 A1 = { a1.x, a1.y, a1.z, a1.w };
            const BoxP0 B0 = { b0.x, b0.y, b0.z, b0.w }; const BoxP1 B1 = { b1.x, b1.y, b1.z, b1.w };
            ok = pair_sat(A0, A1, B0, B1);
        }
        __syncwarp();
        c1 -= nb;
        n_sat += nb;
        const unsigned bal = __ballot_sync(FULL, ok);
        if (ok) W.q2[c2 + __popc(bal & lt)] = (unsigned short)e;
        c2 += __popc(bal);
        if (c2 >= 32) drain_area(W.q2, c2, 32);
    };
    auto drain_stage1 = [&](int nb) {                          // boxes: separating axes next; polygons: straight to the clip
        if (GEOM == GEOM_BOX) drain_sat(nb); else drain_area(W.q1, c1, nb);
    };

    // Stage 1 + queues over rows [r_lo, r_hi) of the item whose rows / columns are staged in W; `cx..ck` are this lane's
    // four columns (boxes: X, Y, r, k of the expanded circumradius test; polygons: bounding box), `diag`: only pairs with
    // column index > row index count (the diagonal block of a chunk).
    auto sweep = [&](int r_lo, int r_hi, bool diag, const float (&cx)[CPL], const float (&cy)[CPL], const float (&cr)[CPL],
                     const float (&ck)[CPL]) {
        unsigned long long CX[CPL / 2], CY[CPL / 2], CR[CPL / 2], CK[CPL / 2];
        if (GEOM == GEOM_BOX) {
#pragma unroll
            for (int h = 0; h < CPL / 2; h++) {
                CX[h] = pk2(cx[2 * h], cx[2 * h + 1]); CY[h] = pk2(cy[2 * h], cy[2 * h + 1]);
                CR[h] = pk2(cr[2 * h], cr[2 * h + 1]); CK[h] = pk2(ck[2 * h], ck[2 * h + 1]);
            }
        }
        n_s1 += (unsigned long long)(r_hi - r_lo) * TN;
        for (int ig = r_lo; ig < r_hi; ig += RG) {
            const int nr = min(RG, r_hi - ig);
            unsigned m = 0;
#pragma unroll
            for (int r = 0; r < RG; r++) {
                if (r < nr) {
                    if (GEOM == GEOM_BOX) {
                        const ulonglong2 ra = *reinterpret_cast<const ulonglong2*>(&W.r2[2 * (ig + r)]);         // {-2X, -2X}, {-2Y, -2Y}
                        const ulonglong2 rb = *reinterpret_cast<const ulonglong2*>(&W.r2[2 * (ig + r) + 1]);     // {-2r, -2r}, {k, k}
#pragma unroll
                        for (int h = 0; h < CPL / 2; h++) {
                            unsigned long long t = fadd2(CK[h], rb.y);
                            t = ffma2(ra.x, CX[h], t);
                            t = ffma2(ra.y, CY[h], t);
                            t = ffma2(rb.x, CR[h], t);
                            m = __funnelshift_l((unsigned)t, m, 1);
                            m = __funnelshift_l((unsigned)(t >> 32), m, 1);
                        }
                    } else {
                        const float4 rbx = W.r2[ig + r];                                                           // {x0, y0, x1, y1}
#pragma unroll
                        for (int k = 0; k < CPL; k++) {
                            // column box {cx, cy, cr, ck} = {x0, y0, x1, y1}; an invalid column has x0 = +inf
                            const bool hit = A.prefilter ? !(cx[k] > rbx.z || cr[k] < rbx.x || cy[k] > rbx.w || ck[k] < rbx.y)
                                                         : (cx[k] < 3.0e38f);
                            m = (m << 1) | (hit ? 1u : 0u);
                        }
                    }
                }
            }
            if (diag && m != 0) {
                unsigned allow = 0;
                for (int r = 0; r < nr; r++) {
#pragma unroll
                    for (int k = 0; k < CPL; k++) allow = (allow << 1) | (((int)lane * CPL + k > ig + r) ? 1u : 0u);
                }
                m &= allow;
            }
            const int cnt = __popc(m);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, d);
                if ((int)lane >= d) incl += t;
            }
            const int tot = __shfl_sync(FULL, incl, 31);
            if (tot) {
                int pos = c1 + incl - cnt;
                const int nbits = nr * CPL;
                while (m) {
                    const int b = 31 - __clz(m);
                    m ^= 1u << b;
                    const unsigned idx = (unsigned)(nbits - 1 - b);
                    W.q1[pos++] = (unsigned short)((((unsigned)ig + (idx >> 2)) << 7) | (lane * CPL + (idx & 3u)));
                }
                c1 += tot;
                while (c1 >= 32) drain_stage1(32);
            }
        }
        if (c1 > 0) drain_stage1(c1);
        if (GEOM == GEOM_BOX && c2 > 0) drain_area(W.q2, c2, c2);
        __syncwarp();
    };

    const int Bmax = A.B;
    // chunk policy.  Full chunks need the fewest rounds (each costs four grid barriers and a serial resolve); when the in-chunk
    // triangles alone are a throughput problem (many long segments: batches), chunks of a quarter of what is left of a segment
    // test ~5x fewer pairs — rows that an earlier chunk's kept rows suppress are never paired with each other — for 1-2 rounds more
    const int chunk_div = A.chunk_div > 0 ? A.chunk_div : ((A.chunk_div == 0 && __ldcg(&A.ctrl[0].work) > A.work_thr) ? 4 : 0);
    int par = 0;
    while (true) {
        Ctrl* C = A.ctrl + par;
        Ctrl* Cn = A.ctrl + (par ^ 1);
        const unsigned n_act = __ldcg(&C->n_act);
        if (n_act == 0) break;
        const int* act = A.act + (size_t)par * A.K;
        int* actn = A.act + (size_t)(par ^ 1) * A.K;

        // ---- phase A: every active segment selects its chunk = the next <= B alive positions from its cursor ----
        // segments are dealt to CTAs first (a segment's position list is 2048 scattered-then-coalesced stores: spread the LSU work)
        for (unsigned i = blockIdx.x + gridDim.x * warp; i < n_act; i += nwarps) {
            const int s = __ldcg(act + i);
            const int cur = __ldcg(A.seg_cur + s), pe = __ldcg(A.seg_pe + s);
            // chunk size of this segment and round: a fraction of what is left of the segment (small chunks waste fewer pair
            // tests on rows that an earlier row of the same chunk suppresses; each round costs four grid barriers)
            int B = Bmax;
            if (chunk_div > 0 && pe - cur > A.chunk_min) B = min(Bmax, max(A.chunk_min, ((pe - cur) / chunk_div + 63) & ~63));
            const int cap = min(B, pe - cur);
            int row_base = 0;
            if (lane == 0) row_base = (int)atomicAdd(&C->rows_used, (unsigned)cap);
            int n = 0, cur_new = pe, myfirst = -1;
            const int wl = (pe - 1) >> 6, wf = cur >> 6;
            // a lane reads two adjacent alive words per step (one 16-byte load); the next step's words are in flight while
            // this step's bits are expanded
            auto load2 = [&](int w0) {
                const int w = w0 + 2 * (int)lane;
                ulonglong2 v = make_ulonglong2(0ull, 0ull);
                if (w <= wl) {
                    v = __ldcg(reinterpret_cast<const ulonglong2*>(A.alive + w));
                    if (w < wf) v.x = 0ull;
                    if (w == wf) v.x &= ~0ull << (cur & 63);
                    if (w + 1 == wf) v.y &= ~0ull << (cur & 63);
                    if (w == wl && (pe & 63)) v.x &= (1ull << (pe & 63)) - 1ull;
                    if (w + 1 == wl && (pe & 63)) v.y &= (1ull << (pe & 63)) - 1ull;
                    if (w + 1 > wl) v.y = 0ull;
                }
                return v;
            };
            const int wstart = wf & ~1;
            ulonglong2 nxt = load2(wstart);
            row_base = __shfl_sync(FULL, row_base, 0);
            // segments are disjoint runs of the position order (every key occupies ONE run: nms.cu, seg_key_of), so the caps sum
            // to <= K.  Malformed keys (a caller whose labels exceed its R3G_NMS_LABEL_BITS promise) could break that: such a
            // segment is dropped rather than allowed to write beyond the position list
            if ((long long)row_base + cap > (long long)A.K) continue;
            for (int w0 = wstart; w0 <= wl; w0 += 64) {
                const ulonglong2 v = nxt;
                if (w0 + 64 <= wl) nxt = load2(w0 + 64);
                const int w = w0 + 2 * (int)lane;
                const int c0 = __popcll(v.x), cnt = c0 + __popcll(v.y);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(FULL, incl, d);
                    if ((int)lane >= d) incl += t;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                const int base = n + incl - cnt;
                int mylast = -1;                                   // position of the row that fills slot B - 1 (if this lane writes it)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    unsigned long long bits = h ? v.y : v.x;
                    int slot = base + (h ? c0 : 0);
                    const int p0w = (w + h) * 64;
                    while (bits && slot < B) {
                        const int t = __ffsll((long long)bits) - 1;
                        bits &= bits - 1ull;
                        A.spos[row_base + slot] = p0w + t;
                        if (slot == 0) myfirst = p0w + t;
                        if (slot == B - 1) mylast = p0w + t;
                        slot++;
                    }
                }
                if (n + total >= B) {                              // chunk full: the cursor moves behind its last row
                    const unsigned who = __ballot_sync(FULL, mylast >= 0);
                    cur_new = __shfl_sync(FULL, mylast, __ffs((int)who) - 1) + 1;
                    n = B;
                    break;
                }
                n += total;
            }
            row_base = __shfl_sync(FULL, row_base, 0);            // (idempotent: every lane holds it after the first pass)
            __syncwarp();
            if (n == 1) {
                if (myfirst >= 0) A.keep_p[myfirst] = 1;                          // alone in its chunk and nothing behind it
            } else if (n >= 2) {
                const int nb = (n + 63) >> 6;
                const int items = items_of(nb) * A.split;
                // the three allocations travel together: lanes 0..2 issue one atomic each
                unsigned long long got = 0ull;
                if (lane == 0) got = atomicAdd(&C->n_ent, 1u);
                else if (lane == 1) got = atomicAdd(&C->words_used, (unsigned long long)n * (unsigned long long)((nb + 1) & ~1));   // even row pitch
                else if (lane == 2) got = atomicAdd(&C->itemsB, (unsigned)items);
                const int e = (int)__shfl_sync(FULL, got, 0);
                const unsigned wbase = (unsigned)__shfl_sync(FULL, got, 1);
                const int itemB0 = (int)__shfl_sync(FULL, got, 2);
                if (lane == 0) {
                    int4* ep = reinterpret_cast<int4*>(A.ent + e);
                    ep[0] = make_int4(s, row_base, n, nb);
                    ep[1] = make_int4((int)wbase, itemB0, cur_new, pe);
                    ep[2] = make_int4(0, 0, 0, 0);
                    ep[3] = make_int4(0, 0, 0, 0);
                }
                for (int k = lane; k < items; k += 32) A.ownerB[itemB0 + k] = e;
            }
            if (cur_new < pe && lane == 0) {
                A.seg_cur[s] = cur_new;
                actn[atomicAdd(&Cn->n_act, 1u)] = s;
            }
        }
        grid_barrier();

        // ---- phase B: chunk-local suppression masks ----
        if (blockIdx.x == 0 && tid == 0) {       // reset what the NEXT round allocates from (nobody touches it until then)
            Cn->n_ent = 0; Cn->rows_used = 0; Cn->kept_used = 0; Cn->itemsB = 0; Cn->itemsD = 0; Cn->ticketB = 0; Cn->ticketD = 0;
            Cn->words_used = 0ull;
            C->n_act = 0;                        // every CTA has read it; the round after next appends to it
        }
        pd = false;
        {
            const unsigned total = __ldcg(&C->itemsB);
            const int split = A.split, rps = 64 / split;
            while (true) {
                unsigned item = 0;
                if (lane == 0) item = atomicAdd(&C->ticketB, 1u);
                item = __shfl_sync(FULL, item, 0);
                if (item >= total) break;
                const int e = __ldcg(A.ownerB + item);
                const int4 e0 = __ldcg(reinterpret_cast<const int4*>(A.ent + e));
                const int4 e1 = __ldcg(reinterpret_cast<const int4*>(A.ent + e) + 1);
                const int row_base = e0.y, n = e0.z, nb = e0.w;
                int li = (int)item - e1.y;
                const int sub = li % split; li /= split;
                int rb = 0;
                while (true) { const int c = (nb - rb + 1) >> 1; if (li < c) break; li -= c; rb++; }
                const int cb0 = rb + 2 * li, ncb = min(2, nb - cb0);
                const int rows = min(64, n - rb * 64);                        // valid rows of the block
                const int r_lo = sub * rps, r_hi = min(rows, r_lo + rps);
                nbw = (nb + 1) & ~1; wb_item = (unsigned)e1.x + (unsigned)(rb * 64) * (unsigned)nbw + (unsigned)cb0;
                // stage rows and columns (gathered through the chunk's position list)
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int t = (int)lane + 32 * h;
                    const int p = __ldcg(A.spos + row_base + rb * 64 + min(t, rows - 1));
                    W.rpos[t] = p;
                    cp16(&W.r0[t], A.p0 + p); cp16(&W.r1[t], A.p1 + p);
                    if (GEOM == GEOM_BOX) { cp16(&W.r2[2 * t], A.p2r + 2 * (size_t)p); cp16(&W.r2[2 * t + 1], A.p2r + 2 * (size_t)p + 1); }
                    else cp16(&W.r2[t], A.p2c + p);
                }
                float cx[CPL], cy[CPL], cr[CPL], ck[CPL];
                const int ncols = min(128, n - cb0 * 64);
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int t = (int)lane * CPL + k;
                    const int p = __ldcg(A.spos + row_base + cb0 * 64 + min(t, ncols - 1));
                    W.cpos[t] = p;
                    cp16(&W.c0[t], A.p0 + p); cp16(&W.c1[t], A.p1 + p);
                    if (t < ncols) {
                        const float4 b = __ldg(A.p2c + p);
                        cx[k] = b.x; cy[k] = b.y; cr[k] = b.z; ck[k] = b.w;
                    } else {
                        cx[k] = (GEOM == GEOM_BOX) ? 0.0f : 3.0e38f; cy[k] = (GEOM == GEOM_BOX) ? 0.0f : 3.0e38f;
                        cr[k] = (GEOM == GEOM_BOX) ? 0.0f : -3.0e38f; ck[k] = (GEOM == GEOM_BOX) ? 3.0e38f : -3.0e38f;
                    }
                }
                for (int k = lane; k < 128; k += 32) W.sm[k] = 0ull;
                cp_wait();
                __syncwarp();
                if (r_lo < r_hi) sweep(r_lo, r_hi, cb0 == rb, cx, cy, cr, ck);
                // pairs of this item still waiting for the restatement now decide into the global words written below
                for (int k = lane; k < c3; k += 32) W.q3[k].w &= ~META_INFLIGHT;
                __syncwarp();
                for (int k = r_lo * 2 + (int)lane; k < r_hi * 2; k += 32) {
                    const int il = k >> 1, c = k & 1;
                    if (c < ncb) A.mask[wb_item + (unsigned)(il * nbw + c)] = W.sm[k];
                }
                __syncwarp();
                while (c3 >= 32) drain_emu(32);
            }
            if (c3 > 0) drain_emu(c3);
        }
        grid_barrier();

        // ---- phase C: greedy scan of each chunk-local mask (one CTA per chunk) ----
        // Software-pipelined in SUPERBLOCKS of 4 x 64 rows, so that one barrier and one global-memory round trip are paid per
        // 256 rows and neither sits on the serial chain.  For superblock S (column window = its 4 blocks + the next 4):
        //   warp 0     : for each of the 4 blocks, chain over the 64 diagonal words (shared memory) with ffs jumps -> kept rows;
        //                lanes 0..7 OR the kept rows' window words into removed[] for the blocks that follow inside the window;
        //   warps 1..7 : prefetch the 256 x 8 window words of superblock S+1 into the other buffer and OR the kept rows of
        //                superblock S-1 into removed[] for the columns BEYOND its window (global mask words).
        {
            constexpr int SB = 4, WIN = 8, SBROWS = 64 * SB;
            typedef unsigned long long WinRow[WIN];
            WinRow* Wd = reinterpret_cast<WinRow*>(rn_smem);                                      // [2][SBROWS]
            unsigned long long* remv = reinterpret_cast<unsigned long long*>(rn_smem + 2 * SBROWS * WIN * 8);   // [32]
            unsigned long long* keptw = remv + 32;                                                // [32]
            int* misc = reinterpret_cast<int*>(keptw + 32);                                       // [16]: 0..2 results, 4..5 kept counts
            int* kpre = misc + 16;                                                                // [32] kept rows before each block
            unsigned short* ksb = reinterpret_cast<unsigned short*>(kpre + 32);                   // [2][SBROWS] kept rows of a superblock
            const unsigned n_ent = __ldcg(&C->n_ent);
            for (unsigned e = blockIdx.x; e < n_ent; e += gridDim.x) {
                const int4 e0 = __ldcg(reinterpret_cast<const int4*>(A.ent + e));
                const int4 e1 = __ldcg(reinterpret_cast<const int4*>(A.ent + e) + 1);
                const int row_base = e0.y, n = e0.z, nb = e0.w, cur_new = e1.z, pe = e1.w;
                const unsigned wbase = (unsigned)e1.x;
                const int nsb = (nb + SB - 1) / SB, np = (nb + 1) & ~1;          // np: words per mask row (even: 16-byte loads)
                constexpr int PAIRS2 = SBROWS * WIN / 2;
                constexpr int PER = (PAIRS2 + (THREADS - 32) - 1) / (THREADS - 32);
                // window of superblock S: every row's 8 words are 4 aligned 16-byte loads; all loads of a thread are issued
                // before anything is stored.  Words left of the diagonal block or right of the chunk were never written: zero.
                auto pf_load = [&](int S, int u, int nt, ulonglong2 (&v)[PER]) {
#pragma unroll
                    for (int j = 0; j < PER; j++) {
                        const int idx = u + j * nt;
                        ulonglong2 x = make_ulonglong2(0ull, 0ull);
                        if (idx < PAIRS2) {
                            const int r = idx >> 2, k2 = idx & 3, bb = S * SB + (r >> 6), col = S * SB + 2 * k2, lr = bb * 64 + (r & 63);
                            if (bb < nb && lr < n && col + 1 >= bb && col < nb) {
#if R3G_NMS_INITCHECK_CLEAN
                                // (diagnostic build for compute-sanitizer --tool initcheck: never touch the unwritten word of a pair)
                                if (col >= bb) x.x = __ldcg(A.mask + wbase + (unsigned)(lr * np + col));
                                if (col + 1 < nb) x.y = __ldcg(A.mask + wbase + (unsigned)(lr * np + col + 1));
#else
                                x = __ldcg(reinterpret_cast<const ulonglong2*>(A.mask + wbase + (unsigned)(lr * np + col)));
                                if (col < bb) x.x = 0ull;
                                if (col + 1 >= nb) x.y = 0ull;
#endif
                            }
                        }
                        v[j] = x;
                    }
                };
                auto pf_store = [&](int bufi, int u, int nt, const ulonglong2 (&v)[PER]) {
#pragma unroll
                    for (int j = 0; j < PER; j++) {
                        const int idx = u + j * nt;
                        if (idx < PAIRS2) *reinterpret_cast<ulonglong2*>(&Wd[bufi * SBROWS + (idx >> 2)][2 * (idx & 3)]) = v[j];
                    }
                };
                __syncthreads();                                                   // the previous chunk is done with shared memory
                if (tid < 64) remv[tid] = 0ull;                                    // removed[] and keptw[]
                {
                    ulonglong2 v[PER];
                    pf_load(0, (int)tid, THREADS, v);
                    pf_store(0, (int)tid, THREADS, v);
                }
                for (int S = 0; S < nsb; S++) {
                    const int bufi = S & 1;
                    __syncthreads();
                    if (warp == 0) {
                        int nks = 0;                                               // kept rows of this superblock so far
                        for (int q = 0; q < SB; q++) {
                            const int bb = S * SB + q;
                            if (bb >= nb) break;
                            const int rows = min(64, n - bb * 64);
                            const unsigned long long vb = (rows == 64) ? ~0ull : ((1ull << rows) - 1ull);
                            const unsigned long long cur0 = remv[bb];
                            unsigned long long acc = 0ull;
                            // the serial chain runs on 32-bit halves (ffs, shift and mask are single instructions there):
                            // rows 0..31 first — their diagonal words also remove rows 32..63 — then rows 32..63
                            const unsigned vlo = (unsigned)vb, vhi = (unsigned)(vb >> 32);
                            unsigned cur_lo = (unsigned)cur0, cur_hi = (unsigned)(cur0 >> 32), kept_lo = 0u, kept_hi = 0u;
                            const WinRow* wq = Wd + bufi * SBROWS + q * 64;
                            unsigned avail = vlo & ~cur_lo;
                            while (avail) {                                        // warp-uniform; one kept row per trip
                                const int t = __ffs((int)avail) - 1;
                                kept_lo |= 1u << t;
                                const unsigned long long d = wq[t][q];
                                acc |= wq[t][lane & (WIN - 1)];                    // every lane loads (no divergence on the chain)
                                cur_lo |= (unsigned)d; cur_hi |= (unsigned)(d >> 32);
                                avail = vlo & ~cur_lo & (0xfffffffeu << t);
                            }
                            avail = vhi & ~cur_hi;
                            while (avail) {
                                const int t = __ffs((int)avail) - 1;
                                kept_hi |= 1u << t;
                                const unsigned long long d = wq[32 + t][q];
                                acc |= wq[32 + t][lane & (WIN - 1)];
                                cur_hi |= (unsigned)(d >> 32);
                                avail = vhi & ~cur_hi & (0xfffffffeu << t);
                            }
                            const unsigned long long kept = ((unsigned long long)kept_hi << 32) | kept_lo;
                            if ((int)lane > q && lane < WIN && S * SB + (int)lane < nb && acc) atomicOr(&remv[S * SB + lane], acc);
                            if (lane == 0) keptw[bb] = kept;
                            // the block's kept rows, compacted (chunk-local row index), for the warps that apply them next step
                            const unsigned klo = (unsigned)kept, khi = (unsigned)(kept >> 32);
                            if ((klo >> lane) & 1u) ksb[bufi * SBROWS + nks + __popc(klo & lt)] = (unsigned short)(bb * 64 + (int)lane);
                            if ((khi >> lane) & 1u) ksb[bufi * SBROWS + nks + __popc(klo) + __popc(khi & lt)] = (unsigned short)(bb * 64 + 32 + (int)lane);
                            nks += __popcll(kept);
                            __syncwarp();
                        }
                        if (lane == 0) misc[4 + bufi] = nks;
                    } else {
                        const int u = (int)tid - 32, nt = THREADS - 32;
                        ulonglong2 pv[PER];
                        if (S + 1 < nsb) pf_load(S + 1, u, nt, pv);               // in flight together with the loads below
                        // kept rows of superblock S-1 x the columns beyond its window: every (row, column) word is one load,
                        // a thread owns one column and every `groups`-th kept row, all its loads in flight together
                        const int first = (S - 1) * SB + WIN;
                        const int ncb = nb - first;
                        if (S > 0 && ncb > 0) {
                            const int nkp = misc[4 + (bufi ^ 1)];
                            const int groups = nt / ncb;                           // ncb <= 24: every column gets >= 9 threads
                            const int g = u / ncb, c = first + (u - g * ncb);
                            if (g < groups) {
                                unsigned long long acc = 0ull;
                                for (int j0 = g; j0 < nkp; j0 += 8 * groups) {
                                    unsigned long long v[8];
#pragma unroll
                                    for (int j = 0; j < 8; j++) {
                                        const int jj = j0 + j * groups;
                                        v[j] = 0ull;
                                        if (jj < nkp) {
                                            const int lr = ksb[(bufi ^ 1) * SBROWS + jj];
                                            v[j] = __ldcg(A.mask + wbase + (unsigned)(lr * np + c));
                                        }
                                    }
#pragma unroll
                                    for (int j = 0; j < 8; j++) acc |= v[j];
                                }
                                if (acc) atomicOr(&remv[c], acc);
                            }
                        }
                        if (S + 1 < nsb) pf_store(bufi ^ 1, u, nt, pv);
                    }
                }
                __syncthreads();
                if (warp == 0) {                                                   // kept counts: prefix over the blocks, allocations
                    const int cntb = ((int)lane < nb) ? __popcll(keptw[lane]) : 0;
                    int incl = cntb;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(FULL, incl, d);
                        if ((int)lane >= d) incl += t;
                    }
                    kpre[lane] = incl - cntb;
                    const int nk = __shfl_sync(FULL, incl, 31);
                    int itemsD = 0, ncg = 0, cg0 = 0;
                    if (cur_new < pe && nk > 0) {
                        cg0 = cur_new >> 7;
                        ncg = ((pe - 1) >> 7) - cg0 + 1;
                        itemsD = ((nk + 63) >> 6) * ncg;
                    }
                    unsigned got = 0u;                                             // the two allocations travel together
                    if (lane == 0) got = atomicAdd(&C->kept_used, (unsigned)nk);
                    else if (lane == 1 && itemsD > 0) got = atomicAdd(&C->itemsD, (unsigned)itemsD);
                    const int kbase = (int)__shfl_sync(FULL, got, 0), itemD0 = (int)__shfl_sync(FULL, got, 1);
                    if (lane == 0) {
                        int4* ep = reinterpret_cast<int4*>(A.ent + e);
                        ep[2] = make_int4(nk, kbase, itemD0, ncg);
                        ep[3] = make_int4(cg0, 0, 0, 0);
                        misc[0] = kbase; misc[1] = itemD0; misc[2] = itemsD;
                    }
                }
                __syncthreads();
                const int kbase = misc[0], itemD0 = misc[1], itemsD = misc[2];
                {
                    constexpr int RPT = B_MAX / THREADS;                           // rows per thread: all position loads in flight together
                    int pp[RPT];
#pragma unroll
                    for (int i = 0; i < RPT; i++) {
                        const int lr = (int)tid + i * THREADS;
                        pp[i] = (lr < n) ? __ldcg(A.spos + row_base + lr) : 0;
                    }
#pragma unroll
                    for (int i = 0; i < RPT; i++) {
                        const int lr = (int)tid + i * THREADS;
                        if (lr < n) {
                            const int bb = lr >> 6, t = lr & 63;
                            const unsigned long long kw = keptw[bb];
                            if ((kw >> t) & 1ull) {
                                A.klist[kbase + kpre[bb] + __popcll(kw & ((1ull << t) - 1ull))] = pp[i];
                                A.keep_p[pp[i]] = 1;
                            }
                        }
                    }
                }
                for (int k = tid; k < itemsD; k += THREADS) A.ownerD[itemD0 + k] = (int)e;
            }
            __syncthreads();
        }
        grid_barrier();

        // ---- phase D: the chunk's kept rows x the later alive candidates of the segment ----
        // (the common small call — every segment fits one chunk — has nothing to apply and no next round: every CTA reads the
        //  same two counters here and leaves without the apply phase's barrier)
        if (__ldcg(&C->itemsD) == 0u && __ldcg(&Cn->n_act) == 0u) {
            if (blockIdx.x == 0 && tid == 0) { A.dbg[4] += 1ull; A.dbg[5] += __ldcg(&C->itemsB); stamp(); }
            break;
        }
        pd = true;
        {
            const unsigned total = __ldcg(&C->itemsD);
            while (true) {
                unsigned item = 0;
                if (lane == 0) item = atomicAdd(&C->ticketD, 1u);
                item = __shfl_sync(FULL, item, 0);
                if (item >= total) break;
                const int e = __ldcg(A.ownerD + item);
                const int4 e1 = __ldcg(reinterpret_cast<const int4*>(A.ent + e) + 1);
                const int4 e2 = __ldcg(reinterpret_cast<const int4*>(A.ent + e) + 2);
                const int4 e3 = __ldcg(reinterpret_cast<const int4*>(A.ent + e) + 3);
                const int cur_new = e1.z, pe = e1.w, nk = e2.x, kbase = e2.y, ncg = e2.w;
                const int li = (int)item - e2.z;
                const int kg = li / ncg, cg = e3.x + (li - kg * ncg);
                cg_item = cg;
                const int rows = min(64, nk - kg * 64);
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int t = (int)lane + 32 * h;
                    const int p = __ldcg(A.klist + kbase + kg * 64 + min(t, rows - 1));
                    W.rpos[t] = p;
                    cp16(&W.r0[t], A.p0 + p); cp16(&W.r1[t], A.p1 + p);
                    if (GEOM == GEOM_BOX) { cp16(&W.r2[2 * t], A.p2r + 2 * (size_t)p); cp16(&W.r2[2 * t + 1], A.p2r + 2 * (size_t)p + 1); }
                    else cp16(&W.r2[t], A.p2c + p);
                }
                const unsigned long long aw = __ldcg(A.alive + (size_t)cg * 2 + (lane >> 4));
                float cx[CPL], cy[CPL], cr[CPL], ck[CPL];
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int t = (int)lane * CPL + k;
                    const int p = cg * 128 + t;
                    const bool ok = p >= cur_new && p < pe && ((aw >> ((lane & 15u) * 4 + k)) & 1ull);
                    const int pc = min(p, A.K - 1);
                    W.cpos[t] = pc;
                    cp16(&W.c0[t], A.p0 + pc); cp16(&W.c1[t], A.p1 + pc);
                    if (ok) {
                        const float4 b = __ldg(A.p2c + pc);
                        cx[k] = b.x; cy[k] = b.y; cr[k] = b.z; ck[k] = b.w;
                    } else {
                        cx[k] = (GEOM == GEOM_BOX) ? 0.0f : 3.0e38f; cy[k] = (GEOM == GEOM_BOX) ? 0.0f : 3.0e38f;
                        cr[k] = (GEOM == GEOM_BOX) ? 0.0f : -3.0e38f; ck[k] = (GEOM == GEOM_BOX) ? 3.0e38f : -3.0e38f;
                    }
                }
                if (lane < 2) W.sm[lane] = 0ull;
                cp_wait();
                __syncwarp();
                sweep(0, rows, false, cx, cy, cr, ck);
                if (lane < 2) {
                    const unsigned long long kill = W.sm[lane];
                    if (kill) atomicAnd(A.alive + (size_t)cg * 2 + lane, ~kill);
                }
                __syncwarp();
                while (c3 >= 32) drain_emu(32);
            }
            if (c3 > 0) drain_emu(c3);
        }
        grid_barrier();
        par ^= 1;
        if (blockIdx.x == 0 && tid == 0) { A.dbg[4] += 1ull; A.dbg[5] += __ldcg(&C->itemsB); A.dbg[6] += __ldcg(&C->itemsD); }
    }
    if (lane == 0) {
        if (n_s1) atomicAdd(A.dbg + 0, n_s1);
        if (n_sat) atomicAdd(A.dbg + 1, n_sat);
        if (n_area) atomicAdd(A.dbg + 2, n_area);
        if (n_emu) atomicAdd(A.dbg + 3, n_emu);
    }
}

}  // namespace rn
}  // namespace r3g
