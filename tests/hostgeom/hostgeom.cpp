// Test-only host build of the PRODUCT geometry headers (csrc/geom.cuh, csrc/emu.cuh) so that the
// fast path can be swept over 10^7 pairs against the oracle on a machine without a GPU.
// Mirrors what the IoU / NMS kernels do per pair: prepare -> circle reject -> SAT -> clamped-boundary
// integral -> (v1: flagged pairs re-evaluated by the restatement).
#include <cstdint>
#include <vector>
#include "geom.cuh"
#include "emu.cuh"

using namespace r3g;

extern "C" __attribute__((visibility("default")))
void hg_iou_matrix(const float* b1, int64_t m, const float* b2, int64_t n, int variant, int mode,
                   float tau, int refit, float* out, int64_t* n_circle_pass, int64_t* n_sat_pass, int64_t* n_emu) {
    std::vector<BoxP0> a0(m), c0(n);
    std::vector<BoxP1> a1(m), c1(n);
    for (int64_t i = 0; i < m; i++) refit ? emu::prep_box_strict(b1 + 5 * i, variant, a0[i], a1[i]) : prep_box(b1 + 5 * i, variant, a0[i], a1[i]);
    for (int64_t j = 0; j < n; j++) refit ? emu::prep_box_strict(b2 + 5 * j, variant, c0[j], c1[j]) : prep_box(b2 + 5 * j, variant, c0[j], c1[j]);
    int64_t nc = 0, ns = 0, ne = 0;
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++) {
            float r = 0.0f;
            if (!circle_reject(a0[i], c0[j])) {
                nc++;
                bool risk;
                PairFrame f;
                if (pair_frame(a0[i], a1[i], c0[j], c1[j], f)) ns++;
                r = pair_overlap(a0[i], a1[i], c0[j], c1[j], variant, mode, tau, risk);
                if (risk) { ne++; r = emu::pair(b1 + 5 * i, b2 + 5 * j, variant, mode); }
            }
            out[i * n + j] = r;
        }
    if (n_circle_pass) *n_circle_pass = nc;
    if (n_sat_pass) *n_sat_pass = ns;
    if (n_emu) *n_emu = ne;
}

// the restatements alone (must agree with the oracle bit-for-bit on the host)
extern "C" __attribute__((visibility("default")))
void hg_emu_matrix(const float* b1, int64_t m, const float* b2, int64_t n, int variant, int mode, float* out) {
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++) out[i * n + j] = emu::pair(b1 + 5 * i, b2 + 5 * j, variant, mode);
}
