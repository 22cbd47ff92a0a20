// TEST INFRASTRUCTURE ONLY — reference v2 (ml_nms_rotated) geometry header, unmodified, behind a C ABI.
//   reference header: r3det/ops/ml_nms_rotated/src/box_iou_rotated_utils.h (torch-free)
// nms_rotated_cpu.cpp of that op does not compile on torch 2.11 (dets.type() dispatch), so the
// 25-line greedy loop (ml_nms_rotated/src/nms_rotated_cpu.cpp:36-57) is restated around the header;
// the sort is a stable descending sort on score (scores in tests are distinct).
#include <algorithm>
#include <cmath>
#include <cassert>
#include <cstdint>
#include <numeric>
#include <vector>
#include R3REF_ML_UTILS_H

// 6-float boxes: x,y,w,h,a,label
extern "C" __attribute__((visibility("default")))
void ref_v2_iou_matrix_f32(const float* b1, int64_t m, const float* b2, int64_t n, float* out) {
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++)
            out[i * n + j] = single_box_iou_rotated<float>(b1 + 6 * i, b2 + 6 * j);
}

extern "C" __attribute__((visibility("default")))
int64_t ref_v2_nms_f32(const float* dets6, const float* scores, int64_t K, float thr, int64_t* keep_out) {
    std::vector<int64_t> order(K);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return scores[a] > scores[b]; });
    std::vector<uint8_t> suppressed(K, 0);
    int64_t num = 0;
    for (int64_t _i = 0; _i < K; _i++) {
        auto i = order[_i];
        if (suppressed[i]) continue;
        keep_out[num++] = i;
        for (int64_t _j = _i + 1; _j < K; _j++) {
            auto j = order[_j];
            if (suppressed[j]) continue;
            auto ovr = single_box_iou_rotated<float>(dets6 + 6 * i, dets6 + 6 * j);
            if (ovr >= thr) suppressed[j] = 1;
        }
    }
    return num;
}
