// TEST INFRASTRUCTURE ONLY — unmodified reference v3 NMS (CPU) behind a C ABI.
//   reference TU: r3det/ops/nms_rotated/src/nms_rotated_cpu.cpp (+ its box_iou_rotated_utils.h)
#include R3REF_NMS_ROTATED_CPU
#include <cstdint>
#include <cstring>

extern "C" __attribute__((visibility("default")))
int64_t ref_v3_nms_f32(const float* dets, const float* scores, int64_t K, float thr, int64_t* keep_out) {
    auto d = at::from_blob(const_cast<float*>(dets), {K, 5}, at::kFloat);
    auto s = at::from_blob(const_cast<float*>(scores), {K}, at::kFloat);
    auto keep = nms_rotated_cpu(d, s, thr).contiguous();
    std::memcpy(keep_out, keep.data_ptr<int64_t>(), sizeof(int64_t) * keep.numel());
    return keep.numel();
}

// the per-pair IoU this NMS uses (nms_rotated's private copy of the header, IoU only)
extern "C" __attribute__((visibility("default")))
void ref_v3nms_iou_matrix_f32(const float* b1, int64_t m, const float* b2, int64_t n, float* out) {
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++)
            out[i * n + j] = single_box_iou_rotated<float>(b1 + 5 * i, b2 + 5 * j);
}
