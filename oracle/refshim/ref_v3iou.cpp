// TEST INFRASTRUCTURE ONLY — unmodified reference v3 IoU (CPU) behind a C ABI.
//   reference TU: r3det/ops/box_iou_rotated/src/box_iou_rotated_cpu.cpp (+ box_iou_rotated_utils.h)
#include R3REF_BOX_IOU_ROTATED_CPU
#include <cstdint>
#include <cstring>

// through the reference's own at::Tensor entry point (box_iou_rotated_cpu.cpp:23-38)
extern "C" __attribute__((visibility("default")))
void ref_v3_iou_matrix_tensor_f32(const float* b1, int64_t m, const float* b2, int64_t n, int iou_or_iof, float* out) {
    auto t1 = at::from_blob(const_cast<float*>(b1), {m, 5}, at::kFloat);
    auto t2 = at::from_blob(const_cast<float*>(b2), {n, 5}, at::kFloat);
    auto r = box_iou_rotated_cpu(t1, t2, iou_or_iof != 0).contiguous();
    std::memcpy(out, r.data_ptr<float>(), sizeof(float) * m * n);
}

// same per-pair function without the per-pair Tensor indexing overhead (fast path for big samples)
extern "C" __attribute__((visibility("default")))
void ref_v3_iou_matrix_f32(const float* b1, int64_t m, const float* b2, int64_t n, int iou_or_iof, float* out) {
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++)
            out[i * n + j] = single_box_iou_rotated<float>(b1 + 5 * i, b2 + 5 * j, iou_or_iof != 0);
}

extern "C" __attribute__((visibility("default")))
void ref_v3_iou_matrix_f64(const double* b1, int64_t m, const double* b2, int64_t n, int iou_or_iof, double* out) {
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++)
            out[i * n + j] = single_box_iou_rotated<double>(b1 + 5 * i, b2 + 5 * j, iou_or_iof != 0);
}
