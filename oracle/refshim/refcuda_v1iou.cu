// Unmodified reference v1 IoU CUDA kernel (r3det/ops/rbbox_geo/src/rbbox_geo_kernel.cu) behind a C ABI.
// Device pointers in, device pointer out.  Returns mean ms per call over `iters` calls (>=1).
#include R3REF_RBBOX_GEO_KERNEL
#include "refcuda_common.h"

R3REF_EXPORT float refcuda_v1_iou_matrix(const float* b1, int64_t m, const float* b2, int64_t n, int iof,
                                         float* out, int iters) {
    int dev = r3ref_current_device();
    auto t1 = r3ref_dev_f32(b1, {m, 5}, dev), t2 = r3ref_dev_f32(b2, {n, 5}, dev);
    torch::Tensor res;
    float ms = r3ref_time_ms(iters, [&] { res = mat_iou_iof_launcher(t1, t2, iof != 0); });
    cudaMemcpy(out, res.data_ptr<float>(), sizeof(float) * m * n, cudaMemcpyDeviceToDevice);
    return ms;
}

R3REF_EXPORT float refcuda_v1_iou_aligned(const float* b1, int64_t n1, const float* b2, int64_t n2, int iof,
                                          float* out, int iters) {
    int dev = r3ref_current_device();
    auto t1 = r3ref_dev_f32(b1, {n1, 5}, dev), t2 = r3ref_dev_f32(b2, {n2, 5}, dev);
    torch::Tensor res;
    float ms = r3ref_time_ms(iters, [&] { res = vec_iou_iof_launcher(t1, t2, iof != 0); });
    cudaMemcpy(out, res.data_ptr<float>(), sizeof(float) * std::max(n1, n2), cudaMemcpyDeviceToDevice);
    return ms;
}
