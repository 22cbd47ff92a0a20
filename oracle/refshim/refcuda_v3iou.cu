// Unmodified reference v3 IoU CUDA kernel (r3det/ops/box_iou_rotated/src/box_iou_rotated_cuda.cu) behind a C ABI.
#include R3REF_BOX_IOU_ROTATED_CUDA
#include "refcuda_common.h"

R3REF_EXPORT float refcuda_v3_iou_matrix(const float* b1, int64_t m, const float* b2, int64_t n, int iou_or_iof,
                                         float* out, int iters) {
    int dev = r3ref_current_device();
    auto t1 = r3ref_dev_f32(b1, {m, 5}, dev), t2 = r3ref_dev_f32(b2, {n, 5}, dev);
    at::Tensor res;
    float ms = r3ref_time_ms(iters, [&] { res = box_iou_rotated_cuda(t1, t2, iou_or_iof != 0); });
    res = res.contiguous();
    cudaMemcpy(out, res.data_ptr<float>(), sizeof(float) * m * n, cudaMemcpyDeviceToDevice);
    return ms;
}
