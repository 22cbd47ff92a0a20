// Unmodified reference polygon NMS CUDA path (r3det/ops/nms_rotated/src/poly_nms_cuda.cu: mask kernel + mask D2H +
// serial CPU scan) behind a C ABI.  dets: K x 9 (8 polygon coordinates + score) DEVICE floats; keep_out is a HOST
// buffer of K int64; returns mean ms per call.
#include "refcuda_common.h"
#include R3REF_POLY_NMS_CUDA
#include <cstring>

R3REF_EXPORT float refcuda_poly_nms(const float* dets, int64_t K, float thr, int64_t* keep_out, int64_t* num_keep, int iters) {
    int dev = r3ref_current_device();
    auto d = r3ref_dev_f32(dets, {K, 9}, dev);
    at::Tensor keep;
    float ms = r3ref_time_ms(iters, [&] { keep = poly_nms_cuda(d, thr); });
    auto kc = keep.to(at::kCPU).contiguous();
    std::memcpy(keep_out, kc.data_ptr<int64_t>(), sizeof(int64_t) * kc.numel());
    *num_keep = kc.numel();
    return ms;
}

// the reference's device function devPolyIoU (poly_nms_cuda.cu:122-143), row by row: p, q (n, 8) DEVICE floats
__global__ void r3ref_poly_iou_kernel(const float* p, const float* q, int64_t n, float* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = devPolyIoU(p + i * 8, q + i * 8);
}

R3REF_EXPORT void refcuda_poly_iou(const float* p, const float* q, int64_t n, float* out) {
    if (n <= 0) return;
    r3ref_poly_iou_kernel<<<(unsigned)((n + 127) / 128), 128>>>(p, q, n, out);
    cudaDeviceSynchronize();
}
