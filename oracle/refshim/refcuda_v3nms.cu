// Unmodified reference v3 NMS CUDA path (r3det/ops/nms_rotated/src/nms_rotated_cuda.cu: mask kernel +
// mask D2H + serial CPU scan) behind a C ABI.  keep_out is a HOST buffer of K int64; returns the count
// through *num_keep and mean ms per call as the return value.
#include R3REF_NMS_ROTATED_CUDA
#include "refcuda_common.h"
#include <cstring>

R3REF_EXPORT float refcuda_v3_nms(const float* dets, const float* scores, int64_t K, float thr,
                                  int64_t* keep_out, int64_t* num_keep, int iters) {
    int dev = r3ref_current_device();
    auto d = r3ref_dev_f32(dets, {K, 5}, dev), s = r3ref_dev_f32(scores, {K}, dev);
    at::Tensor keep;
    float ms = r3ref_time_ms(iters, [&] { keep = nms_rotated_cuda(d, s, thr); });
    auto kc = keep.to(at::kCPU).contiguous();
    std::memcpy(keep_out, kc.data_ptr<int64_t>(), sizeof(int64_t) * kc.numel());
    *num_keep = kc.numel();
    return ms;
}
