// Unmodified reference v1 NMS CUDA path (r3det/ops/rnms/src/rcuda/rnms_kernel.cu: mask kernel + mask D2H + serial
// CPU scan + final index sort) behind a C ABI.  dets: K x 6 (box + score) DEVICE floats; keep_out is a HOST buffer of
// K int64 (ascending original index, as the reference returns it); returns mean ms per call.
#include "refcuda_common.h"
#include R3REF_RNMS_KERNEL
#include <cstring>

R3REF_EXPORT float refcuda_v1_nms(const float* dets, int64_t K, float thr, int64_t* keep_out, int64_t* num_keep, int iters) {
    int dev = r3ref_current_device();
    auto d = r3ref_dev_f32(dets, {K, 6}, dev);
    at::Tensor keep;
    float ms = r3ref_time_ms(iters, [&] { keep = nmsr_cuda(d, thr); });
    auto kc = keep.to(at::kCPU).contiguous();
    std::memcpy(keep_out, kc.data_ptr<int64_t>(), sizeof(int64_t) * kc.numel());
    *num_keep = kc.numel();
    return ms;
}
