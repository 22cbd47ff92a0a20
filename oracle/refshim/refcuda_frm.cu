// Unmodified reference FRM CUDA kernels (r3det/ops/fr/src/feature_refine_kernel.cu) behind a C ABI.
// The Python side of the reference zero-fills the output first (feature_refine_module.py:21,36); the
// timed region includes that fill, as the reference's forward()/backward() do.
#include R3REF_FEATURE_REFINE_KERNEL
#include "refcuda_common.h"

R3REF_EXPORT float refcuda_frm_forward(const float* feat, const float* boxes, int N, int C, int H, int W,
                                       float scale, int points, float* out, int iters) {
    int dev = r3ref_current_device();
    auto f = r3ref_dev_f32(feat, {N, C, H, W}, dev);
    auto b = r3ref_dev_f32(boxes, {(int64_t)N * H * W, 5}, dev);
    auto o = r3ref_dev_f32(out, {N, C, H, W}, dev);
    return r3ref_time_ms(iters, [&] { o.zero_(); FRForwardLauncher(f, b, scale, points, o); });
}

R3REF_EXPORT float refcuda_frm_backward(const float* gout, const float* boxes, int N, int C, int H, int W,
                                        float scale, int points, float* gin, int iters) {
    int dev = r3ref_current_device();
    auto g = r3ref_dev_f32(gout, {N, C, H, W}, dev);
    auto b = r3ref_dev_f32(boxes, {(int64_t)N * H * W, 5}, dev);
    auto o = r3ref_dev_f32(gin, {N, C, H, W}, dev);
    return r3ref_time_ms(iters, [&] { o.zero_(); FRBackwardLauncher(g, b, scale, points, o); });
}
