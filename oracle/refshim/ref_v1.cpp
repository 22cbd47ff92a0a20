// TEST INFRASTRUCTURE ONLY — builds the UNMODIFIED reference host code into oracle/_ref/.
// This TU textually includes the reference's own translation unit where it lies under
// /root/reference (nothing is copied into this repo) and exposes a C ABI over it.
//   reference TU: r3det/ops/rnms/src/rcpu/rnms_cpu.cpp  (v1 geometry :11-221, nmsr_cpu_kernel :223-282)
// The reference has no CPU entry point for the v1 IoU matrix (rbbox_geo is CUDA-only); the loop
// below drives the reference's own host templates exactly as mat_iou_iof_kernel does
// (r3det/ops/rbbox_geo/src/rbbox_geo_kernel.cu:238-267).
#include R3REF_RNMS_CPU   // -DR3REF_RNMS_CPU='"/root/reference/.../rnms_cpu.cpp"'
#include <cstdint>
#include <cstring>

template <typename T>
static T ref_pair_v1(const T* rb1_p, const T* rb2_p, bool iof) {
    Point<T> v1[4], v2[4], u[16];
    rbbox2points(rb1_p, v1);
    rbbox2points(rb2_p, v2);
    int p_cnt = 0;
    p_cnt += vertex_in_rbbox(v1, v2, u + p_cnt);
    p_cnt += vertex_in_rbbox(v2, v1, u + p_cnt);
    p_cnt += rbbox_border_intsec(v1, v2, u + p_cnt);
    if (p_cnt >= 3) {
        T s1 = rb1_p[2] * rb1_p[3];
        T s2 = rb2_p[2] * rb2_p[3];
        T su = area(u, p_cnt);
        su = std::min(su, s1);
        su = std::min(su, s2);
        su = std::max(su, (T)0);
        return iof ? su / s1 : su / (s1 + s2 - su);
    }
    return (T)0;
}

extern "C" __attribute__((visibility("default")))
void ref_v1_iou_matrix_f32(const float* b1, int64_t m, const float* b2, int64_t n, int iof, float* out) {
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++)
            out[i * n + j] = ref_pair_v1<float>(b1 + 5 * i, b2 + 5 * j, iof != 0);
}

extern "C" __attribute__((visibility("default")))
void ref_v1_iou_aligned_f32(const float* b1, int64_t n1, const float* b2, int64_t n2, int iof, float* out) {
    int64_t n = std::max(n1, n2);   // vec_iou_iof_kernel modulo broadcast, rbbox_geo_kernel.cu:278-281
    for (int64_t i = 0; i < n; i++)
        out[i] = ref_pair_v1<float>(b1 + 5 * (i % n1), b2 + 5 * (i % n2), iof != 0);
}

extern "C" __attribute__((visibility("default")))
void ref_v1_iou_matrix_f64(const double* b1, int64_t m, const double* b2, int64_t n, int iof, double* out) {
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++)
            out[i * n + j] = ref_pair_v1<double>(b1 + 5 * i, b2 + 5 * j, iof != 0);
}

// rnms_ext.rnms on CPU tensors: dets (K,6) -> ascending kept indices. Returns count.
extern "C" __attribute__((visibility("default")))
int64_t ref_v1_nms_f32(const float* dets, int64_t K, float thr, int64_t* keep_out) {
    auto t = torch::from_blob(const_cast<float*>(dets), {K, 6}, torch::kFloat32);
    auto keep = rnms_cpu(t, thr).contiguous();
    std::memcpy(keep_out, keep.data_ptr<int64_t>(), sizeof(int64_t) * keep.numel());
    return keep.numel();
}
