// transforms.cu — obb/poly/hbb box transforms under the three angle conventions, one launch per call.
//
// Replaces the torch implementations in r3det/core/bbox/rtransforms.py (each a chain of 20-40 tiny torch
// kernels): poly2obb_v1/v2/v3 :190-277, obb2poly_v1/v2/v3 :367-440, obb2hbb_v1/v2/v3 :443-537,
// hbb2obb_v1/v2/v3 :540-592, obb2xyxy_v1/v2/v3 :595-651, norm_angle :789-805.
// Arithmetic follows the reference expression by expression in FP32 (Python-float constants such as
// np.pi/2 are rounded to FP32 first, as torch does for tensor-scalar ops).
#include "common.cuh"

namespace r3g {

constexpr float PI_F = 3.14159265358979323846f;
constexpr float HALF_PI_F = 1.57079632679489661923f;
constexpr float QUARTER_PI_F = 0.78539816339744830962f;

// torch.remainder: result takes the sign of the divisor
__device__ __forceinline__ float py_remainder(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.0f && ((b < 0.0f) != (m < 0.0f))) m += b;
    return m;
}

__device__ __forceinline__ float norm_angle(float a, int version) {           // rtransforms.py:789-805
    if (version == 2) return py_remainder(a + QUARTER_PI_F, PI_F) - QUARTER_PI_F;
    if (version == 3) return py_remainder(a + HALF_PI_F, PI_F) - HALF_PI_F;
    return a;
}

__device__ __forceinline__ void obb2poly_dev(const float* b, int version, float* p) {
    const float x = b[0], y = b[1], w = b[2], h = b[3], a = b[4];
    const float cosa = cosf(a), sina = sinf(a);
    if (version == 1) {                                                       // :367-388
        const float wx = w / 2 * cosa, wy = w / 2 * sina;
        const float hx = -h / 2 * sina, hy = h / 2 * cosa;
        p[0] = x - wx - hx; p[1] = y - wy - hy;
        p[2] = x + wx - hx; p[3] = y + wy - hy;
        p[4] = x + wx + hx; p[5] = y + wy + hy;
        p[6] = x - wx + hx; p[7] = y - wy + hy;
    } else {                                                                  // :391-440 (v2 == v3)
        const float tlx = -w * 0.5f, tly = -h * 0.5f, brx = w * 0.5f, bry = h * 0.5f;
        const float rx[4] = { tlx, brx, brx, tlx }, ry[4] = { tly, tly, bry, bry };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            p[2 * k] = (cosa * rx[k] + (-sina) * ry[k]) + x;
            p[2 * k + 1] = (sina * rx[k] + cosa * ry[k]) + y;
        }
    }
}

__global__ void obb2poly_kernel(const float* __restrict__ obb, int64_t n, int version, float* __restrict__ poly) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float b[5], p[8];
#pragma unroll
    for (int k = 0; k < 5; k++) b[k] = obb[i * 5 + k];
    obb2poly_dev(b, version, p);
#pragma unroll
    for (int k = 0; k < 8; k++) poly[i * 8 + k] = p[k];
}

__global__ void poly2obb_kernel(const float* __restrict__ poly, int64_t n, int version, float* __restrict__ obb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p[8];
#pragma unroll
    for (int k = 0; k < 8; k++) p[k] = poly[i * 8 + k];
    float o[5];
    if (version == 1) {                                                       // :190-213
        const float cx = (((p[0] + p[2]) + p[4]) + p[6]) / 4.0f;
        const float cy = (((p[1] + p[3]) + p[5]) + p[7]) / 4.0f;
        const float dx01 = p[0] - p[2], dy01 = p[1] - p[3];
        const float dx12 = p[2] - p[4], dy12 = p[3] - p[5];
        const float _w = sqrtf(dx01 * dx01 + dy01 * dy01);
        const float _h = sqrtf(dx12 * dx12 + dy12 * dy12);
        const float th = atan2f(-(p[2] - p[0]), p[3] - p[1]);
        const bool odd = py_remainder(floorf(th / (-HALF_PI_F)), 2.0f) == 0.0f;
        o[0] = cx; o[1] = cy;
        o[2] = odd ? _h : _w;
        o[3] = odd ? _w : _h;
        o[4] = py_remainder(th, -HALF_PI_F);
    } else {                                                                  // :216-277
        const float e1 = sqrtf((p[0] - p[2]) * (p[0] - p[2]) + (p[1] - p[3]) * (p[1] - p[3]));
        const float e2 = sqrtf((p[2] - p[4]) * (p[2] - p[4]) + (p[3] - p[5]) * (p[3] - p[5]));
        const float a1 = atan2f(p[3] - p[1], p[2] - p[0]);
        const float a2 = atan2f(p[7] - p[1], p[6] - p[0]);
        o[0] = (p[0] + p[4]) / 2.0f;
        o[1] = (p[1] + p[5]) / 2.0f;
        o[2] = fmaxf(e1, e2);
        o[3] = fminf(e1, e2);
        o[4] = norm_angle(e1 > e2 ? a1 : a2, version);
    }
#pragma unroll
    for (int k = 0; k < 5; k++) obb[i * 5 + k] = o[k];
}

__device__ __forceinline__ void aabb_of_poly(const float* p, float& xmin, float& ymin, float& xmax, float& ymax) {
    xmin = fminf(fminf(p[0], p[2]), fminf(p[4], p[6])); xmax = fmaxf(fmaxf(p[0], p[2]), fmaxf(p[4], p[6]));
    ymin = fminf(fminf(p[1], p[3]), fminf(p[5], p[7])); ymax = fmaxf(fmaxf(p[1], p[3]), fmaxf(p[5], p[7]));
}

__global__ void obb2hbb_kernel(const float* __restrict__ obb, int64_t n, int version, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float b[5], o[5];
#pragma unroll
    for (int k = 0; k < 5; k++) b[k] = obb[i * 5 + k];
    if (version == 1) {                                                       // :443-462
        const float cosa = cosf(b[4]), sina = sinf(b[4]);
        const float hb_w = cosa * b[2] - sina * b[3];
        const float hb_h = -sina * b[2] + cosa * b[3];
        o[0] = b[0]; o[1] = b[1]; o[2] = hb_h; o[3] = hb_w; o[4] = -HALF_PI_F;
    } else if (version == 2) {                                                // :465-489
        float p[8], xmin, ymin, xmax, ymax;
        obb2poly_dev(b, 2, p);
        aabb_of_poly(p, xmin, ymin, xmax, ymax);
        const float e1 = fabsf(xmax - xmin), e2 = fabsf(ymax - ymin);
        o[0] = (xmax + xmin) / 2.0f; o[1] = (ymax + ymin) / 2.0f;
        if (e1 < e2) { o[2] = e2; o[3] = e1; o[4] = HALF_PI_F; } else { o[2] = e1; o[3] = e2; o[4] = 0.0f; }
    } else {                                                                  // :515-537
        const float c = cosf(b[4]), s = sinf(b[4]);
        const float xb = fabsf(b[2] / 2 * c) + fabsf(b[3] / 2 * s);
        const float yb = fabsf(b[2] / 2 * s) + fabsf(b[3] / 2 * c);
        const float x1 = b[0] - xb, y1 = b[1] - yb, x2 = b[0] + xb, y2 = b[1] + yb;
        const float _x = (x1 + x2) * 0.5f, _y = (y1 + y2) * 0.5f, _w = x2 - x1, _h = y2 - y1;
        o[0] = _x; o[1] = _y;
        if (_w >= _h) { o[2] = _w; o[3] = _h; o[4] = 0.0f; } else { o[2] = _h; o[3] = _w; o[4] = 0.0f - HALF_PI_F; }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) out[i * 5 + k] = o[k];
}

__global__ void hbb2obb_kernel(const float* __restrict__ hbb, int64_t n, int version, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x1 = hbb[i * 4], y1 = hbb[i * 4 + 1], x2 = hbb[i * 4 + 2], y2 = hbb[i * 4 + 3];
    const float x = (x1 + x2) * 0.5f, y = (y1 + y2) * 0.5f, w = x2 - x1, h = y2 - y1;
    float o[5] = { x, y, w, h, 0.0f };
    if (version == 1) { o[2] = h; o[3] = w; o[4] = 0.0f - HALF_PI_F; }                               // :540-554
    else if (!(w >= h)) { o[2] = h; o[3] = w; o[4] = (version == 2) ? HALF_PI_F : 0.0f - HALF_PI_F; } // :557-592
#pragma unroll
    for (int k = 0; k < 5; k++) out[i * 5 + k] = o[k];
}

__global__ void obb2xyxy_kernel(const float* __restrict__ obb, int64_t n, int version, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float b[5], o[4];
#pragma unroll
    for (int k = 0; k < 5; k++) b[k] = obb[i * 5 + k];
    if (version == 1) {                                                       // :595-618
        const float cosa = cosf(b[4]), sina = sinf(b[4]);
        const float dw = cosa * b[2] - sina * b[3];
        const float dh = -sina * b[2] + cosa * b[3];
        o[0] = b[0] - dw / 2; o[1] = b[1] - dh / 2; o[2] = b[0] + dw / 2; o[3] = b[1] + dh / 2;
    } else if (version == 2) {                                                // :621-634
        float p[8];
        obb2poly_dev(b, 2, p);
        aabb_of_poly(p, o[0], o[1], o[2], o[3]);
    } else {                                                                  // :637-651
        const float c = cosf(b[4]), s = sinf(b[4]);
        const float xb = fabsf(b[2] / 2 * c) + fabsf(b[3] / 2 * s);
        const float yb = fabsf(b[2] / 2 * s) + fabsf(b[3] / 2 * c);
        o[0] = b[0] - xb; o[1] = b[1] - yb; o[2] = b[0] + xb; o[3] = b[1] + yb;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) out[i * 4 + k] = o[k];
}

static int check_tf(const char* who, const void* in, int64_t n, int version, const void* out) {
    R3G_REQUIRE(n >= 0, "%s: negative n", who);
    R3G_REQUIRE(version >= 1 && version <= 3, "%s: version must be 1, 2 or 3 (got %d)", who, version);
    if (n == 0) return 1;
    R3G_REQUIRE(in && out, "%s: null pointer", who);
    return 0;
}

}  // namespace r3g

using namespace r3g;

#define R3G_TF_ENTRY(NAME, KERNEL)                                                                   \
    R3G_API int NAME(const float* in, int64_t n, int version, float* out, void* stream) {            \
        int rc = check_tf(#NAME, in, n, version, out);                                               \
        if (rc < 0) return rc;                                                                       \
        if (rc == 1) return R3G_OK;                                                                  \
        KERNEL<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, n, version, out);  \
        R3G_LAUNCH_OK(#KERNEL);                                                                      \
        return R3G_OK;                                                                               \
    }

R3G_TF_ENTRY(r3g_obb2poly_f32, obb2poly_kernel)
R3G_TF_ENTRY(r3g_poly2obb_f32, poly2obb_kernel)
R3G_TF_ENTRY(r3g_obb2hbb_f32, obb2hbb_kernel)
R3G_TF_ENTRY(r3g_hbb2obb_f32, hbb2obb_kernel)
R3G_TF_ENTRY(r3g_obb2xyxy_f32, obb2xyxy_kernel)
