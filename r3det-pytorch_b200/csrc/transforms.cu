#include "common.cuh"
R3G_API int r3g_obb2poly_f32(const float*, int64_t, int, float*, void*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
R3G_API int r3g_poly2obb_f32(const float*, int64_t, int, float*, void*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
R3G_API int r3g_obb2hbb_f32(const float*, int64_t, int, float*, void*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
R3G_API int r3g_hbb2obb_f32(const float*, int64_t, int, float*, void*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
R3G_API int r3g_obb2xyxy_f32(const float*, int64_t, int, float*, void*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
