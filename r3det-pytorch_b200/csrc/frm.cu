#include "common.cuh"
R3G_API int r3g_frm_forward_f32(const float*, const float*, int, int, int, int, float, int, float*, void*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
R3G_API int r3g_frm_backward_workspace_bytes(int, int, int, int, size_t*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
R3G_API int r3g_frm_backward_f32(const float*, const float*, int, int, int, int, float, int, float*, void*, size_t, void*) { r3g::set_error("not built yet"); return R3G_ERR_ARG; }
