// common.cuh — error plumbing and small device helpers shared by the libr3geo translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/r3geo.h"

#define R3G_API extern "C" __attribute__((visibility("default")))

namespace r3g {

void set_error(const char* fmt, ...);      // api.cu; thread-local message for r3g_last_error()
int device_sm_count();                     // cached per device
int current_device_slot();                 // current device index clamped to [0, 64): index for per-device one-time setup

#define R3G_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            r3g::set_error(__VA_ARGS__);               \
            return R3G_ERR_ARG;                        \
        }                                              \
    } while (0)

#define R3G_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            r3g::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));           \
            return R3G_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

#define R3G_LAUNCH_OK(name)                                                            \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            r3g::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));  \
            return R3G_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// streaming (evict-first) stores for write-once outputs that nobody re-reads from L2
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_cs_f1(float* p, float v) { __stcs(p, v); }
#endif

}  // namespace r3g
