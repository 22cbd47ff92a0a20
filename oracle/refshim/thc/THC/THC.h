// TEST/BENCH INFRASTRUCTURE ONLY — stand-in for the THC headers that torch >= 1.11 no longer ships, so that the
// reference's rnms_kernel.cu and poly_nms_cuda.cu compile UNMODIFIED (they use THCCeilDiv, THCudaMalloc / THCudaFree,
// THCudaCheck and `at::globalContext().lazyInitCUDA()` returning a THCState*).  Nothing here touches their arithmetic.
#pragma once
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <cuda_runtime.h>
#include <stdexcept>

struct THCState {};
template <typename T>
__host__ __device__ inline T THCCeilDiv(T a, T b) { return (a + b - 1) / b; }
static inline void* THCudaMalloc(THCState*, size_t n) { void* p = nullptr; cudaMalloc(&p, n); return p; }
static inline void THCudaFree(THCState*, void* p) { cudaFree(p); }
#define THCudaCheck(x)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (x);                                                           \
        if (e__ != cudaSuccess) throw std::runtime_error(cudaGetErrorString(e__));       \
    } while (0)
namespace at {
struct R3ShimContext { THCState* lazyInitCUDA() { static THCState s; return &s; } };
inline R3ShimContext& r3shim_global_context() { static R3ShimContext c; return c; }
}  // namespace at
#define globalContext r3shim_global_context
