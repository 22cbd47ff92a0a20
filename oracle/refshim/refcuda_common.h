// TEST/BENCH INFRASTRUCTURE ONLY — helpers for the CUDA reference shims (oracle/_ref/libref_cuda_*.so).
#pragma once
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>
#include <cstdint>

#define R3REF_EXPORT extern "C" __attribute__((visibility("default")))

static inline at::Tensor r3ref_dev_f32(const float* p, std::vector<int64_t> shape, int dev) {
    return at::from_blob(const_cast<float*>(p), shape, at::TensorOptions().dtype(at::kFloat).device(at::kCUDA, dev));
}
static inline int r3ref_current_device() { int d = 0; cudaGetDevice(&d); return d; }

// Runs fn() `iters` times on the current torch stream and returns the mean milliseconds (CUDA events).
template <typename F>
static inline float r3ref_time_ms(int iters, F fn) {
    cudaStream_t st = at::cuda::getCurrentCUDAStream();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaStreamSynchronize(st);
    cudaDeviceSynchronize();
    // the v1 IoU / FRM reference kernels launch on the legacy default stream; time on stream 0 when asked
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; i++) fn();
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms / (iters > 0 ? iters : 1);
}
