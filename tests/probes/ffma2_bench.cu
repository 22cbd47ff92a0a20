// microbenchmark: issue rate of FFMA vs FFMA2 (packed f32x2) on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int MODE> __global__ void k(float* out, int iters, float x) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; i++) { a[i] = x + i + threadIdx.x; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
    unsigned long long X = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], x, 0.5f);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], X, X);
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* o; cudaMalloc(&o, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(o, 100000, 1.0001f); else k<1><<<148 * 8, 256>>>(o, 100000, 1.0001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double inst = 148.0 * 8 * 8 * 100000.0 * 8;      // warp instructions
            printf("%s: %.3f ms  %.1f G warp-instr/s  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, inst / ms / 1e6, inst * 32 * 2 * (mode ? 2 : 1) / ms / 1e9);
        }
    }
    return 0;
}
