// TMA bring-up probe: 3-D tensor-map loads with the descriptor passed (A) as a top-level __grid_constant__ parameter,
// (B) inside a struct at a fixed index, (C) inside a struct at a dynamic index.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ void load_box(const CUtensorMap* map, float* dst, unsigned long long* bar, int c0, int c1, int c2, unsigned bytes) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 :: "r"(smem_u32(bar)), "r"(0) : "memory");
}
struct Wrap { CUtensorMap m[4]; int pick; };
__global__ void kW(const __grid_constant__ CUtensorMap m, float* out, int c0, int c1, int bytes) {
    __shared__ __align__(128) float buf[64 * 64];
    __shared__ __align__(8) unsigned long long bar;
    load_box(&m, buf, &bar, c0, c1, 1, bytes);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = buf[i];
}
__global__ void kA(const __grid_constant__ CUtensorMap m, float* out) {
    __shared__ __align__(128) float buf[16 * 32];
    __shared__ __align__(8) unsigned long long bar;
    load_box(&m, buf, &bar, 0, 0, 1, 16 * 32 * 4);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = buf[i];
}
__global__ void kB(const __grid_constant__ Wrap w, float* out) {
    __shared__ __align__(128) float buf[16 * 32];
    __shared__ __align__(8) unsigned long long bar;
    load_box(&w.m[1], buf, &bar, 0, 0, 1, 16 * 32 * 4);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = buf[i];
}
__global__ void kC(const __grid_constant__ Wrap w, float* out) {
    __shared__ __align__(128) float buf[16 * 32];
    __shared__ __align__(8) unsigned long long bar;
    int li = 0;
    for (int i = 1; i < 4; i++) if (blockIdx.x + 2 >= (unsigned)i && i <= w.pick) li = i;
    load_box(&w.m[li], buf, &bar, 0, 0, 1, 16 * 32 * 4);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 1; }
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int planes = 4, H = 64, W = 64;
    std::vector<float> h((size_t)planes * H * W);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
    float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 512 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    const cuuint64_t dims[3] = { (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes };
    const cuuint64_t strides[2] = { (cuuint64_t)W * 4, (cuuint64_t)H * W * 4 };
    const cuuint32_t box[3] = { 16, 32, 1 }, es[3] = { 1, 1, 1 };
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    Wrap w; for (int i = 0; i < 4; i++) w.m[i] = m; w.pick = 2;
    if (which >= 10) {
        const int bw = atoi(argv[2]), bh = atoi(argv[3]), c0 = atoi(argv[4]), c1 = atoi(argv[5]);
        const cuuint32_t box2[3] = { (cuuint32_t)bw, (cuuint32_t)bh, 1 };
        CUtensorMap m2;
        r = fn(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode box %d x %d -> %d\n", bw, bh, (int)r);
        kW<<<1, 128>>>(m2, o, c0, c1, bw * bh * 4);
    }
    if (which == 0) kA<<<1, 128>>>(m, o);
    if (which == 1) kB<<<1, 128>>>(w, o);
    if (which == 2) kC<<<1, 128>>>(w, o);
    cudaError_t e = cudaDeviceSynchronize();
    float res[4]; cudaMemcpy(res, o, 16, cudaMemcpyDeviceToHost);
    printf("variant %d: %s; out[0..3] = %g %g %g %g (expect %d ..)\n", which, cudaGetErrorString(e), res[0], res[1], res[2], res[3], H * W);
    return 0;
}
