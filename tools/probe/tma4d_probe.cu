// Development probe: which 4-D TMA box shapes load correctly (box = bx x by x bl x 1 over a (w, h, 6, n) tensor).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int nfloats, int x, int y, int who) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* dst = sm + warp * nfloats;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[warp])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (who < 0 || warp == who) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp])), "r"(nfloats * 4) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(smem_u32(dst)), "l"(&tm), "r"(x), "r"(y + warp), "r"(0), "r"(0), "r"(smem_u32(&bar[warp])) : "memory");
        }
    }
    __syncwarp();
    if (who < 0 || warp == who) {
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar[warp])), "r"(0u) : "memory");
        float s = 0;
        for (int i = lane; i < nfloats; i += 32) s += dst[i];
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) out[warp] = s;
    }
}
int main() {
    const int w = 512, h = 256, n = 2, pitch = 512;
    const long long ls = (long long)pitch * h, imgf = ls * 6;
    float* g; CK(cudaMalloc(&g, imgf * n * 4));
    std::vector<float> hst(imgf * n);
    for (size_t i = 0; i < hst.size(); i++) hst[i] = (float)(i % 977) * 0.001f;
    CK(cudaMemcpy(g, hst.data(), hst.size() * 4, cudaMemcpyHostToDevice));
    float* out; CK(cudaMalloc(&out, 64));
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int shapes[][3] = {{64, 3, 1}, {64, 3, 6}, {64, 1, 6}, {64, 4, 6}, {64, 3, 2}, {72, 3, 6}, {64, 8, 6}};
    for (auto& sh : shapes) for (int who : {0, 1, -1}) for (int x : {0, -2, 58}) {
        CUtensorMap tm;
        const cuuint64_t gdim[4] = {(cuuint64_t)w, (cuuint64_t)h, 6, (cuuint64_t)n};
        const cuuint64_t gstr[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)ls * 4, (cuuint64_t)imgf * 4};
        const cuuint32_t box[4] = {(cuuint32_t)sh[0], (cuuint32_t)sh[1], (cuuint32_t)sh[2], 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = ((Fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int nf = sh[0] * sh[1] * sh[2];
        k<<<1, 128, nf * 4 * 4>>>(tm, out, nf, x, 5, who);
        cudaError_t e = cudaDeviceSynchronize();
        float o[4] = {0}; if (e == cudaSuccess) cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
        printf("box %dx%dx%d who %d x %d: encode %d, run %s  sums %.3f %.3f\n", sh[0], sh[1], sh[2], who, x, (int)r, cudaGetErrorString(e), o[0], o[1]);
        if (e != cudaSuccess) return 0;   // context is dead after a fault
    }
    return 0;
}
