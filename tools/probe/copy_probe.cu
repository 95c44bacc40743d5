// Development probe: DRAM efficiency of the marching blur's access pattern as a pure copy --
// TMA band loads + (A) per-thread STG.64 stores as in the column pass, (B) TMA bulk tensor stores from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/copy_probe tools/probe/copy_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int TW = 128, BH = 32, NSTG = 3;
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool TMA_STORE>
__global__ void __launch_bounds__(256, 2) k_copy(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                                                  float* dst, int w, int h, int pitch, long long img_stride, int seg_rows) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar[NSTG];
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ya = blockIdx.y * seg_rows, img = blockIdx.z;
    const int yb = min(ya + seg_rows, h), nb = (yb - ya + BH - 1) / BH;
    if (tid == 0) { for (int b = 0; b < NSTG; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[b]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto issue = [&](int b) {
        const uint32_t ba = su32(&bar[b % NSTG]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(BH * TW * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(su32(sm + (b % NSTG) * BH * TW)), "l"(&tin), "r"(tx0), "r"(ya + b * BH), "r"(img), "r"(ba) : "memory");
    };
    if (tid == 0) for (int b = 0; b < NSTG && b < nb; b++) issue(b);
    float* out = dst + (long long)img * img_stride;
    for (int b = 0; b < nb; b++) {
        const uint32_t ba = su32(&bar[b % NSTG]);
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(ba), "r"((uint32_t)(b / NSTG) & 1u) : "memory");
        const float* st = sm + (b % NSTG) * BH * TW;
        if (TMA_STORE) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                             ::"l"(&tout), "r"(tx0), "r"(ya + b * BH), "r"(img), "r"(su32(st)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be overwritten again
            }
            __syncthreads();
        } else {
            const int cy = tid >> 6, xp = tid & 63;
            for (int j = 0; j < 8; j++) {
                const int r = cy * 8 + j, gy = ya + b * BH + r, gx = tx0 + 2 * xp;
                if (gy < yb && gx + 1 < w) *reinterpret_cast<float2*>(out + (long long)gy * pitch + gx) = *reinterpret_cast<const float2*>(st + r * TW + 2 * xp);
            }
            __syncthreads();
        }
        if (tid == 0 && b + NSTG < nb) issue(b + NSTG);
    }
}
int main(int argc, char** argv) {
    const int w = 3840, h = 2160, n = argc > 1 ? atoi(argv[1]) : 8, pitch = 3840, seg = argc > 2 ? atoi(argv[2]) : 544;
    const long long img = (long long)pitch * h;
    float *a, *b; CK(cudaMalloc(&a, img * n * 4)); CK(cudaMalloc(&b, img * n * 4));
    CK(cudaMemset(a, 1, img * n * 4));
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap ti, to;
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)img * 4};
    const cuuint32_t box[3] = {TW, BH, 1}, es[3] = {1, 1, 1};
    if (((Fn)fn)(&ti, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1;
    if (((Fn)fn)(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, b, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1;
    const size_t smem = NSTG * BH * TW * 4 + 60000;   // padded so that 2 CTAs fit per SM as in the blur
    CK(cudaFuncSetAttribute(k_copy<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_copy<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid((w + TW - 1) / TW, (h + seg - 1) / seg, n);
    const double bytes = 2.0 * w * h * 4 * n;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k_copy<false><<<grid, 256, smem>>>(ti, to, b, w, h, pitch, img, seg); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("STG stores: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(e0); k_copy<true><<<grid, 256, smem>>>(ti, to, b, w, h, pitch, img, seg); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
        printf("TMA stores: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(e0); cudaMemcpyAsync(b, a, img * n * 4, cudaMemcpyDeviceToDevice); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
        printf("cudaMemcpy: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    }
    return 0;
}
