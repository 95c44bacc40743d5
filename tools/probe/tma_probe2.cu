// Development probe 2: the CUDA programming guide's TMA example (libcu++ wrappers) vs raw PTX, 2-D only.
// usage: tma_probe2 <mode: 0 libcu++ | 1 raw ptx> <BW> <SH> <x> <y>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k_lib(const __grid_constant__ CUtensorMap tm, float* out, int x, int y, int n) {
    extern __shared__ __align__(1024) float smem[];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(smem, &tm, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, n * 4);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = smem[i];
}
__global__ void k_raw(const __grid_constant__ CUtensorMap tm, float* out, int x, int y, int n) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ __align__(8) uint64_t sbar;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&sbar);
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&tm), "r"(x), "r"(y), "r"(bar_a) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(n * 4)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv) {
    int mode = atoi(argv[1]), BW = atoi(argv[2]), SH = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    const int w = 640, h = 426, pitch = 640;
    std::vector<float> hbuf((size_t)pitch * h);
    for (size_t i = 0; i < hbuf.size(); i++) hbuf[i] = (float)(i % 100003);
    float *d, *dout; cudaMalloc(&d, hbuf.size() * 4); cudaMalloc(&dout, BW * SH * 4);
    cudaMemcpy(d, hbuf.data(), hbuf.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)SH}, es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d box %dx%d at (%d,%d): entry %d q %d drv %d rt %d encode %d\n", mode, BW, SH, x, y, (int)ge, (int)q, drv, rt, (int)r);
    if (r) return 2;
    size_t smem = (size_t)BW * SH * 4;
    if (mode == 0) { cudaFuncSetAttribute(k_lib, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_lib<<<1, 128, smem>>>(tm, dout, x, y, BW * SH); }
    else { cudaFuncSetAttribute(k_raw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_raw<<<1, 128, smem>>>(tm, dout, x, y, BW * SH); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e) return 1;
    std::vector<float> o((size_t)BW * SH);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r_ = 0; r_ < SH; r_++) for (int c = 0; c < BW; c++) {
        int gy = y + r_, gx = x + c;
        float exp = (gx < 0 || gx >= w || gy < 0 || gy >= h) ? 0.f : hbuf[(long long)gy * pitch + gx];
        if (o[(size_t)r_ * BW + c] != exp) bad++;
    }
    printf("  mismatches: %d\n", bad);
    return bad != 0;
}
