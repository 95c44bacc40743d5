// Development probe: which tensor-map / kernel variants of the TMA box load run on this GPU.
// usage: tma_probe <rank: 2|3|4> <stride_mode: 0 multiple | 1 non-multiple> <bar: 0 static | 1 dynamic>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BW = 156, SH = 90;
template <int RANK, bool DYNBAR>
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int x, int y, int l, int img) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ __align__(8) uint64_t sbar;
    uint64_t* barp = DYNBAR ? reinterpret_cast<uint64_t*>(smem + BW * SH) : &sbar;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(barp);
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(BW * SH * 4)) : "memory");
        if (RANK == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(dst), "l"(&tm), "r"(x), "r"(y), "r"(l), "r"(img), "r"(bar_a) : "memory");
        else if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(&tm), "r"(x), "r"(y), "r"(img), "r"(bar_a) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(&tm), "r"(x), "r"(y), "r"(bar_a) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < BW * SH; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), smode = atoi(argv[2]), dyn = atoi(argv[3]);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    const int w = 640, h = 426, pitch = 640, layers = 6, nimg = 2;
    long long layer_stride = (long long)pitch * h;
    long long img_stride = layer_stride * layers + (smode ? 32 * 7 : 0);   // non-multiple of the layer stride when smode=1
    std::vector<float> hbuf(img_stride * nimg);
    for (size_t i = 0; i < hbuf.size(); i++) hbuf[i] = (float)(i % 100003);
    float *d, *dout; cudaMalloc(&d, hbuf.size() * 4); cudaMalloc(&dout, BW * SH * 4);
    cudaMemcpy(d, hbuf.data(), hbuf.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)layers, (cuuint64_t)nimg};
    cuuint64_t gstr[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)layer_stride * 4, (cuuint64_t)img_stride * 4};
    cuuint32_t box[4] = {BW, SH, 1, 1}, es[4] = {1, 1, 1, 1};
    int img = 1, l = 2;
    float* base = d;
    if (rank == 3) { gdim[2] = nimg; gstr[1] = (cuuint64_t)img_stride * 4; base = d + l * layer_stride; }
    if (rank == 2) { base = d + l * layer_stride + img * img_stride; }
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, base, gdim, gstr, box, es,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rank %d smode %d dyn %d -> %d\n", rank, smode, dyn, (int)r);
    if (r) return 2;
    size_t smem = BW * SH * 4 + 64;
    int x = -13, y = 115;
#define LAUNCH(R, D) { cudaFuncSetAttribute(k<R, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<R, D><<<1, 256, smem>>>(tm, dout, x, y, l, img); }
    if (rank == 4) { if (dyn) LAUNCH(4, true) else LAUNCH(4, false) }
    else if (rank == 3) { if (dyn) LAUNCH(3, true) else LAUNCH(3, false) }
    else { if (dyn) LAUNCH(2, true) else LAUNCH(2, false) }
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e) return 1;
    std::vector<float> o(BW * SH);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r_ = 0; r_ < SH; r_++) for (int c = 0; c < BW; c++) {
        int gy = y + r_, gx = x + c;
        float exp = (gx < 0 || gx >= w || gy < 0 || gy >= h) ? 0.f : hbuf[img * img_stride + l * layer_stride + (long long)gy * pitch + gx];
        if (o[r_ * BW + c] != exp) bad++;
    }
    printf("  mismatches: %d\n", bad);
    return bad != 0;
}
