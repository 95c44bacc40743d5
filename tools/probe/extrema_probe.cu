// Development probe: k_extrema (generic loads) vs k_extrema_tma on one random octave; compares the masks and times both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -lineinfo -I sift_features_b200/csrc -o tools/probe/extrema_probe tools/probe/extrema_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sb_common.cuh"
#include "sb_pyramid.cuh"
using namespace sb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); return 1; } } while (0)
// layer l = box filter of radius l+1 over hashed noise: smooth across space and layers like a real pyramid, so
// extrema are rare (about 1e-3 of the pixels) instead of 2/27 of them
__global__ void k_fill(float* g, int w, int h, int pitch, long long ls, long long img_floats, int n) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z / N_LAYERS, l = blockIdx.z % N_LAYERS;
    if (x >= w) return;
    const int R = l + 1;
    float acc = 0.f;
    for (int dy = -R; dy <= R; dy++) for (int dx = -R; dx <= R; dx++) {
        uint32_t s = (uint32_t)(x + dx + 64) * 73856093u ^ (uint32_t)(y + dy + 64) * 19349663u ^ (uint32_t)(img + 1) * 83492791u;
        s ^= s >> 13; s *= 0x5bd1e995u; s ^= s >> 15;
        acc += (float)(s >> 8) * (1.0f / 16777216.0f);
    }
    float v = acc / (float)((2 * R + 1) * (2 * R + 1));
    if (y >= 100 && y < 140 && x >= 50 && x < 200) v = 0.25f + 0.01f * l * l;   // a flat patch: the flat-candidate path
    g[img * img_floats + l * ls + (long long)y * pitch + x] = v;
}

int main(int argc, char** argv) {
    const int w = argc > 1 ? atoi(argv[1]) : 3840, h = argc > 2 ? atoi(argv[2]) : 2160, n = argc > 3 ? atoi(argv[3]) : 8;
    const int pitch = (w + 31) / 32 * 32;
    const long long ls = (long long)pitch * h, img_floats = ls * N_LAYERS;
    const int mp = 2 * ex_strips(w);
    const long long mask_words = (long long)mp * h * 3;
    float* g; uint32_t *mA, *mB, *rA, *rB;
    CK(cudaMalloc(&g, img_floats * n * 4));
    CK(cudaMalloc(&mA, mask_words * n * 4)); CK(cudaMalloc(&mB, mask_words * n * 4));
    CK(cudaMalloc(&rA, 3 * h * n * 4)); CK(cudaMalloc(&rB, 3 * h * n * 4));
    k_fill<<<dim3((w + 127) / 128, h, n * N_LAYERS), 128>>>(g, w, h, pitch, ls, img_floats, n);
    CK(cudaDeviceSynchronize());
    CUtensorMap tm;
    {
        typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        const cuuint64_t gdim[4] = {(cuuint64_t)w, (cuuint64_t)h, N_LAYERS, (cuuint64_t)n};
        const cuuint64_t gstr[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)ls * 4, (cuuint64_t)img_floats * 4};
        const cuuint32_t box[4] = {EXT_BOX_W, EXT_RB, N_LAYERS, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = ((Fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode -> %d\n", (int)r);
        if (r) return 1;
    }
    ExtremaParams e{};
    e.gauss = g; e.img_stride = img_floats; e.layer_stride = ls; e.w = w; e.h = h; e.pitch = pitch;
    e.mask_img_stride = mask_words; e.mask_pitch = mp; e.rows_img_stride = 3 * h;
    CK(cudaFuncSetAttribute(k_extrema_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXT_SMEM));
    CK(cudaFuncSetAttribute(k_extrema_tma<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)w * h * 24 * n;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemset(rA, 0, 3 * h * n * 4)); CK(cudaMemset(rB, 0, 3 * h * n * 4));
        CK(cudaMemset(mA, 0xff, mask_words * n * 4)); CK(cudaMemset(mB, 0xff, mask_words * n * 4));
        e.mask = mA; e.rows = rA;
        cudaEventRecord(e0);
        k_extrema<false><<<dim3(ex_strips(w), (h + EX_ROWS * EX_WARPS - 1) / (EX_ROWS * EX_WARPS), n), 32 * EX_WARPS>>>(e);
        cudaEventRecord(e1);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("generic: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
        e.mask = mB; e.rows = rB;
        cudaEventRecord(e0);
        k_extrema_tma<false><<<dim3(ex_strips(w), (h + EXT_ROWS * EX_WARPS - 1) / (EXT_ROWS * EX_WARPS), n), 32 * EX_WARPS, EXT_SMEM>>>(tm, e);
        cudaEventRecord(e1);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        printf("tma:     %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    }
    std::vector<uint32_t> a(mask_words * n), b(mask_words * n), ra(3 * h * n), rb(3 * h * n);
    CK(cudaMemcpy(a.data(), mA, a.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), mB, b.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ra.data(), rA, ra.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rb.data(), rB, rb.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0, badr = 0, cand = 0;
    for (size_t i = 0; i < a.size(); i++) { bad += a[i] != b[i]; cand += __builtin_popcount(a[i]); }
    for (size_t i = 0; i < ra.size(); i++) badr += ra[i] != rb[i];
    printf("mask words differing: %lld of %zu, row counters differing: %lld, candidates (generic): %lld\n", bad, a.size(), badr, cand);
    return bad || badr;
}
