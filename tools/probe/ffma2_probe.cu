// Development probe: throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on this GPU.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b); rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float2 a[8];
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    float2 b = make_float2(s, s * 1.0001f), c = make_float2(0.5f, 0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            else a[i] = ffma2(a[i], b, c);
        }
    }
    float acc = 0;
    for (int i = 0; i < 8; i++) acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 0.999f); else k<1><<<148 * 8, 256>>>(d, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fmas = 148.0 * 8 * 256 * iters * 16;
        printf("mode %d: %.3f ms, %.2f TFMA/s (%.1f FMA/clk/SM at 1.92 GHz)\n", mode, ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.92e9);
    }
    return 0;
}
