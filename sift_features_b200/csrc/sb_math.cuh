// sb_math.cuh -- scalar math whose results must agree bit-for-bit with what the
// reference computes on a Linux host.
//
// The reference crate calls f32::exp (src/lib.rs:705, :862) and f32::powf
// (src/lib.rs:373); on Linux both lower to glibc's expf / powf.  glibc (>= 2.28)
// evaluates them in double with a 32-entry 2^(i/32) table and a cubic (the
// published Szabolcs Nagy / Arm "optimized routines" algorithm).  The functions
// below restate that published algorithm so that device results equal the host
// libm's for the argument ranges this path uses; tests/test_math_port.py checks
// the host compilation of this same header against libm over dense ranges.
//
// The file compiles as plain C++ (host test) and as CUDA (device code).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD static inline
#include <math.h>
#endif

namespace sbm {

// 2^(i/32) as double bits, minus (i << 47) so that adding (k << 47) yields 2^(k/32)
#define SB_EXP2_TAB_INIT                                                                         \
    {0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
     0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
     0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
     0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
     0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
     0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
     0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
     0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL}

#if defined(__CUDACC__)
// device copy of the table; kernels stage it into shared memory (lane-varying index)
static __device__ const uint64_t d_exp2_tab[32] = SB_EXP2_TAB_INIT;
#endif

SB_HD double bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
SB_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
SB_HD double fma_d(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
SB_HD double add_d(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
SB_HD double mul_d(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}

// core shared by expf and powf(2, .): s * (C0 r^3 + C1 r^2 + C2 r + 1) with
// s = 2^(k/32) looked up from `tab` (caller supplies the table: shared memory on
// the device, a static array on the host).
SB_HD float exp2_poly(const uint64_t* tab, uint64_t ki, double r, double c0, double c1, double c2) {
    uint64_t t = tab[ki & 31];
    t += ki << 47;
    double s = bits_to_double(t);
    double z = fma_d(c0, r, c1);
    double r2 = mul_d(r, r);
    double y = fma_d(c2, r, 1.0);
    y = fma_d(z, r2, y);
    y = mul_d(y, s);
    return (float)y;
}

// expf(x) for |x| < 88 (no overflow/underflow handling: the path only feeds
// arguments in [-60, 0]).
SB_HD float expf_glibc(const uint64_t* tab, float x) {
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    double xd = (double)x;
    double z = mul_d(InvLn2N, xd);
    double kd = add_d(z, SHIFT);
    uint64_t ki = double_to_bits(kd);
    kd = add_d(kd, -SHIFT);
    double r = add_d(z, -kd);
    return exp2_poly(tab, ki, r, C0, C1, C2);
}

// powf(2.0f, y) for |y| < 126: glibc's log2 step returns exactly 1.0 for x = 2
// (table entry {invc = 1, logc = 0}, k = 1), leaving exp2 of (double)y.
SB_HD float pow2f_glibc(const uint64_t* tab, float y) {
    const double SHIFT = 0x1.8p+52 / 32.0;
    const double C0 = 0x1.c6af84b912394p-5;
    const double C1 = 0x1.ebfce50fac4f3p-3;
    const double C2 = 0x1.62e42ff0c52d6p-1;
    double xd = (double)y;
    double kd = add_d(xd, SHIFT);
    uint64_t ki = double_to_bits(kd);
    kd = add_d(kd, -SHIFT);
    double r = add_d(xd, -kd);
    return exp2_poly(tab, ki, r, C0, C1, C2);
}

}  // namespace sbm
