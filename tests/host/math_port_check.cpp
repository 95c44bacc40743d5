// Host check of sift_features_b200/csrc/sb_math.cuh against the host libm
// (the functions the reference crate reaches through f32::exp / f32::powf).
// Prints "<name> <tested> <mismatches>" per function; exit 0 always (the
// pytest wrapper asserts on the counts).
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "../../sift_features_b200/csrc/sb_math.cuh"

static const uint64_t TAB[32] = SB_EXP2_TAB_INIT;

int main() {
    // expf on every float in [-60, -2^-20] and a dense sweep near 0
    uint64_t n = 0, bad = 0;
    float lo = -60.0f, hi = -9.5367431640625e-07f;
    uint32_t blo, bhi;
    memcpy(&blo, &hi, 4);  // negative floats: larger magnitude = larger bits
    memcpy(&bhi, &lo, 4);
    for (uint32_t b = blo; b <= bhi; b += 7) {  // stride 7: ~37M samples
        float x;
        memcpy(&x, &b, 4);
        float a = sbm::expf_glibc(TAB, x), c = expf(x);
        n++;
        if (memcmp(&a, &c, 4) != 0) bad++;
    }
    printf("expf %llu %llu\n", (unsigned long long)n, (unsigned long long)bad);
    n = bad = 0;
    float plo = 0.05f, phi = 1.5f;
    memcpy(&blo, &plo, 4);
    memcpy(&bhi, &phi, 4);
    for (uint32_t b = blo; b <= bhi; b += 3) {
        float y;
        memcpy(&y, &b, 4);
        float a = sbm::pow2f_glibc(TAB, y), c = powf(2.0f, y);
        n++;
        if (memcmp(&a, &c, 4) != 0) bad++;
    }
    printf("pow2f %llu %llu\n", (unsigned long long)n, (unsigned long long)bad);
    return 0;
}
