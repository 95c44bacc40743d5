// sb_common.cuh -- layout structs and constants shared by the kernels and the host runtime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace sb {

// src/lib.rs:92-112, 179-193
constexpr int SCALES_PER_OCTAVE = 3;
constexpr int N_LAYERS = SCALES_PER_OCTAVE + 3;      // Gaussian layers per octave
constexpr int N_DOG = SCALES_PER_OCTAVE + 2;
constexpr int IMAGE_BORDER = 5;
constexpr int ORI_BINS = 36;
constexpr int MAX_ORI = ORI_BINS / 2;                // a circular 36-bin histogram has <= 18 strict local maxima
constexpr int DESC_SIZE = 128;
constexpr int MAX_OCT = 16;

// Processing flavour (the crate's `P: Processing` type parameter, src/lib.rs:76-90): which blur / resize arithmetic
// builds the pyramid.  0 = OpenCVProcessing (src/opencv_processing.rs, the one the crate's test and snapshots pin),
// 1 = ImageprocProcessing (src/lib.rs:992-1007, what the crate's sift() defaults to).
constexpr int FL_OPENCV = 0;
constexpr int FL_IMAGEPROC = 1;

// Blur radii of the six Gaussian kernels (index 0 = seed blur, 1..5 = octave layers).
//   OpenCV:    ksize = round(8 sigma + 1) | 1           -> 11, 11, 13, 17, 21, 27 taps
//   imageproc: radius = ceil(2 sigma), sigma as f32     ->  7,  7,  9,  9, 11, 15 taps
__host__ __device__ constexpr int blur_radius(int l, int fl = FL_OPENCV) {
    return fl == FL_OPENCV ? (l == 0 ? 5 : l == 1 ? 5 : l == 2 ? 6 : l == 3 ? 8 : l == 4 ? 10 : 13)
                           : (l == 0 ? 3 : l == 1 ? 3 : l == 2 ? 4 : l == 3 ? 4 : l == 4 ? 5 : 7);
}
constexpr int MAX_TAPS = 27;

// One octave of one image inside the per-image arenas.
//   Gaussian layer l:  gauss + off + l*layer_stride, row pitch `pitch` floats (multiple of 32)
//   extrema mask:      mask + mask_off + ((s-1)*h + y)*mask_pitch words; strip j (60 columns) owns words
//                      2j, 2j+1: bit l of word 2j+k <=> column 60j - 2 + 2l + k (sb_pyramid.cuh, k_extrema)
//   row counters:      rows + row_base + (s-1)*h + y
struct OctLayout {
    int w, h, pitch, mask_pitch;
    int row_base;
    int scanned;  // src/lib.rs:315-317: octaves smaller than 10 px are blurred but never scanned
    long long off, layer_stride, mask_off;
};

struct PyrLayout {
    int n_oct;
    int img_rows;              // row counters per image
    long long img_floats;      // Gaussian arena floats per image
    long long img_mask_words;  // mask words per image
    OctLayout o[MAX_OCT];
};

// candidate key: octave | scale | y:16 | x:16  (natural order == integer order); 64-bit so that seed images of up to
// 65535 pixels per side fit (SB200_MAX_DIM = 8192 input pixels -> 16384-pixel seed image)
typedef unsigned long long CandKey;
__host__ __device__ inline CandKey pack_key(int o, int s, int y, int x) {
    return ((CandKey)o << 34) | ((CandKey)s << 32) | ((CandKey)y << 16) | (CandKey)x;
}
__host__ __device__ inline void unpack_key(CandKey k, int& o, int& s, int& y, int& x) {
    o = (int)(k >> 34);
    s = (int)((k >> 32) & 3u);
    y = (int)((k >> 16) & 0xffffu);
    x = (int)(k & 0xffffu);
}

// refined scale-space point (output of the refinement kernel, one per candidate)
struct Refined {
    float x, y;        // seed-image coordinates: (px + off_x) * 2^octave   (src/lib.rs:376-377)
    float size;        // kp_scale * 2^octave                                (:422)
    float response;    // |contrast|                                         (:357-358)
    float kp_scale;    // sigma in octave pixels                             (:372-374)
    int px, py;        // refined integer position in the octave
    int octave_scale;  // octave << 8 | scale (layer index); -1 when rejected
};

// SiftKeyPoint as kept on the device (src/lib.rs:58-68)
struct DevKeyPoint {
    float x, y, size, angle, response;
    int octave, scale;
    int pad;
};

// KeyPoint as returned to the caller (src/lib.rs:48-56); same layout as sb200_keypoint
struct OutKeyPoint {
    float x, y, size, angle, response;
};

// Debug build (-DSB_BOUNDS_CHECK, tools/bounds_check.sh): the data-dependent global loads of the keypoint kernels
// (refinement, orientation, descriptor) verify their address against the extent of the layer they read and trap with a
// message when it lies outside -- compute-sanitizer is not available on the GPU pool.
#ifdef SB_BOUNDS_CHECK
#define SB_CHECK_LOAD(ptr, base, elems, what)                                                                       \
    do {                                                                                                            \
        if ((ptr) < (base) || (ptr) >= (base) + (elems)) {                                                          \
            printf("SB_BOUNDS_CHECK: %s reads element %lld of a %lld-element layer\n", what,                        \
                   (long long)((ptr) - (base)), (long long)(elems));                                                \
            __trap();                                                                                               \
        }                                                                                                           \
    } while (0)
#else
#define SB_CHECK_LOAD(ptr, base, elems, what) do { } while (0)
#endif

// Programmatic dependent launch (griddepcontrol.wait; SASS ACQBULK / DEPBAR on the grid dependency): the kernels of a
// group are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs may be scheduled while
// the kernel before it in the stream (or graph chain) is still draining; nothing that kernel wrote -- or still reads --
// may be touched before this returns, so it is the first statement of every kernel.  No-op for a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Blackwell packed single precision (SASS FFMA2 / FADD2 / FMUL2): two independent IEEE round-to-nearest
// operations per instruction, i.e. the same bits as the scalar fmaf / + / * -- but half the issue slots,
// which is what bounds the wide-tap blurs.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return d;
}

__device__ __forceinline__ float2 sub2(float2 a, float2 b) {   // FADD2 with a negated operand: a - b, both halves
    float2 d;
    asm("sub.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return d;
}

}  // namespace sb
