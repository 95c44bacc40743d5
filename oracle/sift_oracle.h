/*
 * sift_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference crate's extraction path
 * (tnibler/sift-features, src/lib.rs) with the OpenCV "Processing" flavour
 * (src/opencv_processing.rs) that the crate's only test and its insta
 * snapshots pin.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (sift_features_b200/) never links, imports or calls it.
 *
 * Parity status (flavour A, the default of every function without a flavour argument): PINNED -- against (1) cv2.GaussianBlur / cv2.resize of
 * OpenCV 4.13 bit-for-bit on the pyramid arithmetic, (2) cv2.SIFT_create on
 * identical pixels, (3) the crate's four insta snapshots with the tolerance a
 * different JPEG decoder forces (see tests/test_oracle_golden.py).
 * Flavour B (ImageprocProcessing): PARITY UNPINNED, see below.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).
 */
#ifndef SIFT_ORACLE_H
#define SIFT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_MAX_OCTAVES 24
#define SO_LAYERS 6      /* SCALES_PER_OCTAVE + 3, src/lib.rs:92,221 */
#define SO_DOG_LAYERS 5  /* SCALES_PER_OCTAVE + 2, src/lib.rs:277 */
#define SO_DESC_SIZE 128 /* src/lib.rs:111-112 */

/* src/lib.rs:58-68 (SiftKeyPoint): seed-image coordinates. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, scale;
} so_sift_keypoint;

/* src/lib.rs:48-56 (KeyPoint): input-image coordinates, size = sigma. */
typedef struct {
    float x, y, size, angle, response;
} so_keypoint;

/* candidate = initial discrete extremum, src/lib.rs:324-332 */
typedef struct {
    int32_t octave, scale, y, x;
} so_candidate;

/* src/lib.rs:124-128 (PrecomputedImages) */
typedef struct so_pyramid so_pyramid;

/* ---- Processing flavour A (src/opencv_processing.rs:39-74) ---- */
int so_gaussian_ksize(double sigma);
/* writes ksize taps, returns ksize */
int so_gaussian_taps(double sigma, float* taps, int cap);
void so_gaussian_blur(const float* src, int w, int h, double sigma, float* dst);
void so_resize_linear_2x(const float* src, int w, int h, float* dst /* 2w x 2h */);
void so_resize_nearest_half(const float* src, int w, int h, float* dst /* (w/2) x (h/2) */);

/* ---- Processing flavour B (ImageprocProcessing, src/lib.rs:992-1007): PARITY UNPINNED ----
 * imageproc 0.25 / image 0.25 are not part of the reference tree and no reference test uses this flavour; these
 * functions restate the crates' published algorithms (see the block comment in sift_oracle.c). */
#define SO_PROCESSING_OPENCV 0
#define SO_PROCESSING_IMAGEPROC 1
int so_imageproc_taps(double sigma, float* taps, int cap);
void so_gaussian_blur_imageproc(const float* src, int w, int h, double sigma, float* dst);
void so_resize_triangle_2x(const float* src, int w, int h, float* dst /* 2w x 2h */);
void so_resize_nearest_imageproc(const float* src, int w, int h, float* dst /* (w/2) x (h/2) */);

/* the five per-octave sigmas of src/lib.rs:220-229 (index 1..5) and the seed sigma (:207) */
double so_seed_sigma(void);
double so_octave_sigma(int s);

/* ---- pyramid: src/lib.rs:131-143, 196-279 ---- */
so_pyramid* so_precompute(const uint8_t* gray, int w, int h, int stride);                         /* flavour A */
so_pyramid* so_precompute_flavour(const uint8_t* gray, int w, int h, int stride, int flavour);
void so_pyramid_free(so_pyramid* p);
int so_pyramid_octaves(const so_pyramid* p);
int so_pyramid_width(const so_pyramid* p, int octave);
int so_pyramid_height(const so_pyramid* p, int octave);
const float* so_pyramid_gauss(const so_pyramid* p, int octave, int layer);
const float* so_pyramid_dog(const so_pyramid* p, int octave, int layer);

/* ---- detector: src/lib.rs:281-757 ---- */
/* natural order (octave, scale, y, x); returns total count, writes min(count,cap) */
size_t so_find_candidates(const so_pyramid* p, so_candidate* out, size_t cap);
/* full detector: returns number of SiftKeyPoints (natural order), writes min(n,cap) */
size_t so_find_keypoints(const so_pyramid* p, so_sift_keypoint* out, size_t cap);

/* ---- descriptor: src/lib.rs:785-990 ---- */
void so_compute_descriptor(const float* img, int w, int h, float x, float y, float scale,
                           float orientation_deg, uint8_t out[SO_DESC_SIZE]);

/* ---- assembly: src/lib.rs:147-177 ---- */
/* features_limit < 0 == None.  Returns n; writes min(n,cap) keypoints + descriptors. */
size_t so_sift_with_precomputed(const so_pyramid* p, int64_t features_limit, so_keypoint* kps,
                                uint8_t* desc, size_t cap);
size_t so_sift(const uint8_t* gray, int w, int h, int stride, int64_t features_limit,
               so_keypoint* kps, uint8_t* desc, size_t cap);                                       /* flavour A */
size_t so_sift_flavour(const uint8_t* gray, int w, int h, int stride, int64_t features_limit, int flavour,
                       so_keypoint* kps, uint8_t* desc, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
