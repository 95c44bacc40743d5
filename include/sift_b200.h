/*
 * sift_b200.h -- C ABI of the B200-native SIFT extractor (libsift_b200.so).
 *
 * This is the drop-in boundary for the extraction path of the Rust crate
 * tnibler/sift-features.  Every entry point names the reference interface it
 * replaces (paths relative to the crate root).  Plain pointers and sizes only;
 * no exceptions or aborts cross this boundary -- every call returns a status
 * code and sb200_last_error() explains a failure.
 *
 * Semantics are the crate's (OpenCV-compatible SIFT, OpenCVProcessing flavour of
 * blur/resize, src/opencv_processing.rs):
 *   - images are 8-bit gray, row-major, `stride` bytes between rows
 *     (image::GrayImage, src/lib.rs:71);
 *   - KeyPoint x/y/size are in input-image pixels, size is sigma (half of
 *     OpenCV's KeyPoint::size), angle in degrees (0,360] (src/lib.rs:48-56,164-174);
 *   - descriptors are (n,128) u8 row-major in keypoint order (src/lib.rs:39-46);
 *   - keypoints come in the crate's natural order: octave, scale, y, x of the
 *     initial extremum, then ascending orientation bin (src/lib.rs:155,287-293,
 *     324-334,397); duplicates are kept, exactly as the crate does;
 *   - with a features_limit smaller than the keypoint count the result is the
 *     `limit` strongest by response, strongest first (src/lib.rs:156-161); ties
 *     (unspecified by the crate's sort_unstable) are broken by natural order.
 *
 * Threading: one context per (host thread, device).  Calls on one context are
 * not re-entrant; distinct contexts are independent.
 *
 * There is no CPU fallback: creating a context without a CUDA device fails
 * with SB200_E_CUDA.
 */
#ifndef SIFT_B200_H
#define SIFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_DESC_SIZE 128   /* DESCRIPTOR_SIZE, src/lib.rs:111-112 */
#define SB200_MAX_OCTAVES 16
#define SB200_MAX_DIM 8192    /* max input width/height (seed image <= 16384: an 8K frame, 7680 x 4320, fits) */

/* status codes */
#define SB200_OK 0
#define SB200_E_INVALID 1   /* bad argument (null pointer, zero size, image larger than the context) */
#define SB200_E_CUDA 2      /* CUDA runtime / driver error, no device, out of memory */
#define SB200_E_CAPACITY 3  /* an output buffer / the device-resident result is too small (the host entry points grow
                               the context's own per-candidate arrays transparently and never return this) */
#define SB200_E_STATE 4     /* call needs a prior sb200_precompute / extract on this context */
#define SB200_E_UNSUPPORTED 5 /* an optional component is missing on this machine (nvJPEG for the JPEG entry points) */

typedef struct sb200_ctx sb200_ctx;

/* KeyPoint, src/lib.rs:48-56 */
typedef struct {
    float x, y, size, angle, response;
} sb200_keypoint;

/* SiftKeyPoint, src/lib.rs:58-68 (seed-image coordinates; debugging / parity only) */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, scale;
} sb200_sift_keypoint;

/* arguments of compute_descriptor, src/lib.rs:785-791 */
typedef struct {
    float x, y, scale, orientation;
} sb200_desc_in;

/* one mutual nearest-neighbour pair of sb200_match_descriptors (cv::DMatch of examples/sift-match.rs:30-35):
 * row indices into the query / train descriptor matrices and the squared L2 distance (exact integer) */
typedef struct {
    uint32_t query, train, dist2;
} sb200_dmatch;

/* initial discrete extremum, src/lib.rs:324-332 (parity tests) */
typedef struct {
    int32_t octave, scale, y, x;
} sb200_candidate;

/* SiftResult, src/lib.rs:39-46, for a batch of images.  All pointers refer to
 * pinned host memory owned by the context; they stay valid until the next
 * extract call on the same context or sb200_destroy. */
typedef struct {
    uint64_t n;                     /* total keypoints */
    uint32_t n_images;
    const uint64_t* offsets;        /* n_images+1; image i owns [offsets[i], offsets[i+1]) */
    const sb200_keypoint* keypoints;/* n */
    const uint8_t* descriptors;     /* n x 128 */
} sb200_result;

/* ---- lifetime ---------------------------------------------------------- */

/* Creates a context on CUDA device `device` able to process batches of up to
 * `max_batch` images of up to max_w x max_h pixels per launch group.
 * max_keypoints_per_image = 0 picks a default of max(16384, max_w*max_h/8). */
int sb200_create(int device, uint32_t max_w, uint32_t max_h, uint32_t max_batch,
                 uint32_t max_keypoints_per_image, sb200_ctx** out);
void sb200_destroy(sb200_ctx* ctx);
const char* sb200_last_error(const sb200_ctx* ctx);
const char* sb200_status_string(int status);
/* number of CUDA devices visible, or a negative status */
int sb200_device_count(void);

/* ---- Processing flavour: the type parameter P of sift_with_processing::<P>, src/lib.rs:76-90 ----
 * Which blur / resize arithmetic builds the Gaussian pyramid.  A context starts with SB200_PROCESSING_OPENCV.
 *   SB200_PROCESSING_OPENCV     OpenCVProcessing (src/opencv_processing.rs:39-74): the flavour the crate's test and
 *                               insta snapshots use; pinned bit-for-bit against OpenCV 4.13.
 *   SB200_PROCESSING_IMAGEPROC  ImageprocProcessing (src/lib.rs:992-1007) -- what the crate's plain sift() means
 *                               (src/lib.rs:71-73).  Restated from the published algorithms of imageproc 0.25 /
 *                               image 0.25, whose sources are not part of the reference tree and which no reference
 *                               test exercises: PARITY UNPINNED (bit-exact against this repository's oracle only).
 * The call drops the resident pyramid and takes effect from the next extract / precompute call. */
#define SB200_PROCESSING_OPENCV 0
#define SB200_PROCESSING_IMAGEPROC 1
int sb200_set_processing(sb200_ctx* ctx, int processing);
int sb200_get_processing(const sb200_ctx* ctx);   /* the flavour, or a negative status */

/* ---- optional OpenCV-style post-filters of the host results ----
 * What cv::SIFT::detectAndCompute does after detection and the crate does not (the comparison target of the crate's
 * own bench, benches/sift.rs:99-113 `opencv_sift`): remove_duplicates != 0 drops keypoints that repeat (x, y, size,
 * angle) of another one (KeyPointsFilter::removeDuplicatedSorted; the survivors come out in its sorted order: x, y
 * ascending, size descending, angle ascending); retain_best = n >= 0 keeps, when more than n remain, every keypoint
 * whose response is at least the n-th largest (KeyPointsFilter::retainBest: ties at the boundary survive), in the
 * order the first step left; retain_best < 0 switches it off.  Both are off by default (the crate's behaviour) and
 * apply to the host-result entry points (extract, extract_batch*, extract_precomputed), not to device-resident
 * results.  Independent of features_limit, which is the crate's own truncation (strongest first). */
int sb200_set_postfilter(sb200_ctx* ctx, int remove_duplicates, int64_t retain_best);

/* ---- extraction: sift_with_processing::<P>() with the context's flavour P ----
 * src/lib.rs:71-81.  features_limit < 0 means None.  (The crate's sift() is sift_with_processing::<ImageprocProcessing>;
 * the host mirrors -- Python, C++, Rust -- select that flavour for their sift().) */
int sb200_extract(sb200_ctx* ctx, const uint8_t* gray, uint32_t w, uint32_t h, uint32_t stride,
                  int64_t features_limit, sb200_result* out);

/* n images of identical size, image i at gray + i*image_stride (host memory,
 * pageable or pinned).  n may exceed max_batch: the context pipelines groups
 * of max_batch images (upload of group k+1 overlaps compute of group k). */
int sb200_extract_batch(sb200_ctx* ctx, const uint8_t* gray, uint32_t n, uint32_t w, uint32_t h,
                        uint32_t stride, uint64_t image_stride, int64_t features_limit,
                        sb200_result* out);

/* Same, but `d_gray` is DEVICE memory on the context's device and the results
 * stay in device memory (no host copies in either direction): enqueues the
 * whole pipeline on the context's stream for n <= max_batch images.  Counts
 * are available after sb200_sync via sb200_device_result. */
int sb200_extract_batch_device(sb200_ctx* ctx, const uint8_t* d_gray, uint32_t n, uint32_t w,
                               uint32_t h, uint32_t stride, uint64_t image_stride,
                               int64_t features_limit);
/* The pyramid stages alone (seed, five blurs per octave + decimation, DoG/extrema masks) for a device-resident
 * batch, on alternating slots like sb200_extract_batch_device: precompute_images (src/lib.rs:131-143) for n images
 * at once, and the measurement hook behind bench.py's "pyramid_dog pipelined" figure. */
int sb200_pyramid_batch_device(sb200_ctx* ctx, const uint8_t* d_gray, uint32_t n, uint32_t w, uint32_t h,
                               uint32_t stride, uint64_t image_stride);
/* device-side view of the last sb200_extract_batch_device: per-image keypoint
 * counts are copied to `counts` (n entries, host); d_keypoints / d_descriptors
 * receive the device base pointers of the dense, batch-wide result arrays (image
 * i's keypoints start at the exclusive prefix sum of the counts). */
int sb200_device_result(sb200_ctx* ctx, uint32_t* counts, uint32_t n, const sb200_keypoint** d_keypoints,
                        const uint8_t** d_descriptors, uint32_t* capacity_per_image);
int sb200_sync(sb200_ctx* ctx);

/* ---- staged API: precompute_images() + sift_with_precomputed() ----------
 * src/lib.rs:131-143 and :147-177.  The pyramid stays resident in the context. */
int sb200_precompute(sb200_ctx* ctx, const uint8_t* gray, uint32_t w, uint32_t h, uint32_t stride);
int sb200_extract_precomputed(sb200_ctx* ctx, int64_t features_limit, sb200_result* out);
/* PrecomputedImages accessors (src/lib.rs:124-128): octave count and sizes,
 * one Gaussian layer (0..5) or DoG layer (0..4) copied to host as dense (h,w) f32. */
int sb200_pyramid_info(sb200_ctx* ctx, uint32_t* n_octaves, uint32_t* widths, uint32_t* heights,
                       uint32_t cap);
int sb200_pyramid_layer(sb200_ctx* ctx, uint32_t octave, uint32_t layer, float* out);
int sb200_pyramid_dog(sb200_ctx* ctx, uint32_t octave, uint32_t layer, float* out);

/* ---- parity/debug views of the last single-image extract / precomputed run */
/* candidates in natural order; *n receives the total, up to cap are written */
int sb200_last_candidates(sb200_ctx* ctx, sb200_candidate* out, uint64_t cap, uint64_t* n);
/* SiftKeyPoints (natural order, before any features_limit) */
int sb200_last_sift_keypoints(sb200_ctx* ctx, sb200_sift_keypoint* out, uint64_t cap, uint64_t* n);

/* ---- descriptor only: compute_descriptor(), src/lib.rs:785-990 ----------
 * (the shape benches/descriptor.rs times).  img is a dense f32 image with
 * `stride` floats per row; out receives n x 128 bytes. */
int sb200_compute_descriptors(sb200_ctx* ctx, const float* img, uint32_t w, uint32_t h, uint32_t stride,
                              const sb200_desc_in* kps, uint64_t n, uint8_t* out);
/* device-pointer variant: everything already in device memory, asynchronous */
int sb200_compute_descriptors_device(sb200_ctx* ctx, const float* d_img, uint32_t w, uint32_t h,
                                     uint32_t stride, const sb200_desc_in* d_kps, uint64_t n,
                                     uint8_t* d_out);

/* ---- input prep: the step before the path in the reference's callers ----
 * examples/run-sift.rs:8, examples/sift-match.rs:49, src/lib.rs:1012 turn the decoded image into a GrayImage with the
 * `image` crate's grayscale(): L = (2126 R + 7152 G + 722 B) / 10000 in integer arithmetic (image 0.25; the crate's
 * source is not part of the reference tree, so this formula is restated from its published algorithm).
 * sb200_extract_batch_rgb takes n interleaved 8-bit RGB (channels = 3) or RGBA (channels = 4, alpha ignored) images
 * of `stride` bytes per row, converts them on the device and runs the extraction path on the result. */
int sb200_extract_batch_rgb(sb200_ctx* ctx, const uint8_t* rgb, uint32_t n, uint32_t w, uint32_t h, uint32_t stride,
                            uint64_t image_stride, uint32_t channels, int64_t features_limit, sb200_result* out);
/* the conversion alone (parity view): gray receives w x h bytes */
int sb200_rgb_to_luma(sb200_ctx* ctx, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t stride, uint32_t channels,
                      uint8_t* gray);

/* JPEG input: what `image::open(path)?.grayscale()` + sift() do in examples/run-sift.rs:8-19 and
 * `image::load_from_memory(..).grayscale()` in src/lib.rs:1012, with the decode on the device (nvJPEG batched decode;
 * hardware engines, then GPU Huffman, then nvJPEG's default backend) followed by the same integer luma; a one-component
 * JPEG is its Y plane, as the `image` crate returns it.  Only the compressed bytes cross the host/device boundary.
 * All n streams of one call must have the same frame size (like sb200_extract_batch); lengths in bytes.
 * Decoded pixels are the decoder's: nvJPEG, libjpeg-turbo and the crate's zune-jpeg differ by a grey level or two on
 * the same stream (IDCT rounding, chroma upsampling), so parity through this entry point is a tolerance, and exact only
 * against sb200_decode_jpeg_luma's own pixels.  SB200_E_UNSUPPORTED when libnvjpeg cannot be loaded. */
int sb200_extract_batch_jpeg(sb200_ctx* ctx, const uint8_t* const* jpegs, const uint64_t* lengths, uint32_t n,
                             int64_t features_limit, sb200_result* out);
/* frame size and component count from the header */
int sb200_jpeg_info(sb200_ctx* ctx, const uint8_t* jpeg, uint64_t length, uint32_t* w, uint32_t* h, uint32_t* components);
/* the decode + luma step alone (parity view): gray receives w x h bytes, capacity = size of the buffer */
int sb200_decode_jpeg_luma(sb200_ctx* ctx, const uint8_t* jpeg, uint64_t length, uint8_t* gray, uint64_t capacity);
/* nvJPEG backend that decoded the last group: "hardware", "gpu", "default", or "none" */
const char* sb200_jpeg_backend(const sb200_ctx* ctx);

/* ---- descriptor matching: the step after the path in the reference's examples ----
 * examples/sift-match.rs:30-35 and examples/opencv-cross-match.rs:34-43 hand the (N,128) u8 descriptor
 * matrices to OpenCV's BFMatcher(NORM_L2, crossCheck = true).  sb200_match_descriptors computes the same mutual nearest
 * neighbours on the GPU (u8 x u8 Gram matrix on the tensor cores, exact integer distances): for every query
 * row the train row with the smallest squared L2 distance (smallest index on ties), kept when the query row
 * is in turn the nearest of that train row.  Matches are written in ascending query order; *n_out receives
 * their number (SB200_E_CAPACITY if it exceeds cap; n_query always suffices).  Rows of `out` beyond *n_out, up to
 * min(cap, n_query), are overwritten with unspecified values. */
int sb200_match_descriptors(sb200_ctx* ctx, const uint8_t* query, uint64_t n_query, const uint8_t* train, uint64_t n_train,
                sb200_dmatch* out, uint64_t cap, uint64_t* n_out);
/* device-pointer variant: descriptor matrices already in device memory (16-byte aligned, e.g. the
 * d_descriptors of sb200_device_result); `out` is host memory */
int sb200_match_descriptors_device(sb200_ctx* ctx, const uint8_t* d_query, uint64_t n_query, const uint8_t* d_train,
                       uint64_t n_train, sb200_dmatch* out, uint64_t cap, uint64_t* n_out);

/* ---- multi-GPU: contiguous shards of a batch over several contexts ------
 * One host thread per context; image i goes to context i / ceil(n/n_ctx).  No device-to-device traffic and no
 * collective: images are independent.
 *
 * sb200_extract_batch_multi_parts is the zero-copy form: parts[d] (n_ctx entries) receives context d's own
 * result for its shard -- offsets local to the part, pointers owned by ctxs[d] -- and first_image[d] (n_ctx + 1
 * entries, may be NULL) the index of the part's first image, so image i of part d is image first_image[d] + i of
 * the batch.  The parts in order ARE the batch in image order; nothing is copied on the host.
 *
 * sb200_extract_batch_multi additionally concatenates the parts into one dense result owned by ctxs[0] (a
 * multi-threaded host copy whose duration sb200_last_gather_ms reports). */
int sb200_extract_batch_multi_parts(sb200_ctx* const* ctxs, uint32_t n_ctx, const uint8_t* gray, uint32_t n,
                                    uint32_t w, uint32_t h, uint32_t stride, uint64_t image_stride,
                                    int64_t features_limit, sb200_result* parts, uint64_t* first_image);
int sb200_extract_batch_multi(sb200_ctx* const* ctxs, uint32_t n_ctx, const uint8_t* gray, uint32_t n,
                              uint32_t w, uint32_t h, uint32_t stride, uint64_t image_stride,
                              int64_t features_limit, sb200_result* out);
double sb200_last_gather_ms(const sb200_ctx* ctx);
/* Host wall time (ms) this context's shard took in the last sb200_extract_batch_multi / _multi_parts call: the
 * slowest shard is the call's critical path. */
double sb200_last_shard_ms(const sb200_ctx* ctx);

/* ---- measurement hooks -------------------------------------------------- */
#define SB200_STAGE_SEED 0        /* u8 -> 2x upsample -> seed blur          (src/lib.rs:196-210) */
#define SB200_STAGE_BLUR 1        /* 5 blurs/octave + decimation             (:213-267) */
#define SB200_STAGE_EXTREMA 2     /* DoG + 3x3x3 extrema + ordered compaction (:271-279, 437-506) */
#define SB200_STAGE_REFINE 3      /* refinement, contrast, edge              (:525-653) */
#define SB200_STAGE_ORIENT 4      /* orientation histogram + peaks           (:657-757, 389-431) */
#define SB200_STAGE_DESCRIPTOR 5  /* descriptors + output pack               (:759-990, 164-174) */
#define SB200_STAGE_TOP_BLUR 6    /* the single heaviest launch: 27-tap blur of octave 0 (also counted in BLUR) */
#define SB200_STAGE_COUNT 7
/* With profiling on, every stage of the next extract calls is bracketed by
 * CUDA events on the launching stream(s) (this serialises the stages). */
int sb200_set_profiling(sb200_ctx* ctx, int on);
/* accumulated device milliseconds and kernel launches per stage since the last reset */
int sb200_stage_stats(sb200_ctx* ctx, double* ms, uint64_t* launches, uint32_t cap);
int sb200_reset_stats(sb200_ctx* ctx);
/* per-launch view of the pyramid stages, accumulated like the stage stats while profiling is on: slot
 * octave * 8 + k, where k = 0 is the 2x upsample + seed blur (octave 0 only), k = 1..5 the blur that writes
 * Gaussian layer k, k = 6 the DoG/extrema scan of the octave and k = 7 the fused launch for all the small
 * octaves (filed under the first octave it covers). */
#define SB200_FINE_SLOTS (SB200_MAX_OCTAVES * 8)
int sb200_launch_stats(sb200_ctx* ctx, double* ms, uint64_t* launches, uint32_t cap);
/* total kernel launches issued by this context since creation */
uint64_t sb200_launch_count(const sb200_ctx* ctx);
const char* sb200_stage_name(uint32_t stage);
/* algorithmic bytes of the pyramid + DoG/extrema stages for one w x h image
 * (SURVEY.md section 8d, A(W,H)); stage-wise split in out[0..2] = seed, blur+decimate, extrema */
uint64_t sb200_algorithmic_bytes(uint32_t w, uint32_t h, uint64_t* out, uint32_t cap);

/* CUDA-event timer on the context's main stream (all work of a call is ordered
 * on it): start/stop return immediately, elapsed synchronises. */
int sb200_timer_start(sb200_ctx* ctx);
int sb200_timer_stop(sb200_ctx* ctx);
int sb200_timer_elapsed_ms(sb200_ctx* ctx, float* ms);

/* pinned host memory for callers that want zero-copy uploads */
int sb200_host_alloc(size_t bytes, void** out);
int sb200_host_free(void* p);
/* device memory helpers for callers without their own CUDA allocator (bench/tests) */
int sb200_device_alloc(sb200_ctx* ctx, size_t bytes, void** out);
int sb200_device_free(sb200_ctx* ctx, void* p);
int sb200_memcpy_h2d(sb200_ctx* ctx, void* dst, const void* src, size_t bytes);
int sb200_memcpy_d2h(sb200_ctx* ctx, void* dst, const void* src, size_t bytes);
/* writes `bytes` of a device scratch buffer (>= L2 size) to evict L2 between timed steps */
int sb200_flush_l2(sb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SIFT_B200_H */
