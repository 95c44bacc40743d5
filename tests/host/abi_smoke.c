/* abi_smoke.c -- the C ABI of include/sift_b200.h exercised from a plain C program (no Python, no ctypes): the
 * reference-side binding in its simplest form.  Reads a raw u8 gray image, runs
 *     sb200_create -> sb200_extract -> sb200_extract_batch -> sb200_precompute + sb200_extract_precomputed
 *     -> sb200_compute_descriptors -> sb200_set_processing -> sb200_extract
 * and writes everything it got to a binary file that the pytest wrapper diffs against the ctypes results.
 *
 *     abi_smoke IMAGE.raw WIDTH HEIGHT OUT.bin
 * Output layout (little endian): for each of the 4 result blocks { u64 n; n x sb200_keypoint; n x 128 u8 },
 * then 128 descriptor bytes of the compute_descriptors call.  Exit code 0 on success, 1 + message otherwise. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sift_b200.h"

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int st__ = (call);                                                                           \
        if (st__ != SB200_OK) {                                                                      \
            fprintf(stderr, "%s -> %d (%s): %s\n", #call, st__, sb200_status_string(st__), sb200_last_error(ctx)); \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

static void dump(FILE* f, const sb200_result* r, uint32_t image) {
    uint64_t a = r->offsets[image], b = r->offsets[image + 1], n = b - a;
    fwrite(&n, sizeof n, 1, f);
    fwrite(r->keypoints + a, sizeof(sb200_keypoint), n, f);
    fwrite(r->descriptors + a * SB200_DESC_SIZE, SB200_DESC_SIZE, n, f);
}

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: %s IMAGE.raw WIDTH HEIGHT OUT.bin\n", argv[0]); return 2; }
    const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]);
    uint8_t* img = (uint8_t*)malloc((size_t)w * h * 2);
    FILE* in = fopen(argv[1], "rb");
    if (!in || fread(img, 1, (size_t)w * h, in) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(in);
    memcpy(img + (size_t)w * h, img, (size_t)w * h);   /* a batch of two: the image twice */
    FILE* out = fopen(argv[4], "wb");
    if (!out) return 2;

    sb200_ctx* ctx = NULL;
    if (sb200_device_count() < 1) { fprintf(stderr, "no CUDA device (there is no CPU fallback)\n"); return 1; }
    {
        int st = sb200_create(0, w, h, 2, 0, &ctx);
        if (st != SB200_OK) { fprintf(stderr, "sb200_create -> %d (%s)\n", st, sb200_status_string(st)); return 1; }
    }
    sb200_result r;
    CHECK(sb200_extract(ctx, img, w, h, w, -1, &r));                                   /* block 0: sift_with_processing::<OpenCVProcessing> */
    dump(out, &r, 0);
    CHECK(sb200_extract_batch(ctx, img, 2, w, h, w, (uint64_t)w * h, 50, &r));         /* block 1: batch of two, features_limit 50 */
    if (r.n_images != 2) { fprintf(stderr, "n_images %u\n", r.n_images); return 1; }
    dump(out, &r, 1);
    CHECK(sb200_precompute(ctx, img, w, h, w));                                        /* block 2: staged API */
    uint32_t n_oct = 0;
    CHECK(sb200_pyramid_info(ctx, &n_oct, NULL, NULL, 0));
    if (n_oct < 1) { fprintf(stderr, "n_octaves %u\n", n_oct); return 1; }
    CHECK(sb200_extract_precomputed(ctx, -1, &r));
    dump(out, &r, 0);
    CHECK(sb200_set_processing(ctx, SB200_PROCESSING_IMAGEPROC));                      /* block 3: the crate's sift() */
    CHECK(sb200_extract(ctx, img, w, h, w, -1, &r));
    dump(out, &r, 0);
    CHECK(sb200_set_processing(ctx, SB200_PROCESSING_OPENCV));
    {   /* compute_descriptor on the raw image as f32, benches/descriptor.rs shape */
        float* f = (float*)malloc((size_t)w * h * sizeof(float));
        for (size_t i = 0; i < (size_t)w * h; i++) f[i] = (float)img[i] / 255.0f;
        sb200_desc_in k = {100.0f, 100.0f, 2.1f, 123.0f};
        uint8_t d[SB200_DESC_SIZE];
        CHECK(sb200_compute_descriptors(ctx, f, w, h, w, &k, 1, d));
        fwrite(d, 1, SB200_DESC_SIZE, out);
        free(f);
    }
    fclose(out);
    printf("abi_smoke ok: %llu kernel launches\n", (unsigned long long)sb200_launch_count(ctx));
    sb200_destroy(ctx);
    free(img);
    return 0;
}
