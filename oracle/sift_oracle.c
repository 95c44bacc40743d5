/*
 * sift_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See sift_oracle.h for scope, parity status and who may load this.
 *
 * Arithmetic policy: f32 everywhere the reference uses f32, operations in the
 * order the reference writes them, no contraction (build with
 * -ffp-contract=off); fmaf() only where the OpenCV flavour of blur/resize uses
 * a fused multiply-add (established bit-for-bit against cv2 4.13, see
 * tests/test_oracle_cv2.py).
 */
#include "sift_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* constants: src/lib.rs:92-112, 179-193, 297, 516                     */
/* ------------------------------------------------------------------ */
#define SCALES_PER_OCTAVE 3
static const float CONTRAST_THRESHOLD = 0.04f;
static const float EDGE_THRESHOLD = 10.0f;
static const float ORIENTATION_HISTOGRAM_RADIUS = 1.5f;
#define IMAGE_BORDER 5
#define ORI_BINS 36
static const float LAMBDA_ORI = 1.5f;
static const float LAMBDA_DESCR = 3.0f;
#define N_HIST 4
#define N_BINS 8
static const double SIGMA_IN = 0.5;
static const double SIGMA_MIN = 0.8;
#define INV_DELTA_MIN 2
static const float DELTA_MIN = 0.5f;
static const float LOCALMAX_RATIO = 0.8f;
#define MAX_INTERPOLATION_STEPS 5

/* Rust `as i32` / `as isize` on floats: saturating, NaN -> 0. */
static int64_t rust_f32_as_i64(float v) {
    if (isnan(v)) return 0;
    if (v >= 9.2e18f) return INT64_MAX;
    if (v <= -9.2e18f) return INT64_MIN;
    return (int64_t)v;
}
static int32_t rust_f32_as_i32(float v) {
    if (isnan(v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
/* Rust `as usize`: saturating at 0 for negatives. */
static int64_t rust_f32_as_usize(float v) {
    if (isnan(v) || v <= 0.0f) return 0;
    if (v >= 9.2e18f) return INT64_MAX;
    return (int64_t)v;
}

/* ------------------------------------------------------------------ */
/* Processing flavour A: src/opencv_processing.rs                      */
/* ------------------------------------------------------------------ */

/* opencv_processing.rs:51-57: gaussian_blur_def(src, dst, Size::default(), sigma)
 * => OpenCV picks ksize = cvRound(sigma*4*2+1)|1 for CV_32F. */
int so_gaussian_ksize(double sigma) {
    return ((int)lrint(sigma * 8.0 + 1.0)) | 1;
}

/* OpenCV getGaussianKernel(ksize, sigma, CV_32F): exp(-x^2/(2 sigma^2)) in
 * double, normalised to unit sum, stored as f32. */
int so_gaussian_taps(double sigma, float* taps, int cap) {
    int ks = so_gaussian_ksize(sigma);
    if (ks > cap) return -ks;
    int r = ks / 2;
    double t[64];
    double scale2x = -0.5 / (sigma * sigma);
    double sum = 0.0;
    for (int i = 0; i < ks; i++) {
        double x = (double)(i - r);
        t[i] = exp(scale2x * x * x);
        sum += t[i];
    }
    double inv = 1.0 / sum;
    for (int i = 0; i < ks; i++) taps[i] = (float)(t[i] * inv);
    return ks;
}

/* BORDER_REFLECT_101 (OpenCV BORDER_DEFAULT, opencv_processing.rs:51 *_def) */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

/* Separable filter exactly as OpenCV's vectorised sepFilter2D evaluates it on
 * CV_32F: row pass = left-to-right FMA chain (RowVec_32f), column pass =
 * symmetric-folded FMA chain (SymmColumnVec_32f).  Row pass first. */
void so_gaussian_blur(const float* src, int w, int h, double sigma, float* dst) {
    float k[64];
    int ks = so_gaussian_taps(sigma, k, 64);
    int r = ks / 2;
    float* tmp = (float*)malloc((size_t)w * h * sizeof(float));
    float* pr = (float*)malloc((size_t)(w + 2 * r) * sizeof(float));
    for (int y = 0; y < h; y++) {
        const float* s = src + (size_t)y * w;
        float* t = tmp + (size_t)y * w;
        for (int i = 0; i < w + 2 * r; i++) pr[i] = s[reflect101(i - r, w)];
        /* per pixel: acc = x[-r]*k[0]; acc = fma(x[-r+i], k[i], acc), i = 1..ks-1
         * (loop order swapped so the compiler can vectorise across x; each
         * pixel's chain is unchanged) */
        for (int x = 0; x < w; x++) t[x] = pr[x] * k[0];
        for (int i = 1; i < ks; i++) {
            float ki = k[i];
            const float* q = pr + i;
            for (int x = 0; x < w; x++) t[x] = fmaf(q[x], ki, t[x]);
        }
    }
    for (int y = 0; y < h; y++) {
        float* d = dst + (size_t)y * w;
        const float* c = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) d[x] = c[x] * k[r];
        for (int i = 1; i <= r; i++) {
            const float* a = tmp + (size_t)reflect101(y + i, h) * w;
            const float* b = tmp + (size_t)reflect101(y - i, h) * w;
            float ki = k[r + i];
            for (int x = 0; x < w; x++) d[x] = fmaf(a[x] + b[x], ki, d[x]);
        }
    }
    free(pr);
    free(tmp);
}

/* opencv_processing.rs:66-68: resize(.., INTER_LINEAR) for the exact 2x case
 * of src/lib.rs:201-205.  Half-pixel centres, clamped at the borders; each
 * 1-D lerp is fma(b-a, t, a); horizontal pass, then vertical. */
static void lerp_index_2x(int d, int n, int* s0, int* s1, float* t) {
    /* (d + 0.5) * 0.5 - 0.5 = d/2 - 0.25 */
    int s = (d & 1) ? (d - 1) / 2 : d / 2 - 1;
    float f = (d & 1) ? 0.25f : 0.75f;
    if (s < 0) { s = 0; f = 0.0f; }
    if (s >= n - 1) { s = n - 1; f = 0.0f; }
    *s0 = s;
    *s1 = (s + 1 < n) ? s + 1 : n - 1;
    *t = f;
}

void so_resize_linear_2x(const float* src, int w, int h, float* dst) {
    int W = 2 * w, H = 2 * h;
    float* hbuf = (float*)malloc((size_t)W * h * sizeof(float));
    for (int y = 0; y < h; y++) {
        const float* s = src + (size_t)y * w;
        float* o = hbuf + (size_t)y * W;
        for (int x = 0; x < W; x++) {
            int a, b; float t;
            lerp_index_2x(x, w, &a, &b, &t);
            o[x] = fmaf(s[b] - s[a], t, s[a]);
        }
    }
    for (int y = 0; y < H; y++) {
        int a, b; float t;
        lerp_index_2x(y, h, &a, &b, &t);
        const float* r0 = hbuf + (size_t)a * W;
        const float* r1 = hbuf + (size_t)b * W;
        float* o = dst + (size_t)y * W;
        for (int x = 0; x < W; x++) o[x] = fmaf(r1[x] - r0[x], t, r0[x]);
    }
    free(hbuf);
}

/* opencv_processing.rs:70-72: resize(.., INTER_NEAREST) to (w/2, h/2)
 * (src/lib.rs:247): OpenCV samples floor(dst * src/dst) which is the even
 * source pixel for every size in the halving chain (verified against cv2). */
void so_resize_nearest_half(const float* src, int w, int h, float* dst) {
    int W = w / 2, H = h / 2;
    for (int y = 0; y < H; y++) {
        int sy = (int)floor((double)y * ((double)h / (double)H));
        if (sy > h - 1) sy = h - 1;
        for (int x = 0; x < W; x++) {
            int sx = (int)floor((double)x * ((double)w / (double)W));
            if (sx > w - 1) sx = w - 1;
            dst[(size_t)y * W + x] = src[(size_t)sy * w + sx];
        }
    }
}

/* ------------------------------------------------------------------ */
/* Processing flavour B: ImageprocProcessing, src/lib.rs:992-1007        */
/*                                                                      */
/* PARITY UNPINNED.  The arithmetic lives in two third-party crates     */
/* that are NOT part of /root/reference (Cargo.toml:13-14: image        */
/* ^0.25.2, imageproc ^0.25.0; no Cargo.lock, nothing vendored) and no  */
/* reference test, snapshot or bench exercises this flavour.  What      */
/* follows restates those crates' published algorithms (imageproc       */
/* filter::gaussian_blur_f32 -> separable_filter_equal; image           */
/* imageops::resize -> vertical_sample + horizontal_sample) operation   */
/* by operation; it pins the GPU path to THIS restatement, not to the   */
/* crates' binaries.                                                    */
/* ------------------------------------------------------------------ */

/* imageproc gaussian_kernel_f32(sigma): radius ceil(2 sigma); taps
 * gaussian_pdf(x) = (sigma * sqrt(2 pi)).recip() * exp(-x^2 / (2 sigma^2)), all f32, not renormalised. */
int so_imageproc_taps(double sigma64, float* taps, int cap) {
    float sigma = (float)sigma64; /* src/lib.rs:997: `sigma as f32` */
    int r = (int)ceilf(2.0f * sigma);
    int ks = 2 * r + 1;
    if (ks > cap) return -ks;
    float norm = 1.0f / (sigma * sqrtf(2.0f * 3.14159265358979323846f));
    for (int i = 0; i <= r; i++) {
        float x = (float)i;
        float v = norm * expf(-(x * x) / (2.0f * (sigma * sigma)));
        taps[r + i] = v;
        taps[r - i] = v;
    }
    return ks;
}

/* imageproc separable_filter_equal: horizontal_filter then vertical_filter; each output accumulates
 * acc = acc + pixel * weight from zero over the taps in order (no FMA); out-of-image taps read the clamped edge pixel. */
void so_gaussian_blur_imageproc(const float* src, int w, int h, double sigma, float* dst) {
    float k[64];
    int ks = so_imageproc_taps(sigma, k, 64);
    int r = ks / 2;
    float* tmp = (float*)malloc((size_t)w * h * sizeof(float));
    for (int y = 0; y < h; y++) {
        const float* s = src + (size_t)y * w;
        float* t = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float acc = 0.0f;
            for (int i = 0; i < ks; i++) {
                int xi = x + i - r;
                xi = xi < 0 ? 0 : (xi > w - 1 ? w - 1 : xi);
                acc = acc + s[xi] * k[i];
            }
            t[x] = acc;
        }
    }
    for (int y = 0; y < h; y++) {
        float* d = dst + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float acc = 0.0f;
            for (int i = 0; i < ks; i++) {
                int yi = y + i - r;
                yi = yi < 0 ? 0 : (yi > h - 1 ? h - 1 : yi);
                acc = acc + tmp[(size_t)yi * w + x] * k[i];
            }
            d[x] = acc;
        }
    }
    free(tmp);
}

/* image 0.25 imageops::sample: triangle_kernel / box_kernel and the two 1-D resampling passes, literally.
 * resize() runs vertical_sample first (into an f32 image, unclamped), then horizontal_sample (clamped to the
 * subpixel range, [0, 1] for f32). */
typedef float (*so_kernel_fn)(float);
static float so_triangle_kernel(float x) { return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f; }
static float so_box_kernel(float x) { (void)x; return 1.0f; }

static void so_sample_1d(const float* src, int w, int h, int n_new, int vertical, so_kernel_fn kernel, float support,
                         float* out) {
    const int n = vertical ? h : w;
    const float ratio = (float)n / (float)n_new;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float src_support = support * sratio;
    float ws[64];
    for (int o = 0; o < n_new; o++) {
        float input = ((float)o + 0.5f) * ratio;
        int64_t left = rust_f32_as_i64(floorf(input - src_support));
        if (left < 0) left = 0;
        if (left > n - 1) left = n - 1;
        int64_t right = rust_f32_as_i64(ceilf(input + src_support));
        if (right < left + 1) right = left + 1;
        if (right > n) right = n;
        input = input - 0.5f;
        float sum = 0.0f;
        int cnt = 0;
        for (int64_t i = left; i < right; i++) {
            float wgt = kernel(((float)i - input) / sratio);
            ws[cnt++] = wgt;
            sum += wgt;
        }
        for (int i = 0; i < cnt; i++) ws[i] /= sum;
        if (vertical) {
            for (int x = 0; x < w; x++) {
                float t = 0.0f;
                for (int i = 0; i < cnt; i++) t += src[(size_t)(left + i) * w + x] * ws[i];
                out[(size_t)o * w + x] = t;
            }
        } else {
            for (int y = 0; y < h; y++) {
                float t = 0.0f;
                for (int i = 0; i < cnt; i++) t += src[(size_t)y * w + (left + i)] * ws[i];
                t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t); /* clamp(t, S::DEFAULT_MIN_VALUE, S::DEFAULT_MAX_VALUE) */
                out[(size_t)y * n_new + o] = t;
            }
        }
    }
}

static void so_resize_image_crate(const float* src, int w, int h, int nw, int nh, so_kernel_fn kernel, float support,
                                  float* dst) {
    if (nw <= 0 || nh <= 0) return;
    float* tmp = (float*)malloc((size_t)w * nh * sizeof(float));
    so_sample_1d(src, w, h, nh, 1, kernel, support, tmp);
    so_sample_1d(tmp, w, nh, nw, 0, kernel, support, dst);
    free(tmp);
}

/* src/lib.rs:1001: resize(img, 2w, 2h, FilterType::Triangle) */
void so_resize_triangle_2x(const float* src, int w, int h, float* dst) {
    so_resize_image_crate(src, w, h, 2 * w, 2 * h, so_triangle_kernel, 1.0f, dst);
}
/* src/lib.rs:1005: resize(img, w/2, h/2, FilterType::Nearest) */
void so_resize_nearest_imageproc(const float* src, int w, int h, float* dst) {
    so_resize_image_crate(src, w, h, w / 2, h / 2, so_box_kernel, 0.0f, dst);
}

/* ------------------------------------------------------------------ */
/* pyramid: src/lib.rs:131-143, 196-279                                */
/* ------------------------------------------------------------------ */
struct so_pyramid {
    int n_octaves;
    int w[SO_MAX_OCTAVES], h[SO_MAX_OCTAVES];
    float* gauss[SO_MAX_OCTAVES]; /* (6,h,w) row-major, src/lib.rs:260-264 */
    float* dog[SO_MAX_OCTAVES];   /* (5,h,w), src/lib.rs:271-279 */
};

/* src/lib.rs:207 */
double so_seed_sigma(void) {
    return sqrt(SIGMA_MIN * SIGMA_MIN - SIGMA_IN * SIGMA_IN) * (double)INV_DELTA_MIN;
}

/* src/lib.rs:220-229 */
double so_octave_sigma(int s) {
    double m = pow(2.0, 2.0 / (double)SCALES_PER_OCTAVE);
    /* m.powi(s - 1): LLVM lowers powi to compiler-rt __powidf2 (square-and-multiply) */
    int e = s - 1;
    int recip = e < 0;
    double base = m, a = 1.0;
    int b_ = recip ? -e : e;
    for (;;) {
        if (b_ & 1) a *= base;
        b_ /= 2;
        if (b_ == 0) break;
        base *= base;
    }
    if (recip) a = 1.0 / a;
    double b = a * m;
    return sqrt(b - a) * SIGMA_MIN * (double)INV_DELTA_MIN;
}

so_pyramid* so_precompute(const uint8_t* gray, int w, int h, int stride) {
    return so_precompute_flavour(gray, w, h, stride, SO_PROCESSING_OPENCV);
}

/* precompute_images::<P>, src/lib.rs:131-143, with P selected at run time */
so_pyramid* so_precompute_flavour(const uint8_t* gray, int w, int h, int stride, int flavour) {
    const int B = flavour == SO_PROCESSING_IMAGEPROC;
    so_pyramid* p = (so_pyramid*)calloc(1, sizeof(so_pyramid));
    /* create_seed_image, src/lib.rs:196-210 */
    float* f = (float*)malloc((size_t)w * h * sizeof(float));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            f[(size_t)y * w + x] = (float)gray[(size_t)y * stride + x] / 255.0f; /* img.convert(), :198 */
    int W = w * INV_DELTA_MIN, H = h * INV_DELTA_MIN;
    float* up = (float*)malloc((size_t)W * H * sizeof(float));
    if (B) so_resize_triangle_2x(f, w, h, up);
    else so_resize_linear_2x(f, w, h, up);
    free(f);
    float* seed = (float*)malloc((size_t)W * H * sizeof(float));
    if (B) so_gaussian_blur_imageproc(up, W, H, so_seed_sigma(), seed);
    else so_gaussian_blur(up, W, H, so_seed_sigma(), seed);
    free(up);

    /* src/lib.rs:133-134 */
    int min_axis = W < H ? W : H;
    float lg = log2f((float)min_axis) - 2.0f;
    int64_t no = rust_f32_as_usize(roundf(lg)) + 1;
    if (no > SO_MAX_OCTAVES) no = SO_MAX_OCTAVES;
    p->n_octaves = (int)no;

    /* build_gaussian_scale_space, src/lib.rs:213-267 */
    int cw = W, ch = H;
    float* initial = seed;
    for (int o = 0; o < p->n_octaves; o++) {
        p->w[o] = cw; p->h[o] = ch;
        size_t px = (size_t)cw * ch;
        p->gauss[o] = (float*)malloc(px * SO_LAYERS * sizeof(float));
        memcpy(p->gauss[o], initial, px * sizeof(float));
        free(initial);
        for (int s = 1; s < SO_LAYERS; s++) { /* sigmas.iter().skip(1), :233 */
            if (B) so_gaussian_blur_imageproc(p->gauss[o] + px * (s - 1), cw, ch, so_octave_sigma(s), p->gauss[o] + px * s);
            else so_gaussian_blur(p->gauss[o] + px * (s - 1), cw, ch, so_octave_sigma(s), p->gauss[o] + px * s);
        }
        /* build_dog, :271-279 */
        p->dog[o] = (float*)malloc(px * SO_DOG_LAYERS * sizeof(float));
        for (int s = 0; s < SO_DOG_LAYERS; s++) {
            const float* a = p->gauss[o] + px * (s + 1);
            const float* b = p->gauss[o] + px * s;
            float* d = p->dog[o] + px * s;
            for (size_t i = 0; i < px; i++) d[i] = a[i] - b[i];
        }
        if (o + 1 < p->n_octaves) {
            /* :245-248: layer index len-3 = 3, nearest to (w/2, h/2) */
            int nw = cw / 2, nh = ch / 2;
            initial = (float*)malloc((size_t)(nw > 0 ? nw : 1) * (nh > 0 ? nh : 1) * sizeof(float));
            if (B) so_resize_nearest_imageproc(p->gauss[o] + px * 3, cw, ch, initial);
            else so_resize_nearest_half(p->gauss[o] + px * 3, cw, ch, initial);
            cw = nw; ch = nh;
        } else {
            initial = NULL;
        }
    }
    return p;
}

void so_pyramid_free(so_pyramid* p) {
    if (!p) return;
    for (int o = 0; o < p->n_octaves; o++) { free(p->gauss[o]); free(p->dog[o]); }
    free(p);
}
int so_pyramid_octaves(const so_pyramid* p) { return p->n_octaves; }
int so_pyramid_width(const so_pyramid* p, int o) { return p->w[o]; }
int so_pyramid_height(const so_pyramid* p, int o) { return p->h[o]; }
const float* so_pyramid_gauss(const so_pyramid* p, int o, int l) {
    return p->gauss[o] + (size_t)p->w[o] * p->h[o] * l;
}
const float* so_pyramid_dog(const so_pyramid* p, int o, int l) {
    return p->dog[o] + (size_t)p->w[o] * p->h[o] * l;
}

/* ------------------------------------------------------------------ */
/* detector                                                            */
/* ------------------------------------------------------------------ */
typedef struct {
    const float* d; /* (5,h,w) */
    int w, h;
} dogview;
static inline float D(const dogview* v, int s, int y, int x) {
    return v->d[((size_t)s * v->h + y) * v->w + x];
}

/* Rust f32::total_cmp-based max/min over the 8 ring neighbours
 * (src/lib.rs:439-451, 468-479); for finite inputs plain compares give the
 * same value up to the sign of zero, which `>=` / `<=` cannot see. */
static float ring_max(const dogview* v, int s, int y, int x) {
    float m = D(v, s, y - 1, x - 1);
    float c;
    c = D(v, s, y - 1, x); if (c > m) m = c;
    c = D(v, s, y - 1, x + 1); if (c > m) m = c;
    c = D(v, s, y, x - 1); if (c > m) m = c;
    c = D(v, s, y, x + 1); if (c > m) m = c;
    c = D(v, s, y + 1, x - 1); if (c > m) m = c;
    c = D(v, s, y + 1, x); if (c > m) m = c;
    c = D(v, s, y + 1, x + 1); if (c > m) m = c;
    return m;
}
static float ring_min(const dogview* v, int s, int y, int x) {
    float m = D(v, s, y - 1, x - 1);
    float c;
    c = D(v, s, y - 1, x); if (c < m) m = c;
    c = D(v, s, y - 1, x + 1); if (c < m) m = c;
    c = D(v, s, y, x - 1); if (c < m) m = c;
    c = D(v, s, y, x + 1); if (c < m) m = c;
    c = D(v, s, y + 1, x - 1); if (c < m) m = c;
    c = D(v, s, y + 1, x); if (c < m) m = c;
    c = D(v, s, y + 1, x + 1); if (c < m) m = c;
    return m;
}

/* point_is_local_extremum, src/lib.rs:437-506 (s = centre DoG layer) */
static int point_is_local_extremum(const dogview* v, int s, int x, int y) {
    float threshold = floorf(0.5f * CONTRAST_THRESHOLD / (float)SCALES_PER_OCTAVE); /* :460 => 0 */
    float val = D(v, s, y, x);
    if (fabsf(val) <= threshold) return 0;
    if (val > 0.0f) {
        if (val >= ring_max(v, s, y, x) && val >= ring_max(v, s - 1, y, x) &&
            val >= ring_max(v, s + 1, y, x)) {
            float p = D(v, s - 1, y, x), n = D(v, s + 1, y, x);
            return val >= fmaxf(p, n);
        }
    } else {
        if (val <= ring_min(v, s, y, x) && val <= ring_min(v, s - 1, y, x) &&
            val <= ring_min(v, s + 1, y, x)) {
            float p = D(v, s - 1, y, x), n = D(v, s + 1, y, x);
            return val <= fminf(p, n);
        }
    }
    return 0;
}

typedef struct {
    int scale, x, y;
    float offset_scale, offset_x, offset_y;
} interp_result;

/* interpolate_extremum, src/lib.rs:525-603 */
static int interpolate_extremum(const dogview* v, int scale, int x, int y, interp_result* out) {
    for (int it = 0; it < MAX_INTERPOLATION_STEPS; it++) {
        int p = scale - 1, c = scale, n = scale + 1;
        float g1 = (D(v, n, y, x) - D(v, p, y, x)) / 2.f;
        float g2 = (D(v, c, y + 1, x) - D(v, c, y - 1, x)) / 2.f;
        float g3 = (D(v, c, y, x + 1) - D(v, c, y, x - 1)) / 2.f;

        float value2x = D(v, c, y, x) * 2.f;
        float h11 = D(v, n, y, x) + D(v, p, y, x) - value2x;
        float h12 = (D(v, n, y + 1, x) - D(v, n, y - 1, x) - D(v, p, y + 1, x) + D(v, p, y - 1, x)) / 4.f;
        float h13 = (D(v, n, y, x + 1) - D(v, n, y, x - 1) - D(v, p, y, x + 1) + D(v, p, y, x - 1)) / 4.f;
        float h22 = D(v, c, y + 1, x) + D(v, c, y - 1, x) - value2x;
        float h33 = D(v, c, y, x + 1) + D(v, c, y, x - 1) - value2x;
        float h23 = (D(v, c, y + 1, x + 1) - D(v, c, y + 1, x - 1) - D(v, c, y - 1, x + 1) +
                     D(v, c, y - 1, x - 1)) / 4.f;

        float det = h11 * h22 * h33 - h11 * h23 * h23 - h12 * h12 * h33 + 2.f * h12 * h13 * h23 -
                    h13 * h13 * h22;
        float hinv11 = (h22 * h33 - h23 * h23) / det;
        float hinv12 = (h13 * h23 - h12 * h33) / det;
        float hinv13 = (h12 * h23 - h13 * h22) / det;
        float hinv22 = (h11 * h33 - h13 * h13) / det;
        float hinv23 = (h12 * h13 - h11 * h23) / det;
        float hinv33 = (h11 * h22 - h12 * h12) / det;

        float offset_scale = -(hinv11 * g1 + hinv12 * g2 + hinv13 * g3);
        float offset_x = -(hinv13 * g1 + hinv23 * g2 + hinv33 * g3);
        float offset_y = -(hinv12 * g1 + hinv22 * g2 + hinv23 * g3);

        if (fabsf(offset_scale) < 0.5f && fabsf(offset_x) < 0.5f && fabsf(offset_y) < 0.5f) {
            out->scale = scale; out->x = x; out->y = y;
            out->offset_scale = offset_scale; out->offset_x = offset_x; out->offset_y = offset_y;
            return 1;
        }
        /* :588-590 -- `as isize` saturates, NaN -> 0; an overflowing add would
         * panic (debug) or wrap to an out-of-range usize (release): reject. */
        int64_t rx = rust_f32_as_i64(roundf(offset_x));
        int64_t ry = rust_f32_as_i64(roundf(offset_y));
        int64_t rs = rust_f32_as_i64(roundf(offset_scale));
        const int64_t BIG = (int64_t)1 << 40;
        if (rx > BIG || rx < -BIG || ry > BIG || ry < -BIG || rs > BIG || rs < -BIG) return 0;
        int64_t nx = (int64_t)x + rx, ny = (int64_t)y + ry, ns = (int64_t)scale + rs;
        if (ns < 1 || ns > SCALES_PER_OCTAVE || nx < IMAGE_BORDER || nx >= v->w - IMAGE_BORDER ||
            ny < IMAGE_BORDER || ny >= v->h - IMAGE_BORDER)
            return 0;
        x = (int)nx; y = (int)ny; scale = (int)ns;
    }
    return 0;
}

/* extremum_contrast, src/lib.rs:606-626 (s = centre layer of the slice) */
static float extremum_contrast(const dogview* v, int s, int x, int y, float os, float ox, float oy) {
    float g1 = (D(v, s + 1, y, x) - D(v, s - 1, y, x)) / 2.f;
    float g2 = (D(v, s, y + 1, x) - D(v, s, y - 1, x)) / 2.f;
    float g3 = (D(v, s, y, x + 1) - D(v, s, y, x - 1)) / 2.f;
    float interp = os * g1 + oy * g2 + ox * g3;
    return D(v, s, y, x) + interp / 2.f;
}

/* extremum_is_on_edge, src/lib.rs:630-653 */
static int extremum_is_on_edge(const dogview* v, int s, int x, int y) {
    float val2x = D(v, s, y, x) * 2.0f;
    float h11 = D(v, s, y + 1, x) + D(v, s, y - 1, x) - val2x;
    float d22 = D(v, s, y, x + 1) + D(v, s, y, x - 1) - val2x;
    float h12 = (D(v, s, y + 1, x + 1) - D(v, s, y + 1, x - 1) - D(v, s, y - 1, x + 1) +
                 D(v, s, y - 1, x - 1)) / 4.f;
    float tr = d22 + h11;
    float det = d22 * h11 - h12 * h12;
    if (det <= 0.f) return 1;
    float e1 = EDGE_THRESHOLD + 1.0f;
    return (tr * tr * EDGE_THRESHOLD) > (e1 * e1) * det; /* (C+1).powi(2) * det */
}

/* gradient_direction_histogram, src/lib.rs:657-757; hist has ORI_BINS entries */
static void gradient_direction_histogram(const float* img, int w, int h, int x, int y, int radius,
                                         float sigma, float* hist) {
    const int n_bins = ORI_BINS;
    float grad_weight_scale = -1.0f / (2.0f * sigma * sigma);
    float raw_hist[ORI_BINS + 4];
    for (int i = 0; i < n_bins + 4; i++) raw_hist[i] = 0.0f;
    const float PI32 = 3.14159265358979323846f;
    float bin_angle_step = (float)n_bins / (PI32 * 2.f);
    for (int yp = -radius; yp <= radius; yp++) {
        if (yp <= -y) continue;
        int64_t yi = (int64_t)y + yp;
        if (yi <= 0 || yi >= h - 1) continue;
        for (int xp = -radius; xp <= radius; xp++) {
            if (xp <= -x) continue;
            int64_t xi = (int64_t)x + xp;
            if (xi <= 0 || xi >= w - 1) continue;
            float dx = img[yi * w + xi + 1] - img[yi * w + xi - 1];
            float dy = img[(yi - 1) * w + xi] - img[(yi + 1) * w + xi];
            float wexp = (float)(yp * yp + xp * xp) * grad_weight_scale;
            float weight = expf(wexp);
            float mag = sqrtf(dx * dx + dy * dy);
            float ori = (float)atan2((double)dy, (double)dx);
            float raw_bin = bin_angle_step * ori;
            int bin = rust_f32_as_i32(roundf(raw_bin));
            if (bin >= n_bins) bin -= n_bins;
            else if (bin < 0) bin += n_bins;
            raw_hist[bin + 2] += weight * mag;
        }
    }
    raw_hist[1] = raw_hist[n_bins + 1];
    raw_hist[0] = raw_hist[n_bins];
    raw_hist[n_bins + 2] = raw_hist[2];
    raw_hist[n_bins + 3] = raw_hist[3];
    for (int i = 2; i < n_bins + 2; i++) {
        hist[i - 2] = (raw_hist[i - 2] + raw_hist[i + 2]) * (1.f / 16.f) +
                      (raw_hist[i - 1] + raw_hist[i + 1]) * (4.f / 16.f) + raw_hist[i] * 6.f / 16.f;
    }
}

static int cand_scan_ok(int w, int h) { return !(h < 2 * IMAGE_BORDER || w < 2 * IMAGE_BORDER); }

size_t so_find_candidates(const so_pyramid* p, so_candidate* out, size_t cap) {
    size_t n = 0;
    for (int o = 0; o < p->n_octaves; o++) {
        dogview v = {p->dog[o], p->w[o], p->h[o]};
        if (!cand_scan_ok(v.w, v.h)) continue; /* src/lib.rs:315-317 */
        for (int s = 1; s <= SCALES_PER_OCTAVE; s++)
            for (int y = IMAGE_BORDER; y < v.h - IMAGE_BORDER; y++)
                for (int x = IMAGE_BORDER; x < v.w - IMAGE_BORDER; x++)
                    if (point_is_local_extremum(&v, s, x, y)) {
                        if (n < cap) { so_candidate c = {o, s, y, x}; out[n] = c; }
                        n++;
                    }
    }
    return n;
}

/* find_keypoints + find_extrema_in_dog_img, src/lib.rs:281-435 */
size_t so_find_keypoints(const so_pyramid* p, so_sift_keypoint* out, size_t cap) {
    size_t n = 0;
    for (int o = 0; o < p->n_octaves; o++) {
        dogview v = {p->dog[o], p->w[o], p->h[o]};
        if (!cand_scan_ok(v.w, v.h)) continue;
        for (int s = 1; s <= SCALES_PER_OCTAVE; s++) {
            for (int y0 = IMAGE_BORDER; y0 < v.h - IMAGE_BORDER; y0++) {
                for (int x0 = IMAGE_BORDER; x0 < v.w - IMAGE_BORDER; x0++) {
                    if (!point_is_local_extremum(&v, s, x0, y0)) continue;
                    interp_result r;
                    if (!interpolate_extremum(&v, s, x0, y0, &r)) continue;
                    float contrast = fabsf(extremum_contrast(&v, r.scale, r.x, r.y, r.offset_scale,
                                                             r.offset_x, r.offset_y));
                    if (contrast * (float)SCALES_PER_OCTAVE <= CONTRAST_THRESHOLD) continue; /* :360 */
                    if (extremum_is_on_edge(&v, r.scale, r.x, r.y)) continue;              /* :365 */

                    float octave_scale_factor = ldexpf(1.0f, o); /* 2_f32.powi(octave), :369 */
                    float kp_scale = (float)SIGMA_MIN *
                                     powf(2.f, ((float)r.scale + r.offset_scale) / (float)SCALES_PER_OCTAVE) *
                                     2.f; /* :372-374 */
                    float kp_x = ((float)r.x + r.offset_x) * octave_scale_factor;
                    float kp_y = ((float)r.y + r.offset_y) * octave_scale_factor;
                    int radius = rust_f32_as_i32(roundf(3.f * ORIENTATION_HISTOGRAM_RADIUS * kp_scale)); /* :380 */
                    float hist[ORI_BINS];
                    gradient_direction_histogram(so_pyramid_gauss(p, o, r.scale), v.w, v.h, r.x, r.y,
                                                 radius, LAMBDA_ORI * kp_scale, hist);
                    float histogram_max = hist[0];
                    for (int k = 1; k < ORI_BINS; k++) if (hist[k] > histogram_max) histogram_max = hist[k];
                    float localmax_threshold = histogram_max * LOCALMAX_RATIO;
                    for (int k = 0; k < ORI_BINS; k++) { /* :397-431 */
                        int km = k > 0 ? k - 1 : ORI_BINS - 1;
                        int kp = k < ORI_BINS - 1 ? k + 1 : 0;
                        if (hist[k] > hist[km] && hist[k] > hist[kp] && hist[k] >= localmax_threshold) {
                            float interp = (hist[km] - hist[kp]) / (hist[km] - 2.0f * hist[k] + hist[kp]);
                            float bin = (float)k + 0.5f * interp;
                            if (bin < 0.0f) bin = (float)ORI_BINS + bin;
                            else if (bin >= (float)ORI_BINS) bin = bin - (float)ORI_BINS;
                            float kp_angle = 360.0f - (360.0f / (float)ORI_BINS) * bin;
                            if (n < cap) {
                                so_sift_keypoint kpt;
                                kpt.x = kp_x; kpt.y = kp_y;
                                kpt.size = kp_scale * octave_scale_factor;
                                kpt.response = contrast;
                                kpt.octave = o; kpt.scale = r.scale; kpt.angle = kp_angle;
                                out[n] = kpt;
                            }
                            n++;
                        }
                    }
                }
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* descriptor: src/lib.rs:785-990                                      */
/* ------------------------------------------------------------------ */
void so_compute_descriptor(const float* img, int width, int height, float xf, float yf, float scale,
                           float orientation, uint8_t out[SO_DESC_SIZE]) {
    const int n_hist = N_HIST, n_bins = N_BINS;
    int64_t x = rust_f32_as_usize(roundf(xf));
    int64_t y = rust_f32_as_usize(roundf(yf));
    const float BIN_ANGLE_STEP = (float)N_BINS / 360.0f;
    float hist_width = LAMBDA_DESCR * scale;
    int radius = rust_f32_as_i32(roundf(LAMBDA_DESCR * scale * sqrtf(2.0f) * (float)(n_hist + 1) * 0.5f));
    const float RADS_PER_DEG = 3.14159265358979323846f / 180.0f; /* f32::to_radians */
    float rad = orientation * RADS_PER_DEG;
    float sin_ori = sinf(rad), cos_ori = cosf(rad);
    float sin_s = sin_ori / hist_width, cos_s = cos_ori / hist_width;

    float hist[(N_HIST + 2) * (N_HIST + 2) * N_BINS];
    memset(hist, 0, sizeof(hist));
    const float weight_scale = -2.f / (float)(n_hist * n_hist);
    const double DEG_PER_RAD = 180.0 / 3.14159265358979323846; /* f64::to_degrees */

    for (int yw = -radius; yw <= radius; yw++) {
        for (int xw = -radius; xw <= radius; xw++) {
            float col_rot = (float)xw * cos_s - (float)yw * sin_s;
            float row_rot = (float)xw * sin_s + (float)yw * cos_s;
            float row_bin = row_rot + (float)(n_hist / 2);
            float col_bin = col_rot + (float)(n_hist / 2);
            int64_t ay = y + yw, ax = x + xw; /* `y as i32 + y_in_window` */
            if (!(row_bin > -0.5f && row_bin < (float)n_hist + 0.5f && col_bin > -0.5f &&
                  col_bin < (float)n_hist + 0.5f && ay > 0 && ay < height - 1 && ax > 0 && ax < width - 1))
                continue;
            float dx = img[ay * width + ax + 1] - img[ay * width + ax - 1];
            float dy = img[(ay - 1) * width + ax] - img[(ay + 1) * width + ax];
            float wgt = col_rot * col_rot + row_rot * row_rot;
            float weight = expf(wgt * weight_scale);
            double deg = atan2((double)dy, (double)dx) * DEG_PER_RAD;
            float orient = (float)fmod(deg + 360.0, 360.0) - orientation;
            float mag = sqrtf(dx * dx + dy * dy);

            /* :890-948 */
            float rb = row_bin - 0.5f, cb = col_bin - 0.5f;
            mag = mag * weight;
            float obin = orient * BIN_ANGLE_STEP;
            float row_floor = floorf(rb), col_floor = floorf(cb), ori_floor = floorf(obin);
            float row_frac = rb - row_floor, col_frac = cb - col_floor, ori_frac = obin - ori_floor;
            float c1 = mag * row_frac, c0 = mag - c1;
            float c11 = c1 * col_frac, c10 = c1 - c11;
            float c01 = c0 * col_frac, c00 = c0 - c01;
            float c111 = c11 * ori_frac, c110 = c11 - c111;
            float c101 = c10 * ori_frac, c100 = c10 - c101;
            float c011 = c01 * ori_frac, c010 = c01 - c011;
            float c001 = c00 * ori_frac, c000 = c00 - c001;
            int r1 = (int)rust_f32_as_usize(row_floor + 1.f), cc1 = (int)rust_f32_as_usize(col_floor + 1.f);
            int r2 = (int)rust_f32_as_usize(row_floor + 2.f), cc2 = (int)rust_f32_as_usize(col_floor + 2.f);
            float of = ori_floor;
            if (of < 0.f) of += (float)n_bins;
            else if (of >= (float)n_bins) of -= (float)n_bins;
            int o0 = (int)rust_f32_as_usize(of);
            if (o0 >= n_bins) o0 = n_bins - 1; /* unreachable for in-range angles; the crate would panic */
            int o1 = (o0 + 1 >= n_bins) ? 0 : o0 + 1;
#define HI(r, c, o) (((r) * (N_HIST + 2) + (c)) * N_BINS + (o))
            hist[HI(r1, cc1, o0)] += c000;
            hist[HI(r1, cc1, o1)] += c001;
            hist[HI(r1, cc2, o0)] += c010;
            hist[HI(r1, cc2, o1)] += c011;
            hist[HI(r2, cc1, o0)] += c100;
            hist[HI(r2, cc1, o1)] += c101;
            hist[HI(r2, cc2, o0)] += c110;
            hist[HI(r2, cc2, o1)] += c111;
        }
    }
    /* :951: hist[1..-1, 1..-1, :] flattened row-major */
    float flat[SO_DESC_SIZE];
    for (int r = 0; r < n_hist; r++)
        for (int c = 0; c < n_hist; c++)
            for (int o = 0; o < n_bins; o++)
                flat[(r * n_hist + c) * n_bins + o] = hist[HI(r + 1, c + 1, o)];
#undef HI
    /* :957-962: sum of squares in chunks of 4 */
    float acc = 0.f;
    for (int i = 0; i < SO_DESC_SIZE; i += 4) {
        float s = 0.f;
        for (int j = 0; j < 4; j++) s += flat[i + j] * flat[i + j];
        acc = (i == 0) ? s : acc + s;
    }
    float l2_uncapped = sqrtf(acc);
    float cap = l2_uncapped * 0.2f;
    for (int i = 0; i < SO_DESC_SIZE; i++) flat[i] = fminf(flat[i], cap); /* f32::min */
    acc = 0.f;
    for (int i = 0; i < SO_DESC_SIZE; i += 4) {
        float s = 0.f;
        for (int j = 0; j < 4; j++) s += flat[i + j] * flat[i + j];
        acc = (i == 0) ? s : acc + s;
    }
    float l2_capped = sqrtf(acc);
    float l2_normalizer = 512.0f / fmaxf(l2_capped, FLT_EPSILON);
    for (int i = 0; i < SO_DESC_SIZE; i++) {
        int32_t q = rust_f32_as_i32(roundf(flat[i] * l2_normalizer));
        out[i] = q > 255 ? 255 : (uint8_t)q; /* `x as u8` on i32 truncates; q >= 0 here */
    }
}

/* ------------------------------------------------------------------ */
/* assembly: src/lib.rs:147-177                                        */
/* ------------------------------------------------------------------ */
typedef struct { so_sift_keypoint k; size_t idx; } ranked_kp;
static int cmp_response_desc(const void* a, const void* b) {
    const ranked_kp* ka = (const ranked_kp*)a; const ranked_kp* kb = (const ranked_kp*)b;
    /* kp2.response.total_cmp(&kp1.response); responses are non-negative finite */
    if (ka->k.response > kb->k.response) return -1;
    if (ka->k.response < kb->k.response) return 1;
    /* sort_unstable leaves tie order unspecified; the oracle breaks ties by natural index */
    return ka->idx < kb->idx ? -1 : (ka->idx > kb->idx ? 1 : 0);
}

size_t so_sift_with_precomputed(const so_pyramid* p, int64_t features_limit, so_keypoint* kps,
                                uint8_t* desc, size_t cap) {
    size_t n = so_find_keypoints(p, NULL, 0);
    so_sift_keypoint* sk = (so_sift_keypoint*)malloc((n ? n : 1) * sizeof(so_sift_keypoint));
    so_find_keypoints(p, sk, n);
    if (features_limit >= 0 && (size_t)features_limit < n) {
        ranked_kp* rk = (ranked_kp*)malloc(n * sizeof(ranked_kp));
        for (size_t i = 0; i < n; i++) { rk[i].k = sk[i]; rk[i].idx = i; }
        qsort(rk, n, sizeof(ranked_kp), cmp_response_desc);
        n = (size_t)features_limit;
        for (size_t i = 0; i < n; i++) sk[i] = rk[i].k;
        free(rk);
    }
    for (size_t i = 0; i < n && i < cap; i++) {
        const so_sift_keypoint* k = &sk[i];
        /* compute_descriptors, :759-782 */
        float angle = 360.0f - k->angle;
        float f = ldexpf(1.0f, -k->octave);
        if (desc)
            so_compute_descriptor(so_pyramid_gauss(p, k->octave, k->scale), p->w[k->octave],
                                  p->h[k->octave], k->x * f, k->y * f, k->size * f, angle,
                                  desc + i * SO_DESC_SIZE);
        if (kps) {
            kps[i].x = k->x * DELTA_MIN; kps[i].y = k->y * DELTA_MIN; kps[i].size = k->size * DELTA_MIN;
            kps[i].angle = k->angle; kps[i].response = k->response;
        }
    }
    free(sk);
    return n;
}

size_t so_sift(const uint8_t* gray, int w, int h, int stride, int64_t features_limit, so_keypoint* kps,
               uint8_t* desc, size_t cap) {
    return so_sift_flavour(gray, w, h, stride, features_limit, SO_PROCESSING_OPENCV, kps, desc, cap);
}

/* sift_with_processing::<P>, src/lib.rs:76-81 */
size_t so_sift_flavour(const uint8_t* gray, int w, int h, int stride, int64_t features_limit, int flavour,
                       so_keypoint* kps, uint8_t* desc, size_t cap) {
    so_pyramid* p = so_precompute_flavour(gray, w, h, stride, flavour);
    size_t n = so_sift_with_precomputed(p, features_limit, kps, desc, cap);
    so_pyramid_free(p);
    return n;
}
