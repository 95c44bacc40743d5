// sb_pyramid.cuh -- Gaussian scale space, DoG + extrema, ordered candidate compaction.
//
// Arithmetic contract (what makes the pyramid bit-identical to the oracle and to
// OpenCV's vectorised GaussianBlur/resize that the reference's OpenCVProcessing
// calls, src/opencv_processing.rs:20-28,51-57):
//   * 2x upsample (src/lib.rs:201-205): per axis lerp = fma(b - a, t, a), t in {0, .25, .75},
//     horizontal then vertical, source index clamped at the borders;
//   * blur (src/lib.rs:209, 233-236): row pass first, acc = x[-R]*k[0], then
//     acc = fma(x[-R+i], k[i], acc) for i = 1..2R; column pass acc = c[0]*k[R], then
//     acc = fma(c[+i] + c[-i], k[R+i], acc) for i = 1..R; BORDER_REFLECT_101;
//   * decimation (src/lib.rs:245-248): even rows / even columns of layer 3;
//   * DoG (src/lib.rs:271-279): G[l+1] - G[l], recomputed on the fly (never stored).
// The translation unit is compiled with --fmad=false: every FMA below is explicit.
//
// Flavour FL_IMAGEPROC (template parameter FL = 1) swaps in the arithmetic of the crate's default Processing
// (src/lib.rs:992-1007: imageproc::filter::gaussian_blur_f32, image::imageops::resize) as restated by the oracle
// from the published algorithms of imageproc 0.25 / image 0.25 -- PARITY UNPINNED: those crates' sources are not
// part of the reference tree and no reference test exercises this flavour:
//   * blur: taps exp(-x^2 / 2 sigma^2) / (sigma sqrt(2 pi)) in f32, radius ceil(2 sigma), NOT renormalised; horizontal
//     pass then vertical pass, each acc = 0, acc = acc + x[i] * k[i] left to right / top to bottom (a multiply and an
//     add, no FMA, no symmetric folding); border = clamp to the edge pixel;
//   * 2x upsample (resize, FilterType::Triangle): vertical pass then horizontal pass, out = a * wa + b * wb with
//     (wa, wb) in {(.25, .75), (.75, .25)} and the clamped edge weights renormalised to (1, 0); result clamped to [0, 1];
//   * decimation (resize, FilterType::Nearest): source pixel floor((d + 0.5) * (n_src / n_dst)) evaluated in f32 (the odd
//     pixels 2d + 1), result clamped to [0, 1].
#pragma once
#include <cuda.h>

#include "sb_common.cuh"

namespace sb {

// taps of the six Gaussian kernels, [kernel][0..2R]; filled by the host at context creation
__constant__ float c_taps[N_LAYERS][32];
// the same taps duplicated into both halves of a 64-bit operand for the packed f32x2 arithmetic
__constant__ float2 c_taps2[N_LAYERS][32];
// the same for flavour FL_IMAGEPROC
__constant__ float c_taps_b[N_LAYERS][32];
__constant__ float2 c_taps2_b[N_LAYERS][32];

template <int FL>
__device__ __forceinline__ float tap(const int l, const int i) { return FL == FL_OPENCV ? c_taps[l][i] : c_taps_b[l][i]; }
template <int FL>
__device__ __forceinline__ float2 tap2(const int l, const int i) { return FL == FL_OPENCV ? c_taps2[l][i] : c_taps2_b[l][i]; }

// acc + a * k per half with a separately rounded product (the imageproc flavour's accumulation).  Scalar intrinsics:
// ptxas contracts a mul.rn.f32x2 followed by an add.rn.f32x2 into one FFMA2 (observed in SASS; the explicit rounding
// modifier does not stop it for the packed forms), which __fmul_rn / __fadd_rn are documented never to allow.
__device__ __forceinline__ float2 madd_unfused2(const float2 acc, const float2 a, const float2 k) {
    return make_float2(__fadd_rn(acc.x, __fmul_rn(a.x, k.x)), __fadd_rn(acc.y, __fmul_rn(a.y, k.y)));
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}
// source index of a filter tap that falls outside the image: BORDER_REFLECT_101 (OpenCV) / clamp to the edge (imageproc)
template <int FL>
__device__ __forceinline__ int border_index(const int i, const int n) {
    return FL == FL_OPENCV ? reflect101(i, n) : min(max(i, 0), n - 1);
}

// source indices / weight of the 2x bilinear upsample for destination index d (n source samples)
__device__ __forceinline__ void lerp2x(int d, int n, int& a, int& b, float& t) {
    int s = (d & 1) ? ((d - 1) >> 1) : ((d >> 1) - 1);
    float f = (d & 1) ? 0.25f : 0.75f;
    if (s < 0) { s = 0; f = 0.0f; }
    if (s >= n - 1) { s = n - 1; f = 0.0f; }
    a = s;
    b = (s + 1 < n) ? s + 1 : n - 1;
    t = f;
}
__device__ __forceinline__ int lerp2x_lo(int d, int n) {
    int s = (d & 1) ? ((d - 1) >> 1) : ((d >> 1) - 1);
    return s < 0 ? 0 : (s > n - 1 ? n - 1 : s);
}

template <int L, int FL = FL_OPENCV>
struct BlurCfg {
    static constexpr int R = blur_radius(L, FL);
    static constexpr int TW = 128, TH = 64;      // output tile
    static constexpr int SH = TH + 2 * R;        // staged rows
    static constexpr int SW = TW + 2 * R;        // staged columns
    static constexpr int WIN = 8 + 2 * R;        // inputs feeding 8 consecutive row-pass outputs
    static constexpr int NV4 = (WIN + 3) / 4;    // as float4 loads
    static constexpr int SPITCH = 164;           // == 4 (mod 32): conflict-free LDS.128 across rows
    static constexpr int IPITCH = 132;           // == 4 (mod 32)
    static constexpr int PY = 16;                // rows per column-pass thread
    static constexpr int THREADS = 256;
    static constexpr int IN_H = 40, IN_W = 72;   // seed only: staged input tile (aliases `inter`)
    static constexpr size_t SMEM = (size_t)SH * (SPITCH + IPITCH) * sizeof(float);
    static_assert(120 + 4 * NV4 <= SPITCH, "row-pass window overruns the stage pitch");
    static_assert(IN_H * (IN_W + SW) <= SH * IPITCH, "input tile + horizontal buffer do not fit the aliased buffer");
    static_assert(L != 0 || ((TH + 2 * R) / 2 + 2 <= IN_H && (TW + 2 * R) / 2 + 2 <= IN_W), "input tile too small");
};

struct BlurParams {
    const float* src;        // source layer, image 0
    float* dst;              // destination layer, image 0
    long long img_stride;    // floats between consecutive images' arenas
    int w, h, pitch;         // layer size (for the seed: 2W x 2H)
    float* dec;              // layer 0 of the next octave (decimated copy), or nullptr
    int dec_w, dec_h, dec_pitch;
    const uint8_t* in;       // seed only: u8 input, image 0
    long long in_img_stride; // bytes between images
    int in_w, in_h, in_stride;
};

// One kernel for all six Gaussian blurs.  L selects the tap set; SEED fuses
// u8 -> f32/255 -> 2x bilinear upsample in front of the blur (create_seed_image,
// src/lib.rs:196-210); DECIMATE also writes the even pixels into the next octave.
template <int L, bool SEED, bool DECIMATE, int FL = FL_OPENCV>
__global__ void __launch_bounds__(256, 2) k_blur(const BlurParams p) {
    pdl_wait();
    static_assert(FL == FL_OPENCV || (!SEED && !DECIMATE), "the imageproc flavour upsamples and decimates in its own kernels");
    using C = BlurCfg<L, FL>;
    constexpr int R = C::R;
    extern __shared__ __align__(16) float smem[];
    float* stage = smem;
    float* inter = smem + C::SH * C::SPITCH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * C::TW, ty0 = blockIdx.y * C::TH;
    const long long img = blockIdx.z;
    const int w = p.w, h = p.h;
    // staged rows / columns that feed at least one real output of this tile
    const int rows_needed = min(C::TH, h - ty0) + 2 * R;
    const int cols_needed = min(C::TW, w - tx0) + 2 * R;

    if (SEED) {
        // ---- stage the u8 input tile as f32 / 255 (ConvertBuffer, src/lib.rs:198) ----
        float* in_tile = inter;
        const int ylo = max(ty0 - R, 0), yhi = min(ty0 + C::TH + R, h) - 1;
        const int xlo = max(tx0 - R, 0), xhi = min(tx0 + C::TW + R, w) - 1;
        const int iy0 = lerp2x_lo(ylo, p.in_h), ix0 = lerp2x_lo(xlo, p.in_w);
        const int iy1 = min(lerp2x_lo(yhi, p.in_h) + 1, p.in_h - 1);
        const int ix1 = min(lerp2x_lo(xhi, p.in_w) + 1, p.in_w - 1);
        const int ih = iy1 - iy0 + 1, iw = ix1 - ix0 + 1;
        const uint8_t* in = p.in + img * p.in_img_stride;
        for (int idx = tid; idx < ih * iw; idx += C::THREADS) {
            int r = idx / iw, c = idx - r * iw;
            float v = (float)in[(long long)(iy0 + r) * p.in_stride + ix0 + c];
            in_tile[r * C::IN_W + c] = v / 255.0f;
        }
        __syncthreads();
        // horizontal lerp of every staged input row (OpenCV resize: horizontal pass first):
        // hbuf[r][col] = fma(I[r][bx] - I[r][ax], fx, I[r][ax]) for the SW staged columns
        float* hbuf = in_tile + C::IN_H * C::IN_W;  // IN_H x SW, also inside `inter`
        for (int col = tid & 127; col < C::SW; col += 128) {
            if (col < cols_needed) {
                const int x = reflect101(tx0 - R + col, w);
                int ax, bx;
                float fx;
                lerp2x(x, p.in_w, ax, bx, fx);
                for (int r = tid >> 7; r < ih; r += C::THREADS / 128) {
                    const float a = in_tile[r * C::IN_W + ax - ix0], b = in_tile[r * C::IN_W + bx - ix0];
                    hbuf[r * C::SW + col] = fmaf(b - a, fx, a);
                }
            }
        }
        __syncthreads();
        // vertical lerp into the stage: U[y][x] = fma(H[by][x] - H[ay][x], fy, H[ay][x])
        for (int col = tid & 127; col < C::SW; col += 128) {
            const bool col_ok = col < cols_needed;
            for (int row = tid >> 7; row < C::SH; row += C::THREADS / 128) {
                float v = 0.0f;
                if (col_ok && row < rows_needed) {
                    const int y = reflect101(ty0 - R + row, h);
                    int ay, by;
                    float fy;
                    lerp2x(y, p.in_h, ay, by, fy);
                    const float a = hbuf[(ay - iy0) * C::SW + col], b = hbuf[(by - iy0) * C::SW + col];
                    v = fmaf(b - a, fy, a);
                }
                stage[row * C::SPITCH + col] = v;
            }
        }
    } else {
        const float* src = p.src + img * p.img_stride;
        for (int idx = tid; idx < C::SH * C::SW; idx += C::THREADS) {
            int row = idx / C::SW, col = idx - row * C::SW;
            float v = 0.0f;
            if (row < rows_needed && col < cols_needed) {
                int gy = border_index<FL>(ty0 - R + row, h);
                int gx = border_index<FL>(tx0 - R + col, w);
                v = __ldg(src + (long long)gy * p.pitch + gx);
            }
            stage[row * C::SPITCH + col] = v;
        }
    }
    __syncthreads();

    // ---- row pass: warp = 32 staged rows x one 8-pixel segment ----
    {
        constexpr int NGROUPS = (C::SH + 31) / 32;
        constexpr int NSEG = C::TW / 8;
        for (int task = warp; task < NGROUPS * NSEG; task += C::THREADS / 32) {
            const int g = task / NSEG, seg = task - g * NSEG;
            const int row = g * 32 + lane;
            if (row < rows_needed && row < C::SH) {
                float win[4 * C::NV4];
                const float4* sp = reinterpret_cast<const float4*>(stage + row * C::SPITCH + seg * 8);
#pragma unroll
                for (int v = 0; v < C::NV4; v++) {
                    float4 q = sp[v];
                    win[4 * v + 0] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
                }
                float out[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float acc = win[j] * tap<FL>(L, 0);
#pragma unroll
                    for (int i = 1; i <= 2 * R; i++)
                        acc = FL == FL_OPENCV ? fmaf(win[j + i], tap<FL>(L, i), acc) : __fadd_rn(acc, __fmul_rn(win[j + i], tap<FL>(L, i)));
                    out[j] = acc;
                }
                float4* ip = reinterpret_cast<float4*>(inter + row * C::IPITCH + seg * 8);
                ip[0] = make_float4(out[0], out[1], out[2], out[3]);
                ip[1] = make_float4(out[4], out[5], out[6], out[7]);
            }
        }
    }
    __syncthreads();

    // ---- column pass: thread = one column x PY consecutive rows ----
    {
        float* dst = p.dst + img * p.img_stride;
        float* dec = DECIMATE ? p.dec + img * p.img_stride : nullptr;
        constexpr int NCH = C::TH / C::PY;
        for (int task = tid; task < C::TW * NCH; task += C::THREADS) {
            const int cy = task / C::TW, x = task - cy * C::TW;
            const int y0 = cy * C::PY;
            const int gx = tx0 + x;
            if (ty0 + y0 >= h) continue;  // whole chunk below the image (warp-uniform)
            float c[C::PY + 2 * R];
#pragma unroll
            for (int j = 0; j < C::PY + 2 * R; j++) c[j] = inter[(y0 + j) * C::IPITCH + x];
#pragma unroll
            for (int j = 0; j < C::PY; j++) {
                float acc;
                if (FL == FL_OPENCV) {
                    acc = c[j + R] * tap<FL>(L, R);
#pragma unroll
                    for (int i = 1; i <= R; i++) acc = fmaf(c[j + R + i] + c[j + R - i], tap<FL>(L, R + i), acc);
                } else {
                    acc = c[j] * tap<FL>(L, 0);
#pragma unroll
                    for (int i = 1; i <= 2 * R; i++) acc = __fadd_rn(acc, __fmul_rn(c[j + i], tap<FL>(L, i)));
                }
                const int gy = ty0 + y0 + j;
                if (gy < h && gx < w) {
                    dst[(long long)gy * p.pitch + gx] = acc;
                    if (DECIMATE && !(gy & 1) && !(gx & 1)) {
                        int dy = gy >> 1, dx = gx >> 1;
                        if (dy < p.dec_h && dx < p.dec_w) dec[(long long)dy * p.dec_pitch + dx] = acc;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// u8 -> f32 / 255 -> 2x bilinear upsample (first half of create_seed_image, src/lib.rs:196-205) as an
// elementwise kernel for octave sizes that take the TMA blur; the seed blur then runs as k_blur_tma<0>.
// A thread owns the 2 x 4 output block rows {2y+1, 2y+2}, columns {4k .. 4k+3}: both rows interpolate
// between input rows y and y+1, the columns between input columns 2k-1 .. 2k+2 (clamped at the border,
// where the lerp degenerates to fma(0, t, a) = a exactly as the clamped reference index does).
// ---------------------------------------------------------------------------
struct UpsampleParams {
    const uint8_t* in;
    long long in_img_stride;
    int in_w, in_h, in_stride;
    float* dst;              // 2W x 2H f32, image 0
    long long img_stride;
    int pitch;
};

constexpr int UPS_ROWS = 8;   // input row pairs per CTA (amortises the table set-up and the CTA launch)

// One 2 x 4 block of the upsampled image: output rows {2y+1, 2y+2} (y = -1 yields output row 0 only), output columns
// {4k .. 4k+3}, from input rows (y, y+1) and input columns 2k-1 .. 2k+2, all clamped at the border (where the lerp
// degenerates to fma(0, t, a) = a exactly as the reference's clamped index does).  s_norm = v / 255 for the 256 pixel
// values.  Horizontal pass first, then vertical (OpenCV's resize order).
// (split into the eight pixel loads and the arithmetic so that a caller can have the loads of several blocks in flight)
__device__ __forceinline__ void upsample_load(const uint8_t* __restrict__ in, const int W, const int H, const int in_stride,
                                              const int k, const int y, uint32_t raw[8]) {
    const int y0 = max(y, 0), y1 = min(y + 1, H - 1);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int x = min(max(2 * k - 1 + c, 0), W - 1);
        raw[c] = in[(long long)y0 * in_stride + x];
        raw[4 + c] = in[(long long)y1 * in_stride + x];
    }
}
__device__ __forceinline__ void upsample_math(const uint32_t raw[8], const float* __restrict__ s_norm, const int k, float o[2][4]) {
    float a[2][4];
#pragma unroll
    for (int c = 0; c < 4; c++) { a[0][c] = s_norm[raw[c]]; a[1][c] = s_norm[raw[4 + c]]; }
    // horizontal pass: out col 4k = lerp(c0, c1, .75), 4k+1 = lerp(c1, c2, .25), 4k+2 = lerp(c1, c2, .75), 4k+3 = lerp(c2, c3, .25)
    float hrow[2][4];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        hrow[r][0] = fmaf(a[r][1] - a[r][0], 0.75f, a[r][0]);
        hrow[r][1] = fmaf(a[r][2] - a[r][1], 0.25f, a[r][1]);
        hrow[r][2] = fmaf(a[r][2] - a[r][1], 0.75f, a[r][1]);
        hrow[r][3] = fmaf(a[r][3] - a[r][2], 0.25f, a[r][2]);
    }
    // the reference clamps the source index and zeroes the weight at the borders (column 0 and 2W-1)
    if (k == 0) { hrow[0][0] = a[0][1]; hrow[1][0] = a[1][1]; }   // 2k-1 < 0: both taps are column 0
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const float f = r ? 0.75f : 0.25f;
#pragma unroll
        for (int c = 0; c < 4; c++) o[r][c] = fmaf(hrow[1][c] - hrow[0][c], f, hrow[0][c]);
    }
}
__device__ __forceinline__ void upsample_block(const uint8_t* __restrict__ in, const int W, const int H, const int in_stride,
                                               const float* __restrict__ s_norm, const int k, const int y, float o[2][4]) {
    uint32_t raw[8];
    upsample_load(in, W, H, in_stride, k, y, raw);
    upsample_math(raw, s_norm, k, o);
}

__global__ void __launch_bounds__(256) k_upsample2x(const UpsampleParams p) {
    pdl_wait();
    // v / 255 for the 256 possible pixel values (the IEEE division itself, done once per CTA instead of 8x per thread)
    __shared__ float s_norm[256];
    s_norm[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;      // output columns 4k .. 4k+3
    const long long img = blockIdx.z;
    const int W = p.in_w, H = p.in_h;
    if (4 * k >= 2 * W) return;
    const uint8_t* in = p.in + img * p.in_img_stride;
    float* dst = p.dst + img * p.img_stride;
    const bool full = (4 * k + 3) < 2 * W;
#pragma unroll 2
  for (int yy = 0; yy < UPS_ROWS; yy++) {
    const int y = (int)blockIdx.y * UPS_ROWS + yy - 1;        // input row pair (y, y+1); y = -1 yields output row 0
    if (y >= H) break;
    float o[2][4];
    upsample_block(in, W, H, p.in_stride, s_norm, k, y, o);
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int Y = 2 * y + 1 + r;
        if (Y < 0 || Y >= 2 * H) continue;
        float* q = dst + (long long)Y * p.pitch + 4 * k;
        if (full) *reinterpret_cast<float4*>(q) = make_float4(o[r][0], o[r][1], o[r][2], o[r][3]);
        else for (int c = 0; c < 4 && 4 * k + c < 2 * W; c++) q[c] = o[r][c];
    }
  }
}

// ---------------------------------------------------------------------------
// Flavour FL_IMAGEPROC: the two resizes of the crate's default Processing as kernels of their own.
//
// k_upsample2x_b -- image::imageops::resize(.., FilterType::Triangle) for the exact 2x case of src/lib.rs:201-205,
// fused with the u8 -> f32 / 255 conversion (:198).  The image crate samples vertically first (into an f32 image),
// then horizontally; for 2x the normalised triangle weights are (.25, .75) / (.75, .25) and (1, 0) at the clamped
// ends, each output is the sum of at most two products accumulated from zero, and the final value is clamped to
// [0, 1].  A thread owns one output row and four output columns {4k .. 4k+3} (input columns 2k-1 .. 2k+2).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float tri2x(const float a, const float b, const bool odd, const bool at_end) {
    // output 2k (odd = false): a = in[k-1], b = in[k];  output 2k+1 (odd = true): a = in[k], b = in[k+1];
    // at_end: the outer neighbour does not exist and the surviving weight renormalises to exactly 1
    if (at_end) return odd ? a : b;
    return odd ? a * 0.75f + b * 0.25f : a * 0.25f + b * 0.75f;
}

__global__ void __launch_bounds__(256) k_upsample2x_b(const UpsampleParams p) {
    pdl_wait();
    __shared__ float s_norm[256];
    s_norm[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;      // output columns 4k .. 4k+3
    const int Y = blockIdx.y;                                  // output row
    const long long img = blockIdx.z;
    const int W = p.in_w, H = p.in_h;
    if (4 * k >= 2 * W) return;
    const uint8_t* in = p.in + img * p.in_img_stride;
    // vertical pass for the four input columns: rows (ky - 1, ky) for an even output row, (ky, ky + 1) for an odd one
    const int ky = Y >> 1;
    const bool yodd = Y & 1;
    const int ya = yodd ? ky : max(ky - 1, 0), yb = yodd ? min(ky + 1, H - 1) : ky;
    const bool yend = yodd ? (ky == H - 1) : (ky == 0);
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int x = min(max(2 * k - 1 + c, 0), W - 1);
        v[c] = tri2x(s_norm[in[(long long)ya * p.in_stride + x]], s_norm[in[(long long)yb * p.in_stride + x]], yodd, yend);
    }
    // horizontal pass: v[c] is input column 2k - 1 + c
    float o[4];
    o[0] = tri2x(v[0], v[1], false, 2 * k == 0);
    o[1] = tri2x(v[1], v[2], true, 2 * k == W - 1);
    o[2] = tri2x(v[1], v[2], false, false);          // output 4k+2 exists only when input column 2k+1 does
    o[3] = tri2x(v[2], v[3], true, 2 * k + 1 == W - 1);
    float* q = p.dst + img * p.img_stride + (long long)Y * p.pitch + 4 * k;
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (4 * k + c < 2 * W) q[c] = fminf(fmaxf(o[c], 0.0f), 1.0f);
}

// k_decimate_b -- image::imageops::resize(.., FilterType::Nearest) to (w/2, h/2) (src/lib.rs:245-248): the box
// kernel with support 0 keeps the one source pixel floor((d + 0.5) * (n_src / n_dst)), all in f32, clamped to the
// image; vertical pass, horizontal pass, clamp to [0, 1].
struct DecimateParams {
    const float* src;   // layer 3 of the octave, image 0
    float* dst;         // layer 0 of the next octave, image 0
    long long img_stride;
    int w, h, pitch, dw, dh, dpitch;
};
__global__ void __launch_bounds__(256) k_decimate_b(const DecimateParams p) {
    pdl_wait();
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    const long long img = blockIdx.z;
    if (dx >= p.dw) return;
    const float ry = (float)p.h / (float)p.dh, rx = (float)p.w / (float)p.dw;
    const int sy = min(max((int)floorf(((float)dy + 0.5f) * ry), 0), p.h - 1);
    const int sx = min(max((int)floorf(((float)dx + 0.5f) * rx), 0), p.w - 1);
    const float v = p.src[img * p.img_stride + (long long)sy * p.pitch + sx];
    p.dst[img * p.img_stride + (long long)dy * p.dpitch + dx] = fminf(fmaxf(v, 0.0f), 1.0f);
}

// ---------------------------------------------------------------------------
// TMA variant of the blur for octaves of at least TMA_MIN_DIM pixels per side: the
// (TH+2R) x BW input box (halo included) is fetched by ONE cp.async.bulk.tensor
// (UTMALDG) issued by an elected thread and completed on an mbarrier, instead of
// ~54 dependent global loads per thread.  Out-of-image box elements arrive as
// zeros; tiles on the image border then patch their halo with BORDER_REFLECT_101
// copies taken from the box itself.  Same row/column arithmetic as k_blur.
// The tensor map is 4-D: (x, y, layer, image) over one octave of the slot's arena.
// ---------------------------------------------------------------------------
constexpr int TMA_MIN_DIM = 32;
// output tile width per tap set (measured on B200): the narrow taps run faster with 64-column tiles
// (3 CTAs per SM), the 21- and 27-tap kernels with 128-column tiles (less horizontal halo per output)
__host__ __device__ constexpr int blur_tile_w(int l) { return l <= 3 ? 64 : 128; }

template <int L>
struct TmaCfg {
    static constexpr int R = blur_radius(L);
    static constexpr int RA = (R + 3) / 4 * 4;   // left halo of the box: the box's first column must be 16-byte aligned
    static constexpr int XO = RA - R;            // box column of the first element the filter needs
    static constexpr int TW = blur_tile_w(L), TH = 64;
    static constexpr int SH = TH + 2 * R;
    static constexpr int SW = TW + 2 * R;
    static constexpr int WIN = XO + 8 + 2 * R;   // box floats read for 8 consecutive row-pass outputs (aligned start)
    static constexpr int NV4 = (WIN + 3) / 4;
    // box width: multiple of 4 floats (16 B) with BW/4 odd, so that 8 consecutive rows start in 8
    // distinct 4-bank groups (conflict-free LDS.128 with lanes <-> rows)
    static constexpr int BW_MIN = (XO + SW > TW - 8 + 4 * NV4) ? XO + SW : TW - 8 + 4 * NV4;
    static constexpr int BW4 = (BW_MIN + 3) / 4;
    static constexpr int BW = 4 * ((BW4 % 2) ? BW4 : BW4 + 1);
    static constexpr int IPITCH = TW + 4;        // == 4 (mod 32) for TW = 64, 128
    static constexpr int PY = 8;                 // rows per column-pass thread (which owns two adjacent columns)
    static constexpr int THREADS = 256;
    // the box arrives as NBAND row bands of BAND rows, each on its own mbarrier, so the row pass of the first
    // 32 staged rows starts while the rest of the box is still in flight
    static constexpr int BAND = 32;
    static constexpr int NBAND = (SH + BAND - 1) / BAND;
    static constexpr int SHP = NBAND * BAND;       // staged rows incl. padding of the last band
    static constexpr uint32_t BAND_BYTES = (uint32_t)BAND * BW * sizeof(float);
    static constexpr size_t SMEM = ((size_t)SHP * BW + (size_t)SH * IPITCH) * sizeof(float);
    static constexpr int CTAS_PER_SM = (TW == 64) ? 3 : 2;
    static_assert(BW >= XO + SW && (BW / 4) % 2 == 1 && TW - 8 + 4 * NV4 <= BW, "box width");
    static_assert(BAND_BYTES % 128 == 0 && BW <= 256, "TMA box limits / 128-byte aligned band destinations");
    static_assert(IPITCH % 32 == 4, "inter pitch");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int L, bool DECIMATE>
__global__ void __launch_bounds__(256, TmaCfg<L>::CTAS_PER_SM) k_blur_tma(const __grid_constant__ CUtensorMap tmap, const BlurParams p,
                                                      const int src_layer) {
    pdl_wait();
    using C = TmaCfg<L>;
    constexpr int R = C::R;
    extern __shared__ __align__(1024) float smem_tma[];  // own symbol: `smem` above is declared 16-byte aligned
    __shared__ __align__(8) uint64_t bar[C::NBAND];
    float* stage = smem_tma;                     // SHP x BW, dense (TMA box layout)
    float* inter = smem_tma + C::SHP * C::BW;    // SH x IPITCH
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * C::TW, ty0 = blockIdx.y * C::TH;
    const int img = blockIdx.z;
    const int w = p.w, h = p.h;
    const int rows_needed = min(C::TH, h - ty0) + 2 * R;
    const int cols_needed = min(C::TW, w - tx0) + 2 * R;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < C::NBAND; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < C::NBAND; b++) {
            const uint32_t bar_a = smem_u32(&bar[b]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(C::BAND_BYTES) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                ::"r"(smem_u32(stage + b * C::BAND * C::BW)), "l"(&tmap), "r"(tx0 - C::RA), "r"(ty0 - R + b * C::BAND),
                  "r"(src_layer), "r"(img), "r"(bar_a)
                : "memory");
        }
    }
    const bool border_tile = tx0 - R < 0 || ty0 - R < 0 || tx0 + C::TW + R > w || ty0 + C::TH + R > h;
    auto wait_band = [&](int b) {
        uint32_t done = 0;
        const uint32_t bar_a = smem_u32(&bar[b]);
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
        }
    };
    if (border_tile) {  // the halo patch below reads across bands: wait for the whole box
#pragma unroll
        for (int b = 0; b < C::NBAND; b++) wait_band(b);
    }
    // BORDER_REFLECT_101 for tiles that stick out of the image (block-uniform branch).
    // Box element (row, XO + col) holds image pixel (ty0 - R + row, tx0 - R + col).
    if (border_tile) {
        for (int idx = tid; idx < C::SH * C::SW; idx += C::THREADS) {
            const int row = idx / C::SW, col = idx - row * C::SW;
            if (row < rows_needed && col < cols_needed) {
                const int gy = ty0 - R + row, gx = tx0 - R + col;
                if (gy < 0 || gy >= h || gx < 0 || gx >= w) {
                    const int ry = reflect101(gy, h) - (ty0 - R), rx = reflect101(gx, w) - (tx0 - R);
                    // the source pixel is inside the image, hence inside the box
                    stage[row * C::BW + C::XO + col] = stage[ry * C::BW + C::XO + rx];
                }
            }
        }
        __syncthreads();
    }

    // ---- row pass: warp = 32 staged rows x one 8-pixel segment ----
    {
        constexpr int NGROUPS = (C::SH + 31) / 32;
        constexpr int NSEG = C::TW / 8;
        for (int task = warp; task < NGROUPS * NSEG; task += C::THREADS / 32) {
            const int g = task / NSEG, seg = task - g * NSEG;
            const int row = g * 32 + lane;
            if (!border_tile) wait_band(g);   // warp-uniform; a no-op once the band has landed
            if (row < rows_needed && row < C::SH) {
                float win[4 * C::NV4];
                const float4* sp = reinterpret_cast<const float4*>(stage + row * C::BW + seg * 8);
#pragma unroll
                for (int v = 0; v < C::NV4; v++) {
                    float4 q = sp[v];
                    win[4 * v + 0] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
                }
                // packed arithmetic: outputs (2jp, 2jp+1) share one FFMA2 chain; the operand pair for tap i starts
                // at window index XO + 2jp + i, which is an aligned register pair of `win` when even and comes
                // from the one-float-shifted copy `sh` when odd
                float2 sh[2 * C::NV4];
#pragma unroll
                for (int m = 0; 2 * m + 2 < 4 * C::NV4; m++) sh[m] = make_float2(win[2 * m + 1], win[2 * m + 2]);
                float2 acc[4];
#pragma unroll
                for (int jp = 0; jp < 4; jp++) {
#pragma unroll
                    for (int i = 0; i <= 2 * R; i++) {
                        const int sidx = C::XO + 2 * jp + i;
                        const float2 in = (sidx & 1) ? sh[sidx >> 1] : make_float2(win[sidx], win[sidx + 1]);
                        acc[jp] = (i == 0) ? mul2(in, c_taps2[L][0]) : fma2(in, c_taps2[L][i], acc[jp]);
                    }
                }
                float4* ip = reinterpret_cast<float4*>(inter + row * C::IPITCH + seg * 8);
                ip[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
                ip[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
            }
        }
    }
    __syncthreads();

    // ---- column pass: thread = two adjacent columns x PY consecutive rows, packed f32x2 ----
    {
        float* dst = p.dst + (long long)img * p.img_stride;
        float* dec = DECIMATE ? p.dec + (long long)img * p.img_stride : nullptr;
        constexpr int NCH = C::TH / C::PY;
        constexpr int NXP = C::TW / 2;
        for (int task = tid; task < NXP * NCH; task += C::THREADS) {
            const int cy = task / NXP, xp = task - cy * NXP;
            const int y0 = cy * C::PY;
            const int gx = tx0 + 2 * xp;
            if (ty0 + y0 >= h) continue;  // whole chunk below the image (warp-uniform)
            float2 c[C::PY + 2 * R];
#pragma unroll
            for (int j = 0; j < C::PY + 2 * R; j++)
                c[j] = *reinterpret_cast<const float2*>(inter + (y0 + j) * C::IPITCH + 2 * xp);
#pragma unroll
            for (int j = 0; j < C::PY; j++) {
                float2 acc = mul2(c[j + R], c_taps2[L][R]);
#pragma unroll
                for (int i = 1; i <= R; i++) acc = fma2(add2(c[j + R + i], c[j + R - i]), c_taps2[L][R + i], acc);
                const int gy = ty0 + y0 + j;
                if (gy < h && gx < w) {
                    float* q = dst + (long long)gy * p.pitch + gx;
                    if (gx + 1 < w) *reinterpret_cast<float2*>(q) = acc;
                    else q[0] = acc.x;
                    if (DECIMATE && !(gy & 1)) {   // gx is even
                        const int dy = gy >> 1, dx = gx >> 1;
                        if (dy < p.dec_h && dx < p.dec_w) dec[(long long)dy * p.dec_pitch + dx] = acc.x;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Marching variant of the TMA blur (the one the pipeline uses for octaves of at least TMA_MIN_DIM pixels per
// side).  A CTA owns a 128-column strip of one vertical segment of the image and walks down it in bands of 32
// rows: while band b+1 is in flight (cp.async.bulk.tensor, two stage buffers, one mbarrier each) the CTA runs
// the row pass of band b into a ring of row-pass results and the column pass of the 32 output rows whose
// (2R+1)-row window is complete.  Compared with independent (64 + 2R)-row tiles this computes every row-pass
// result once instead of (64 + 2R) / 64 times, keeps one band in flight per CTA at all times, and needs a
// single __syncthreads per band.
//   * ring: three 32-row slots (band b in slot b % 3) followed by a mirror of the first 2R rows of slot 0, so
//     the rows a column-pass thread reads (band b and the first 2R rows of band b+1) are always contiguous and
//     every smem offset is an immediate;
//   * step j = { issue TMA of band j+NSTG; row pass of band j+1; column pass of band j-1; barrier };
//   * BORDER_REFLECT_101: left/right halo columns are patched in the stage buffer (edge strips only, as in
//     k_blur_tma); rows above / below the image are row-passed from the stage row they mirror, or copied from
//     the ring when that row belongs to the previous band.
// Same arithmetic, operation for operation, as k_blur / k_blur_tma.
// ---------------------------------------------------------------------------
// strip width per tap set (measured on B200): the 17- and 21-tap kernels run faster as four 128-thread CTAs per SM
// on 64-column strips, the others as two 256-thread CTAs on 128-column strips
#ifndef SB_MARCH_TW
#define SB_MARCH_TW(l) (((l) == 3 || (l) == 4) ? 64 : 128)
#endif
__host__ __device__ constexpr int march_tile_w(int l) { return SB_MARCH_TW(l); }
// consecutive outputs per row-pass task: 16 (eight independent FMA chains per lane, one task per warp and band)
// measured 4-7 % faster than 8 on every tap set
#ifndef SB_MARCH_BH
#define SB_MARCH_BH(l) 32
#endif
#ifndef SB_MARCH_PY
#define SB_MARCH_PY(l) 8
#endif
#ifndef SB_MARCH_CTAS
#define SB_MARCH_CTAS(l) 0   // resident CTAs per SM the kernel is compiled for; 0: 512 threads' worth
#endif
#ifndef SB_MARCH_SEG
#define SB_MARCH_SEG(l) 16
#endif

template <int L, int FL = FL_OPENCV>
struct MarchCfg {
    static constexpr int R = blur_radius(L, FL);
    static constexpr int RA = (R + 3) / 4 * 4;   // left halo of the box: a TMA box starts on a 16-byte boundary
    static constexpr int XO = RA - R;            // box column of the first element the filter needs
    static constexpr int TW = march_tile_w(L), BH = SB_MARCH_BH(L);   // strip width, rows per band
    static constexpr int SW = TW + 2 * R;
    static constexpr int SEG = SB_MARCH_SEG(L);  // consecutive outputs of one row-pass task (one lane): SEG / 2 FMA chains
    static constexpr int WIN = XO + SEG + 2 * R; // box floats read for them (aligned start)
    static constexpr int NV4 = (WIN + 3) / 4;
    static constexpr int BW_MIN = (XO + SW > TW - SEG + 4 * NV4) ? XO + SW : TW - SEG + 4 * NV4;
    static constexpr int BW4 = (BW_MIN + 3) / 4;
    static constexpr int BW = 4 * ((BW4 % 2) ? BW4 : BW4 + 1);   // BW/4 odd: conflict-free LDS.128 with lanes <-> rows
    static constexpr int IPITCH = TW + 4;        // == 4 (mod 32)
    static constexpr int PY = SB_MARCH_PY(L);    // output rows per column-pass thread (which owns two adjacent columns)
    static constexpr int THREADS = (TW / 2) * (BH / PY);   // one column-pass task per thread
    static constexpr int CTAS_PER_SM = SB_MARCH_CTAS(L) ? SB_MARCH_CTAS(L) : 512 / THREADS;
    // stage buffers (bands in flight + the one being filtered): the narrow tap sets are memory-bound and need two
    // bands in flight per CTA to hide the TMA latency behind their short steps; the wide ones have no room for a
    // third buffer next to two resident CTAs and do not need it
    static constexpr int NSTG = (L <= 2 && BH == 32) ? 3 : 2;
    static constexpr int RING_ROWS = 3 * BH + 2 * R;   // three band slots + a mirror of the first 2R rows of slot 0
    static constexpr uint32_t BAND_BYTES = (uint32_t)BH * BW * sizeof(float);
#ifndef SB_MARCH_PAD
#define SB_MARCH_PAD 0
#endif
    static constexpr size_t SMEM = (size_t)NSTG * BAND_BYTES + (size_t)RING_ROWS * IPITCH * sizeof(float) + SB_MARCH_PAD;
    static_assert(BW >= XO + SW && (BW / 4) % 2 == 1 && TW - SEG + 4 * NV4 <= BW && BW <= 256, "box width");
    static_assert(BAND_BYTES % 128 == 0, "128-byte aligned stage buffers");
    static_assert(2 * R + 1 <= BH && PY + 2 * R <= 2 * BH && BH % 32 == 0, "window spans at most two bands");
    static_assert((TW / 2) * (BH / PY) == THREADS, "one column-pass task per thread");
};

// SEEDF != 0 (the seed blur of the OpenCV flavour, L = 0): the input bands are not loaded but COMPUTED -- the CTA upsamples
// the u8 input into its stage buffers (upsample_block: the arithmetic of k_upsample2x, 2 x 4 output blocks aligned with the
// band: band rows start at an odd row of the upsampled image and box columns at a multiple of 4) -- so the upsampled image
// is never written to and read back from HBM (12.25 -> 4.25 bytes per seed pixel).  Columns / rows outside the image are
// left to the border logic below, exactly as for a loaded band; the results are bit-identical to the two-kernel seed.
//   SEEDF == 1: all threads produce band j+1 into ONE stage buffer, barrier, row pass of j+1 and column pass of j-1,
//     barrier.  Measured SLOWER than the two kernels (the u8 loads of the produce phase are exposed and the phases
//     serialise: 731 against 563 us per 32-image 1080p group).
//   SEEDF == 2 (what the pipeline uses): warp-specialised -- SEED_PRODUCERS extra threads (four warps) do nothing but
//     produce bands, up to NSTG ahead of the eight filter warps, which run the unchanged marching code; a `full` mbarrier
//     per stage buffer takes the place of the TMA's (one arrival per producer warp), an `empty` one hands the buffer back,
//     and the filter warps synchronise among themselves on a named barrier.  A producer thread owns a group of four box
//     columns and a third of the band's 16 row pairs, loads its input bytes in one round and interpolates every input row
//     horizontally once.  450 against 565 us (two kernels) per 32-image 1080p group.
#ifndef SB_SEED_PRODUCERS
#define SB_SEED_PRODUCERS 128
#endif
#ifndef SB_SEED_PREFETCH
#define SB_SEED_PREFETCH 0   // L1 prefetch of the next band's input lines by the producer warps: measured 2 % slower
#endif
#ifndef SB_SEED_LUT
#define SB_SEED_LUT 0        // 1: v / 255 from the shared-memory table (random banks: conflicts) instead of norm255()
#endif
// v / 255 for an 8-bit v, the correctly rounded f32 quotient without a division: one Newton step on v * RN(1/255)
// (verified equal to the IEEE division for all 256 values; the bit-exact pyramid tests would catch any other)
__device__ __forceinline__ float norm255(const uint32_t v) {
    const float f = (float)v, r = 1.0f / 255.0f;
    const float q = f * r;
    return fmaf(fmaf(-q, 255.0f, f), r, q);
}
constexpr int SEED_PRODUCERS = SB_SEED_PRODUCERS;
template <int L, bool DECIMATE, int FL = FL_OPENCV, int SEEDF = 0>
__global__ void __launch_bounds__(MarchCfg<L, FL>::THREADS + (SEEDF == 2 ? SEED_PRODUCERS : 0), MarchCfg<L, FL>::CTAS_PER_SM)
k_blur_march(const __grid_constant__ CUtensorMap tmap, const BlurParams p, const int src_layer, const int bands_per_cta,
             const int strips, const long long total_bands) {
    pdl_wait();
    static_assert(FL == FL_OPENCV || !DECIMATE, "the imageproc flavour decimates in its own kernel");
    static_assert(!SEEDF || (L == 0 && FL == FL_OPENCV && !DECIMATE), "the fused seed is layer 0 of the OpenCV flavour");
    using C = MarchCfg<L, FL>;
    static_assert(!SEEDF || (C::RA % 4 == 0 && C::BW % 4 == 0 && C::TW % 4 == 0 && C::BH % 2 == 0 && (C::R & 1) == 1 && C::THREADS >= 256),
                  "upsample blocks are aligned with the band");
    // barrier among the filter warps (all threads of the CTA unless producer warps exist)
    auto csync = [&]() {
        if constexpr (SEEDF == 2) asm volatile("bar.sync 1, %0;" ::"n"(C::THREADS) : "memory");
        else __syncthreads();
    };
    constexpr int R = C::R;
    constexpr int STAGE_FLOATS = C::BH * C::BW, SLOT_FLOATS = C::BH * C::IPITCH;
    extern __shared__ __align__(1024) float smem_march[];
    __shared__ __align__(8) uint64_t bar[C::NSTG];
    __shared__ __align__(8) uint64_t empty_bar[SEEDF == 2 ? C::NSTG : 1];
    float* const stage = smem_march;                               // NSTG x BH x BW (TMA box layout)
    float* const inter = smem_march + C::NSTG * STAGE_FLOATS;      // RING_ROWS x IPITCH
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = p.w, h = p.h;
    __shared__ float s_norm[SEEDF ? 256 : 1];   // v / 255 for the 256 pixel values (the IEEE division, once per CTA)
    if (SEEDF && tid < 256) s_norm[tid] = (float)tid / 255.0f;
    // Work distribution: the (strip, image) columns of the launch, each cut into bands of BH output rows, form one
    // sequence of total_bands bands; CTA c owns the bands [c * bands_per_cta, (c + 1) * bands_per_cta) of it and marches
    // down every piece of a column that falls into its range.  Every CTA gets the same number of bands whatever the
    // image height, strip count and batch size are, so a launch has no partially filled last wave.
    const int nb = (h + C::BH - 1) / C::BH;                        // bands per column
    long long gb, gb_end;
    if (bands_per_cta > 0) {
        gb = (long long)blockIdx.x * bands_per_cta;
        gb_end = min(gb + bands_per_cta, total_bands);
    } else {
        // aligned mode: every column is cut into k = -bands_per_cta pieces of (almost) equal length and a CTA owns exactly
        // one piece -- one pipeline start per CTA, and the CTAs of neighbouring strips walk the same rows at the same time
        const int k = -bands_per_cta;
        const int unit = blockIdx.x / k, pc = blockIdx.x - unit * k;
        gb = (long long)unit * nb + (long long)pc * nb / k;
        gb_end = (long long)unit * nb + (long long)(pc + 1) * nb / k;
    }
    bool first_piece = true;
    while (gb < gb_end) {
    const int unit = (int)(gb / nb), band0 = (int)(gb - (long long)unit * nb);
    const int take = (int)min((long long)(nb - band0), gb_end - gb);
    gb += take;
    const int tx0 = (unit % strips) * C::TW;
    const int img = unit / strips;
    const int ya = band0 * C::BH;
    const int yb = min(ya + take * C::BH, h);
    const int n_out = (yb - ya + C::BH - 1) / C::BH;           // output bands
    const int n_in = (yb - ya + 2 * R + C::BH - 1) / C::BH;    // input bands (n_out or n_out + 1)
    const int in0 = ya - R;                                    // first input row of band 0
    const bool hedge = tx0 - R < 0 || tx0 + C::TW + R > w;
    if (SEEDF != 1 && tid == 0) {   // fresh barriers per piece (every load of the previous piece has been waited for)
#pragma unroll
        for (int b = 0; b < C::NSTG; b++) {
            if (!first_piece) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar[b])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[b])), "n"(SEEDF == 2 ? SEED_PRODUCERS / 32 : 1));
            if (SEEDF == 2) {
                if (!first_piece) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&empty_bar[b])) : "memory");
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty_bar[b])));
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    first_piece = false;
    __syncthreads();   // (all threads of the CTA, producer warps included)
    auto issue = [&](const int b, const int stg) {   // band b into stage buffer stg == b % NSTG
        const uint32_t bar_a = smem_u32(&bar[stg]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(C::BAND_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(stage + stg * STAGE_FLOATS)), "l"(&tmap), "r"(tx0 - C::RA), "r"(in0 + b * C::BH),
              "r"(src_layer), "r"(img), "r"(bar_a)
            : "memory");
    };
    if (!SEEDF && tid == 0) {
#pragma unroll
        for (int b = 0; b < C::NSTG; b++)
            if (b < n_in) issue(b, b);
    }
    // SEEDF: input band b, upsampled from the u8 image straight into stage buffer 0.  Band row pair rp holds the rows
    // {2y+1, 2y+2} of the upsampled image with y = (band_y0 - 1) / 2 + rp (band_y0 is odd), box column group g its
    // columns {4k .. 4k+3} with k = (tx0 - RA) / 4 + g.
    auto produce = [&](const int b, float* const dst, const int t0, const int nt) {
        const int ypair0 = (in0 + b * C::BH - 1) >> 1;
        const int kg0 = (tx0 - C::RA) >> 2;
        const uint8_t* const in = p.in + (long long)img * p.in_img_stride;
        constexpr int NG = C::BW / 4;
        for (int t = t0; t < (C::BH / 2) * NG; t += nt) {
            const int rp = t / NG, g = t - rp * NG;
            const int y = ypair0 + rp, k = kg0 + g;
            if (k < 0 || 4 * k >= w || y < -1 || y >= p.in_h) continue;   // wholly outside the image: never read
            float o[2][4];
            upsample_block(in, p.in_w, p.in_h, p.in_stride, s_norm, k, y, o);
            float4* const q = reinterpret_cast<float4*>(dst + (2 * rp) * C::BW + 4 * g);
            q[0] = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
            q[C::BW / 4] = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
        }
    };
    auto mbar_wait = [&](uint64_t* const b, const uint32_t parity) {
        const uint32_t a = smem_u32(b);
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done) : "r"(a), "r"(parity) : "memory");
        }
    };
    if constexpr (SEEDF == 2) {
        if (tid >= C::THREADS) {   // producer warps: band after band of this piece, at most NSTG ahead of the filter warps
            for (int b = 0; b < n_in; b++) {
                const int stg = b % C::NSTG;
                if (b >= C::NSTG) {   // the producers run ahead: wait politely (a tight poll takes issue slots from the filter warps)
                    const uint32_t a = smem_u32(&empty_bar[stg]), parity = (uint32_t)((b / C::NSTG - 1) & 1);
                    for (;;) {
                        uint32_t done;
                        asm volatile(
                            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                            : "=r"(done) : "r"(a), "r"(parity) : "memory");
                        if (done) break;
                        __nanosleep(256);
                    }
                }
                {   // A producer thread owns one group of four box columns and walks down a third of the band's 16 row
                    // pairs: every input row is loaded and interpolated horizontally once and serves the pair above and
                    // the pair below it; all (at most 28) pixel loads of a thread are in flight together.
                    float* const dst = stage + stg * STAGE_FLOATS;
                    const int ypair0 = (in0 + b * C::BH - 1) >> 1;
                    const uint8_t* const in = p.in + (long long)img * p.in_img_stride;
                    constexpr int NG = C::BW / 4;
                    static_assert(C::BH == 32 && 3 * NG <= SEED_PRODUCERS, "three producer threads per column group");
                    const int pt = tid - C::THREADS;
                    const int seg = pt / NG, g = pt - seg * NG;
                    const int k = ((tx0 - C::RA) >> 2) + g;
                    if (seg < 3 && k >= 0 && 4 * k < w) {
                        const int W = p.in_w, H = p.in_h;
                        const int rp0 = seg == 0 ? 0 : (seg == 1 ? 6 : 11), np = seg == 0 ? 6 : 5;   // row pairs [rp0, rp0 + np)
                        uint32_t raw[7][4];
                        if (k >= 1 && 2 * k + 2 < W) {   // interior columns: four consecutive bytes from one row address
#pragma unroll
                            for (int r = 0; r < 7; r++) {   // input rows of the pairs: pair rp reads rows rp and rp + 1 (clamped)
                                const int yr = min(max(ypair0 + rp0 + min(r, np), 0), H - 1);
                                const uint8_t* const row = in + ((long long)yr * p.in_stride + (2 * k - 1));
#pragma unroll
                                for (int c = 0; c < 4; c++) raw[r][c] = row[c];
                            }
                        } else {
                            int xo[4];
#pragma unroll
                            for (int c = 0; c < 4; c++) xo[c] = min(max(2 * k - 1 + c, 0), W - 1);
#pragma unroll
                            for (int r = 0; r < 7; r++) {
                                const int yr = min(max(ypair0 + rp0 + min(r, np), 0), H - 1);
                                const uint8_t* const row = in + (long long)yr * p.in_stride;
#pragma unroll
                                for (int c = 0; c < 4; c++) raw[r][c] = row[xo[c]];
                            }
                        }
#if SB_SEED_PREFETCH
                        if (b + 1 < n_in) {   // the same columns one band further down: pull their lines into L1 for the next round
#pragma unroll
                            for (int r = 1; r < 7; r++) {
                                const int yr = min(max(ypair0 + C::BH / 2 + rp0 + min(r, np), 0), H - 1);
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(in + ((long long)yr * p.in_stride + max(2 * k - 1, 0))));
                            }
                        }
#endif
                        float2 hr[7][2];   // horizontally interpolated rows, as the column pairs the packed vertical pass takes
#pragma unroll
                        for (int r = 0; r < 7; r++) {
#if SB_SEED_LUT
                            const float a0 = s_norm[raw[r][0]], a1 = s_norm[raw[r][1]], a2 = s_norm[raw[r][2]], a3 = s_norm[raw[r][3]];
#else
                            const float a0 = norm255(raw[r][0]), a1 = norm255(raw[r][1]), a2 = norm255(raw[r][2]), a3 = norm255(raw[r][3]);
#endif
                            hr[r][0].x = k == 0 ? a1 : fmaf(a1 - a0, 0.75f, a0);   // 2k-1 < 0: both taps are column 0
                            hr[r][0].y = fmaf(a2 - a1, 0.25f, a1);
                            hr[r][1].x = fmaf(a2 - a1, 0.75f, a1);
                            hr[r][1].y = fmaf(a3 - a2, 0.25f, a2);
                        }
                        const float2 f25 = make_float2(0.25f, 0.25f), f75 = make_float2(0.75f, 0.75f);
#pragma unroll
                        for (int q = 0; q < 6; q++) {
                            const int y = ypair0 + rp0 + q;
                            if (q < np && y >= -1 && y < H) {
                                // fma(bottom - top, f, top) per output, two columns per packed operation
                                const float2 d0 = sub2(hr[q + 1][0], hr[q][0]), d1 = sub2(hr[q + 1][1], hr[q][1]);
                                const float2 u0 = fma2(d0, f25, hr[q][0]), u1 = fma2(d1, f25, hr[q][1]);
                                const float2 v0 = fma2(d0, f75, hr[q][0]), v1 = fma2(d1, f75, hr[q][1]);
                                float4* const out = reinterpret_cast<float4*>(dst + (2 * (rp0 + q)) * C::BW + 4 * g);
                                out[0] = make_float4(u0.x, u0.y, u1.x, u1.y);
                                out[C::BW / 4] = make_float4(v0.x, v0.y, v1.x, v1.y);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[stg])) : "memory");
            }
            continue;   // next piece (its barriers are initialised behind the CTA-wide barrier above)
        }
    }

    // ---- row pass of input band b (stage buffer stg, its mbarrier at `parity`; ring slot `slot`):
    //      warp = the band's 32 rows x one 8-pixel segment ----
    auto row_pass = [&](const int b, const int stg_, const uint32_t parity, const int slot) {
        const int stg = SEEDF == 1 ? 0 : stg_;
        if (SEEDF != 1) mbar_wait(&bar[stg], parity);
        float* const st = stage + stg * STAGE_FLOATS;
        const int band_y0 = in0 + b * C::BH;
        if (hedge) {  // block-uniform: box element (row, XO + col) holds image pixel (band_y0 + row, tx0 - R + col)
            // only the R columns left of the image and the first R columns right of it feed a stored output: a warp
            // takes one such column, its lanes the rows of the band
            const int nl = max(R - tx0, 0);                       // strip columns [0, nl) lie left of the image
            const int cr0 = w - tx0 + R;                          // first strip column right of the image
            const int np = nl + max(min(C::SW, cr0 + R) - cr0, 0);
            for (int c = warp; c < np; c += C::THREADS / 32) {
                const int col = c < nl ? c : cr0 + (c - nl);
                const int rx = border_index<FL>(tx0 - R + col, w) - (tx0 - R);
                if (rx >= 0 && rx < C::SW) {
                    for (int row = lane; row < C::BH; row += 32) st[row * C::BW + C::XO + col] = st[row * C::BW + C::XO + rx];
                }
            }
            csync();
        }
        float* const ib = inter + slot * SLOT_FLOATS;
        const bool vedge = band_y0 < 0 || band_y0 + C::BH > h;                      // block-uniform
        const int rows_needed = yb + R - band_y0;   // >= BH except in the last band
        constexpr int NSEG = C::TW / C::SEG;
#pragma unroll 1
        for (int task = warp; task < NSEG * (C::BH / 32); task += C::THREADS / 32) {   // task = 32 rows x one segment
            const int seg = task % NSEG;
            const int row = (task / NSEG) * 32 + lane;
            if (row >= rows_needed) continue;
            int srow = row;         // stage row this lane filters
            bool copy_prev = false; // the mirrored row belongs to the previous band: copy its row-pass result from the ring
            if (vedge) {
                const int yy = border_index<FL>(band_y0 + row, h);
                srow = yy - band_y0;
                copy_prev = srow < 0;
            }
            const float* const srcrow = st + srow * C::BW;
            float* const dstrow = ib + row * C::IPITCH;
            float4 o[C::SEG / 4];
            if (copy_prev) {
                const int pslot = slot == 0 ? 2 : slot - 1;
                const float4* src = reinterpret_cast<const float4*>(inter + pslot * SLOT_FLOATS + (srow + C::BH) * C::IPITCH + seg * C::SEG);
#pragma unroll
                for (int v = 0; v < C::SEG / 4; v++) o[v] = src[v];
            } else {
                float win[4 * C::NV4];
                const float4* sp = reinterpret_cast<const float4*>(srcrow + seg * C::SEG);
#pragma unroll
                for (int v = 0; v < C::NV4; v++) {
                    float4 q = sp[v];
                    win[4 * v + 0] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
                }
                // outputs (2jp, 2jp+1) share one accumulator pair.  The operands of tap i start at window index
                // XO + 2jp + i: when that is even they are an aligned register pair and the step is one packed
                // FFMA2; when it is odd the pair straddles two registers pairs, and two scalar FFMAs on the halves
                // of the accumulator cost less than assembling the shifted pair (same FMA-pipe cycles, no MOVs)
                float2 acc[C::SEG / 2];
#pragma unroll
                for (int jp = 0; jp < C::SEG / 2; jp++) {
#pragma unroll
                    for (int i = 0; i <= 2 * R; i++) {
                        const int sidx = C::XO + 2 * jp + i;
                        // the kernel is symmetric (tap i == tap 2R - i, bit for bit): naming every tap by its upper-half
                        // index keeps the set of constants the row and the column pass touch to R + 1 values
                        const int ti = i < R ? 2 * R - i : i;
                        if ((sidx & 1) == 0) {
                            const float2 in = make_float2(win[sidx], win[sidx + 1]);
                            if (i == 0) acc[jp] = mul2(in, tap2<FL>(L, ti));
                            else if (FL == FL_OPENCV) acc[jp] = fma2(in, tap2<FL>(L, ti), acc[jp]);
                            else acc[jp] = madd_unfused2(acc[jp], in, tap2<FL>(L, ti));
                        } else if (i == 0) {
                            acc[jp] = make_float2(win[sidx] * tap<FL>(L, ti), win[sidx + 1] * tap<FL>(L, ti));
                        } else if (FL == FL_OPENCV) {
                            acc[jp].x = fmaf(win[sidx], tap<FL>(L, ti), acc[jp].x);
                            acc[jp].y = fmaf(win[sidx + 1], tap<FL>(L, ti), acc[jp].y);
                        } else {
                            acc[jp].x = __fadd_rn(acc[jp].x, __fmul_rn(win[sidx], tap<FL>(L, ti)));
                            acc[jp].y = __fadd_rn(acc[jp].y, __fmul_rn(win[sidx + 1], tap<FL>(L, ti)));
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < C::SEG / 4; v++) o[v] = make_float4(acc[2 * v].x, acc[2 * v].y, acc[2 * v + 1].x, acc[2 * v + 1].y);
            }
            float4* const ip = reinterpret_cast<float4*>(dstrow + seg * C::SEG);
#pragma unroll
            for (int v = 0; v < C::SEG / 4; v++) ip[v] = o[v];
            if (slot == 0 && row < 2 * R) {  // mirror of the head of slot 0 behind slot 2
#pragma unroll
                for (int v = 0; v < C::SEG / 4; v++) ip[3 * SLOT_FLOATS / 4 + v] = o[v];
            }
        }
    };

    // ---- column pass: thread = two adjacent columns x PY consecutive rows of the band, packed f32x2 ----
    const int cy = tid / (C::TW / 2), xp = tid - cy * (C::TW / 2);
    const int y0 = cy * C::PY;
    const int gx = tx0 + 2 * xp;
    const bool x_full = gx + 1 < w;
    float* qband = p.dst + (long long)img * p.img_stride + (long long)(ya + y0) * p.pitch + gx;   // first output row of band 0
    float* dband = nullptr;
    if (DECIMATE) dband = p.dec + (long long)img * p.img_stride + (long long)((ya + y0) >> 1) * p.dec_pitch + (gx >> 1);
    const float* const cbase = inter + y0 * C::IPITCH + 2 * xp;
    // output band jb (ring slot `slot`; band jb+1 follows contiguously: slot 3 mirrors slot 0)
    auto col_pass = [&](const int jb, const int slot) {
        const int gy0 = ya + jb * C::BH + y0;
        if (gy0 < yb && gx < w) {  // else: whole chunk below the segment (warp-uniform) / columns right of the image
            const float* const base = cbase + slot * SLOT_FLOATS;
            float2 c[C::PY + 2 * R];
#pragma unroll
            for (int j = 0; j < C::PY + 2 * R; j++) c[j] = *reinterpret_cast<const float2*>(base + j * C::IPITCH);
            // all PY outputs first (PY independent FADD2 -> FFMA2 chains the scheduler can interleave), stores after
            float2 out[C::PY];
#pragma unroll
            for (int j = 0; j < C::PY; j++) {
                float2 acc;
                if (FL == FL_OPENCV) {
                    acc = mul2(c[j + R], tap2<FL>(L, R));
#pragma unroll
                    for (int i = 1; i <= R; i++) acc = fma2(add2(c[j + R + i], c[j + R - i]), tap2<FL>(L, R + i), acc);
                } else {   // top to bottom, a multiply and an add per tap (taps named by their upper-half index)
                    acc = mul2(c[j], tap2<FL>(L, 2 * R));
#pragma unroll
                    for (int i = 1; i <= 2 * R; i++) acc = madd_unfused2(acc, c[j + i], tap2<FL>(L, i < R ? 2 * R - i : i));
                }
                out[j] = acc;
            }
            float* q = qband;
            if (x_full && gy0 + C::PY <= yb) {
#pragma unroll
                for (int j = 0; j < C::PY; j++) {
                    *reinterpret_cast<float2*>(q) = out[j];
                    q += p.pitch;
                }
            } else {
#pragma unroll
                for (int j = 0; j < C::PY; j++) {
                    if (gy0 + j < yb) {
                        if (x_full) *reinterpret_cast<float2*>(q) = out[j];
                        else q[0] = out[j].x;
                    }
                    q += p.pitch;
                }
            }
            if (DECIMATE) {   // ya, y0 and gx are even: output row gy0 + j is even for even j
                const int dx = gx >> 1;
#pragma unroll
                for (int j = 0; j < C::PY; j += 2) {
                    const int dy = (gy0 + j) >> 1;
                    if (gy0 + j < yb && dy < p.dec_h && dx < p.dec_w) dband[(long long)(j >> 1) * p.dec_pitch] = out[j].x;
                }
            }
        }
        qband += (long long)C::BH * p.pitch;
        if (DECIMATE) dband += (long long)(C::BH / 2) * p.dec_pitch;
    };

    if (SEEDF == 1) { produce(0, stage, tid, C::THREADS); csync(); }
    row_pass(0, 0, 0u, 0);
    csync();
    int rslot = 1, cslot = 0;   // ring slots of band j+1 (row pass) and band j-1 (column pass)
    int fstg = 0;               // stage buffer of band j (free again), == j % NSTG
    int rstg = 1;               // stage buffer of band j+1, == (j+1) % NSTG
    uint32_t rpar = 0;          // ... and the phase of its mbarrier, == ((j+1) / NSTG) & 1
    for (int j = 0; j <= n_out; j++) {
        // band j was row-passed before the last barrier: its stage buffer is free for band j+NSTG
        if (!SEEDF && tid == 0 && j + C::NSTG < n_in) issue(j + C::NSTG, fstg);
        if (SEEDF == 2 && tid == 0)   // hand the buffer back to the producer warps
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[fstg])) : "memory");
        fstg = fstg == C::NSTG - 1 ? 0 : fstg + 1;
        if (SEEDF == 1 && j + 1 < n_in) { produce(j + 1, stage, tid, C::THREADS); csync(); }   // (band j was row-passed before the last barrier)
        if (j + 1 < n_in) row_pass(j + 1, rstg, rpar, rslot);
        if (rstg == C::NSTG - 1) { rstg = 0; rpar ^= 1u; } else rstg++;
        rslot = rslot == 2 ? 0 : rslot + 1;
        if (j >= 1) {
            col_pass(j - 1, cslot);
            cslot = cslot == 2 ? 0 : cslot + 1;
        }
        csync();
    }
    }   // pieces
}

// ---------------------------------------------------------------------------
// DoG + 3x3x3 extrema (build_dog + point_is_local_extremum, src/lib.rs:271-279,
// 437-506) for the three scales of one octave in one pass over its six Gaussian
// layers.  A warp owns a strip of EX_SPAN = 60 output columns: lane l holds the
// column pair 60*strip - 2 + 2l, +1 (one 8-byte load per layer and row), so the
// lanes 1..30 find both horizontal neighbours of their columns in the adjacent
// lanes (two shuffles per DoG layer, no edge cases) and lanes 0 / 31 only carry
// the halo.  The warp walks down EX_ROWS rows keeping, per DoG layer, the
// horizontal 3-max / 3-min of the two previous rows in registers.
// Output per (scale, row, strip): two ballot words {B0, B1}, bit l of Bk <=> column
// 60*strip - 2 + 2l + k -- a raster-ordered bit mask, so candidate order is
// deterministic -- plus a per-row population count.
// ---------------------------------------------------------------------------
constexpr int EX_ROWS = 32;
constexpr int EX_WARPS = 4;
constexpr int EX_SPAN = 60;  // output columns per warp
__host__ __device__ constexpr int ex_strips(int w) { return (w + EX_SPAN - 1) / EX_SPAN; }

struct ExtremaParams {
    const float* gauss;       // octave base (layer 0), image 0
    long long img_stride;     // floats
    long long layer_stride;
    int w, h, pitch;
    uint32_t* mask;           // octave mask base, image 0
    long long mask_img_stride;
    int mask_pitch;           // words per (scale, row): 2 * ex_strips(w)
    uint32_t* rows;           // octave row counters, image 0
    int rows_img_stride;
};

// sm_100 three-input min/max (FMNMX3)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// rolling three-row state of one warp: the DoG values of rows (c-1, c, c+1) for the lane's two columns
struct ExState {
    float2 d[3][N_DOG];
};

struct ExWarp {
    const float* g;            // layer 0, row 0 at this lane's (clamped) column pair
    uint32_t* mask;            // lanes 0..2: word pair of (scale lane+1, row 0) of this strip
    uint32_t* rows;            // lanes 0..2: row counter of (scale lane+1, row 0)
    int ls;                    // layer stride in floats
    int h, pitch, mask_pitch, lane;
    uint32_t cmask0, cmask1;   // warp-wide: lanes whose column 0 / 1 can hold candidates (IMAGE_BORDER, halo lanes excluded)
    bool coherent;             // the layers were written earlier in this launch (k_tail): no read-only-path loads
};

__device__ __forceinline__ const float2* ex_addr(const float* base, const int off) {
    const float2* q;   // base + 4 * off as one IMAD.WIDE
    asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(q) : "r"(off), "l"(base));
    return q;
}

__device__ __forceinline__ void ex_load(const ExWarp& W, const int r, float2* gv) {
    const int roff = min(max(r, 0), W.h - 1) * W.pitch;
#pragma unroll
    for (int l = 0; l < N_LAYERS; l++) {
        const float2* q = ex_addr(W.g, roff + l * W.ls);
        gv[l] = W.coherent ? *q : __ldg(q);
    }
}

// (v > 0 and v >= Mx) or (v < 0 and v <= mn) where Mx / mn are the max / min of a neighbourhood that contains v
// (so >= / <= can only hold with equality): one select and three compares, no branches
__device__ __forceinline__ bool ex_is_extremum(const float v, const float Mx, const float mn) {
    const float t = v > 0.0f ? Mx : mn;
    return (v == t) & (v != 0.0f);
}

// Evaluates the centre row c = r-1 once rows r-2, r-1, r sit in slots A, B, K of the state (K a compile-time
// constant thanks to the 3x unrolled callers, so the state never moves between registers): vertical
// 3-max / 3-min per DoG layer, then across the three layers of each scale, then across the three columns
// (neighbour lanes by shuffle) -- 44 three-input min/max and 12 shuffles per column pair.
template <int K, bool KEEP_FLAT>
__device__ __forceinline__ void ex_eval(const ExWarp& W, const ExState& S, const int c) {
    constexpr int A = (K + 1) % 3, B = (K + 2) % 3;  // rows c-1 and c
    float2 VM[N_DOG], Vm[N_DOG];   // column-wise max / min over the three rows
#pragma unroll
    for (int l = 0; l < N_DOG; l++) {
        VM[l] = make_float2(fmax3(S.d[A][l].x, S.d[B][l].x, S.d[K][l].x), fmax3(S.d[A][l].y, S.d[B][l].y, S.d[K][l].y));
        Vm[l] = make_float2(fmin3(S.d[A][l].x, S.d[B][l].x, S.d[K][l].x), fmin3(S.d[A][l].y, S.d[B][l].y, S.d[K][l].y));
    }
    uint32_t mine0 = 0, mine1 = 0;  // lane s-1 keeps the ballots of scale s
    if ((c >= IMAGE_BORDER) & (c < W.h - IMAGE_BORDER)) {  // warp-uniform: rows inside the border margin hold no candidates
#pragma unroll
    for (int s = 1; s <= SCALES_PER_OCTAVE; s++) {
        const float2 v = S.d[B][s];
        // max / min over rows and layers, per column; then over the three columns (v itself included)
        const float2 XM = make_float2(fmax3(VM[s - 1].x, VM[s].x, VM[s + 1].x), fmax3(VM[s - 1].y, VM[s].y, VM[s + 1].y));
        const float2 Xm = make_float2(fmin3(Vm[s - 1].x, Vm[s].x, Vm[s + 1].x), fmin3(Vm[s - 1].y, Vm[s].y, Vm[s + 1].y));
        const float lM = __shfl_up_sync(0xffffffffu, XM.y, 1), rM = __shfl_down_sync(0xffffffffu, XM.x, 1);
        const float lm = __shfl_up_sync(0xffffffffu, Xm.y, 1), rm = __shfl_down_sync(0xffffffffu, Xm.x, 1);
        const bool e0 = ex_is_extremum(v.x, fmax3(lM, XM.x, XM.y), fmin3(lm, Xm.x, Xm.y));
        const bool e1 = ex_is_extremum(v.y, fmax3(XM.x, XM.y, rM), fmin3(Xm.x, Xm.y, rm));
        uint32_t b0 = __ballot_sync(0xffffffffu, e0) & W.cmask0, b1 = __ballot_sync(0xffffffffu, e1) & W.cmask1;
        // A candidate whose three DoG layers are each spatially constant over its 3x3 window has zero
        // spatial derivatives (h12 = h13 = h22 = h33 = h23 = 0, g2 = g3 = 0), so interpolate_extremum
        // computes det = 0, every cofactor quotient is 0/0 = NaN, the NaN offsets never pass `abs() < 0.5`,
        // round(NaN) as isize = 0 keeps the point in place, and after MAX_INTERPOLATION_STEPS it returns
        // None (src/lib.rs:545-602).  Such points can never become keypoints, and saturated / constant
        // image regions produce them for every pixel, so the pipeline drops them here.  KEEP_FLAT = true
        // reproduces the reference's full candidate list for the parity view (sb200_last_candidates).
        if (!KEEP_FLAT && (b0 | b1)) {  // warp-uniform and rare: most rows of a strip hold no extremum
            bool f0 = true, f1 = true;
#pragma unroll
            for (int l = s - 1; l <= s + 1; l++) {
                // column k of the lane is vertically constant in layer l
                const bool c0 = VM[l].x == Vm[l].x, c1 = VM[l].y == Vm[l].y;
                const float lv = __shfl_up_sync(0xffffffffu, VM[l].y, 1), rv = __shfl_down_sync(0xffffffffu, VM[l].x, 1);
                const bool lc = __shfl_up_sync(0xffffffffu, (int)c1, 1) != 0, rc = __shfl_down_sync(0xffffffffu, (int)c0, 1) != 0;
                const bool mid = c0 & c1 & (VM[l].x == VM[l].y);
                f0 &= mid & lc & (lv == VM[l].x);
                f1 &= mid & rc & (rv == VM[l].y);
            }
            b0 = __ballot_sync(0xffffffffu, e0 & !f0) & W.cmask0;
            b1 = __ballot_sync(0xffffffffu, e1 & !f1) & W.cmask1;
        }
        if (W.lane == s - 1) { mine0 = b0; mine1 = b1; }
    }
    }
    if (W.lane < SCALES_PER_OCTAVE && c < W.h) {
        *reinterpret_cast<uint2*>(W.mask + (long long)c * W.mask_pitch) = make_uint2(mine0, mine1);
        if (mine0 | mine1) atomicAdd(W.rows + c, (uint32_t)(__popc(mine0) + __popc(mine1)));
    }
}

// Generic-load variant (octaves too small for a tensor map): processes image row r (already in gv) into
// slot K, issues the loads of row r+1 as soon as gv is consumed, and (EVAL) evaluates the centre row r-1.
template <int K, bool EVAL, bool KEEP_FLAT>
__device__ __forceinline__ void ex_step(const ExWarp& W, ExState& S, const int r, float2* gv) {
#pragma unroll
    for (int l = 0; l < N_DOG; l++) S.d[K][l] = sub2(gv[l + 1], gv[l]);
    ex_load(W, r + 1, gv);  // in flight while this row is evaluated (row index clamped: harmless past the end)
    if (EVAL) ex_eval<K, KEEP_FLAT>(W, S, r - 1);
}

__device__ __forceinline__ void ex_setup(ExWarp& W, const ExtremaParams& p, const int strip, const int lane, const long long img) {
    W.h = p.h; W.pitch = p.pitch; W.mask_pitch = p.mask_pitch; W.lane = lane;
    W.ls = (int)p.layer_stride;
    W.coherent = false;
    const int x0 = strip * EX_SPAN - 2 + 2 * lane;
    // column pairs outside [0, pitch-2] load a clamped pair instead: only halo lanes and columns inside the
    // IMAGE_BORDER margin can be affected, and neither they nor their neighbours can be candidates
    W.g = p.gauss + img * p.img_stride + min(max(x0, 0), p.pitch - 2);
    const int sl = min(lane, SCALES_PER_OCTAVE - 1);  // lanes 0..2 write the words of scales 1..3
    W.mask = p.mask + img * p.mask_img_stride + (long long)sl * W.h * W.mask_pitch + 2 * strip;
    W.rows = p.rows + img * p.rows_img_stride + sl * W.h;
    const bool inner = (lane >= 1) & (lane <= 30);
    W.cmask0 = __ballot_sync(0xffffffffu, inner & (x0 >= IMAGE_BORDER) & (x0 < p.w - IMAGE_BORDER));
    W.cmask1 = __ballot_sync(0xffffffffu, inner & (x0 + 1 >= IMAGE_BORDER) & (x0 + 1 < p.w - IMAGE_BORDER));
}

// one warp: strip `strip`, centre rows [y0, y0 + rows) of image `img`
template <bool KEEP_FLAT>
__device__ __forceinline__ void ex_strip(const ExtremaParams& p, const int strip, const int y0, const int rows,
                                         const int lane, const long long img, const bool coherent) {
    ExWarp W;
    ex_setup(W, p, strip, lane, img);
    W.coherent = coherent;
    ExState S;
    const int r_end = min(y0 + rows, W.h);  // centre rows [y0, r_end): image rows y0-1 .. r_end
    float2 gv[N_LAYERS];
    ex_load(W, y0 - 1, gv);
    ex_step<0, false, KEEP_FLAT>(W, S, y0 - 1, gv);
    ex_step<1, false, KEEP_FLAT>(W, S, y0, gv);
    for (int r = y0 + 1; r <= r_end; r += 3) {
        ex_step<2, true, KEEP_FLAT>(W, S, r, gv);
        if (r + 1 <= r_end) ex_step<0, true, KEEP_FLAT>(W, S, r + 1, gv);
        if (r + 2 <= r_end) ex_step<1, true, KEEP_FLAT>(W, S, r + 2, gv);
    }
}

template <bool KEEP_FLAT>
__global__ void __launch_bounds__(32 * EX_WARPS, 5) k_extrema(const ExtremaParams p) {
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = (blockIdx.y * EX_WARPS + warp) * EX_ROWS;
    if (y0 >= p.h) return;
    ex_strip<KEEP_FLAT>(p, blockIdx.x, y0, EX_ROWS, lane, blockIdx.z, false);
}

// ---------------------------------------------------------------------------
// TMA-fed variant for octaves that own a tensor map (>= TMA_MIN_DIM per side).  Every warp is its
// own pipeline: an elected lane streams (68 columns x 3 rows x 6 layers) boxes of its strip through a ring of
// EXT_NS shared-memory stages (cp.async.bulk.tensor.4d completing on one mbarrier per stage), the warp consumes
// a stage as three row steps -- the rotation period of the three-row state, so a stage is one trip of the
// unrolled loop -- and hands it back for the box EXT_NS stages ahead.  The loads no longer occupy registers or
// issue slots, and each warp keeps (EXT_NS - 1) x 4.5 KB in flight, which is what lets the kernel follow the
// HBM stream.  Out-of-image box elements arrive as zeros; they only ever feed rows / columns inside the
// IMAGE_BORDER margin, which cannot hold candidates.
// ---------------------------------------------------------------------------
constexpr int EXT_RB = 3;                         // rows per stage
constexpr int EXT_NS = 2;                         // stages per warp
constexpr int EXT_STEPS = 12;                     // stages per warp pass
constexpr int EXT_ROWS = EXT_RB * EXT_STEPS - 2;  // centre rows per warp (two halo rows)
// a TMA box must start on a 16-byte boundary of the row: the box begins 4 columns left of the strip (60j - 4)
// and is 68 columns wide, the lane's pair sits at box column 2 + 2*lane
constexpr int EXT_BOX_W = 68;
constexpr int EXT_BOX_X0 = 4;
constexpr uint32_t EXT_STAGE_BYTES = N_LAYERS * EXT_RB * EXT_BOX_W * sizeof(float);   // bytes one box delivers
constexpr int EXT_STAGE_FLOATS = (EXT_STAGE_BYTES + 127) / 128 * 128 / sizeof(float);  // stage stride (128-byte aligned)
constexpr size_t EXT_SMEM = (size_t)EX_WARPS * EXT_NS * EXT_STAGE_FLOATS * sizeof(float);

template <int K, bool EVAL, bool KEEP_FLAT>
__device__ __forceinline__ void ext_step(const ExWarp& W, ExState& S, const float2* sp, const int r) {
    float2 gv[N_LAYERS];
#pragma unroll
    for (int l = 0; l < N_LAYERS; l++) gv[l] = sp[(l * EXT_RB + K) * (EXT_BOX_W / 2)];  // sp: lane's pair in row 0, layer 0
#pragma unroll
    for (int l = 0; l < N_DOG; l++) S.d[K][l] = sub2(gv[l + 1], gv[l]);
    if (EVAL) ex_eval<K, KEEP_FLAT>(W, S, r - 1);
}

template <bool KEEP_FLAT>
__global__ void __launch_bounds__(32 * EX_WARPS, 5) k_extrema_tma(const __grid_constant__ CUtensorMap tmap, const ExtremaParams p) {
    pdl_wait();
    extern __shared__ __align__(1024) float ext_smem[];
    __shared__ __align__(8) uint64_t bar[EX_WARPS][EXT_NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = (blockIdx.y * EX_WARPS + warp) * EXT_ROWS;
    if (y0 >= p.h) return;  // warps are independent: no CTA-wide barrier below
    const int img = blockIdx.z;
    ExWarp W;
    ex_setup(W, p, blockIdx.x, lane, img);
    float* ring = ext_smem + warp * EXT_NS * EXT_STAGE_FLOATS;
    // image rows y0-1 .. min(y0 + EXT_ROWS, h), three per stage
    const int n_stages = (min(y0 + EXT_ROWS, p.h) - y0 + 2 + EXT_RB - 1) / EXT_RB;
    const int bx = blockIdx.x * EX_SPAN - EXT_BOX_X0;
    auto issue = [&](const int s, const int slot) {
        const uint32_t bar_a = smem_u32(&bar[warp][slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(EXT_STAGE_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(ring + slot * EXT_STAGE_FLOATS)), "l"(&tmap), "r"(bx), "r"(y0 - 1 + s * EXT_RB), "r"(0), "r"(img),
              "r"(bar_a)
            : "memory");
    };
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < EXT_NS; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[warp][b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int b = 0; b < EXT_NS; b++)
            if (b < n_stages) issue(b, b);
    }
    __syncwarp();
    ExState S;
    int slot = 0;
    uint32_t phase = 0;
    for (int s = 0; s < n_stages; s++) {
        {
            const uint32_t bar_a = smem_u32(&bar[warp][slot]);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                    : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
            }
        }
        const float2* sp = reinterpret_cast<const float2*>(ring + slot * EXT_STAGE_FLOATS) + (EXT_BOX_X0 - 2) / 2 + lane;
        const int r = y0 - 1 + s * EXT_RB;
        if (s == 0) {
            ext_step<0, false, KEEP_FLAT>(W, S, sp, r);
            ext_step<1, false, KEEP_FLAT>(W, S, sp, r + 1);
        } else {
            ext_step<0, true, KEEP_FLAT>(W, S, sp, r);
            ext_step<1, true, KEEP_FLAT>(W, S, sp, r + 1);
        }
        ext_step<2, true, KEEP_FLAT>(W, S, sp, r + 2);
        __syncwarp();  // every lane has consumed the stage: hand it back to the copy engine
        if (lane == 0 && s + EXT_NS < n_stages) issue(s + EXT_NS, slot);
        if (++slot == EXT_NS) { slot = 0; phase ^= 1u; }
    }
}

// ---------------------------------------------------------------------------
// Tail of the pyramid: every octave small enough for a whole layer to sit in shared memory (tail_fits) is
// processed by ONE CTA per image in ONE launch -- five blurs, decimation into the next octave and the
// DoG/extrema scan, octave after octave -- instead of six launch-latency-bound kernels per octave.  Same
// arithmetic as k_blur (row pass left-to-right FMA chain, column pass symmetric-folded chain,
// BORDER_REFLECT_101) and the same extrema code as k_extrema (generic loads, coherent path: the layers were
// written by this CTA).
// In octaves this small nearly every pixel is a border pixel (a 27-tap window on a 60 x 33 layer), and resolving
// BORDER_REFLECT_101 per tap made the kernel instruction-bound on index arithmetic (one CTA: 92 us per image, a
// quarter of a single image's pyramid).  Both shared-memory layers are therefore PADDED: `a` (the blur's source)
// carries TAIL_RMAX mirrored columns on either side of every row, `b` (the row-pass result) TAIL_RMAX mirrored rows
// above and below, filled by a short pass of their own (one border_index per halo element instead of one per tap),
// so that both filter passes are plain sliding windows.  Four block barriers per blur instead of two, each phase a
// fraction of the old ones.  All six layers of the octave also stay in shared memory (dense, even pitch), and the
// extrema scan reads them there instead of through L2: its warps walk down a chain of dependent row loads.
// ---------------------------------------------------------------------------
constexpr int TAIL_MAX_PX = 2304;     // e.g. 64 x 36: larger octaves keep one CTA busy for too long
constexpr int TAIL_RMAX = 13;         // widest blur radius (27 taps)
constexpr int TAIL_A_FLOATS = 4096;   // (w + 2 RMAX) x h
constexpr int TAIL_B_FLOATS = 4096;   // w x (h + 2 RMAX)
#ifndef SB_TAIL_THREADS
#define SB_TAIL_THREADS 512
#endif
constexpr int TAIL_THREADS = SB_TAIL_THREADS;
constexpr int TAIL_L_FLOATS = 2560;   // one dense layer at an even pitch, (w + 1 & ~1) x h: the extrema scan's copy
constexpr size_t TAIL_SMEM = (size_t)(TAIL_A_FLOATS + TAIL_B_FLOATS + N_LAYERS * TAIL_L_FLOATS) * sizeof(float);
static_assert(blur_radius(N_LAYERS - 1) <= TAIL_RMAX, "halo of the widest blur");
__host__ __device__ constexpr bool tail_fits(const long long w, const long long h) {
    return w * h <= TAIL_MAX_PX && (w + 2 * TAIL_RMAX) * h <= TAIL_A_FLOATS && w * (h + 2 * TAIL_RMAX) <= TAIL_B_FLOATS &&
           ((w + 1) / 2 * 2) * h <= TAIL_L_FLOATS;
}

struct TailParams {
    PyrLayout L;
    int o_first;                 // first octave of the tail
    float* gauss;                // Gaussian arena, image 0
    uint32_t* mask;              // mask arena, image 0
    uint32_t* rows;              // row counters, image 0
};

// halo columns [-R, 0) and [w, w + R) of every row of `a` (pitch w + 2 RMAX, column 0 at offset RMAX): mirrored
// (BORDER_REFLECT_101) or clamped to the edge pixel (imageproc flavour)
template <int FL>
__device__ __forceinline__ void tail_halo_cols(float* __restrict__ a, const int R, const int w, const int h) {
    const int wp = w + 2 * TAIL_RMAX;
    for (int idx = threadIdx.x; idx < 2 * R * h; idx += TAIL_THREADS) {
        const int y = idx / (2 * R), k = idx - y * (2 * R);
        const int col = k < R ? k - R : w + (k - R);
        float* const row = a + y * wp + TAIL_RMAX;
        row[col] = row[border_index<FL>(col, w)];
    }
}

template <int LI, int R_NEXT, int FL>
__device__ __forceinline__ void tail_blur(float* __restrict__ a /* smem, padded columns: source, then result */,
                                          float* __restrict__ b /* smem, padded rows */,
                                          float* __restrict__ lay /* smem, dense copy of the result at pitch ps */,
                                          const int ps, float* __restrict__ dst, float* __restrict__ dec, const int w, const int h,
                                          const int pitch, const int dec_w, const int dec_h, const int dec_pitch) {
    constexpr int R = blur_radius(LI, FL);
    const int n = w * h, wp = w + 2 * TAIL_RMAX;
    for (int idx = threadIdx.x; idx < n; idx += TAIL_THREADS) {
        const int y = idx / w, x = idx - y * w;
        const float* q = a + y * wp + (TAIL_RMAX - R) + x;
        float acc = q[0] * tap<FL>(LI, 0);
#pragma unroll
        for (int i = 1; i <= 2 * R; i++)
            acc = FL == FL_OPENCV ? fmaf(q[i], tap<FL>(LI, i), acc) : __fadd_rn(acc, __fmul_rn(q[i], tap<FL>(LI, i)));
        b[idx + TAIL_RMAX * w] = acc;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * R * w; idx += TAIL_THREADS) {   // mirrored rows [-R, 0) and [h, h + R)
        const int k = idx / w, x = idx - k * w;
        const int row = k < R ? k - R : h + (k - R);
        b[(row + TAIL_RMAX) * w + x] = b[(border_index<FL>(row, h) + TAIL_RMAX) * w + x];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < n; idx += TAIL_THREADS) {
        const int y = idx / w, x = idx - y * w;
        const float* c = b + idx + TAIL_RMAX * w;
        float acc;
        if (FL == FL_OPENCV) {
            acc = c[0] * tap<FL>(LI, R);
#pragma unroll
            for (int i = 1; i <= R; i++) acc = fmaf(c[i * w] + c[-i * w], tap<FL>(LI, R + i), acc);
        } else {   // top to bottom, a multiply and an add per tap
            acc = c[-R * w] * tap<FL>(LI, 0);
#pragma unroll
            for (int i = 1; i <= 2 * R; i++) acc = __fadd_rn(acc, __fmul_rn(c[(i - R) * w], tap<FL>(LI, i)));
        }
        a[y * wp + TAIL_RMAX + x] = acc;
        lay[y * ps + x] = acc;
        dst[(long long)y * pitch + x] = acc;
        if (dec && !(y & 1) && !(x & 1)) {
            const int dy = y >> 1, dx = x >> 1;
            if (dy < dec_h && dx < dec_w) dec[(long long)dy * dec_pitch + dx] = acc;
        }
    }
    __syncthreads();
    if (R_NEXT > 0) {   // the next blur's halo columns
        tail_halo_cols<FL>(a, R_NEXT, w, h);
        __syncthreads();
    }
}

template <bool KEEP_FLAT, int FL = FL_OPENCV>
__global__ void __launch_bounds__(TAIL_THREADS) k_tail(const TailParams p) {
    pdl_wait();
    extern __shared__ __align__(16) float tail_smem[];
    float* const a = tail_smem;                    // (w + 2 RMAX) x h
    float* const b = tail_smem + TAIL_A_FLOATS;    // w x (h + 2 RMAX)
    float* const lay = b + TAIL_B_FLOATS;          // N_LAYERS x TAIL_L_FLOATS: the octave's layers, dense
    const long long img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* const gimg = p.gauss + img * p.L.img_floats;
    for (int o = p.o_first; o < p.L.n_oct; o++) {
        const OctLayout& ol = p.L.o[o];
        const int w = ol.w, h = ol.h;
        if (w < 1 || h < 1) break;
        float* const g0 = gimg + ol.off;
        // layer 0: written by the previous octave's decimation (an earlier launch for o_first, this CTA otherwise)
        const int ps = (w + 1) & ~1;   // even pitch of the dense copies: the scan loads aligned column pairs
        for (int idx = tid; idx < w * h; idx += TAIL_THREADS) {
            const int y = idx / w, x = idx - y * w;
            const float v = g0[(long long)y * ol.pitch + x];
            a[y * (w + 2 * TAIL_RMAX) + TAIL_RMAX + x] = v;
            lay[y * ps + x] = v;
        }
        if (ps != w) {   // the padding column (never a candidate, never a candidate's neighbour): defined values all the same
            for (int idx = tid; idx < N_LAYERS * h; idx += TAIL_THREADS) {
                const int l = idx / h, y = idx - l * h;
                lay[l * TAIL_L_FLOATS + y * ps + w] = 0.0f;
            }
        }
        __syncthreads();
        tail_halo_cols<FL>(a, blur_radius(1, FL), w, h);
        __syncthreads();
        float* dec = nullptr;
        int dw = 0, dh = 0, dp = 0;
        if (o + 1 < p.L.n_oct && p.L.o[o + 1].w >= 1 && p.L.o[o + 1].h >= 1) {
            dec = gimg + p.L.o[o + 1].off; dw = p.L.o[o + 1].w; dh = p.L.o[o + 1].h; dp = p.L.o[o + 1].pitch;
        }
        tail_blur<1, blur_radius(2, FL), FL>(a, b, lay + 1 * TAIL_L_FLOATS, ps, g0 + 1 * ol.layer_stride, nullptr, w, h, ol.pitch, 0, 0, 0);
        tail_blur<2, blur_radius(3, FL), FL>(a, b, lay + 2 * TAIL_L_FLOATS, ps, g0 + 2 * ol.layer_stride, nullptr, w, h, ol.pitch, 0, 0, 0);
        tail_blur<3, blur_radius(4, FL), FL>(a, b, lay + 3 * TAIL_L_FLOATS, ps, g0 + 3 * ol.layer_stride, FL == FL_OPENCV ? dec : nullptr,
                                             w, h, ol.pitch, dw, dh, dp);
        if (FL != FL_OPENCV && dec) {
            // the imageproc flavour's Nearest resize of layer 3 (k_decimate_b): source pixel floor((d + 0.5) * (n_src /
            // n_dst)) in f32, result clamped to [0, 1]; `a` holds layer 3 until the column pass of the next blur, two
            // barriers from here
            const float ry = (float)h / (float)dh, rx = (float)w / (float)dw;
            for (int idx = tid; idx < dw * dh; idx += TAIL_THREADS) {
                const int dy = idx / dw, dx = idx - dy * dw;
                const int sy = min(max((int)floorf(((float)dy + 0.5f) * ry), 0), h - 1);
                const int sx = min(max((int)floorf(((float)dx + 0.5f) * rx), 0), w - 1);
                dec[(long long)dy * dp + dx] = fminf(fmaxf(a[sy * (w + 2 * TAIL_RMAX) + TAIL_RMAX + sx], 0.0f), 1.0f);
            }
        }
        tail_blur<4, blur_radius(5, FL), FL>(a, b, lay + 4 * TAIL_L_FLOATS, ps, g0 + 4 * ol.layer_stride, nullptr, w, h, ol.pitch, 0, 0, 0);
        tail_blur<5, 0, FL>(a, b, lay + 5 * TAIL_L_FLOATS, ps, g0 + 5 * ol.layer_stride, nullptr, w, h, ol.pitch, 0, 0, 0);
        if (ol.scanned) {
            ExtremaParams e;
            // the layers where this CTA keeps them in shared memory (generic loads; image stride 0: every CTA its own)
            e.gauss = lay; e.img_stride = 0; e.layer_stride = TAIL_L_FLOATS;
            e.w = w; e.h = h; e.pitch = ps;
            e.mask = p.mask + ol.mask_off; e.mask_img_stride = p.L.img_mask_words; e.mask_pitch = ol.mask_pitch;
            e.rows = p.rows + ol.row_base; e.rows_img_stride = p.L.img_rows;
            // short row blocks: the scan is a chain of dependent row loads, so the CTA's 16 warps want many short tasks
            constexpr int TROWS = 6;
            const int strips = ex_strips(w), blocks = (h + TROWS - 1) / TROWS;
            for (int t = warp; t < strips * blocks; t += TAIL_THREADS / 32) {
                const int strip = t % strips, y0 = (t / strips) * TROWS;
                ex_strip<KEEP_FLAT>(e, strip, y0, TROWS, lane, img, true);
            }
        }
        __syncthreads();
    }
}

// Exclusive scan of the per-row candidate counts of one image (one CTA per image).
__global__ void __launch_bounds__(1024) k_rowscan(const uint32_t* __restrict__ rows, uint32_t* __restrict__ rowoff,
                                                   int n_rows, uint32_t* __restrict__ cand_count) {
    pdl_wait();
    __shared__ uint32_t wsum[32];
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    rows += (long long)img * n_rows;
    rowoff += (long long)img * n_rows;
    const int per = (n_rows + 1023) / 1024;
    const int start = tid * per, end = min(start + per, n_rows);
    uint32_t local = 0;
    for (int i = start; i < end; i++) local += rows[i];
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = wsum[lane], s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        wsum[lane] = s - v;  // exclusive warp offsets
        if (lane == 31) cand_count[img] = s;
    }
    __syncthreads();
    uint32_t run = wsum[warp] + incl - local;
    for (int i = start; i < end; i++) {
        uint32_t c = rows[i];
        rowoff[i] = run;
        run += c;
    }
}

// Ordered compaction: warp per (scale, row) entry; writes packed candidate keys in
// raster order at the row's scanned offset => natural order of src/lib.rs:287-293,324-332.
// A lane takes one strip's word pair {B0, B1} and emits its columns in ascending order
// (bit l of B0 is column 60*strip - 2 + 2l, bit l of B1 the column after it).
__global__ void __launch_bounds__(256) k_compact(const PyrLayout L, const uint32_t* __restrict__ mask,
                                                  const uint32_t* __restrict__ rows,
                                                  const uint32_t* __restrict__ rowoff, CandKey* __restrict__ keys,
                                                  uint32_t cap) {
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ridx = blockIdx.x * 8 + warp;
    const long long img = blockIdx.y;
    if (ridx >= L.img_rows) return;
    int o = 0;
    while (o + 1 < L.n_oct && ridx >= L.o[o + 1].row_base) o++;
    const OctLayout& ol = L.o[o];
    const int local = ridx - ol.row_base;
    const int s = local / ol.h + 1, y = local - (s - 1) * ol.h;
    const uint2* words = reinterpret_cast<const uint2*>(mask + img * L.img_mask_words + ol.mask_off + (long long)local * ol.mask_pitch);
    const int ns = ol.mask_pitch >> 1;
    // the row's count, its scanned offset and its first 32 mask word pairs are independent loads: all three are in
    // flight before the count decides whether the row has anything to emit (the kernel is a chain of round trips)
    const uint32_t cnt = rows[img * L.img_rows + ridx];
    uint32_t pos0 = rowoff[img * L.img_rows + ridx];
    uint2 first = (lane < ns) ? words[lane] : make_uint2(0u, 0u);
    if (cnt == 0) return;
    CandKey* out = keys + img * (long long)cap;
    for (int sb = 0; sb < ns; sb += 32) {
        uint2 word = sb == 0 ? first : ((sb + lane < ns) ? words[sb + lane] : make_uint2(0u, 0u));
        uint32_t c = __popc(word.x) + __popc(word.y), incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        uint32_t pos = pos0 + incl - c;
        const int xbase = (sb + lane) * EX_SPAN - 2;
        while (word.x | word.y) {
            const int l0 = word.x ? __ffs(word.x) - 1 : 32, l1 = word.y ? __ffs(word.y) - 1 : 32;
            int x;
            if (l0 <= l1) { x = xbase + 2 * l0; word.x &= word.x - 1; }
            else { x = xbase + 2 * l1 + 1; word.y &= word.y - 1; }
            if (pos < cap) out[pos] = pack_key(o, s, y, x);
            pos++;
        }
        pos0 += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// dense copy of a DoG layer for the PrecomputedImages.dog accessor (src/lib.rs:126)
__global__ void k_dog_layer(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int w,
                            int h, int pitch) {
    pdl_wait();
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < w && y < h) out[(long long)y * w + x] = b[(long long)y * pitch + x] - a[(long long)y * pitch + x];
}

}  // namespace sb
