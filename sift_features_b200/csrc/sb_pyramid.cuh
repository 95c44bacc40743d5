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
#pragma once
#include "sb_common.cuh"

namespace sb {

// taps of the six Gaussian kernels, [kernel][0..2R]; filled by the host at context creation
__constant__ float c_taps[N_LAYERS][32];

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
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

template <int L>
struct BlurCfg {
    static constexpr int R = blur_radius(L);
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
    static_assert(IN_H * IN_W <= SH * IPITCH, "input tile does not fit the aliased buffer");
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
template <int L, bool SEED, bool DECIMATE>
__global__ void __launch_bounds__(256, 2) k_blur(const BlurParams p) {
    using C = BlurCfg<L>;
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
        for (int idx = tid; idx < C::SH * C::SW; idx += C::THREADS) {
            int row = idx / C::SW, col = idx - row * C::SW;
            float v = 0.0f;
            if (row < rows_needed && col < cols_needed) {
                int y = reflect101(ty0 - R + row, h);
                int x = reflect101(tx0 - R + col, w);
                int ay, by, ax, bx;
                float fy, fx;
                lerp2x(y, p.in_h, ay, by, fy);
                lerp2x(x, p.in_w, ax, bx, fx);
                const float* r0 = in_tile + (ay - iy0) * C::IN_W - ix0;
                const float* r1 = in_tile + (by - iy0) * C::IN_W - ix0;
                float p00 = r0[ax], p01 = r0[bx], p10 = r1[ax], p11 = r1[bx];
                float h0 = fmaf(p01 - p00, fx, p00);
                float h1 = fmaf(p11 - p10, fx, p10);
                v = fmaf(h1 - h0, fy, h0);
            }
            stage[row * C::SPITCH + col] = v;
        }
    } else {
        const float* src = p.src + img * p.img_stride;
        for (int idx = tid; idx < C::SH * C::SW; idx += C::THREADS) {
            int row = idx / C::SW, col = idx - row * C::SW;
            float v = 0.0f;
            if (row < rows_needed && col < cols_needed) {
                int gy = reflect101(ty0 - R + row, h);
                int gx = reflect101(tx0 - R + col, w);
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
                    float acc = win[j] * c_taps[L][0];
#pragma unroll
                    for (int i = 1; i <= 2 * R; i++) acc = fmaf(win[j + i], c_taps[L][i], acc);
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
                float acc = c[j + R] * c_taps[L][R];
#pragma unroll
                for (int i = 1; i <= R; i++) acc = fmaf(c[j + R + i] + c[j + R - i], c_taps[L][R + i], acc);
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
// DoG + 3x3x3 extrema (build_dog + point_is_local_extremum, src/lib.rs:271-279,
// 437-506) for the three scales of one octave in one pass over its six Gaussian
// layers.  A warp owns 32 consecutive columns (one mask word) and walks down
// EX_ROWS rows keeping, per DoG layer, the horizontal 3-max / 3-min of the two
// previous rows in registers.  Output: one ballot word per (scale, row, 32
// columns) -- a raster-ordered bit mask, so candidate order is deterministic --
// plus a per-row population count.
// ---------------------------------------------------------------------------
constexpr int EX_ROWS = 16;
constexpr int EX_WARPS = 4;

struct ExtremaParams {
    const float* gauss;       // octave base (layer 0), image 0
    long long img_stride;     // floats
    long long layer_stride;
    int w, h, pitch;
    uint32_t* mask;           // octave mask base, image 0
    long long mask_img_stride;
    int mask_pitch;
    uint32_t* rows;           // octave row counters, image 0
    int rows_img_stride;
};

__global__ void __launch_bounds__(32 * EX_WARPS) k_extrema(const ExtremaParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x;
    const int y0 = (blockIdx.y * EX_WARPS + warp) * EX_ROWS;
    const long long img = blockIdx.z;
    const int w = p.w, h = p.h;
    if (y0 >= h) return;
    const float* g = p.gauss + img * p.img_stride;
    uint32_t* mask = p.mask + img * p.mask_img_stride;
    uint32_t* rows = p.rows + img * p.rows_img_stride;

    const int x = strip * 32 + lane;
    const int xc = min(x, w - 1);
    // lane 0 fetches the column left of the strip, lane 31 the column right of it
    const int xe = (lane == 0) ? max(strip * 32 - 1, 0) : min(strip * 32 + 32, w - 1);
    const bool edge_lane = (lane == 0) || (lane == 31);
    const bool x_ok = (x >= IMAGE_BORDER) && (x < w - IMAGE_BORDER);

    float hmaxA[N_DOG], hmaxB[N_DOG], hminA[N_DOG], hminB[N_DOG], vB[N_DOG];
#pragma unroll
    for (int l = 0; l < N_DOG; l++) { hmaxA[l] = hmaxB[l] = hminA[l] = hminB[l] = vB[l] = 0.0f; }

    const int r_end = min(y0 + EX_ROWS, h);  // centre rows [y0, r_end)
    for (int r = y0 - 1; r <= r_end; r++) {
        const int rc = min(max(r, 0), h - 1);
        const float* row = g + (long long)rc * p.pitch;
        float gv[N_LAYERS], ge[N_LAYERS];
#pragma unroll
        for (int l = 0; l < N_LAYERS; l++) {
            gv[l] = __ldg(row + l * p.layer_stride + xc);
            ge[l] = edge_lane ? __ldg(row + l * p.layer_stride + xe) : 0.0f;
        }
        float d[N_DOG], hmx[N_DOG], hmn[N_DOG];
#pragma unroll
        for (int l = 0; l < N_DOG; l++) {
            d[l] = gv[l + 1] - gv[l];
            float e = ge[l + 1] - ge[l];
            float left = __shfl_up_sync(0xffffffffu, d[l], 1);
            float right = __shfl_down_sync(0xffffffffu, d[l], 1);
            if (lane == 0) left = e;
            if (lane == 31) right = e;
            hmx[l] = fmaxf(fmaxf(left, d[l]), right);
            hmn[l] = fminf(fminf(left, d[l]), right);
        }
        if (r >= y0 + 1) {
            const int c = r - 1;  // centre row: A = c-1, B = c, new = c+1
            float M[N_DOG], m[N_DOG];
#pragma unroll
            for (int l = 0; l < N_DOG; l++) {
                M[l] = fmaxf(fmaxf(hmaxA[l], hmaxB[l]), hmx[l]);
                m[l] = fminf(fminf(hminA[l], hminB[l]), hmn[l]);
            }
            const bool ok = x_ok && (c >= IMAGE_BORDER) && (c < h - IMAGE_BORDER);
#pragma unroll
            for (int s = 1; s <= SCALES_PER_OCTAVE; s++) {
                const float v = vB[s];
                const float Mx = fmaxf(fmaxf(M[s - 1], M[s]), M[s + 1]);
                const float mn = fminf(fminf(m[s - 1], m[s]), m[s + 1]);
                const bool ext = ok && ((v > 0.0f && v >= Mx) || (v < 0.0f && v <= mn));
                const uint32_t bits = __ballot_sync(0xffffffffu, ext);
                if (lane == 0) {
                    const long long ri = (long long)(s - 1) * h + c;
                    mask[ri * p.mask_pitch + strip] = bits;
                    if (bits) atomicAdd(rows + ri, (uint32_t)__popc(bits));
                }
            }
        }
#pragma unroll
        for (int l = 0; l < N_DOG; l++) {
            hmaxA[l] = hmaxB[l]; hmaxB[l] = hmx[l];
            hminA[l] = hminB[l]; hminB[l] = hmn[l];
            vB[l] = d[l];
        }
    }
}

// Exclusive scan of the per-row candidate counts of one image (one CTA per image).
__global__ void __launch_bounds__(1024) k_rowscan(const uint32_t* __restrict__ rows, uint32_t* __restrict__ rowoff,
                                                   int n_rows, uint32_t* __restrict__ cand_count) {
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
__global__ void __launch_bounds__(256) k_compact(const PyrLayout L, const uint32_t* __restrict__ mask,
                                                  const uint32_t* __restrict__ rows,
                                                  const uint32_t* __restrict__ rowoff, uint32_t* __restrict__ keys,
                                                  uint32_t cap) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ridx = blockIdx.x * 8 + warp;
    const long long img = blockIdx.y;
    if (ridx >= L.img_rows) return;
    const uint32_t cnt = rows[img * L.img_rows + ridx];
    if (cnt == 0) return;
    int o = 0;
    while (o + 1 < L.n_oct && ridx >= L.o[o + 1].row_base) o++;
    const OctLayout& ol = L.o[o];
    const int local = ridx - ol.row_base;
    const int s = local / ol.h + 1, y = local - (s - 1) * ol.h;
    const uint32_t* words = mask + img * L.img_mask_words + ol.mask_off + (long long)local * ol.mask_pitch;
    uint32_t pos0 = rowoff[img * L.img_rows + ridx];
    uint32_t* out = keys + img * (long long)cap;
    const int nw = (ol.w + 31) >> 5;
    for (int wb = 0; wb < nw; wb += 32) {
        uint32_t word = (wb + lane < nw) ? words[wb + lane] : 0u;
        uint32_t c = __popc(word), incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        uint32_t pos = pos0 + incl - c;
        while (word) {
            int b = __ffs(word) - 1;
            word &= word - 1;
            if (pos < cap) out[pos] = pack_key(o, s, y, (wb + lane) * 32 + b);
            pos++;
        }
        pos0 += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// dense copy of a DoG layer for the PrecomputedImages.dog accessor (src/lib.rs:126)
__global__ void k_dog_layer(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int w,
                            int h, int pitch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < w && y < h) out[(long long)y * w + x] = b[(long long)y * pitch + x] - a[(long long)y * pitch + x];
}

}  // namespace sb
