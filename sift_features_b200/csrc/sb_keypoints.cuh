// sb_keypoints.cuh -- refinement, orientation assignment, ordering, descriptors.
//
// All arithmetic follows src/lib.rs operation by operation in f32 (the file is
// compiled with --fmad=false, IEEE division and square root); the transcendental
// calls the crate makes through libm (f32::exp, f32::powf, f64::atan2, f32::sin_cos)
// are reproduced by sb_math.cuh or by double-precision evaluation rounded once.
#pragma once
#include "sb_common.cuh"
#include "sb_math.cuh"

namespace sb {

struct KpParams {
    PyrLayout L;
    const float* gauss;        // Gaussian arena, image 0
    const CandKey* keys;       // [img][cap] packed candidate keys (natural order)
    const uint32_t* cand_count;// [img]
    uint32_t cap;              // candidate capacity per image
    Refined* refined;          // [img][cap]
    uint32_t* n_ori;           // [img][cap] orientations per candidate
    float* angles;             // [img][cap][MAX_ORI]
    uint32_t* kp_off;          // [img][cap] exclusive scan of n_ori
    uint32_t* kp_count;        // [img]
    DevKeyPoint* kps;          // [img][kcap] SiftKeyPoints in natural order
    uint32_t kcap;
};

__device__ __forceinline__ float pow2i(int e) { return __int_as_float((127 + e) << 23); }

// atan2(y, x) in (-pi, pi]: minimax polynomial of atan on [0,1] (|error| < 3.3e-7 rad) plus octant folding.
// Used where the result only selects a bin (with an exact f64 re-evaluation near bin boundaries) or feeds a
// continuous interpolation weight.
__device__ __forceinline__ float fast_atan2_rad(const float y, const float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = (mx > 0.f) ? __fdividef(mn, mx) : 0.f;
    const float s = a * a;
    float p = 0x1.be6ae0p-8f;
    p = fmaf(p, s, -0x1.134924p-5f);
    p = fmaf(p, s, 0x1.462378p-4f);
    p = fmaf(p, s, -0x1.0f04d4p-3f);
    p = fmaf(p, s, 0x1.95aa00p-3f);
    p = fmaf(p, s, -0x1.552b7cp-2f);
    p = fmaf(p, s, 0x1.ffff7ep-1f);
    float r = p * a;                                   // [0, pi/4]
    if (ay > ax) r = 1.57079632679489661923f - r;      // [0, pi/2]
    if (x < 0.f) r = 3.14159265358979323846f - r;      // [0, pi]
    return (y < 0.f) ? -r : r;
}

// DoG value (src/lib.rs:275): layer l of the octave whose Gaussian layer 0 is `g`
struct DogView {
    const float* g;
    long long ls;
    int pitch;
    __device__ __forceinline__ float operator()(int l, int y, int x) const {
        const float* p = g + (long long)l * ls + (long long)y * pitch + x;
        SB_CHECK_LOAD(p, g, N_LAYERS * ls, "k_refine DoG (lower layer)");
        SB_CHECK_LOAD(p + ls, g, N_LAYERS * ls, "k_refine DoG (upper layer)");
        return __ldg(p + ls) - __ldg(p);
    }
};

// interpolate_extremum + extremum_contrast + extremum_is_on_edge + keypoint geometry
// (src/lib.rs:334-380, 525-653).  One thread per candidate.
__global__ void __launch_bounds__(128) k_refine(const KpParams P) {
    pdl_wait();
    __shared__ uint64_t s_tab[32];
    if (threadIdx.x < 32) s_tab[threadIdx.x] = sbm::d_exp2_tab[threadIdx.x];
    __syncthreads();
    const long long img = blockIdx.y;
    const uint32_t n = min(P.cand_count[img], P.cap);
    const CandKey* keys = P.keys + img * (long long)P.cap;
    Refined* out = P.refined + img * (long long)P.cap;
    const float* gimg = P.gauss + img * P.L.img_floats;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int o, scale, y, x;
        unpack_key(keys[i], o, scale, y, x);
        const OctLayout& ol = P.L.o[o];
        DogView D{gimg + ol.off, ol.layer_stride, ol.pitch};
        const int w = ol.w, h = ol.h;
        bool ok = false;
        float off_s = 0.f, off_x = 0.f, off_y = 0.f;
        for (int it = 0; it < 5; it++) {  // MAX_INTERPOLATION_STEPS, src/lib.rs:516
            const int p = scale - 1, c = scale, nn = scale + 1;
            const float vn = D(nn, y, x), vp = D(p, y, x), vc = D(c, y, x);
            const float cyp = D(c, y + 1, x), cym = D(c, y - 1, x);
            const float cxp = D(c, y, x + 1), cxm = D(c, y, x - 1);
            const float g1 = (vn - vp) / 2.f;
            const float g2 = (cyp - cym) / 2.f;
            const float g3 = (cxp - cxm) / 2.f;
            const float value2x = vc * 2.f;
            const float h11 = vn + vp - value2x;
            const float h12 = (D(nn, y + 1, x) - D(nn, y - 1, x) - D(p, y + 1, x) + D(p, y - 1, x)) / 4.f;
            const float h13 = (D(nn, y, x + 1) - D(nn, y, x - 1) - D(p, y, x + 1) + D(p, y, x - 1)) / 4.f;
            const float h22 = cyp + cym - value2x;
            const float h33 = cxp + cxm - value2x;
            const float h23 = (D(c, y + 1, x + 1) - D(c, y + 1, x - 1) - D(c, y - 1, x + 1) + D(c, y - 1, x - 1)) / 4.f;
            const float det = h11 * h22 * h33 - h11 * h23 * h23 - h12 * h12 * h33 + 2.f * h12 * h13 * h23 -
                              h13 * h13 * h22;
            const float hinv11 = (h22 * h33 - h23 * h23) / det;
            const float hinv12 = (h13 * h23 - h12 * h33) / det;
            const float hinv13 = (h12 * h23 - h13 * h22) / det;
            const float hinv22 = (h11 * h33 - h13 * h13) / det;
            const float hinv23 = (h12 * h13 - h11 * h23) / det;
            const float hinv33 = (h11 * h22 - h12 * h12) / det;
            off_s = -(hinv11 * g1 + hinv12 * g2 + hinv13 * g3);
            off_x = -(hinv13 * g1 + hinv23 * g2 + hinv33 * g3);
            off_y = -(hinv12 * g1 + hinv22 * g2 + hinv23 * g3);
            if (fabsf(off_s) < 0.5f && fabsf(off_x) < 0.5f && fabsf(off_y) < 0.5f) { ok = true; break; }
            // src/lib.rs:588-599: move by round() (half away from zero; NaN -> 0;
            // an unrepresentable step can only land outside the image => reject)
            float rx = roundf(off_x), ry = roundf(off_y), rs = roundf(off_s);
            if (isnan(rx)) rx = 0.f;
            if (isnan(ry)) ry = 0.f;
            if (isnan(rs)) rs = 0.f;
            if (fabsf(rx) > 1e5f || fabsf(ry) > 1e5f || fabsf(rs) > 1e5f) break;
            const int nx = x + (int)rx, ny = y + (int)ry, ns = scale + (int)rs;
            if (ns < 1 || ns > SCALES_PER_OCTAVE || nx < IMAGE_BORDER || nx >= w - IMAGE_BORDER ||
                ny < IMAGE_BORDER || ny >= h - IMAGE_BORDER)
                break;
            x = nx; y = ny; scale = ns;
        }
        Refined r;
        r.octave_scale = -1;
        r.x = r.y = r.size = r.response = r.kp_scale = 0.f;
        r.px = x; r.py = y;
        if (ok) {
            // extremum_contrast, src/lib.rs:606-626 (gradient recomputed at the final point)
            const float g1 = (D(scale + 1, y, x) - D(scale - 1, y, x)) / 2.f;
            const float cyp = D(scale, y + 1, x), cym = D(scale, y - 1, x);
            const float cxp = D(scale, y, x + 1), cxm = D(scale, y, x - 1);
            const float vc = D(scale, y, x);
            const float g2 = (cyp - cym) / 2.f;
            const float g3 = (cxp - cxm) / 2.f;
            const float interp = off_s * g1 + off_y * g2 + off_x * g3;
            const float contrast = fabsf(vc + interp / 2.f);
            bool keep = !(contrast * 3.f <= 0.04f);  // src/lib.rs:360
            if (keep) {
                // extremum_is_on_edge, src/lib.rs:630-653
                const float val2x = vc * 2.0f;
                const float h11 = cyp + cym - val2x;
                const float d22 = cxp + cxm - val2x;
                const float h12 = (D(scale, y + 1, x + 1) - D(scale, y + 1, x - 1) - D(scale, y - 1, x + 1) +
                                   D(scale, y - 1, x - 1)) / 4.f;
                const float tr = d22 + h11;
                const float det = d22 * h11 - h12 * h12;
                if (det <= 0.f) keep = false;
                else if ((tr * tr * 10.0f) > (11.0f * 11.0f) * det) keep = false;
            }
            if (keep) {
                const float osf = pow2i(o);
                const float kp_scale = 0.8f * sbm::pow2f_glibc(s_tab, ((float)scale + off_s) / 3.f) * 2.f;
                r.x = ((float)x + off_x) * osf;
                r.y = ((float)y + off_y) * osf;
                r.size = kp_scale * osf;
                r.response = contrast;
                r.kp_scale = kp_scale;
                r.octave_scale = (o << 8) | scale;
            }
        }
        out[i] = r;
    }
}

// gradient_direction_histogram + peak extraction (src/lib.rs:380-431, 657-757).
// One warp per candidate.  The histogram is accumulated in the reference's raster
// order per bin (the owner lane of a bin adds its samples in ascending sample
// order), so the smoothed histogram, the peak set and the angles are the
// reference's f32 results, not an approximation of them.
constexpr int ORI_WARPS = 8;

// largest img with off[img] <= j (off is an exclusive prefix sum with off[n] = total)
__device__ __forceinline__ int find_image(const uint32_t* __restrict__ off, const int n, const uint32_t j) {
    int lo = 0, hi = n;   // invariant: off[lo] <= j < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= j) lo = mid; else hi = mid;
    }
    return lo;
}

// Launch-wide work list pulled in chunks from an atomic counter: lane 0 issues the atomic for the NEXT chunk before
// the warp processes the current one and the result is only broadcast afterwards, so its latency is hidden.
// With little work per warp (single images) a chunk is one item and nothing is requested ahead, so no warp sits on
// work that an idle one could do.
struct WorkQueue {
    uint32_t* counter;
    uint32_t pending;   // lane 0: first item of the chunk requested last
    uint32_t chunk;     // items per request
    bool ahead;         // request the next chunk before processing the current one
    __device__ __forceinline__ WorkQueue(uint32_t* c, const uint32_t total, const uint32_t warps, const uint32_t max_chunk)
        : counter(c), pending(0u) {
        ahead = total > 8u * warps * max_chunk;
        chunk = ahead ? max_chunk : 1u;
    }
    __device__ __forceinline__ void request(const int lane) {
        if (lane == 0) pending = atomicAdd(counter, chunk);
    }
    __device__ __forceinline__ uint32_t take() const { return __shfl_sync(0xffffffffu, pending, 0); }
    // first item of this warp's next chunk; call once per chunk
    __device__ __forceinline__ uint32_t next(const int lane, const bool first_call) {
        if (first_call || !ahead) request(lane);
        const uint32_t v = take();
        if (ahead) request(lane);   // in flight while the chunk is processed
        return v;
    }
};

#ifndef SB_ORI_CHUNK
#define SB_ORI_CHUNK 8
#endif
constexpr int ORI_CHUNK = SB_ORI_CHUNK;

// The candidates of all images of the group form one work list (cand_off = exclusive prefix sum of the per-image
// counts); warps pull chunks of it from an atomic counter, so the launch is balanced whatever the per-image and
// per-candidate cost, and its grid is exactly the number of resident CTAs.
__global__ void __launch_bounds__(32 * ORI_WARPS) k_orient(const KpParams P, const uint32_t* __restrict__ cand_off,
                                                            const int n_img, uint32_t* __restrict__ work) {
    pdl_wait();
    __shared__ uint64_t s_tab[32];
    __shared__ float s_val[ORI_WARPS][32];
    __shared__ float s_raw[ORI_WARPS][40];
    __shared__ float s_hist[ORI_WARPS][ORI_BINS];
    if (threadIdx.x < 32) s_tab[threadIdx.x] = sbm::d_exp2_tab[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float PI32 = 3.14159265358979323846f;
    const float bin_angle_step = 36.0f / (PI32 * 2.f);  // src/lib.rs:718
    const uint32_t total = __ldg(cand_off + n_img);

    WorkQueue wq(work, total, gridDim.x * ORI_WARPS, ORI_CHUNK);
    for (uint32_t first = wq.next(lane, true); first < total; first = wq.next(lane, false)) {
      for (uint32_t j = first; j < min(first + wq.chunk, total); j++) {
        const int img = find_image(cand_off, n_img, j);
        const uint32_t i = j - __ldg(cand_off + img);
        const Refined* refined = P.refined + img * (long long)P.cap;
        uint32_t* n_ori = P.n_ori + img * (long long)P.cap;
        float* angles = P.angles + img * (long long)P.cap * MAX_ORI;
        const float* gimg = P.gauss + img * P.L.img_floats;
        const Refined r = refined[i];
        if (r.octave_scale < 0) {
            if (lane == 0) n_ori[i] = 0;
            continue;
        }
        const int o = r.octave_scale >> 8, layer = r.octave_scale & 255;
        const OctLayout& ol = P.L.o[o];
        const float* I = gimg + ol.off + (long long)layer * ol.layer_stride;
        asm volatile("" : "+l"(I));   // a plain 64-bit pointer from here on (not base + layer offset re-derived per load)
        const int w = ol.w, h = ol.h, pitch = ol.pitch;
        const int x = r.px, y = r.py;
        const int radius = (int)roundf(3.f * 1.5f * r.kp_scale);  // src/lib.rs:380
        const float sigma = 1.5f * r.kp_scale;                    // LAMBDA_ORI * kp_scale, :386
        const float gws = -1.0f / (2.0f * sigma * sigma);         // :667
        const int side = 2 * radius + 1, total = side * side;
        // the raw histogram accumulates in shared memory (s_raw[2 .. 2 + ORI_BINS), the layout the smoothing reads)
        s_raw[warp][lane + 2] = 0.f;
        if (lane < ORI_BINS - 32) s_raw[warp][32 + lane + 2] = 0.f;
        // sample `idx` of the (2r+1)^2 window in raster order: pixel loads are issued one batch of 32 ahead.
        // The lane's (row, column) advance by 32 samples per batch: at most two row wraps (side >= 17), no division.
        int yq_n = lane / side, xq_n = lane - yq_n * side;   // window coordinates of the lane's next sample
        auto fetch = [&](const int idx, int& yp, int& xp, float4& q) -> bool {
            yp = yq_n - radius; xp = xq_n - radius;
            const int yi = y + yp, xi = x + xp;
            const bool in = idx < total && yi > 0 && yi < h - 1 && xi > 0 && xi < w - 1;
            if (in) {
                const float* c = I + (yi * pitch + xi);
                asm volatile("" : "+l"(c));   // one address; the neighbours are pointer +- pitch
                const ptrdiff_t dp = pitch;
                SB_CHECK_LOAD(c - dp, I, (long long)h * pitch, "k_orient (row above)");
                SB_CHECK_LOAD(c + dp, I, (long long)h * pitch, "k_orient (row below)");
                q = make_float4(__ldg(c + 1), __ldg(c - 1), __ldg(c - dp), __ldg(c + dp));
            }
            xq_n += 32;
            if (xq_n >= side) { xq_n -= side; yq_n++; }
            if (xq_n >= side) { xq_n -= side; yq_n++; }
            return in;
        };
        int yp_n, xp_n;
        float4 q_n = make_float4(0.f, 0.f, 0.f, 0.f);
        bool in_n = fetch(lane, yp_n, xp_n, q_n);
        for (int base = 0; base < total; base += 32) {
            const int yp = yp_n, xp = xp_n;
            const float4 q = q_n;
            const bool in = in_n;
            if (base + 32 < total) in_n = fetch(base + 32 + lane, yp_n, xp_n, q_n);
            int bin = -1;
            float val = 0.f;
            if (in) {
                {
                    const float dx = q.x - q.y;
                    const float dy = q.z - q.w;
                    const float wexp = (float)(yp * yp + xp * xp) * gws;
                    const float weight = sbm::expf_glibc(s_tab, wexp);
                    const float mag = sqrtf(dx * dx + dy * dy);
                    // bin = round(36/(2pi) * (atan2_f64(dy,dx) as f32)), src/lib.rs:715-726.
                    // f32 atan2 decides unless the scaled angle is close to a rounding
                    // boundary; then the f64 evaluation the reference uses decides.
                    float raw = bin_angle_step * fast_atan2_rad(dy, dx);   // |error| < 4e-7 rad => < 3e-6 bins
                    if (fabsf(fabsf(raw - truncf(raw)) - 0.5f) < 1e-4f)   // > 25x the approximation error
                        raw = bin_angle_step * (float)atan2((double)dy, (double)dx);
                    bin = (int)roundf(raw);
                    if (bin >= ORI_BINS) bin -= ORI_BINS;
                    else if (bin < 0) bin += ORI_BINS;
                    val = weight * mag;
                }
            }
            // ordered accumulation: publish values, group lanes by bin; the lowest lane of every group adds the
            // group's values to the bin in lane order (= the reference's raster order), groups have distinct bins
            s_val[warp][lane] = val;
            __syncwarp();
            const uint32_t grp = __match_any_sync(0xffffffffu, bin);
            if (bin >= 0 && lane == __ffs(grp) - 1) {
                float hsum = s_raw[warp][bin + 2];
                uint32_t m = grp;
                while (m) { const int j = __ffs(m) - 1; m &= m - 1; hsum += s_val[warp][j]; }
                s_raw[warp][bin + 2] = hsum;
            }
            __syncwarp();
        }
        // raw_hist with 2-bin circular padding, src/lib.rs:742-749
        if (lane == 0) {
            s_raw[warp][1] = s_raw[warp][ORI_BINS + 1];
            s_raw[warp][0] = s_raw[warp][ORI_BINS];
            s_raw[warp][ORI_BINS + 2] = s_raw[warp][2];
            s_raw[warp][ORI_BINS + 3] = s_raw[warp][3];
        }
        __syncwarp();
        // smoothing, src/lib.rs:751-755
        float hmax_local = -1.f;
        for (int k = lane; k < ORI_BINS; k += 32) {
            const float* rw = s_raw[warp] + k + 2;
            const float hv = (rw[-2] + rw[2]) * (1.f / 16.f) + (rw[-1] + rw[1]) * (4.f / 16.f) + rw[0] * 6.f / 16.f;
            s_hist[warp][k] = hv;
            hmax_local = fmaxf(hmax_local, hv);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) hmax_local = fmaxf(hmax_local, __shfl_xor_sync(0xffffffffu, hmax_local, d));
        __syncwarp();
        const float thr = hmax_local * 0.8f;  // src/lib.rs:394
        // peaks, ascending bin order (src/lib.rs:397-431)
        uint32_t count = 0;
        for (int kb = 0; kb < ORI_BINS; kb += 32) {
            const int k = kb + lane;
            bool peak = false;
            float ang = 0.f;
            if (k < ORI_BINS) {
                const int km = k > 0 ? k - 1 : ORI_BINS - 1;
                const int kp = k < ORI_BINS - 1 ? k + 1 : 0;
                const float hk = s_hist[warp][k], hm = s_hist[warp][km], hp = s_hist[warp][kp];
                if (hk > hm && hk > hp && hk >= thr) {
                    peak = true;
                    const float interp = (hm - hp) / (hm - 2.0f * hk + hp);
                    float b = (float)k + 0.5f * interp;
                    if (b < 0.0f) b = 36.0f + b;
                    else if (b >= 36.0f) b = b - 36.0f;
                    ang = 360.0f - (360.0f / 36.0f) * b;
                }
            }
            const uint32_t pm = __ballot_sync(0xffffffffu, peak);
            if (peak) {
                const uint32_t slot = count + __popc(pm & ((1u << lane) - 1u));
                if (slot < MAX_ORI) angles[(long long)i * MAX_ORI + slot] = ang;
            }
            count += __popc(pm);
        }
        if (lane == 0) n_ori[i] = min(count, (uint32_t)MAX_ORI);
        __syncwarp();
      }
    }
}

// exclusive scan of n_ori over the candidates of one image (one CTA per image)
__global__ void __launch_bounds__(1024) k_kpscan(const KpParams P) {
    pdl_wait();
    __shared__ uint32_t wsum[32];
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = min(P.cand_count[img], P.cap);
    const uint32_t* cnt = P.n_ori + (long long)img * P.cap;
    uint32_t* off = P.kp_off + (long long)img * P.cap;
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t start = min(tid * per, n), end = min(start + per, n);
    uint32_t local = 0;
    for (uint32_t i = start; i < end; i++) local += cnt[i];
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
        wsum[lane] = s - v;
        if (lane == 31) P.kp_count[img] = s;
    }
    __syncthreads();
    uint32_t run = wsum[warp] + incl - local;
    for (uint32_t i = start; i < end; i++) {
        uint32_t c = cnt[i];
        off[i] = run;
        run += c;
    }
}

// SiftKeyPoint records in natural order (src/lib.rs:419-427)
__global__ void __launch_bounds__(256) k_emit(const KpParams P) {
    pdl_wait();
    const long long img = blockIdx.y;
    const uint32_t n = min(P.cand_count[img], P.cap);
    const Refined* refined = P.refined + img * (long long)P.cap;
    const uint32_t* n_ori = P.n_ori + img * (long long)P.cap;
    const uint32_t* kp_off = P.kp_off + img * (long long)P.cap;
    const float* angles = P.angles + img * (long long)P.cap * MAX_ORI;
    DevKeyPoint* kps = P.kps + img * (long long)P.kcap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = n_ori[i];
        if (!c) continue;
        const Refined r = refined[i];
        const uint32_t base = kp_off[i];
        for (uint32_t k = 0; k < c; k++) {
            if (base + k >= P.kcap) break;
            DevKeyPoint kp;
            kp.x = r.x; kp.y = r.y; kp.size = r.size; kp.response = r.response;
            kp.angle = angles[(long long)i * MAX_ORI + k];
            kp.octave = r.octave_scale >> 8; kp.scale = r.octave_scale & 255; kp.pad = 0;
            kps[base + k] = kp;
        }
    }
}

// features_limit (src/lib.rs:156-161): stable LSD radix sort of the keypoint
// indices by descending response; ties keep natural order.  One CTA per image.
__global__ void __launch_bounds__(1024) k_sort_response(const DevKeyPoint* __restrict__ kps_all,
                                                         const uint32_t* __restrict__ kp_count, uint32_t kcap,
                                                         uint32_t* __restrict__ scratch /* [img][4*kcap] */,
                                                         uint32_t* __restrict__ order /* [img][kcap] */) {
    pdl_wait();
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_wcnt[32][257];
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = min(kp_count[img], kcap);
    const DevKeyPoint* kps = kps_all + (long long)img * kcap;
    uint32_t* keyA = scratch + (long long)img * 4 * kcap;
    uint32_t* idxA = keyA + kcap;
    uint32_t* keyB = idxA + kcap;
    uint32_t* idxB = keyB + kcap;
    for (uint32_t i = tid; i < n; i += 1024) {
        keyA[i] = ~__float_as_uint(kps[i].response);  // responses are >= 0: bit pattern is monotonic
        idxA[i] = i;
    }
    __syncthreads();
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += 1024) atomicAdd(&s_hist[(keyA[i] >> shift) & 255u], 1u);
        __syncthreads();
        if (warp == 0) {  // exclusive scan of 256 bins: 8 per lane
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = s_hist[lane * 8 + k]; sum += v[k]; }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { s_hist[lane * 8 + k] = run; run += v[k]; }
        }
        __syncthreads();
        for (uint32_t c0 = 0; c0 < n; c0 += 1024) {
            for (int k = tid; k < 32 * 257; k += 1024) (&s_wcnt[0][0])[k] = 0;
            __syncthreads();
            const uint32_t i = c0 + tid;
            const bool valid = i < n;
            const uint32_t key = valid ? keyA[i] : 0u, id = valid ? idxA[i] : 0u;
            const uint32_t dg = valid ? ((key >> shift) & 255u) : 256u;
            const uint32_t grp = __match_any_sync(0xffffffffu, dg);
            const uint32_t rank = __popc(grp & ((1u << lane) - 1u));
            if (valid && lane == __ffs(grp) - 1) s_wcnt[warp][dg] = __popc(grp);
            __syncthreads();
            if (tid < 256) {
                uint32_t run = s_hist[tid];
                for (int wv = 0; wv < 32; wv++) {
                    uint32_t t = s_wcnt[wv][tid];
                    s_wcnt[wv][tid] = run;
                    run += t;
                }
                s_hist[tid] = run;
            }
            __syncthreads();
            if (valid) {
                const uint32_t pos = s_wcnt[warp][dg] + rank;
                keyB[pos] = key;
                idxB[pos] = id;
            }
            __syncthreads();
        }
        uint32_t* t;
        t = keyA; keyA = keyB; keyB = t;
        t = idxA; idxA = idxB; idxB = t;
        __syncthreads();
    }
    uint32_t* out = order + (long long)img * kcap;
    for (uint32_t i = tid; i < n; i += 1024) out[i] = idxA[i];
}

// ---------------------------------------------------------------------------
// compute_descriptor (src/lib.rs:785-990) + output pack (src/lib.rs:164-174).
// One warp per keypoint.
//   * Only the inner 4x4 cells of the reference's (6,6,8) histogram survive the crop at :951, so
//     only those 128 bins are accumulated.
//   * No atomics: the warp keeps DESC_COPIES = 16 lane-private copies of the 128 bins in shared memory,
//     copy k at word bin*DESC_COPIES + k, so a half-warp's read-modify-writes always hit 16 distinct
//     banks.  Lanes l and l+16 share a copy WITHOUT ever touching the same word in the same instruction:
//     a sample feeds two adjacent orientation bins, one even and one odd; the lower half-warp adds its
//     even-bin parts first and its odd-bin parts second, the upper half-warp the other way round, so all 32
//     lanes are active in all eight read-modify-writes of a sample.  The copies are summed
//     at the end (f32 sums in a different order than the reference's raster order: the +-1 byte
//     tolerance of the north star covers it).  The sample geometry (rotation, bins, cell indices, fractions)
//     is the reference's f32 arithmetic; magnitude, Gaussian weight and gradient angle use single-instruction
//     SFU approximations (relative error ~1e-6), far below the u8 quantisation step.
//   * Window samples that fall outside the rotated 4x4 grid (about half) are never visited: the lanes
//     compute, in parallel, the column span of every window row that can intersect the grid and a running
//     sample count (one packed word per non-empty row in shared memory); every lane then walks that table with
//     its own cursor, 32 samples forward per step, so the expensive gradient /
//     exp / atan2 / trilinear part always runs with 32 active lanes.
//   * The four pixel loads of the next 32 samples are issued before the arithmetic of the current 32.
// ---------------------------------------------------------------------------
#ifndef SB_DESC_WARPS
#define SB_DESC_WARPS 8
#endif
#ifndef SB_DESC_COPIES
#define SB_DESC_COPIES 16
#endif
#ifndef SB_DESC_L2_PREFETCH
#define SB_DESC_L2_PREFETCH 1
#endif
#ifndef SB_DESC_EXACT_GEOM
#define SB_DESC_EXACT_GEOM 0
#endif
#ifndef SB_DESC_PAIR
#define SB_DESC_PAIR 2   // 1: two samples per lane and step, packed f32x2 arithmetic (descriptor_sample2); 2: + two register sets
#endif
#ifndef SB_DESC_LIST_PAIR
#define SB_DESC_LIST_PAIR 1   // the same choice for compute_descriptor on caller-supplied keypoints (k_descriptor_list)
#endif
#ifndef SB_DESC_LIST_MINB
#define SB_DESC_LIST_MINB 3
#endif
#ifndef SB_DESC_MINB
#define SB_DESC_MINB 3   // resident CTAs per SM the extraction kernel is compiled for (shared memory allows 3)
#endif
constexpr int DESC_WARPS = SB_DESC_WARPS;
constexpr int DESC_COPIES = SB_DESC_COPIES;
constexpr int DESC_MAXROWS = 256;  // window rows: 2 * radius + 1 with radius <= 127 (+ the sentinel entry)
static_assert(DESC_COPIES == 16, "the parity-split scatter pairs lanes l and l + 16 on one copy");
constexpr int DESC_CELL_WORDS = 8 * DESC_COPIES;               // one spatial cell: 8 orientation bins x copies
constexpr int DESC_HIST_WORDS = 16 * DESC_CELL_WORDS;          // the 4x4 cells the crop at :951 keeps
// per warp: histogram copies, row table u16[256] = (span length << 8) | first column  (+ 64 words of slack)
constexpr int DESC_SMEM_WORDS = DESC_HIST_WORDS + DESC_MAXROWS / 2 + DESC_MAXROWS / 4;
constexpr size_t DESC_SMEM_BYTES = 256 + (size_t)DESC_WARPS * DESC_SMEM_WORDS * sizeof(float);

struct DescTarget {
    const float* img;  // layer base
    int w, h, pitch;
    float x, y, scale, orientation;  // arguments of compute_descriptor
};

struct DescGeom {
    const float* img;
    int w, h, pitch, x, y, radius;
    float sin_s, cos_s, orientation;
    int park;         // pixel offset a lane without a sample loads around: the keypoint's pixel, clamped into the interior
    float ori_bins;   // orientation in histogram bins (8 per turn)
    float wscale;     // -(1/8) log2(e) / hist_width^2: exponent of the Gaussian weight per squared window distance
};

// single-instruction SFU approximations (flush-to-zero: no denormal rescaling sequences around the MUFU)
__device__ __forceinline__ float rcp_approx(const float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_approx(const float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(const float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// atan2(y, x) in orientation-bin units (8 bins per turn), [0, 8]: minimax polynomial of atan on [0,1]
// (|error| < 3.3e-7 rad, coefficients pre-multiplied by 8 / 2pi) plus octant folding.  The reference's f64 atan2
// only feeds the trilinear orientation weights here (continuous in the angle, taken modulo 8 bins by the caller), so
// a 5e-7 bin error is far below the u8 quantisation step of the descriptor.
__device__ __forceinline__ float fast_atan2_bins(const float y, const float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    // the floor keeps the flush-to-zero reciprocal finite at 0 / 0 (quotient 0 * finite = 0) and for denormal gradients
    // (deep inside constant regions), whose direction is then arbitrary -- their magnitude, < 1e-30 against the
    // >= 1e-3 of any keypoint's neighbourhood, cannot move a descriptor byte
    const float a = mn * rcp_approx(fmaxf(mx, 1e-30f));
    const float s = a * a;
    float p = 0x1.1c32bcp-7f;
    p = fmaf(p, s, -0x1.5e8130p-5f);
    p = fmaf(p, s, 0x1.9f40a4p-4f);
    p = fmaf(p, s, -0x1.59126ap-3f);
    p = fmaf(p, s, 0x1.0240f6p-2f);
    p = fmaf(p, s, -0x1.b26416p-2f);
    p = fmaf(p, s, 0x1.45f2b4p+0f);
    float r = p * a;                 // [0, 1]: one bin is pi/4
    if (ay > ax) r = 2.0f - r;       // [0, 2]
    if (x < 0.f) r = 4.0f - r;       // [0, 4]
    if (y < 0.f) r = 8.0f - r;       // (4, 8]
    return r;
}

// one queued sample per active lane: gradient, weight, angle, trilinear split, accumulation
struct DescPix { float xp, xm, ym, yp; };  // I[y][x+1], I[y][x-1], I[y-1][x], I[y+1][x]

struct DescAt { int yw, xw; };              // window coordinates of a lane's sample (relative to the keypoint pixel)

// A lane without a sample is parked in the row table's sentinel entry: window columns >= 255 - radius, which are
// outside the rotated 4x4 grid for every orientation (|column| > radius = 2.5 sqrt(2) cells), so nothing of it is
// accumulated whatever pixels it carries -- it loads the neighbourhood of the keypoint's own pixel, clamped into the
// image interior (DescGeom::park), instead of branching.
__device__ __forceinline__ DescPix descriptor_fetch(const DescGeom& G, const DescAt at, const bool active) {
    const int i = active ? (G.y + at.yw) * G.pitch + (G.x + at.xw)   // a layer holds < 2^31 floats: 32-bit offsets
                         : G.park;
    // one 64-bit address, kept opaque so that the neighbours are pointer +- pitch (two adds each) instead of
    // three more base + 64-bit index computations
    const float* pc = G.img + i;
    asm volatile("" : "+l"(pc));
    const ptrdiff_t pitch = G.pitch;
    SB_CHECK_LOAD(pc - pitch, G.img, (long long)G.h * G.pitch, "descriptor_fetch (row above)");
    SB_CHECK_LOAD(pc + pitch, G.img, (long long)G.h * G.pitch, "descriptor_fetch (row below)");
    SB_CHECK_LOAD(pc - 1, G.img, (long long)G.h * G.pitch, "descriptor_fetch (left)");
    SB_CHECK_LOAD(pc + 1, G.img, (long long)G.h * G.pitch, "descriptor_fetch (right)");
    DescPix q;
    q.xp = __ldg(pc + 1); q.xm = __ldg(pc - 1);
    q.ym = __ldg(pc - pitch); q.yp = __ldg(pc + pitch);
    return q;
}

// one sample per lane: gradient, weight, angle, trilinear split, accumulation.  Branch-free up to the
// accumulation; the parts of a sample that fall outside the 4x4 grid (all of it for a sample that fails the exact
// membership test) are predicated off read-modify-write by read-modify-write, and the four cells sit at constant
// offsets from one base address per orientation bin.
__device__ __forceinline__ void descriptor_sample(const DescGeom& G, const DescAt at, const DescPix px, float* hist,
                                                  const int lane) {
    // geometry (src/lib.rs:818-833): row_bin - 0.5 and col_bin - 0.5 of the rotated, scaled sample position.
    // SB_DESC_EXACT_GEOM = 1 keeps the reference's operation sequence (mul, mul, add, + 2, - 0.5); the default
    // contracts it into two FMAs per coordinate (results within one rounding of the reference's; the trilinear
    // weights are continuous across cell boundaries, so a sample that lands an ulp on the other side of one moves
    // 1e-7 of its magnitude between two cells).
    const int yw = at.yw, xw = at.xw;
    const float fx = (float)xw, fy = (float)yw;
#if SB_DESC_EXACT_GEOM
    const float col_rot = fx * G.cos_s - fy * G.sin_s;
    const float row_rot = fx * G.sin_s + fy * G.cos_s;
    const float rb = (row_rot + 2.0f) - 0.5f, cbn = (col_rot + 2.0f) - 0.5f;
    const float wexp = fmaf(col_rot, col_rot, row_rot * row_rot) * (-0.125f * 1.44269504088896341f);   // exp(-2/4^2 * ..), :859
#else
    const float rb = fmaf(fx, G.sin_s, fmaf(fy, G.cos_s, 1.5f));
    const float cbn = fmaf(fx, G.cos_s, fmaf(fy, -G.sin_s, 1.5f));
    const float wexp = fmaf(fx, fx, fy * fy) * G.wscale;   // |rotated position|^2 = (x^2 + y^2) / hist_width^2
#endif
    // The membership test of src/lib.rs:834-837, -0.5 < row_bin, col_bin < 4.5, needs no instruction of its own:
    // it holds exactly when floor(row_bin - 0.5) and floor(col_bin - 0.5) lie in -1..3 (at the one value where the
    // two differ, row_bin == -0.5, the sample's share of every kept cell is c1 = mag * 0 = +0), and a sample whose
    // floors are outside that range has no valid cell below, so none of it is accumulated -- which is also what
    // happens to lanes without a sample (parked in the sentinel row).  (The image-bounds half of the test,
    // :838-841, is enforced by the span construction.)
    const float dx = px.xp - px.xm;
    const float dy = px.ym - px.yp;
    // magnitude, Gaussian weight and angle: fast approximations (relative error ~1e-6), see header comment
    const float d2 = fmaf(dx, dx, dy * dy);
    // sqrt(d2) = d2 * rsqrt(d2); the floor keeps the flush-to-zero rsqrt finite at d2 = 0 (product 0) and for
    // gradients below 1e-18, which weigh nothing next to a keypoint's neighbourhood
    const float root = d2 * rsqrt_approx(fmaxf(d2, 1e-36f));
    const float mag = root * ex2_approx(wexp);
    const float obin = fast_atan2_bins(dy, dx) - G.ori_bins;                       // :871, in bins; in [-8, 8]
    const float row_floor = floorf(rb), col_floor = floorf(cbn), ori_floor = floorf(obin);
    const float row_frac = rb - row_floor, col_frac = cbn - col_floor, ori_frac = obin - ori_floor;
    // spatial half of the trilinear split exactly as src/lib.rs:906-913
    const float c1 = mag * row_frac, c0 = mag - c1;
    float sp[4];                                    // cells (r1, q1), (r1, q1+1), (r1+1, q1), (r1+1, q1+1)
    sp[3] = c1 * col_frac; sp[2] = c1 - sp[3];      // c11, c10
    sp[1] = c0 * col_frac; sp[0] = c0 - sp[1];      // c01, c00
    // the reference adds into cells (row_floor+1 .. +2, col_floor+1 .. +2) of its 6x6 grid and keeps
    // cells 1..4 (:951): in inner-grid terms the first cell is (r1, q1) in -1..3
    const int r1 = (int)row_floor, q1 = (int)col_floor;
    const int oi = (int)ori_floor;                  // in [-16, 16): orientation bins o0 = oi mod 8 and o0 + 1 mod 8 (:926-938)
    // Orientation half (:914-919): bin o0 gets c - c * ori_frac, bin o0 + 1 gets c * ori_frac.  One of the two bins is
    // even, the other odd: E = (oi + 1) & 6, O = (oi & 7) | 1.  Lanes 0-15 (hl = 0) add to E first and O second,
    // lanes 16-31 (hl = 1) to O first and E second, so the two lanes that share a histogram copy never touch the same
    // word in the same instruction.  The first bin is o0 exactly when hl == (oi & 1); its weight is then 1 - ori_frac
    // (c * (1 - f) instead of the reference's c - c * f: one rounding apart), else ori_frac (the reference's products).
    const int hl = lane >> 4;
    const int oF = ((oi + 1 - hl) & (6 | hl)) | hl;
    const int oS = ((oi + hl) & (7 - hl)) | (1 - hl);
    const float wF = ((oi ^ hl) & 1) ? ori_frac : 1.0f - ori_frac;
    // the four spatial cells sit at constant offsets from cell (r1, q1); parts that fall outside the 4x4 grid are
    // predicated off (the base may then point outside the histogram -- it is only dereferenced for cells inside).
    // Cell (r, q) is inside iff both indices are in 0..3, i.e. (r | q) has no bit above bit 1.
    const int r2 = r1 + 1, q2 = q1 + 1;
    const bool ok[4] = {((r1 | q1) & ~3) == 0, ((r1 | q2) & ~3) == 0, ((r2 | q1) & ~3) == 0, ((r2 | q2) & ~3) == 0};
    constexpr int CO[4] = {0, DESC_CELL_WORDS, 4 * DESC_CELL_WORDS, 5 * DESC_CELL_WORDS};
    float* const cellb = hist + (lane & (DESC_COPIES - 1)) + (r1 * 4 + q1) * DESC_CELL_WORDS;
    float* const bF = cellb + oF * DESC_COPIES;
    float* const bS = cellb + oS * DESC_COPIES;
    float vF[4], vS[4], old[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { vF[k] = sp[k] * wF; vS[k] = sp[k] - vF[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) old[k] = bF[CO[k]];     // (left unset for parts outside the grid)
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) bF[CO[k]] = old[k] + vF[k];
    __syncwarp();   // the other half-warp's first-bin stores precede this half's second-bin loads
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) old[k] = bS[CO[k]];
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) bS[CO[k]] = old[k] + vS[k];
    __syncwarp();
}

// ---- two samples per lane and step (SB_DESC_PAIR) ----
// The sample step is bound by instruction issue, and about a quarter of its instructions are plain f32 adds, multiplies
// and FMAs.  A lane that takes TWO samples per step -- its samples of two consecutive batches of 32 -- runs those as
// packed f32x2 operations (SASS FFMA2 / FADD2 / FMUL2: the same IEEE results as the scalar instructions, half the issue
// slots); the SFU approximations, the float -> int conversions, the octant folding and the shared-memory accumulation
// stay per sample.  Operation for operation the arithmetic of descriptor_sample, and the two samples are accumulated one
// after the other, so every histogram word receives its contributions in the same order: bit-identical descriptors.
__device__ __forceinline__ float2 dup2(const float v) { return make_float2(v, v); }

__device__ __forceinline__ float2 fast_atan2_bins2(const float2 y, const float2 x) {
    const float ax0 = fabsf(x.x), ay0 = fabsf(y.x), ax1 = fabsf(x.y), ay1 = fabsf(y.y);
    const float mx0 = fmaxf(ax0, ay0), mn0 = fminf(ax0, ay0), mx1 = fmaxf(ax1, ay1), mn1 = fminf(ax1, ay1);
    const float2 a = mul2(make_float2(mn0, mn1), make_float2(rcp_approx(fmaxf(mx0, 1e-30f)), rcp_approx(fmaxf(mx1, 1e-30f))));
    const float2 s = mul2(a, a);
    float2 p = dup2(0x1.1c32bcp-7f);
    p = fma2(p, s, dup2(-0x1.5e8130p-5f));
    p = fma2(p, s, dup2(0x1.9f40a4p-4f));
    p = fma2(p, s, dup2(-0x1.59126ap-3f));
    p = fma2(p, s, dup2(0x1.0240f6p-2f));
    p = fma2(p, s, dup2(-0x1.b26416p-2f));
    p = fma2(p, s, dup2(0x1.45f2b4p+0f));
    float2 r = mul2(p, a);
    if (ay0 > ax0) r.x = 2.0f - r.x;
    if (x.x < 0.f) r.x = 4.0f - r.x;
    if (y.x < 0.f) r.x = 8.0f - r.x;
    if (ay1 > ax1) r.y = 2.0f - r.y;
    if (x.y < 0.f) r.y = 4.0f - r.y;
    if (y.y < 0.f) r.y = 8.0f - r.y;
    return r;
}

// the accumulation half of descriptor_sample for one sample: spatial parts sp-split into first / second orientation bin
__device__ __forceinline__ void descriptor_scatter(float* hist, const int lane, const float row_floor, const float col_floor,
                                                   const float ori_floor, const float vF0, const float vF1, const float vF2,
                                                   const float vF3, const float vS0, const float vS1, const float vS2,
                                                   const float vS3) {
    const int r1 = (int)row_floor, q1 = (int)col_floor;
    const int oi = (int)ori_floor;
    const int hl = lane >> 4;
    const int oF = ((oi + 1 - hl) & (6 | hl)) | hl;
    const int oS = ((oi + hl) & (7 - hl)) | (1 - hl);
    const int r2 = r1 + 1, q2 = q1 + 1;
    const bool ok[4] = {((r1 | q1) & ~3) == 0, ((r1 | q2) & ~3) == 0, ((r2 | q1) & ~3) == 0, ((r2 | q2) & ~3) == 0};
    constexpr int CO[4] = {0, DESC_CELL_WORDS, 4 * DESC_CELL_WORDS, 5 * DESC_CELL_WORDS};
    float* const cellb = hist + (lane & (DESC_COPIES - 1)) + (r1 * 4 + q1) * DESC_CELL_WORDS;
    float* const bF = cellb + oF * DESC_COPIES;
    float* const bS = cellb + oS * DESC_COPIES;
    const float vF[4] = {vF0, vF1, vF2, vF3}, vS[4] = {vS0, vS1, vS2, vS3};
    float old[4];
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) old[k] = bF[CO[k]];
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) bF[CO[k]] = old[k] + vF[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) old[k] = bS[CO[k]];
#pragma unroll
    for (int k = 0; k < 4; k++) if (ok[k]) bS[CO[k]] = old[k] + vS[k];
    __syncwarp();
}

__device__ __forceinline__ void descriptor_sample2(const DescGeom& G, const DescAt atA, const DescAt atB, const DescPix pA,
                                                   const DescPix pB, float* hist, const int lane) {
    const float2 fx = make_float2((float)atA.xw, (float)atB.xw), fy = make_float2((float)atA.yw, (float)atB.yw);
    const float2 sin2 = dup2(G.sin_s), cos2 = dup2(G.cos_s), nsin2 = dup2(-G.sin_s);
    const float2 rb = fma2(fx, sin2, fma2(fy, cos2, dup2(1.5f)));
    const float2 cbn = fma2(fx, cos2, fma2(fy, nsin2, dup2(1.5f)));
    const float2 wexp = mul2(fma2(fx, fx, mul2(fy, fy)), dup2(G.wscale));
    const float2 dx = sub2(make_float2(pA.xp, pB.xp), make_float2(pA.xm, pB.xm));
    const float2 dy = sub2(make_float2(pA.ym, pB.ym), make_float2(pA.yp, pB.yp));
    const float2 d2 = fma2(dx, dx, mul2(dy, dy));
    const float2 root = mul2(d2, make_float2(rsqrt_approx(fmaxf(d2.x, 1e-36f)), rsqrt_approx(fmaxf(d2.y, 1e-36f))));
    const float2 mag = mul2(root, make_float2(ex2_approx(wexp.x), ex2_approx(wexp.y)));
    const float2 obin = sub2(fast_atan2_bins2(dy, dx), dup2(G.ori_bins));
    const float2 row_floor = make_float2(floorf(rb.x), floorf(rb.y));
    const float2 col_floor = make_float2(floorf(cbn.x), floorf(cbn.y));
    const float2 ori_floor = make_float2(floorf(obin.x), floorf(obin.y));
    const float2 row_frac = sub2(rb, row_floor), col_frac = sub2(cbn, col_floor), ori_frac = sub2(obin, ori_floor);
    const float2 c1 = mul2(mag, row_frac), c0 = sub2(mag, c1);
    float2 sp[4];
    sp[3] = mul2(c1, col_frac); sp[2] = sub2(c1, sp[3]);
    sp[1] = mul2(c0, col_frac); sp[0] = sub2(c0, sp[1]);
    // weight of the orientation bin a lane adds first: ori_frac or 1 - ori_frac (see descriptor_sample)
    const float2 omf = sub2(dup2(1.0f), ori_frac);
    const int hl = lane >> 4;
    const float2 wF = make_float2((((int)ori_floor.x ^ hl) & 1) ? ori_frac.x : omf.x,
                                  (((int)ori_floor.y ^ hl) & 1) ? ori_frac.y : omf.y);
    float2 vF[4], vS[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { vF[k] = mul2(sp[k], wF); vS[k] = sub2(sp[k], vF[k]); }
    descriptor_scatter(hist, lane, row_floor.x, col_floor.x, ori_floor.x, vF[0].x, vF[1].x, vF[2].x, vF[3].x,
                       vS[0].x, vS[1].x, vS[2].x, vS[3].x);
    descriptor_scatter(hist, lane, row_floor.y, col_floor.y, ori_floor.y, vF[0].y, vF[1].y, vF[2].y, vF[3].y,
                       vS[0].y, vS[1].y, vS[2].y, vS[3].y);
}

// Window radius limit of the row table (8-bit spans): scale <= DESC_MAX_SCALE.  The extraction path stays below
// scale 3.6 (radius 38); compute_descriptor on caller-supplied keypoints reports larger scales as invalid
// arguments (SB200_E_INVALID) instead of computing something the crate would not.
constexpr int DESC_MAX_RADIUS = 127;
__host__ __device__ __forceinline__ int descriptor_radius(const float scale) {
    return (int)roundf(fminf(3.0f * scale * sqrtf(2.0f) * 5.0f * 0.5f, 1e6f));   // src/lib.rs:800
}

// returns false (warp-uniform, nothing written) when the keypoint's window exceeds DESC_MAX_RADIUS
template <int PAIR>
__device__ __forceinline__ bool descriptor_warp(const DescTarget t, float* wsm /* smem [DESC_SMEM_WORDS] */, int lane,
                                                uint8_t* out /* 128 B */) {
    float* hist = wsm;
    uint16_t* row_tab = reinterpret_cast<uint16_t*>(wsm + DESC_HIST_WORDS);   // [DESC_MAXROWS] (span length << 8) | (xlo + radius)
    {
        float4* h4 = reinterpret_cast<float4*>(hist);
        for (int k = lane; k < DESC_HIST_WORDS / 4; k += 32) h4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    DescGeom G;
    G.img = t.img;
    asm volatile("" : "+l"(G.img));   // a plain 64-bit pointer from here on (not base + layer offset re-derived per use)
    G.w = t.w; G.h = t.h; G.pitch = t.pitch; G.orientation = t.orientation;
    G.ori_bins = t.orientation * (8.0f / 360.0f);
    // `x.round() as usize` (src/lib.rs:796-797): saturating, negative -> 0
    const float xr = roundf(t.x), yr = roundf(t.y);
    G.x = xr > 0.f ? (int)fminf(xr, 1e9f) : 0;
    G.y = yr > 0.f ? (int)fminf(yr, 1e9f) : 0;
    // (a caller-supplied keypoint may sit on the border or outside the image: its own neighbourhood is not loadable)
    G.park = min(max(G.y, 1), max(G.h - 2, 1)) * G.pitch + min(max(G.x, 1), max(G.w - 2, 1));
    const float hist_width = 3.0f * t.scale;
    G.radius = descriptor_radius(t.scale);  // :800 (<= 38 on the extraction path)
    if (G.radius > DESC_MAX_RADIUS || !(t.scale > 0.f)) return false;
#if SB_DESC_L2_PREFETCH
    // The window's pixels come from DRAM (a group's pyramids are far larger than L2) and the sample loop only looks one
    // batch of 32 samples ahead, which no longer covers a DRAM round trip (ncu: 30 % of the stall samples on the
    // prefetched pixels).  Pull the bounding box of the window into L2 now, while the row table is built: a few
    // prefetches per lane, no registers held, and the loop's loads then hit L2.
    {
        const int x0 = max(G.x - G.radius - 1, 0), x1 = min(G.x + G.radius + 1, G.w - 1);
        const int y0 = max(G.y - G.radius - 1, 0), y1 = min(G.y + G.radius + 1, G.h - 1);
        for (int r = y0 + lane; r <= y1; r += 32) {
            const float* row = G.img + (long long)r * G.pitch;
            const uintptr_t a = reinterpret_cast<uintptr_t>(row + x0) & ~(uintptr_t)127, b = reinterpret_cast<uintptr_t>(row + x1);
            for (uintptr_t q = a; q <= b; q += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
    }
#endif
    const float rad = t.orientation * (3.14159265358979323846f / 180.0f);            // f32::to_radians
    double sd, cd;
    sincos((double)rad, &sd, &cd);  // libm sinf/cosf are (nearly always) correctly rounded: round once from f64
    G.sin_s = (float)sd / hist_width;
    G.cos_s = (float)cd / hist_width;
    G.wscale = (-0.125f * 1.44269504088896341f) / (hist_width * hist_width);
    // conservative per-row column range of the rotated 5x5-cell square: x*k + t in (-2.5, 2.5)
    const float inv_s = fabsf(G.sin_s) > 1e-7f ? 1.0f / G.sin_s : 0.f;
    const float inv_c = fabsf(G.cos_s) > 1e-7f ? 1.0f / G.cos_s : 0.f;
    const int xw_min = max(-G.radius, 1 - G.x), xw_max = min(G.radius, G.w - 2 - G.x);
    const uint32_t lt = (1u << lane) - 1u;
    // ---- row table: lanes take window rows in parallel.  Entry k describes the k-th non-empty row (the
    //      non-empty rows are consecutive: convex square, convex image): span length and first column.  At
    //      most a couple of samples per row fail the exact test, which the sample step repeats. ----
    uint32_t n_rows = 0, total = 0, yq_first = 0;  // warp-uniform
    const int nwin = 2 * G.radius + 1;
    for (int rb = 0; rb < nwin; rb += 32) {
        const int yq = rb + lane, yw = yq - G.radius;
        const int ay = G.y + yw;
        int xlo = 0, len = 0;
        if (yq < nwin && ay > 0 && ay < G.h - 1) {
            const float ys = (float)yw * G.sin_s, yc = (float)yw * G.cos_s;
            float lo = (float)xw_min, hi = (float)xw_max;
            if (inv_s != 0.f) {  // row_rot = x*sin_s + yc
                const float a = (-2.5f - yc) * inv_s, b = (2.5f - yc) * inv_s;
                lo = fmaxf(lo, floorf(fminf(a, b)));
                hi = fminf(hi, ceilf(fmaxf(a, b)));
            }
            if (inv_c != 0.f) {  // col_rot = x*cos_s - ys
                const float a = (-2.5f + ys) * inv_c, b = (2.5f + ys) * inv_c;
                lo = fmaxf(lo, floorf(fminf(a, b)));
                hi = fminf(hi, ceilf(fmaxf(a, b)));
            }
            xlo = (int)lo;
            len = max((int)hi - xlo + 1, 0);
        }
        const uint32_t ne = __ballot_sync(0xffffffffu, len > 0);
        uint32_t incl = (uint32_t)len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (len > 0) {
            const uint32_t k = n_rows + __popc(ne & lt);
            row_tab[k] = (uint16_t)((len << 8) | (xlo + G.radius));
        }
        if (n_rows == 0 && ne) yq_first = rb + __ffs(ne) - 1;
        n_rows += __popc(ne);
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
    // sentinel behind the last row: a span no cursor can leave, whose columns (>= 255 - radius > radius) lie outside
    // the rotated grid for every orientation -- a lane that has run out of samples parks there and accumulates nothing
    if (lane == 0) row_tab[n_rows] = 0xffffu;
    __syncwarp();
    // ---- samples, 32 at a time: every lane walks the table with its own cursor (row k, offset xo in the row),
    //      32 samples forward per batch -- one or two short rows at most, no search ----
    uint32_t cur_k = 0, cur_xo = lane, cur_w = row_tab[0];
    const int y_first_w = (int)yq_first - G.radius;
    auto lookup = [&](const uint32_t advance, DescAt& at) -> bool {
        cur_xo += advance;
        while (cur_xo >= (cur_w >> 8)) {
            cur_xo -= cur_w >> 8;
            cur_k++;
            cur_w = row_tab[cur_k];
        }
        at.yw = y_first_w + (int)cur_k;
        at.xw = (int)((cur_w & 255u) + cur_xo) - G.radius;
        return cur_k < n_rows;
    };
    if constexpr (PAIR == 2) {
    if (total) {
        // two batches of 32 samples per step; a cursor is never moved more than 63 samples past the last one, so it
        // cannot leave the 255-column sentinel entry
        // two register sets take turns (the loop body is unrolled twice): the pixels of the next pair of batches are
        // loaded into the set the current pair does not use, so nothing is moved between registers
        DescAt a0, b0, a1, b1;
        DescPix pa0, pb0, pa1, pb1;
        bool act = lookup(0, a0);
        pa0 = descriptor_fetch(G, a0, act);
        act = lookup(32, b0);
        pb0 = descriptor_fetch(G, b0, act);
        for (uint32_t base = 0; base < total; base += 128) {
            const bool more = base + 64 < total;
            if (more) {
                act = lookup(32, a1);
                pa1 = descriptor_fetch(G, a1, act);
                act = lookup(32, b1);
                pb1 = descriptor_fetch(G, b1, act);
            }
            descriptor_sample2(G, a0, b0, pa0, pb0, hist, lane);
            if (!more) break;
            if (base + 128 < total) {
                act = lookup(32, a0);
                pa0 = descriptor_fetch(G, a0, act);
                act = lookup(32, b0);
                pb0 = descriptor_fetch(G, b0, act);
            }
            descriptor_sample2(G, a1, b1, pa1, pb1, hist, lane);
        }
    }
    } else if constexpr (PAIR == 1) {
    if (total) {
        // two batches of 32 samples per step; a cursor is never moved more than 63 samples past the last one, so it
        // cannot leave the 255-column sentinel entry
        DescAt at_a, at_b;
        bool act = lookup(0, at_a);
        DescPix p_a = descriptor_fetch(G, at_a, act);
        act = lookup(32, at_b);
        DescPix p_b = descriptor_fetch(G, at_b, act);
        for (uint32_t base = 0; base < total; base += 64) {
            const DescAt ca = at_a, cb = at_b;
            const DescPix qa = p_a, qb = p_b;
            if (base + 64 < total) {
                act = lookup(32, at_a);
                p_a = descriptor_fetch(G, at_a, act);
                act = lookup(32, at_b);
                p_b = descriptor_fetch(G, at_b, act);
            }
            descriptor_sample2(G, ca, cb, qa, qb, hist, lane);
        }
    }
    } else {
    if (total) {
        DescAt at_next;
        bool a_next = lookup(0, at_next);
        DescPix p_next = descriptor_fetch(G, at_next, a_next);
        for (uint32_t base = 0; base < total; base += 32) {
            const DescAt at = at_next;
            const DescPix px = p_next;
            if (base + 32 < total) {
                a_next = lookup(32, at_next);
                p_next = descriptor_fetch(G, at_next, a_next);
            }
            descriptor_sample(G, at, px, hist, lane);
        }
    }
    }
    __syncwarp();
    // sum the private copies: lane owns flat[4*lane .. 4*lane+3]; rotated start => conflict-free reads
    float f[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float* hb = hist + (4 * lane + k) * DESC_COPIES;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DESC_COPIES; c++) acc += hb[(c + lane) & (DESC_COPIES - 1)];
        f[k] = acc;
    }
    // l2 norm in chunks of four, chunks added in order (src/lib.rs:957-962)
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) s += f[k] * f[k];
    float acc = 0.f;
    for (int j = 0; j < 32; j++) {
        const float sj = __shfl_sync(0xffffffffu, s, j);
        acc = (j == 0) ? sj : acc + sj;
    }
    const float cap = sqrtf(acc) * 0.2f;
#pragma unroll
    for (int k = 0; k < 4; k++) f[k] = fminf(f[k], cap);
    s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) s += f[k] * f[k];
    acc = 0.f;
    for (int j = 0; j < 32; j++) {
        const float sj = __shfl_sync(0xffffffffu, s, j);
        acc = (j == 0) ? sj : acc + sj;
    }
    const float norm = 512.0f / fmaxf(sqrtf(acc), 1.1920929e-07f);
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float q = roundf(f[k] * norm);  // f >= 0; NaN (never produced) would map to 0 like `as i32`
        const int qi = (q > 255.f) ? 255 : (q >= 0.f ? (int)q : 0);
        packed |= (uint32_t)qi << (8 * k);
    }
    reinterpret_cast<uint32_t*>(out)[lane] = packed;
    __syncwarp();
    return true;
}

struct DescParams {
    PyrLayout L;
    const float* gauss;         // Gaussian arena, image 0
    const DevKeyPoint* kps;     // [img][kcap] natural order
    const uint32_t* kp_count;   // [img]
    const uint32_t* order;      // [img][kcap] response-sorted permutation (used when limit < count)
    uint32_t kcap;
    long long limit;            // features_limit, < 0 == None
    const uint32_t* out_count;  // [img] keypoints returned per image
    const uint32_t* out_off;    // [img] offset of image's first keypoint in the dense outputs
    OutKeyPoint* out_kps;       // dense, batch-wide
    uint8_t* out_desc;          // dense, batch-wide, 128 B per keypoint
};

// per-image output counts (features_limit applied, src/lib.rs:156-161) and their
// exclusive scan over the images of the batch; out_off[n_img] = total.  One CTA.
__global__ void __launch_bounds__(1024) k_out_offsets(const uint32_t* __restrict__ kp_count, uint32_t kcap,
                                                       long long limit, int n_img, uint32_t* __restrict__ out_count,
                                                       uint32_t* __restrict__ out_off) {
    pdl_wait();
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < n_img; base += 1024) {
        const int i = base + tid;
        uint32_t c = 0;
        if (i < n_img) {
            c = min(kp_count[i], kcap);
            if (limit >= 0 && (unsigned long long)limit < c) c = (uint32_t)limit;
            out_count[i] = c;
        }
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t v = wsum[lane], sc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, sc, d);
                if (lane >= d) sc += t;
            }
            wsum[lane] = sc - v;
        }
        __syncthreads();
        const uint32_t run = s_run;
        if (i < n_img) out_off[i] = run + wsum[warp] + incl - c;
        __syncthreads();
        if (tid == 1023) s_run = run + wsum[31] + incl;
        __syncthreads();
    }
    if (tid == 0) out_off[n_img] = s_run;
}

// compute_descriptors (src/lib.rs:759-782) over the keypoint list of each image and
// the final KeyPoint records (src/lib.rs:164-174).
#ifndef SB_DESC_CHUNK
#define SB_DESC_CHUNK 2
#endif
constexpr int DESC_CHUNK = SB_DESC_CHUNK;

// The output keypoints of all images of the group form one work list (out_off is their exclusive prefix sum);
// warps pull chunks from an atomic counter (see k_orient).
__global__ void __launch_bounds__(32 * DESC_WARPS, SB_DESC_MINB) k_descriptor(const DescParams P, const int n_img,
                                                                uint32_t* __restrict__ work) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned char desc_smem[];  // DESC_SMEM_BYTES, dynamic (> 48 KB)
    uint64_t* s_tab = reinterpret_cast<uint64_t*>(desc_smem);
    float (*s_hist)[DESC_SMEM_WORDS] = reinterpret_cast<float (*)[DESC_SMEM_WORDS]>(desc_smem + 256);
    if (threadIdx.x < 32) s_tab[threadIdx.x] = sbm::d_exp2_tab[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t total = __ldg(P.out_off + n_img);
    WorkQueue wq(work, total, gridDim.x * DESC_WARPS, DESC_CHUNK);
    for (uint32_t first = wq.next(lane, true); first < total; first = wq.next(lane, false)) {
        for (uint32_t g = first; g < min(first + wq.chunk, total); g++) {
            const int img = find_image(P.out_off, n_img, g);
            const uint32_t j = g - __ldg(P.out_off + img);
            const uint32_t n = min(P.kp_count[img], P.kcap);
            const bool limited = P.out_count[img] < n;
            const DevKeyPoint kp = (P.kps + img * (long long)P.kcap)[limited ? (P.order + img * (long long)P.kcap)[j] : j];
            const OctLayout& ol = P.L.o[kp.octave];
            DescTarget t;
            t.img = P.gauss + img * P.L.img_floats + ol.off + (long long)kp.scale * ol.layer_stride;
            t.w = ol.w; t.h = ol.h; t.pitch = ol.pitch;
            const float f = pow2i(-kp.octave);  // 2_f32.powi(-octave), src/lib.rs:768
            t.x = kp.x * f; t.y = kp.y * f; t.scale = kp.size * f;
            t.orientation = 360.0f - kp.angle;   // :766
            descriptor_warp<SB_DESC_PAIR>(t, s_hist[warp], lane, P.out_desc + (size_t)g * DESC_SIZE);
            if (lane == 0) {
                OutKeyPoint o;  // DELTA_MIN = 0.5 undoes the seed upsampling, src/lib.rs:168-170
                o.x = kp.x * 0.5f; o.y = kp.y * 0.5f; o.size = kp.size * 0.5f;
                o.angle = kp.angle; o.response = kp.response;
                P.out_kps[g] = o;
            }
        }
    }
}

// compute_descriptor on caller-supplied keypoints and a dense f32 image
// (benches/descriptor.rs:18-32 shape; src/lib.rs:785).
struct DescIn { float x, y, scale, orientation; };
__global__ void __launch_bounds__(32 * DESC_WARPS, SB_DESC_LIST_MINB) k_descriptor_list(const float* __restrict__ img, int w, int h,
                                                                      int pitch, const DescIn* __restrict__ kps,
                                                                      unsigned long long n, uint8_t* __restrict__ out,
                                                                      uint32_t* __restrict__ err) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned char desc_smem[];  // DESC_SMEM_BYTES, dynamic (> 48 KB)
    uint64_t* s_tab = reinterpret_cast<uint64_t*>(desc_smem);
    float (*s_hist)[DESC_SMEM_WORDS] = reinterpret_cast<float (*)[DESC_SMEM_WORDS]>(desc_smem + 256);
    if (threadIdx.x < 32) s_tab[threadIdx.x] = sbm::d_exp2_tab[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned long long j = (unsigned long long)blockIdx.x * DESC_WARPS + warp; j < n;
         j += (unsigned long long)gridDim.x * DESC_WARPS) {
        const DescIn k = kps[j];
        DescTarget t;
        t.img = img; t.w = w; t.h = h; t.pitch = pitch;
        t.x = k.x; t.y = k.y; t.scale = k.scale; t.orientation = k.orientation;
        if (!descriptor_warp<SB_DESC_LIST_PAIR>(t, s_hist[warp], lane, out + j * DESC_SIZE)) {
            reinterpret_cast<uint32_t*>(out + j * DESC_SIZE)[lane] = 0u;
            if (lane == 0) atomicOr(err, 1u);   // reported as SB200_E_INVALID by the next synchronising call
        }
    }
}

}  // namespace sb
