// sift_b200.cu -- host runtime + C ABI (include/sift_b200.h) of the B200-native SIFT extractor.
//
// One context = one device, two "slots" (stream + device arenas) so that the
// upload / compute / download of consecutive image groups overlap.  Every image
// of a group is processed by the same launches (grid.z / grid.y = image), the
// per-image results are written densely in image order, and the only host
// synchronisation per group is the read of the per-image keypoint counts.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -shared ...
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <functional>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sift_b200.h"
#include "sb_common.cuh"
#include "sb_jpeg.h"
#include "sb_keypoints.cuh"
#include "sb_match.cuh"
#include "sb_pyramid.cuh"

using namespace sb;

static_assert(sizeof(OutKeyPoint) == sizeof(sb200_keypoint), "keypoint layout");
static_assert(sizeof(DescIn) == sizeof(sb200_desc_in), "desc_in layout");
static_assert(sizeof(MatchOut) == sizeof(sb200_dmatch), "match layout");

namespace {

// groups in flight per context: upload / pyramid / keypoint stages of consecutive groups overlap across slots
#ifndef SB_SLOTS
#define SB_SLOTS 2
#endif
constexpr int N_SLOTS = SB_SLOTS;

const char* kStageNames[SB200_STAGE_COUNT] = {"seed", "blur", "extrema", "refine", "orient", "descriptor", "top_blur"};

constexpr int N_SIDE_MAX = 4;
struct Slot {
    int index = 0;
    cudaStream_t stream = nullptr;
    // layers 4, 5 and the extrema scan of octave o run on side stream o % n_side while the main stream already builds
    // the next octave (which only needs layer 3); with several side streams the side work of consecutive octaves
    // overlaps too -- on one side stream it, not the chain of octaves, was the critical path of a single image
    cudaStream_t side[N_SIDE_MAX] = {nullptr};
    cudaEvent_t ev_fork[MAX_OCT] = {nullptr};
    cudaEvent_t ev_join[N_SIDE_MAX] = {nullptr};
    cudaEvent_t ev_counts = nullptr;
    cudaEvent_t ev_upload = nullptr;   // the last host->device copy out of the slot's pinned staging buffer
    // input
    uint8_t* d_in = nullptr;
    uint8_t* h_in = nullptr;  // pinned staging for pageable callers
    size_t in_cap = 0;
    uint8_t* d_rgb = nullptr;  // interleaved RGB / RGBA input of the colour entry point (allocated on first use)
    uint8_t* h_rgb = nullptr;
    size_t rgb_cap = 0;
    // pyramid
    float* d_gauss = nullptr;
    uint32_t* d_mask = nullptr;
    uint32_t* d_rows = nullptr;  // [B][img_rows] counters, followed by [B] cand_count, kp_count
    uint32_t* d_rowoff = nullptr;
    uint32_t* d_counts = nullptr;  // cand_count[B], kp_count[B], out_count[B], out_off[B+1]
    uint32_t* d_sched = nullptr;   // work counters [4], scratch counts [B], candidate prefix sums [B+1]
    // candidates / keypoints
    CandKey* d_keys = nullptr;
    Refined* d_refined = nullptr;
    uint32_t* d_nori = nullptr;
    float* d_angles = nullptr;
    uint32_t* d_kpoff = nullptr;
    DevKeyPoint* d_kps = nullptr;
    uint32_t* d_sort = nullptr;
    uint32_t* d_order = nullptr;
    OutKeyPoint* d_out_kps = nullptr;
    uint8_t* d_out_desc = nullptr;
    uint32_t* h_counts = nullptr;  // pinned mirror of d_counts
    // CUDA graphs of the whole per-group kernel sequence, keyed by its launch parameters (a few shapes per slot)
    struct PipeGraph {
        uint32_t n, w, h, stride; uint64_t img_stride; const uint8_t* d_in; bool detect;
        cudaGraphExec_t exec; uint64_t launches, stage_launches[SB200_STAGE_COUNT]; uint64_t last_use;
    };
    std::vector<PipeGraph> graphs;
    // state of the group in flight
    uint32_t n_imgs = 0;
    uint64_t first_img = 0;
    int64_t limit = -1;
    bool busy = false;
    bool staged = false;          // the slot's pinned staging buffer already holds the group starting at staged_first
    uint64_t staged_first = 0;
};

// chunks of bitstreams in flight: JPEG_AHEAD chunks are decoding (round-robin over JPEG_STREAMS streams, so that
// nvJPEG's latency-bound Huffman kernels of consecutive chunks overlap) while the groups of an earlier chunk run
#ifndef SB_JPEG_AHEAD
#define SB_JPEG_AHEAD 3
#endif
constexpr int JPEG_AHEAD = SB_JPEG_AHEAD;
constexpr int JPEG_STAGES = JPEG_AHEAD + 1;
// one decode stream per stage buffer: stage buffer, nvJPEG state ("lane") and stream of chunk c are all c % JPEG_STAGES,
// so an nvJPEG state is only ever used on one stream
constexpr int JPEG_STREAMS = JPEG_STAGES;

// stage buffer of the JPEG entry points: decoded pixels (Y or interleaved RGB) of one chunk of bitstreams
struct JpegStage {
    uint8_t* d = nullptr;
    size_t cap = 0;
    cudaEvent_t decoded = nullptr;          // recorded on the decode stream after the chunk
    cudaEvent_t read[SB_SLOTS] = {nullptr};  // recorded on a slot's stream after it read its group out of the buffer
    bool read_valid[SB_SLOTS] = {false};
};

struct StageEvents {
    cudaEvent_t a, b;
    int stage;
};

}  // namespace

struct sb200_ctx {
    int device = 0;
    uint32_t max_w = 0, max_h = 0, max_batch = 0, cap = 0;
    std::string err;
    Slot slot[N_SLOTS];
    // layout of the current image size
    PyrLayout L{};
    uint32_t cur_w = 0, cur_h = 0;
    // TMA descriptors of the Gaussian arenas: [slot][octave][destination layer 1..5]
    CUtensorMap tmap[N_SLOTS][MAX_OCT][N_LAYERS];
    CUtensorMap tmap_m[N_SLOTS][MAX_OCT][N_LAYERS];  // marching blur: (BW x 32) boxes
    int flavour = FL_OPENCV;                   // Processing flavour of the pyramid (sb200_set_processing)
    bool march = true;                         // SB200_BLUR=tile selects the independent-tile TMA blur (debugging aid)
    int seg_rows_override = 0;                 // SB200_SEG_ROWS: fixed segment height of the marching blur (tests)
    int pieces_override = 0;                   // SB200_PIECES: pieces per column of the aligned distribution (experiments)
    int seed_fused = 2;                        // the seed blur upsamples its own input bands: 2 = producer warps (default), 1 = in
                                               // line (SB200_SEED=fused), 0 = separate upsample kernel (SB200_SEED=split)
    bool tail = true;                          // SB200_TAIL=0: per-layer launches for the small octaves too (debugging aid)
    bool use_graphs = true;                    // SB200_GRAPHS=0: plain stream launches
    bool fork_octaves = true;                  // SB200_FORK=0: every kernel of a group on one stream
    int n_side = 3;                            // side streams per slot (SB200_SIDES=1..N_SIDE_MAX)
    uint64_t graph_clock = 0;
    CUtensorMap tmap_ex[N_SLOTS][MAX_OCT];  // [slot][octave]: (68 x 3 x 6) boxes of the extrema scan
    bool tmap_ok[MAX_OCT] = {false};
    void* encode_fn = nullptr;  // cuTensorMapEncodeTiled
    // capacities the arenas were sized for
    long long gauss_floats_cap = 0, mask_words_cap = 0;
    int rows_cap = 0;
    // results (pinned, grow-only)
    uint64_t* h_offsets = nullptr;
    size_t offsets_cap = 0;
    sb200_keypoint* h_kps = nullptr;
    uint8_t* h_desc = nullptr;
    size_t res_cap = 0;
    uint64_t res_n = 0;
    double last_gather_ms = 0.0;   // host time of the dense gather of the last sb200_extract_batch_multi
    double last_shard_ms = 0.0;    // host time this context's shard of the last multi-device call took
    // optional OpenCV-style post-filters of the host results (sb200_set_postfilter)
    bool pf_dedup = false;
    int64_t pf_retain = -1;
    // staged API state
    bool have_pyramid = false;
    bool have_single = false;  // slot[0] holds the intermediates of a single-image run
    int64_t last_limit = -1;
    int dev_rr = 0;     // slot used by the next sb200_extract_batch_device
    int last_slot = 0;  // slot holding the most recent run (debug / device-result views)
    // descriptor-only scratch
    float* d_dimg = nullptr;
    size_t dimg_cap = 0;
    DescIn* d_dkps = nullptr;
    uint8_t* d_ddesc = nullptr;
    size_t dkps_cap = 0;
    uint32_t* d_err = nullptr;   // device-side argument errors of k_descriptor_list, read by the next synchronising call
    bool err_pending = false;
    // JPEG input (nvJPEG, loaded on first use): bitstreams are decoded a chunk at a time on their own streams into
    // stage buffers, so that the decode of the next chunks overlaps the extraction of chunk c
    JpegDecoder jpeg;
    JpegStage jstage[JPEG_STAGES];
    cudaStream_t jstream[JPEG_STREAMS] = {nullptr};
    uint32_t jpeg_chunk = 128;   // SB200_JPEG_CHUNK; nvJPEG decodes Huffman on the GPU for batches of >= 100 streams
    // matcher scratch (grow-only)
    uint8_t* d_mdesc[2] = {nullptr, nullptr};        // query / train descriptors
    uint32_t* d_mnorm[2] = {nullptr, nullptr};
    uint32_t* d_mnbp[2] = {nullptr, nullptr};
    unsigned long long* d_mbest[2] = {nullptr, nullptr};
    size_t m_cap[2] = {0, 0};       // rows d_mnorm / d_mnbp / d_mbest are sized for
    size_t mdesc_cap[2] = {0, 0};   // rows d_mdesc is sized for (only the host entry point uses it)
    MatchOut* d_mout = nullptr;
    size_t mout_cap = 0;
    uint32_t* d_mcount = nullptr;
    // measurement
    bool profiling = false;
    std::vector<StageEvents> pending;
    std::vector<cudaEvent_t> ev_pool;
    double stage_ms[SB200_STAGE_COUNT] = {0};
    // per-launch view of the pyramid stages (profiling mode): slot octave * 8 + k, k = 0 upsample + seed blur,
    // 1..5 blur of layer k, 6 extrema scan, 7 fused tail (filed under its first octave)
    double fine_ms[SB200_FINE_SLOTS] = {0};
    uint64_t fine_launches[SB200_FINE_SLOTS] = {0};
    uint64_t stage_launches[SB200_STAGE_COUNT] = {0};
    uint64_t launches = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
    int sm_count = 148;
    int ori_ctas = 148, desc_ctas = 148;   // resident CTAs of k_orient / k_descriptor (their grids)
};

namespace {

int fail(sb200_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, SB200_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),      \
                        __FILE__, __LINE__);                                                             \
    } while (0)

// ---- layout (src/lib.rs:133-134, 245-248) ---------------------------------
int octave_count(uint32_t W2, uint32_t H2) {
    const uint32_t min_axis = std::min(W2, H2);
    const float lg = log2f((float)min_axis) - 2.0f;
    const float r = roundf(lg);
    long long n = (r > 0.f ? (long long)r : 0) + 1;  // `as usize` saturates at 0
    if (n > MAX_OCT) n = MAX_OCT;
    return (int)n;
}

PyrLayout make_layout(uint32_t w, uint32_t h) {
    PyrLayout L{};
    int cw = (int)w * 2, ch = (int)h * 2;
    L.n_oct = octave_count((uint32_t)cw, (uint32_t)ch);
    long long off = 0, moff = 0;
    int rows = 0;
    for (int o = 0; o < L.n_oct; o++) {
        OctLayout& ol = L.o[o];
        ol.w = cw; ol.h = ch;
        ol.pitch = (std::max(cw, 1) + 31) / 32 * 32;
        ol.mask_pitch = 2 * ex_strips(std::max(cw, 1));
        ol.layer_stride = (long long)ol.pitch * std::max(ch, 1);
        ol.off = off;
        off += ol.layer_stride * N_LAYERS;
        ol.mask_off = moff;
        moff += (long long)ol.mask_pitch * std::max(ch, 1) * SCALES_PER_OCTAVE;
        ol.row_base = rows;
        rows += std::max(ch, 1) * SCALES_PER_OCTAVE;
        ol.scanned = !(ch < 2 * IMAGE_BORDER || cw < 2 * IMAGE_BORDER);
        cw /= 2; ch /= 2;
    }
    L.img_floats = off;
    L.img_mask_words = moff;
    L.img_rows = rows;
    return L;
}

// OpenCV getGaussianKernel(ksize, sigma, CV_32F) (what gaussian_blur_def builds,
// src/opencv_processing.rs:51-57): exp(-x^2 / 2 sigma^2) in double, normalised, stored f32.
void gaussian_taps(double sigma, float* out, int& ksize) {
    ksize = ((int)lrint(sigma * 8.0 + 1.0)) | 1;
    const int r = ksize / 2;
    double t[64], sum = 0.0;
    const double scale2x = -0.5 / (sigma * sigma);
    for (int i = 0; i < ksize; i++) {
        const double x = (double)(i - r);
        t[i] = exp(scale2x * x * x);
        sum += t[i];
    }
    const double inv = 1.0 / sum;
    for (int i = 0; i < ksize; i++) out[i] = (float)(t[i] * inv);
}

// imageproc 0.25 gaussian_blur_f32(img, sigma as f32) (src/lib.rs:997): kernel radius ceil(2 sigma), taps
// gaussian_pdf(x) = (sigma * sqrt(2 pi)).recip() * exp(-x^2 / (2 sigma^2)) in f32, NOT renormalised.  Restated from
// the published algorithm (the imageproc sources are not part of the reference tree): parity unpinned.
void imageproc_taps(double sigma64, float* out, int& ksize) {
    const float sigma = (float)sigma64;
    const int r = (int)ceilf(2.0f * sigma);
    ksize = 2 * r + 1;
    const float norm = 1.0f / (sigma * sqrtf(2.0f * 3.14159265358979323846f));
    for (int i = 0; i <= r; i++) {
        const float x = (float)i;
        const float v = norm * expf(-(x * x) / (2.0f * (sigma * sigma)));
        out[r + i] = v;
        out[r - i] = v;
    }
}

// sigmas of src/lib.rs:207 and :220-229
double seed_sigma() { return sqrt(0.8 * 0.8 - 0.5 * 0.5) * 2.0; }
double octave_sigma(int s) {
    const double m = pow(2.0, 2.0 / 3.0);
    int e = s - 1;
    const bool recip = e < 0;
    int b = recip ? -e : e;
    double base = m, a = 1.0;
    for (;;) {  // powi == square-and-multiply
        if (b & 1) a *= base;
        b /= 2;
        if (b == 0) break;
        base *= base;
    }
    if (recip) a = 1.0 / a;
    const double bb = a * m;
    return sqrt(bb - a) * 0.8 * 2.0;
}

// ---- stage timing ----------------------------------------------------------
cudaEvent_t get_event(sb200_ctx* ctx) {
    if (!ctx->ev_pool.empty()) {
        cudaEvent_t e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct StageScope {
    sb200_ctx* ctx;
    cudaStream_t st;
    int stage;
    cudaEvent_t a = nullptr;
    StageScope(sb200_ctx* c, cudaStream_t s, int stg) : ctx(c), st(s), stage(stg) {
        if (ctx->profiling && stage >= 0) {
            a = get_event(ctx);
            cudaEventRecord(a, st);
        }
    }
    ~StageScope() {
        if (ctx->profiling && stage >= 0) {
            cudaEvent_t b = get_event(ctx);
            cudaEventRecord(b, st);
            ctx->pending.push_back({a, b, stage});
        }
    }
};

void drain_stage_events(sb200_ctx* ctx) {
    for (auto& p : ctx->pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            if (p.stage < SB200_STAGE_COUNT) ctx->stage_ms[p.stage] += ms;
            else { ctx->fine_ms[p.stage - SB200_STAGE_COUNT] += ms; ctx->fine_launches[p.stage - SB200_STAGE_COUNT]++; }
        }
        ctx->ev_pool.push_back(p.a);
        ctx->ev_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

inline void count_launch(sb200_ctx* ctx, int stage, int n = 1) {
    ctx->launches += n;
    ctx->stage_launches[stage] += n;
}

// ---- allocation --------------------------------------------------------------
template <class T>
cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}

// per-candidate / per-keypoint arrays of a slot, sized by the context's current capacity
int alloc_cand_arrays(sb200_ctx* ctx, Slot& s) {
    const size_t B = ctx->max_batch, cap = ctx->cap;
    CU(dalloc(&s.d_keys, cap * B));
    CU(dalloc(&s.d_refined, cap * B));
    CU(dalloc(&s.d_nori, cap * B));
    CU(dalloc(&s.d_angles, cap * B * MAX_ORI));
    CU(dalloc(&s.d_kpoff, cap * B));
    CU(dalloc(&s.d_kps, cap * B));
    CU(dalloc(&s.d_sort, 4 * cap * B));
    CU(dalloc(&s.d_order, cap * B));
    CU(dalloc(&s.d_out_kps, cap * B));
    CU(dalloc(&s.d_out_desc, cap * B * DESC_SIZE));
    return SB200_OK;
}

void free_cand_arrays(Slot& s) {
    cudaFree(s.d_keys); cudaFree(s.d_refined); cudaFree(s.d_nori); cudaFree(s.d_angles); cudaFree(s.d_kpoff);
    cudaFree(s.d_kps); cudaFree(s.d_sort); cudaFree(s.d_order); cudaFree(s.d_out_kps); cudaFree(s.d_out_desc);
    s.d_keys = nullptr; s.d_refined = nullptr; s.d_nori = nullptr; s.d_angles = nullptr; s.d_kpoff = nullptr;
    s.d_kps = nullptr; s.d_sort = nullptr; s.d_order = nullptr; s.d_out_kps = nullptr; s.d_out_desc = nullptr;
}

int alloc_slot(sb200_ctx* ctx, Slot& s) {
    const size_t B = ctx->max_batch, cap = ctx->cap;
    CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    for (auto& q : s.side) CU(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&s.ev_counts, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.ev_upload, cudaEventDisableTiming));
    for (auto& e : s.ev_fork) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : s.ev_join) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s.in_cap = (size_t)ctx->max_w * ctx->max_h * B;
    CU(dalloc(&s.d_in, s.in_cap));
    CU(cudaHostAlloc((void**)&s.h_in, std::max<size_t>(s.in_cap, 1), cudaHostAllocDefault));
    CU(dalloc(&s.d_gauss, (size_t)ctx->gauss_floats_cap * B));
    CU(dalloc(&s.d_mask, (size_t)ctx->mask_words_cap * B));
    CU(dalloc(&s.d_rows, (size_t)ctx->rows_cap * B));
    CU(dalloc(&s.d_rowoff, (size_t)ctx->rows_cap * B));
    CU(dalloc(&s.d_counts, 4 * B + 1));
    CU(dalloc(&s.d_sched, 2 * B + 5));
    CU(cudaHostAlloc((void**)&s.h_counts, (4 * B + 1) * sizeof(uint32_t), cudaHostAllocDefault));
    return alloc_cand_arrays(ctx, s);
}

void free_slot(Slot& s) {
    for (auto& g : s.graphs) cudaGraphExecDestroy(g.exec);
    s.graphs.clear();
    if (s.stream) cudaStreamDestroy(s.stream);
    for (auto q : s.side) if (q) cudaStreamDestroy(q);
    if (s.ev_counts) cudaEventDestroy(s.ev_counts);
    if (s.ev_upload) cudaEventDestroy(s.ev_upload);
    for (auto e : s.ev_fork) if (e) cudaEventDestroy(e);
    for (auto e : s.ev_join) if (e) cudaEventDestroy(e);
    cudaFree(s.d_in); cudaFreeHost(s.h_in); cudaFree(s.d_rgb); cudaFreeHost(s.h_rgb); cudaFree(s.d_gauss); cudaFree(s.d_mask); cudaFree(s.d_rows);
    cudaFree(s.d_rowoff); cudaFree(s.d_counts); cudaFree(s.d_sched); cudaFreeHost(s.h_counts);
    free_cand_arrays(s);
    s = Slot{};
}

int ensure_result_capacity(sb200_ctx* ctx, uint64_t need_kp, uint64_t need_imgs) {
    if (need_imgs + 1 > ctx->offsets_cap) {
        size_t ncap = std::max<size_t>(need_imgs + 1, ctx->offsets_cap * 2 + 16);
        uint64_t* p = nullptr;
        CU(cudaHostAlloc((void**)&p, ncap * sizeof(uint64_t), cudaHostAllocDefault));
        if (ctx->h_offsets) {
            memcpy(p, ctx->h_offsets, ctx->offsets_cap * sizeof(uint64_t));
            cudaFreeHost(ctx->h_offsets);
        }
        ctx->h_offsets = p;
        ctx->offsets_cap = ncap;
    }
    if (need_kp > ctx->res_cap) {
        // pending async copies into the old buffers must land before they move
        for (auto& s : ctx->slot) if (s.stream) CU(cudaStreamSynchronize(s.stream));
        size_t ncap = std::max<size_t>(need_kp, ctx->res_cap * 2 + 4096);
        sb200_keypoint* k = nullptr;
        uint8_t* d = nullptr;
        CU(cudaHostAlloc((void**)&k, ncap * sizeof(sb200_keypoint), cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&d, ncap * SB200_DESC_SIZE, cudaHostAllocDefault));
        if (ctx->h_kps) {
            memcpy(k, ctx->h_kps, ctx->res_n * sizeof(sb200_keypoint));
            memcpy(d, ctx->h_desc, ctx->res_n * SB200_DESC_SIZE);
            cudaFreeHost(ctx->h_kps);
            cudaFreeHost(ctx->h_desc);
        }
        ctx->h_kps = k;
        ctx->h_desc = d;
        ctx->res_cap = ncap;
    }
    return SB200_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int LI>
int encode_one(sb200_ctx* ctx, int slot, int o) {
    using C = TmaCfg<LI>;
    const OctLayout& ol = ctx->L.o[o];
    const cuuint64_t gdim[4] = {(cuuint64_t)ol.w, (cuuint64_t)ol.h, (cuuint64_t)N_LAYERS, (cuuint64_t)ctx->max_batch};
    const cuuint64_t gstr[3] = {(cuuint64_t)ol.pitch * 4, (cuuint64_t)ol.layer_stride * 4,
                                (cuuint64_t)ctx->L.img_floats * 4};
    const cuuint32_t box[4] = {(cuuint32_t)C::BW, (cuuint32_t)C::BAND, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)ctx->encode_fn)(&ctx->tmap[slot][o][LI], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                                 ctx->slot[slot].d_gauss + ol.off, gdim, gstr, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SB200_E_CUDA, "cuTensorMapEncodeTiled failed (%d) octave %d layer %d", (int)r, o, LI);
    return SB200_OK;
}

template <int LI, int FL>
int encode_march(sb200_ctx* ctx, int slot, int o) {
    using C = MarchCfg<LI, FL>;
    const OctLayout& ol = ctx->L.o[o];
    const cuuint64_t gdim[4] = {(cuuint64_t)ol.w, (cuuint64_t)ol.h, (cuuint64_t)N_LAYERS, (cuuint64_t)ctx->max_batch};
    const cuuint64_t gstr[3] = {(cuuint64_t)ol.pitch * 4, (cuuint64_t)ol.layer_stride * 4,
                                (cuuint64_t)ctx->L.img_floats * 4};
    const cuuint32_t box[4] = {(cuuint32_t)C::BW, (cuuint32_t)C::BH, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)ctx->encode_fn)(&ctx->tmap_m[slot][o][LI], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                                 ctx->slot[slot].d_gauss + ol.off, gdim, gstr, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SB200_E_CUDA, "cuTensorMapEncodeTiled failed (%d) octave %d layer %d (march)", (int)r, o, LI);
    return SB200_OK;
}

int encode_extrema(sb200_ctx* ctx, int slot, int o) {
    const OctLayout& ol = ctx->L.o[o];
    const cuuint64_t gdim[4] = {(cuuint64_t)ol.w, (cuuint64_t)ol.h, (cuuint64_t)N_LAYERS, (cuuint64_t)ctx->max_batch};
    const cuuint64_t gstr[3] = {(cuuint64_t)ol.pitch * 4, (cuuint64_t)ol.layer_stride * 4,
                                (cuuint64_t)ctx->L.img_floats * 4};
    const cuuint32_t box[4] = {(cuuint32_t)EXT_BOX_W, (cuuint32_t)EXT_RB, (cuuint32_t)N_LAYERS, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)ctx->encode_fn)(&ctx->tmap_ex[slot][o], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                                 ctx->slot[slot].d_gauss + ol.off, gdim, gstr, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SB200_E_CUDA, "cuTensorMapEncodeTiled failed (%d) octave %d (extrema)", (int)r, o);
    return SB200_OK;
}

int build_tensor_maps(sb200_ctx* ctx) {
    for (int o = 0; o < MAX_OCT; o++) ctx->tmap_ok[o] = false;
    if (!ctx->encode_fn || getenv("SB200_NO_TMA")) return SB200_OK;  // env switch: debugging aid only
    for (int o = 0; o < ctx->L.n_oct; o++) {
        const OctLayout& ol = ctx->L.o[o];
        if (ol.w < TMA_MIN_DIM || ol.h < TMA_MIN_DIM) continue;
        for (int sl = 0; sl < N_SLOTS; sl++) {
            int r;
            if (ctx->flavour == FL_OPENCV) {
                if ((o == 0 && (r = encode_one<0>(ctx, sl, o))) || (r = encode_one<1>(ctx, sl, o)) || (r = encode_one<2>(ctx, sl, o)) || (r = encode_one<3>(ctx, sl, o)) ||
                    (r = encode_one<4>(ctx, sl, o)) || (r = encode_one<5>(ctx, sl, o)) || (r = encode_extrema(ctx, sl, o)) ||
                    (o == 0 && (r = encode_march<0, FL_OPENCV>(ctx, sl, o))) || (r = encode_march<1, FL_OPENCV>(ctx, sl, o)) || (r = encode_march<2, FL_OPENCV>(ctx, sl, o)) ||
                    (r = encode_march<3, FL_OPENCV>(ctx, sl, o)) || (r = encode_march<4, FL_OPENCV>(ctx, sl, o)) || (r = encode_march<5, FL_OPENCV>(ctx, sl, o)))
                    return r;
            } else {
                if ((r = encode_extrema(ctx, sl, o)) ||
                    (o == 0 && (r = encode_march<0, FL_IMAGEPROC>(ctx, sl, o))) || (r = encode_march<1, FL_IMAGEPROC>(ctx, sl, o)) || (r = encode_march<2, FL_IMAGEPROC>(ctx, sl, o)) ||
                    (r = encode_march<3, FL_IMAGEPROC>(ctx, sl, o)) || (r = encode_march<4, FL_IMAGEPROC>(ctx, sl, o)) || (r = encode_march<5, FL_IMAGEPROC>(ctx, sl, o)))
                    return r;
            }
        }
        ctx->tmap_ok[o] = true;
    }
    return SB200_OK;
}

int set_image_size(sb200_ctx* ctx, uint32_t w, uint32_t h) {
    if (w == 0 || h == 0) return fail(ctx, SB200_E_INVALID, "empty image (%ux%u)", w, h);
    if (w > ctx->max_w || h > ctx->max_h)
        return fail(ctx, SB200_E_INVALID, "image %ux%u larger than the context's %ux%u", w, h, ctx->max_w, ctx->max_h);
    if (w != ctx->cur_w || h != ctx->cur_h) {
        PyrLayout L = make_layout(w, h);
        if (L.img_floats > ctx->gauss_floats_cap || L.img_mask_words > ctx->mask_words_cap || L.img_rows > ctx->rows_cap)
            return fail(ctx, SB200_E_INVALID, "internal: layout of %ux%u exceeds the arena", w, h);
        ctx->L = L;
        ctx->cur_w = w;
        ctx->cur_h = h;
        ctx->have_pyramid = ctx->have_single = false;
        int rc = build_tensor_maps(ctx);
        if (rc) return rc;
    }
    return SB200_OK;
}

// ---- launches -----------------------------------------------------------------
// Launch of a kernel of the group pipeline.  With programmatic dependent launch on (default; SB200_PDL=0 turns it off)
// the launch carries cudaLaunchAttributeProgrammaticStreamSerialization: the kernel may be scheduled while its
// predecessor in the stream -- inside a captured graph: on its chain -- is still draining, and waits for it in its first
// statement (pdl_wait, griddepcontrol.wait).  Measured: 2-4 % of a single image's latency (the ~35 dependent launches
// on its critical path), nothing on batches.  No kernel triggers its dependents early -- measured without effect on a
// single image, and a dependent's waiting CTAs would sit on SM slots the other group in flight could use: the trigger
// is the implicit one at the exit of the predecessor's last CTA.
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SB200_PDL"); return !(e && !strcmp(e, "0")); }();
    return on;
}
template <typename... KA, typename... A>
void klaunch(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);   // errors surface at the enqueue functions' cudaGetLastError
}

template <int LI, bool SEED, bool DEC>
int set_blur_attr(sb200_ctx* ctx) {
    CU(cudaFuncSetAttribute(k_blur<LI, SEED, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)BlurCfg<LI>::SMEM));
    return SB200_OK;
}

template <int LI, bool SEED, bool DEC>
void launch_blur(cudaStream_t st, const BlurParams& p, uint32_t n) {
    using C = BlurCfg<LI>;
    dim3 grid((p.w + C::TW - 1) / C::TW, (p.h + C::TH - 1) / C::TH, n);
    klaunch(k_blur<LI, SEED, DEC>, dim3(grid), dim3(C::THREADS), C::SMEM, st, p);
}

template <int LI, bool DEC>
int set_tma_attr(sb200_ctx* ctx) {
    CU(cudaFuncSetAttribute(k_blur_tma<LI, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaCfg<LI>::SMEM));
    CU(cudaFuncSetAttribute(k_blur_tma<LI, DEC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return SB200_OK;
}

template <int LI, bool DEC>
void launch_blur_tma(cudaStream_t st, const CUtensorMap& tm, const BlurParams& p, uint32_t n, int src_layer) {
    using C = TmaCfg<LI>;
    dim3 grid((p.w + C::TW - 1) / C::TW, (p.h + C::TH - 1) / C::TH, n);
    klaunch(k_blur_tma<LI, DEC>, dim3(grid), dim3(C::THREADS), C::SMEM, st, tm, p, src_layer);
}

template <int LI, bool DEC, int FL = FL_OPENCV, int SEEDF = 0>
int set_march_attr(sb200_ctx* ctx) {
    CU(cudaFuncSetAttribute(k_blur_march<LI, DEC, FL, SEEDF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MarchCfg<LI, FL>::SMEM));
    CU(cudaFuncSetAttribute(k_blur_march<LI, DEC, FL, SEEDF>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return SB200_OK;
}

template <int LI>
int set_imageproc_attrs(sb200_ctx* ctx) {
    CU(cudaFuncSetAttribute(k_blur<LI, false, false, FL_IMAGEPROC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)BlurCfg<LI, FL_IMAGEPROC>::SMEM));
    return set_march_attr<LI, false, FL_IMAGEPROC>(ctx);
}

// Launch shape of the marching blur (k_blur_march): the (strip, image) columns of the launch are cut into bands of 32
// output rows.
//   * A launch of less than about 1.5 waves of the machine's resident-CTA slots (the small octaves, single images) deals
//     the bands of all columns out as ONE sequence in equal runs, one run per slot: a CTA marches the rest of one column
//     and the head of the next, every slot gets the same number of bands whatever the image height, strip count and batch
//     size are, and a launch that cannot fill the machine at all gives every CTA a single band.
//   * A launch of several waves cuts every column into k pieces of (almost) equal length, one piece per CTA (see below):
//     one pipeline start per CTA, the CTAs of neighbouring strips walk the same rows at the same time (their halos hit
//     in L2, their rows share DRAM pages), and the hardware's CTA scheduler balances the waves.
constexpr int MARCH_MIN_BANDS = 2;
constexpr int MARCH_TARGET_BANDS = 17;
template <int LI, bool DEC, int FL = FL_OPENCV, int SEEDF = 0>
void launch_blur_march(sb200_ctx* ctx, cudaStream_t st, const CUtensorMap& tm, const BlurParams& p, uint32_t n, int src_layer) {
    using C = MarchCfg<LI, FL>;
    const int strips = (p.w + C::TW - 1) / C::TW;
    const long long nb = (p.h + C::BH - 1) / C::BH;
    const long long total = (long long)strips * n * nb;
    const long long slots = (long long)C::CTAS_PER_SM * ctx->sm_count;
    const long long waves = std::max<long long>(1, (total + MARCH_TARGET_BANDS * slots / 2) / (MARCH_TARGET_BANDS * slots));
    // a launch that cannot fill the machine anyway (single images, the small octaves) is latency bound: one band per CTA
    const long long min_bands = total <= slots ? 1 : std::min<long long>(MARCH_MIN_BANDS, nb);
    long long per = std::max<long long>((total + waves * slots - 1) / (waves * slots), min_bands);
    if (ctx->seg_rows_override > 0) per = std::max(1, ctx->seg_rows_override / C::BH);   // SB200_SEG_ROWS (tests)
    long long grid = (total + per - 1) / per;
    if (ctx->seg_rows_override <= 0 && (waves >= 2 || ctx->pieces_override > 0)) {
        // A launch of several waves: every column is cut into k pieces of (almost) equal length, one piece per CTA, so
        // that a CTA starts its pipeline once and the CTAs of neighbouring strips walk the same rows at the same time.
        // Launch time measured on B200 ~ (columns * k / slots + 0.5) * (nb / k + 1.45) band steps (the half wave is the
        // tail, 1.45 bands the start of a piece): minimal at k = sqrt(0.5 * nb * slots / (1.45 * columns)).
        long long k = (long long)std::llround(std::sqrt(0.5 * (double)nb * (double)slots / (1.45 * (double)strips * n)));
        if (ctx->pieces_override > 0) k = ctx->pieces_override;   // SB200_PIECES (experiments)
        k = std::max<long long>(1, std::min<long long>(k, nb));
        per = -k;
        grid = (long long)strips * n * k;
    }
    klaunch(k_blur_march<LI, DEC, FL, SEEDF>, dim3((unsigned)grid), dim3(C::THREADS + (SEEDF == 2 ? SEED_PRODUCERS : 0)), C::SMEM, st, tm, p, src_layer, (int)per, strips, total);
}

// one blur of the imageproc flavour: marching TMA kernel for octaves that own a tensor map, generic tiles otherwise
template <int LI>
void launch_blur_imageproc(sb200_ctx* ctx, Slot& s, cudaStream_t st, int o, const BlurParams& p, uint32_t n, int src_layer) {
    if (ctx->tmap_ok[o] && ctx->march) {
        launch_blur_march<LI, false, FL_IMAGEPROC>(ctx, st, ctx->tmap_m[s.index][o][LI], p, n, src_layer);
    } else {
        using C = BlurCfg<LI, FL_IMAGEPROC>;
        dim3 grid((p.w + C::TW - 1) / C::TW, (p.h + C::TH - 1) / C::TH, n);
        klaunch(k_blur<LI, false, false, FL_IMAGEPROC>, dim3(grid), dim3(C::THREADS), C::SMEM, st, p);
    }
}

template <bool KEEP_FLAT>
void launch_extrema(sb200_ctx* ctx, cudaStream_t st, int slot, int o, const ExtremaParams& e, uint32_t n) {
    const OctLayout& ol = ctx->L.o[o];
    if (ctx->tmap_ok[o]) {
        dim3 grid(ex_strips(ol.w), (ol.h + EXT_ROWS * EX_WARPS - 1) / (EXT_ROWS * EX_WARPS), n);
        klaunch(k_extrema_tma<KEEP_FLAT>, dim3(grid), dim3(32 * EX_WARPS), EXT_SMEM, st, ctx->tmap_ex[slot][o], e);
    } else {
        dim3 grid(ex_strips(ol.w), (ol.h + EX_ROWS * EX_WARPS - 1) / (EX_ROWS * EX_WARPS), n);
        klaunch(k_extrema<KEEP_FLAT>, dim3(grid), dim3(32 * EX_WARPS), 0, st, e);
    }
}

// first octave of the fused tail (k_tail): every octave from there on fits a shared-memory buffer
int tail_first_octave(const sb200_ctx* ctx) {
    if (!ctx->tail) return ctx->L.n_oct;
    int o = ctx->L.n_oct;
    while (o > 1 && tail_fits(ctx->L.o[o - 1].w, ctx->L.o[o - 1].h)) o--;   // octave 0 always takes the seed path
    return o;
}

// The pyramid in the arithmetic of the crate's default Processing (ImageprocProcessing, src/lib.rs:992-1007): the same
// stages as below with the imageproc tap sets / clamp borders, the image crate's two resizes as kernels of their own,
// the small octaves through the fused tail kernel in its imageproc instantiation.  The extrema scan and everything after it do not
// depend on the flavour.
int enqueue_pyramid_imageproc(sb200_ctx* ctx, Slot& s, uint32_t n, uint32_t w, uint32_t h, uint32_t in_stride,
                              uint64_t in_img_stride, const uint8_t* d_in) {
    const PyrLayout& L = ctx->L;
    cudaStream_t st = s.stream;
    CU(cudaMemsetAsync(s.d_rows, 0, (size_t)L.img_rows * n * sizeof(uint32_t), st));
    auto blur_params = [&](int o, int src_layer, int dst_layer) {
        const OctLayout& ol = L.o[o];
        BlurParams p{};
        p.src = s.d_gauss + ol.off + (long long)src_layer * ol.layer_stride;
        p.dst = s.d_gauss + ol.off + (long long)dst_layer * ol.layer_stride;
        p.img_stride = L.img_floats;
        p.w = ol.w; p.h = ol.h; p.pitch = ol.pitch;
        return p;
    };
    {
        StageScope sc(ctx, st, SB200_STAGE_SEED);
        StageScope fine(ctx, st, SB200_STAGE_COUNT + 0);
        // upsample into layer 5 of octave 0 (free until the last blur of the octave overwrites it), seed blur into layer 0
        UpsampleParams u{};
        u.in = d_in; u.in_img_stride = (long long)in_img_stride;
        u.in_w = (int)w; u.in_h = (int)h; u.in_stride = (int)in_stride;
        u.dst = s.d_gauss + L.o[0].off + 5 * L.o[0].layer_stride;
        u.img_stride = L.img_floats; u.pitch = L.o[0].pitch;
        dim3 grid((((int)w + 1) / 2 + 255) / 256, 2 * h, n);
        klaunch(k_upsample2x_b, dim3(grid), dim3(256), 0, st, u);
        launch_blur_imageproc<0>(ctx, s, st, 0, blur_params(0, 5, 0), n, 5);
        count_launch(ctx, SB200_STAGE_SEED, 2);
    }
    // layers 4 and 5 and the extrema scan of an octave on a side stream, as in enqueue_pyramid
    const bool fork = ctx->fork_octaves && !ctx->profiling;
    uint32_t forked = 0;
    cudaStream_t const st_main = st;
    const int o_tail = tail_first_octave(ctx);
    for (int o = 0; o < o_tail; o++) {
        const OctLayout& ol = L.o[o];
        if (ol.w < 1 || ol.h < 1) continue;
        st = st_main;
        {
            StageScope sc(ctx, st, SB200_STAGE_BLUR);
            for (int l = 1; l < N_LAYERS; l++) {
                if (fork && l == 4) {
                    const int q = o % ctx->n_side;
                    CU(cudaEventRecord(s.ev_fork[o], st_main));
                    CU(cudaStreamWaitEvent(s.side[q], s.ev_fork[o], 0));
                    st = s.side[q];
                    forked |= 1u << q;
                }
                {
                    StageScope fine(ctx, st, SB200_STAGE_COUNT + o * 8 + l);
                    const BlurParams p = blur_params(o, l - 1, l);
                    switch (l) {
                        case 1: launch_blur_imageproc<1>(ctx, s, st, o, p, n, 0); break;
                        case 2: launch_blur_imageproc<2>(ctx, s, st, o, p, n, 1); break;
                        case 3: launch_blur_imageproc<3>(ctx, s, st, o, p, n, 2); break;
                        case 4: launch_blur_imageproc<4>(ctx, s, st, o, p, n, 3); break;
                        default: launch_blur_imageproc<5>(ctx, s, st, o, p, n, 4); break;
                    }
                    count_launch(ctx, SB200_STAGE_BLUR);
                }
                if (l == 3 && o + 1 < L.n_oct && L.o[o + 1].w >= 1 && L.o[o + 1].h >= 1) {
                    DecimateParams d{};
                    d.src = s.d_gauss + ol.off + 3 * ol.layer_stride;
                    d.dst = s.d_gauss + L.o[o + 1].off;
                    d.img_stride = L.img_floats;
                    d.w = ol.w; d.h = ol.h; d.pitch = ol.pitch;
                    d.dw = L.o[o + 1].w; d.dh = L.o[o + 1].h; d.dpitch = L.o[o + 1].pitch;
                    klaunch(k_decimate_b, dim3((d.dw + 255) / 256, d.dh, n), dim3(256), 0, st, d);
                    count_launch(ctx, SB200_STAGE_BLUR);
                }
            }
        }
        if (ol.scanned) {
            StageScope sc(ctx, st, SB200_STAGE_EXTREMA);
            StageScope fine(ctx, st, SB200_STAGE_COUNT + o * 8 + 6);
            ExtremaParams e{};
            e.gauss = s.d_gauss + ol.off;
            e.img_stride = L.img_floats;
            e.layer_stride = ol.layer_stride;
            e.w = ol.w; e.h = ol.h; e.pitch = ol.pitch;
            e.mask = s.d_mask + ol.mask_off;
            e.mask_img_stride = L.img_mask_words;
            e.mask_pitch = ol.mask_pitch;
            e.rows = s.d_rows + ol.row_base;
            e.rows_img_stride = L.img_rows;
            launch_extrema<false>(ctx, st, s.index, o, e, n);
            count_launch(ctx, SB200_STAGE_EXTREMA);
        }
    }
    st = st_main;
    if (o_tail < L.n_oct && L.o[o_tail].w >= 1 && L.o[o_tail].h >= 1) {
        // blurs + decimation + extrema of all the remaining (small) octaves: one CTA per image, one launch
        StageScope sc(ctx, st, SB200_STAGE_BLUR);
        StageScope fine(ctx, st, SB200_STAGE_COUNT + o_tail * 8 + 7);
        TailParams t{};
        t.L = L; t.o_first = o_tail;
        t.gauss = s.d_gauss; t.mask = s.d_mask; t.rows = s.d_rows;
        klaunch(k_tail<false, FL_IMAGEPROC>, dim3(n), dim3(TAIL_THREADS), TAIL_SMEM, st, t);
        count_launch(ctx, SB200_STAGE_BLUR);
    }
    for (int q = 0; q < N_SIDE_MAX; q++) {
        if (!(forked >> q & 1)) continue;
        CU(cudaEventRecord(s.ev_join[q], s.side[q]));
        CU(cudaStreamWaitEvent(st_main, s.ev_join[q], 0));
    }
    CU(cudaGetLastError());
    return SB200_OK;
}

// Gaussian scale space + DoG/extrema masks for the n images staged in slot.d_in
int enqueue_pyramid(sb200_ctx* ctx, Slot& s, uint32_t n, uint32_t w, uint32_t h, uint32_t in_stride,
                    uint64_t in_img_stride, const uint8_t* d_in) {
    if (ctx->flavour == FL_IMAGEPROC) return enqueue_pyramid_imageproc(ctx, s, n, w, h, in_stride, in_img_stride, d_in);
    const PyrLayout& L = ctx->L;
    cudaStream_t st = s.stream;
    CU(cudaMemsetAsync(s.d_rows, 0, (size_t)L.img_rows * n * sizeof(uint32_t), st));
    {
        StageScope sc(ctx, st, SB200_STAGE_SEED);
        StageScope fine(ctx, st, SB200_STAGE_COUNT + 0);
        BlurParams p{};
        p.dst = s.d_gauss + L.o[0].off;
        p.img_stride = L.img_floats;
        p.w = L.o[0].w; p.h = L.o[0].h; p.pitch = L.o[0].pitch;
        p.in = d_in; p.in_img_stride = (long long)in_img_stride;
        p.in_w = (int)w; p.in_h = (int)h; p.in_stride = (int)in_stride;
        if (ctx->tmap_ok[0] && ctx->march && ctx->seed_fused) {
            // the marching seed blur upsamples its input bands itself (k_blur_march<0, .., SEEDF>): one launch and the
            // upsampled image never exists in HBM.  Producer warps (SEEDF = 2) are the default; SB200_SEED=fused takes the
            // in-line variant, SB200_SEED=split the separate upsample kernel below (both slower, kept as cross-checks)
            if (ctx->seed_fused == 2) launch_blur_march<0, false, FL_OPENCV, 2>(ctx, st, ctx->tmap_m[s.index][0][0], p, n, 5);
            else launch_blur_march<0, false, FL_OPENCV, 1>(ctx, st, ctx->tmap_m[s.index][0][0], p, n, 5);
            count_launch(ctx, SB200_STAGE_SEED);
        } else if (ctx->tmap_ok[0]) {
            // upsample into layer 5 of octave 0 (free until the last blur of the octave overwrites it),
            // then the seed blur as a TMA blur from layer 5 into layer 0
            UpsampleParams u{};
            u.in = d_in; u.in_img_stride = (long long)in_img_stride;
            u.in_w = (int)w; u.in_h = (int)h; u.in_stride = (int)in_stride;
            u.dst = s.d_gauss + L.o[0].off + 5 * L.o[0].layer_stride;
            u.img_stride = L.img_floats; u.pitch = L.o[0].pitch;
            dim3 grid((((int)w + 1) / 2 + 255) / 256, (h + 1 + UPS_ROWS - 1) / UPS_ROWS, n);
            klaunch(k_upsample2x, dim3(grid), dim3(256), 0, st, u);
            if (ctx->march) launch_blur_march<0, false>(ctx, st, ctx->tmap_m[s.index][0][0], p, n, 5);
            else launch_blur_tma<0, false>(st, ctx->tmap[s.index][0][0], p, n, 5);
            count_launch(ctx, SB200_STAGE_SEED, 2);
        } else {
            launch_blur<0, true, false>(st, p, n);
            count_launch(ctx, SB200_STAGE_SEED);
        }
    }
    const int o_tail = tail_first_octave(ctx);
    // Octave o+1 starts from the decimated layer 3 of octave o: layers 4 and 5 and the extrema scan of octave o are
    // forked onto the slot's side stream so that the chain of octaves (the critical path of a small batch) does not
    // wait for them.  Stage timing needs one stream.
    const bool fork = ctx->fork_octaves && !ctx->profiling;
    uint32_t forked = 0;   // side streams in use
    cudaStream_t const st_main = st;
    for (int o = 0; o < o_tail; o++) {
        const OctLayout& ol = L.o[o];
        if (ol.w < 1 || ol.h < 1) continue;
        st = st_main;
        {
            StageScope sc(ctx, st, SB200_STAGE_BLUR);
            for (int l = 1; l < N_LAYERS; l++) {
                if (fork && l == 4) {
                    const int q = o % ctx->n_side;
                    CU(cudaEventRecord(s.ev_fork[o], st_main));
                    CU(cudaStreamWaitEvent(s.side[q], s.ev_fork[o], 0));
                    st = s.side[q];
                    forked |= 1u << q;
                }
                BlurParams p{};
                p.src = s.d_gauss + ol.off + (long long)(l - 1) * ol.layer_stride;
                p.dst = s.d_gauss + ol.off + (long long)l * ol.layer_stride;
                p.img_stride = L.img_floats;
                p.w = ol.w; p.h = ol.h; p.pitch = ol.pitch;
                const bool dec = (l == 3) && (o + 1 < L.n_oct) && L.o[o + 1].w >= 1 && L.o[o + 1].h >= 1;
                // the heaviest single launch gets its own event pair (roofline of the dominant kernel)
                StageScope top(ctx, st, (o == 0 && l == 5) ? SB200_STAGE_TOP_BLUR : -1);
                StageScope fine(ctx, st, SB200_STAGE_COUNT + o * 8 + l);
                if (dec) {
                    p.dec = s.d_gauss + L.o[o + 1].off;
                    p.dec_w = L.o[o + 1].w; p.dec_h = L.o[o + 1].h; p.dec_pitch = L.o[o + 1].pitch;
                }
                if (ctx->tmap_ok[o] && ctx->march) {
                    const CUtensorMap& tm = ctx->tmap_m[s.index][o][l];
                    switch (l) {
                        case 1: launch_blur_march<1, false>(ctx, st, tm, p, n, 0); break;
                        case 2: launch_blur_march<2, false>(ctx, st, tm, p, n, 1); break;
                        case 3:
                            if (dec) launch_blur_march<3, true>(ctx, st, tm, p, n, 2);
                            else launch_blur_march<3, false>(ctx, st, tm, p, n, 2);
                            break;
                        case 4: launch_blur_march<4, false>(ctx, st, tm, p, n, 3); break;
                        default: launch_blur_march<5, false>(ctx, st, tm, p, n, 4); break;
                    }
                } else if (ctx->tmap_ok[o]) {
                    const CUtensorMap& tm = ctx->tmap[s.index][o][l];
                    switch (l) {
                        case 1: launch_blur_tma<1, false>(st, tm, p, n, 0); break;
                        case 2: launch_blur_tma<2, false>(st, tm, p, n, 1); break;
                        case 3:
                            if (dec) launch_blur_tma<3, true>(st, tm, p, n, 2);
                            else launch_blur_tma<3, false>(st, tm, p, n, 2);
                            break;
                        case 4: launch_blur_tma<4, false>(st, tm, p, n, 3); break;
                        default: launch_blur_tma<5, false>(st, tm, p, n, 4); break;
                    }
                } else {
                    switch (l) {
                        case 1: launch_blur<1, false, false>(st, p, n); break;
                        case 2: launch_blur<2, false, false>(st, p, n); break;
                        case 3:
                            if (dec) launch_blur<3, false, true>(st, p, n);
                            else launch_blur<3, false, false>(st, p, n);
                            break;
                        case 4: launch_blur<4, false, false>(st, p, n); break;
                        default: launch_blur<5, false, false>(st, p, n); break;
                    }
                }
                count_launch(ctx, SB200_STAGE_BLUR);
                if (o == 0 && l == 5) ctx->stage_launches[SB200_STAGE_TOP_BLUR]++;
            }
        }
        if (ol.scanned) {
            StageScope sc(ctx, st, SB200_STAGE_EXTREMA);
            StageScope fine(ctx, st, SB200_STAGE_COUNT + o * 8 + 6);
            ExtremaParams e{};
            e.gauss = s.d_gauss + ol.off;
            e.img_stride = L.img_floats;
            e.layer_stride = ol.layer_stride;
            e.w = ol.w; e.h = ol.h; e.pitch = ol.pitch;
            e.mask = s.d_mask + ol.mask_off;
            e.mask_img_stride = L.img_mask_words;
            e.mask_pitch = ol.mask_pitch;
            e.rows = s.d_rows + ol.row_base;
            e.rows_img_stride = L.img_rows;
            launch_extrema<false>(ctx, st, s.index, o, e, n);
            count_launch(ctx, SB200_STAGE_EXTREMA);
        }
    }
    st = st_main;
    if (o_tail < L.n_oct && L.o[o_tail].w >= 1 && L.o[o_tail].h >= 1) {
        // blurs + decimation + extrema of all the remaining (small) octaves: one CTA per image, one launch
        StageScope sc(ctx, st, SB200_STAGE_BLUR);
        StageScope fine(ctx, st, SB200_STAGE_COUNT + o_tail * 8 + 7);
        TailParams t{};
        t.L = L; t.o_first = o_tail;
        t.gauss = s.d_gauss; t.mask = s.d_mask; t.rows = s.d_rows;
        klaunch(k_tail<false>, dim3(n), dim3(TAIL_THREADS), TAIL_SMEM, st, t);
        count_launch(ctx, SB200_STAGE_BLUR);
    }
    for (int q = 0; q < N_SIDE_MAX; q++) {   // join: the candidate scan needs every octave's mask
        if (!(forked >> q & 1)) continue;
        CU(cudaEventRecord(s.ev_join[q], s.side[q]));
        CU(cudaStreamWaitEvent(st_main, s.ev_join[q], 0));
    }
    CU(cudaGetLastError());
    return SB200_OK;
}

KpParams make_kp_params(sb200_ctx* ctx, Slot& s) {
    KpParams P{};
    P.L = ctx->L;
    P.gauss = s.d_gauss;
    P.keys = s.d_keys;
    P.cand_count = s.d_counts;
    P.cap = ctx->cap;
    P.refined = s.d_refined;
    P.n_ori = s.d_nori;
    P.angles = s.d_angles;
    P.kp_off = s.d_kpoff;
    P.kp_count = s.d_counts + ctx->max_batch;
    P.kps = s.d_kps;
    P.kcap = ctx->cap;
    return P;
}

// candidates -> keypoints -> descriptors for the n pyramids resident in the slot
int enqueue_detect(sb200_ctx* ctx, Slot& s, uint32_t n, int64_t limit) {
    const PyrLayout& L = ctx->L;
    cudaStream_t st = s.stream;
    const uint32_t B = ctx->max_batch;
    uint32_t* cand_count = s.d_counts;
    uint32_t* kp_count = s.d_counts + B;
    uint32_t* out_count = s.d_counts + 2 * B;
    uint32_t* out_off = s.d_counts + 3 * B;
    KpParams P = make_kp_params(ctx, s);
    const int gx = std::max(1, std::min(4 * ctx->sm_count, (int)(8 * ctx->sm_count / std::max(1u, n)) + 1));
    uint32_t* work = s.d_sched;                // [0]: orientation work list, [1]: descriptor work list
    uint32_t* cand_off = s.d_sched + 4 + B;    // exclusive prefix sum of the per-image candidate counts, [n] = total
    CU(cudaMemsetAsync(work, 0, 4 * sizeof(uint32_t), st));
    {
        StageScope sc(ctx, st, SB200_STAGE_EXTREMA);
        klaunch(k_rowscan, dim3(n), dim3(1024), 0, st, s.d_rows, s.d_rowoff, L.img_rows, cand_count);
        dim3 grid((L.img_rows + 7) / 8, n);
        klaunch(k_compact, dim3(grid), dim3(256), 0, st, L, s.d_mask, s.d_rows, s.d_rowoff, s.d_keys, ctx->cap);
        count_launch(ctx, SB200_STAGE_EXTREMA, 2);
    }
    {
        StageScope sc(ctx, st, SB200_STAGE_REFINE);
        klaunch(k_refine, dim3(gx, n), dim3(128), 0, st, P);
        count_launch(ctx, SB200_STAGE_REFINE);
    }
    {
        StageScope sc(ctx, st, SB200_STAGE_ORIENT);
        klaunch(k_out_offsets, dim3(1), dim3(1024), 0, st, cand_count, ctx->cap, -1LL, (int)n, s.d_sched + 4, cand_off);
        klaunch(k_orient, dim3(ctx->ori_ctas), dim3(32 * ORI_WARPS), 0, st, P, cand_off, (int)n, work);
        klaunch(k_kpscan, dim3(n), dim3(1024), 0, st, P);
        klaunch(k_emit, dim3(gx, n), dim3(256), 0, st, P);
        count_launch(ctx, SB200_STAGE_ORIENT, 4);
    }
    {
        StageScope sc(ctx, st, SB200_STAGE_DESCRIPTOR);
        if (limit >= 0) {
            klaunch(k_sort_response, dim3(n), dim3(1024), 0, st, s.d_kps, kp_count, ctx->cap, s.d_sort, s.d_order);
            count_launch(ctx, SB200_STAGE_DESCRIPTOR);
        }
        klaunch(k_out_offsets, dim3(1), dim3(1024), 0, st, kp_count, ctx->cap, (long long)limit, (int)n, out_count, out_off);
        DescParams D{};
        D.L = L;
        D.gauss = s.d_gauss;
        D.kps = s.d_kps;
        D.kp_count = kp_count;
        D.order = s.d_order;
        D.kcap = ctx->cap;
        D.limit = limit;
        D.out_count = out_count;
        D.out_off = out_off;
        D.out_kps = s.d_out_kps;
        D.out_desc = s.d_out_desc;
        klaunch(k_descriptor, dim3(ctx->desc_ctas), dim3(32 * DESC_WARPS), DESC_SMEM_BYTES, st, D, (int)n, work + 1);
        count_launch(ctx, SB200_STAGE_DESCRIPTOR, 2);
    }
    CU(cudaGetLastError());
    return SB200_OK;
}

// The whole kernel sequence of one group (pyramid + detection) -- about fifty launches whose grids and arguments
// depend only on (n, w, h, input pointer / strides, limit or not) -- is captured once per such shape into a CUDA graph
// and replayed: one launch per group on the host, and no per-kernel launch latency between the many short kernels of
// the small octaves on the device.  Stage timing (profiling) needs events between the kernels: plain launches then.
int run_pipeline(sb200_ctx* ctx, Slot& s, uint32_t n, uint32_t w, uint32_t h, uint32_t stride, uint64_t img_stride,
                 const uint8_t* d_in, int64_t limit, bool detect = true) {
    if (!ctx->use_graphs || ctx->profiling) {
        int rc = enqueue_pyramid(ctx, s, n, w, h, stride, img_stride, d_in);
        if (rc || !detect) return rc;
        return enqueue_detect(ctx, s, n, limit);
    }
    if (limit >= 0) {   // features_limit is a kernel argument (and adds the sort): run directly
        int rc = enqueue_pyramid(ctx, s, n, w, h, stride, img_stride, d_in);
        if (rc || !detect) return rc;
        return enqueue_detect(ctx, s, n, limit);
    }
    Slot::PipeGraph* hit = nullptr;
    for (auto& g : s.graphs)
        if (g.n == n && g.w == w && g.h == h && g.stride == stride && g.img_stride == img_stride && g.d_in == d_in &&
            g.detect == detect)
            hit = &g;
    if (!hit) {
        const uint64_t l0 = ctx->launches;
        uint64_t sl0[SB200_STAGE_COUNT];
        for (int i = 0; i < SB200_STAGE_COUNT; i++) sl0[i] = ctx->stage_launches[i];
        CU(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_pyramid(ctx, s, n, w, h, stride, img_stride, d_in);
        if (!rc && detect) rc = enqueue_detect(ctx, s, n, limit);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(s.stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(ctx, SB200_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        Slot::PipeGraph g{};
        g.n = n; g.w = w; g.h = h; g.stride = stride; g.img_stride = img_stride; g.d_in = d_in; g.detect = detect;
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(ctx, SB200_E_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e));
        g.launches = ctx->launches - l0;
        for (int i = 0; i < SB200_STAGE_COUNT; i++) g.stage_launches[i] = ctx->stage_launches[i] - sl0[i];
        // the capture itself launched nothing: undo the counts, the replay below adds them
        ctx->launches = l0;
        for (int i = 0; i < SB200_STAGE_COUNT; i++) ctx->stage_launches[i] = sl0[i];
        if (s.graphs.size() >= 6) {   // evict the least recently used shape
            size_t k = 0;
            for (size_t i = 1; i < s.graphs.size(); i++) if (s.graphs[i].last_use < s.graphs[k].last_use) k = i;
            cudaGraphExecDestroy(s.graphs[k].exec);
            s.graphs.erase(s.graphs.begin() + k);
        }
        s.graphs.push_back(g);
        hit = &s.graphs.back();
    }
    hit->last_use = ++ctx->graph_clock;
    CU(cudaGraphLaunch(hit->exec, s.stream));
    ctx->launches += hit->launches;
    for (int i = 0; i < SB200_STAGE_COUNT; i++) ctx->stage_launches[i] += hit->stage_launches[i];
    return SB200_OK;
}

// Host-side copies of more than a few megabytes (packing pageable input into the pinned staging buffers, the dense
// gather of the multi-device entry point) are cut into chunks that a few short-lived threads pull from a shared
// counter: one core moves ~10 GB/s, a 32-image 1080p group is 66 MB and the results of 8192 VGA images 1.5 GB.
struct CopyJob { void* dst; const void* src; size_t bytes; };
void parallel_copy(const std::vector<CopyJob>& jobs) {
    constexpr size_t CHUNK = (size_t)2 << 20;
    struct Piece { char* d; const char* s; size_t n; };
    std::vector<Piece> pieces;
    size_t total = 0;
    for (const auto& j : jobs)
        for (size_t o = 0; o < j.bytes; o += CHUNK) {
            pieces.push_back({(char*)j.dst + o, (const char*)j.src + o, std::min(CHUNK, j.bytes - o)});
            total += pieces.back().n;
        }
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::min<size_t>({pieces.size(), (size_t)std::max(1u, std::min(hw ? hw : 4u, 16u)), total / ((size_t)3 << 20) + 1});
    if (const char* e = getenv("SB200_COPY_THREADS")) {   // experiments; 0 skips the copy altogether (timing only)
        if (atoi(e) == 0) return;
        nt = std::min<size_t>(pieces.size(), (size_t)std::max(1, atoi(e)));
    }
    if (nt <= 1) {
        for (auto& q : pieces) memcpy(q.d, q.s, q.n);
        return;
    }
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t i = next.fetch_add(1); i < pieces.size(); i = next.fetch_add(1)) memcpy(pieces[i].d, pieces[i].s, pieces[i].n);
    };
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

bool is_device_accessible_host(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// interleaved 8-bit RGB(A) -> luma, the integer arithmetic of the `image` crate's grayscale() that the reference's
// callers run before sift() (examples/run-sift.rs:8, src/lib.rs:1012): (2126 R + 7152 G + 722 B) / 10000
__global__ void __launch_bounds__(256) k_luma(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, size_t n_px,
                                               int channels) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const uint8_t* p = rgb + i * channels;
    gray[i] = (uint8_t)((2126u * p[0] + 7152u * p[1] + 722u * p[2]) / 10000u);
}

// staging buffers of the colour entry points, allocated on first use
int ensure_rgb(sb200_ctx* ctx, Slot& s, size_t bytes) {
    if (s.rgb_cap >= bytes) return SB200_OK;
    CU(cudaStreamSynchronize(s.stream));
    cudaFree(s.d_rgb); cudaFreeHost(s.h_rgb);
    s.d_rgb = nullptr; s.h_rgb = nullptr; s.rgb_cap = 0;
    const size_t cap = (size_t)ctx->max_w * ctx->max_h * 4 * ctx->max_batch;
    CU(dalloc(&s.d_rgb, cap));
    CU(cudaHostAlloc((void**)&s.h_rgb, cap, cudaHostAllocDefault));
    s.rgb_cap = cap;
    return SB200_OK;
}

// where the pixels of a call come from: packed host / pinned images (gray, RGB, RGBA) or JPEG bitstreams
struct Source {
    const uint8_t* img = nullptr;
    uint32_t stride = 0;
    uint64_t image_stride = 0;
    uint32_t channels = 1;
    const uint8_t* const* jpeg = nullptr;   // non-null: n bitstreams, all of the same frame size
    const size_t* jpeg_len = nullptr;
    bool jpeg_rgb = false;                  // some stream has three components: decode to RGB, then luma
    uint32_t n = 0;                         // images of the call
    uint32_t chunk = 0;                     // bitstreams per decode batch (groups never straddle chunks)
};

int jpeg_stage_ready(sb200_ctx* ctx, size_t bytes) {
    for (auto& js : ctx->jstream)
        if (!js) CU(cudaStreamCreateWithFlags(&js, cudaStreamNonBlocking));
    for (auto& g : ctx->jstage) {
        if (!g.decoded) {
            CU(cudaEventCreateWithFlags(&g.decoded, cudaEventDisableTiming));
            for (auto& e : g.read) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (g.cap < bytes) {
            for (auto& js : ctx->jstream) CU(cudaStreamSynchronize(js));
            for (auto& t : ctx->slot) CU(cudaStreamSynchronize(t.stream));
            cudaFree(g.d); g.d = nullptr; g.cap = 0;
            CU(dalloc(&g.d, bytes));
            g.cap = bytes;
        }
    }
    return SB200_OK;
}

// decodes chunk c of the call's bitstreams into stage buffer c % JPEG_STAGES on decode stream c % JPEG_STREAMS
// (device-side decode: only the compressed bytes cross the host/device boundary)
int decode_chunk(sb200_ctx* ctx, const Source& src, uint32_t c, uint32_t w, uint32_t h) {
    JpegStage& g = ctx->jstage[c % JPEG_STAGES];
    cudaStream_t js = ctx->jstream[c % JPEG_STREAMS];
    const uint32_t first = c * src.chunk, n = std::min(src.chunk, src.n - first);
    const size_t bpp = src.jpeg_rgb ? 3 : 1, img_bytes = (size_t)w * h * bpp;
    for (int k = 0; k < N_SLOTS; k++)   // groups of the chunk that used this buffer before have to be done with it
        if (g.read_valid[k]) { CU(cudaStreamWaitEvent(js, g.read[k], 0)); g.read_valid[k] = false; }
    std::vector<uint8_t*> dst(n);
    for (uint32_t i = 0; i < n; i++) dst[i] = g.d + i * img_bytes;
    std::string err;
    if (!ctx->jpeg.decode((int)(c % JPEG_STAGES), src.jpeg + first, src.jpeg_len + first, n, src.jpeg_rgb, dst.data(),
                          (size_t)w * bpp, js, err))
        return fail(ctx, SB200_E_INVALID, "%s", err.c_str());
    CU(cudaEventRecord(g.decoded, js));
    return SB200_OK;
}

// a group's pixels out of the stage buffer into the slot's input buffer (luma on the way for colour streams)
int fetch_decoded(sb200_ctx* ctx, Slot& s, const Source& src, uint64_t first_img, uint32_t n, uint32_t w, uint32_t h) {
    const uint32_t c = (uint32_t)(first_img / src.chunk);
    if (first_img % src.chunk == 0) {
        // first group of a chunk: keep JPEG_AHEAD chunks decoding beyond this one (at the start of the call, this
        // chunk and the ones after it) while this chunk's groups run
        for (uint32_t d = (c == 0 ? 0 : c + JPEG_AHEAD); d <= c + JPEG_AHEAD; d++) {
            if (d != 0 && (uint64_t)d * src.chunk >= src.n) break;
            int rc = decode_chunk(ctx, src, d, w, h);
            if (rc) return rc;
        }
    }
    JpegStage& g = ctx->jstage[c % JPEG_STAGES];
    const size_t px = (size_t)w * h, bpp = src.jpeg_rgb ? 3 : 1;
    const uint8_t* from = g.d + (first_img - (uint64_t)c * src.chunk) * px * bpp;
    CU(cudaStreamWaitEvent(s.stream, g.decoded, 0));
    if (src.jpeg_rgb) {
        k_luma<<<(unsigned)((px * n + 255) / 256), 256, 0, s.stream>>>(from, s.d_in, px * n, 3);
        ctx->launches++;
    } else {
        CU(cudaMemcpyAsync(s.d_in, from, px * n, cudaMemcpyDeviceToDevice, s.stream));
    }
    const int k = (int)(&s - ctx->slot);
    CU(cudaEventRecord(g.read[k], s.stream));
    g.read_valid[k] = true;
    return SB200_OK;
}

// upload (or decode) + full pipeline + async read-back of the counts for one group; channels = 1 (gray) or 3 / 4
// (RGB / RGBA, converted to luma on the device)
// Pageable input: packs the group's pixels into the slot's pinned staging buffer (host work only).  The buffer is free
// as soon as the previous copy OUT of it is over -- long before the slot's previous group has finished on the GPU --,
// so the batch loop calls this for the next group while it waits for that previous group, and by the time the slot is
// collected the upload can be issued at once, as for pinned input.
int stage_group(sb200_ctx* ctx, Slot& s, const Source& src, uint64_t first_img, uint32_t n, uint32_t w, uint32_t h) {
    if (src.jpeg) return SB200_OK;
    const uint32_t channels = src.channels, stride = src.stride;
    const uint64_t image_stride = src.image_stride;
    const uint8_t* img = src.img + first_img * image_stride;
    if (is_device_accessible_host(img)) return SB200_OK;
    const size_t rowb = (size_t)w * channels;   // bytes per packed row
    if (channels > 1) {
        int rc = ensure_rgb(ctx, s, rowb * h * ctx->max_batch);
        if (rc) return rc;
    }
    uint8_t* const h_up = channels > 1 ? s.h_rgb : s.h_in;
    const bool contiguous = (image_stride == (uint64_t)stride * h);
    CU(cudaEventSynchronize(s.ev_upload));
    std::vector<CopyJob> jobs;
    if (stride == rowb) {
        if (contiguous) jobs.push_back({h_up, img, rowb * h * n});
        else for (uint32_t i = 0; i < n; i++) jobs.push_back({h_up + (size_t)i * rowb * h, img + i * image_stride, rowb * h});
    } else {
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t y = 0; y < h; y++)
                jobs.push_back({h_up + ((size_t)i * h + y) * rowb, img + i * image_stride + (size_t)y * stride, rowb});
    }
    parallel_copy(jobs);
    s.staged = true;
    s.staged_first = first_img;
    return SB200_OK;
}

int launch_group(sb200_ctx* ctx, Slot& s, const Source& src, uint64_t first_img, uint32_t n, uint32_t w, uint32_t h,
                 int64_t limit) {
    cudaStream_t st = s.stream;
    if (src.jpeg) {
        int rc = fetch_decoded(ctx, s, src, first_img, n, w, h);
        if (rc) return rc;
    } else {
        const uint32_t channels = src.channels, stride = src.stride;
        const uint64_t image_stride = src.image_stride;
        const uint8_t* img = src.img + first_img * image_stride;
        const size_t rowb = (size_t)w * channels;   // bytes per packed row
        if (channels > 1) {
            int rc = ensure_rgb(ctx, s, rowb * h * ctx->max_batch);
            if (rc) return rc;
        }
        uint8_t* const d_up = channels > 1 ? s.d_rgb : s.d_in;
        uint8_t* const h_up = channels > 1 ? s.h_rgb : s.h_in;
        const bool contiguous = (image_stride == (uint64_t)stride * h);
        if (is_device_accessible_host(img)) {
            if (contiguous && stride == rowb) {   // densely packed batch: one linear copy
                CU(cudaMemcpyAsync(d_up, img, rowb * h * n, cudaMemcpyHostToDevice, st));
            } else if (contiguous) {
                CU(cudaMemcpy2DAsync(d_up, rowb, img, stride, rowb, (size_t)h * n, cudaMemcpyHostToDevice, st));
            } else {
                for (uint32_t i = 0; i < n; i++)
                    CU(cudaMemcpy2DAsync(d_up + (size_t)i * rowb * h, rowb, img + i * image_stride, stride, rowb, h,
                                         cudaMemcpyHostToDevice, st));
            }
        } else {
            // pageable memory: packed into the slot's pinned staging buffer (stage_group, normally already done by
            // the batch loop while the slot's previous group was still running), then one async copy
            const size_t total = rowb * h * n;
            if (!(s.staged && s.staged_first == first_img) && stride == rowb && contiguous && total <= ((size_t)12 << 20)) {
                // a small densely packed group nobody staged ahead (a single image, the first group of a call): packed
                // and uploaded piece by piece on this thread, so that the copy engine moves piece k while piece k + 1 is
                // packed -- the upload costs the packing time alone
                CU(cudaEventSynchronize(s.ev_upload));
                constexpr size_t PIECE = (size_t)512 << 10;
                for (size_t o = 0; o < total; o += PIECE) {
                    const size_t nb = std::min(PIECE, total - o);
                    memcpy(h_up + o, img + o, nb);
                    CU(cudaMemcpyAsync(d_up + o, h_up + o, nb, cudaMemcpyHostToDevice, st));
                }
            } else {
                if (!(s.staged && s.staged_first == first_img)) {
                    int rc = stage_group(ctx, s, src, first_img, n, w, h);
                    if (rc) return rc;
                }
                CU(cudaMemcpyAsync(d_up, h_up, total, cudaMemcpyHostToDevice, st));
            }
            s.staged = false;
            CU(cudaEventRecord(s.ev_upload, st));
        }
        if (channels > 1) {
            const size_t n_px = (size_t)w * h * n;
            k_luma<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>(s.d_rgb, s.d_in, n_px, (int)channels);
            ctx->launches++;
        }
    }
    int rc = run_pipeline(ctx, s, n, w, h, w, (uint64_t)w * h, s.d_in, limit);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s.h_counts, s.d_counts, (4 * (size_t)ctx->max_batch + 1) * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(s.ev_counts, st));
    s.n_imgs = n;
    s.first_img = first_img;
    s.limit = limit;
    s.busy = true;
    return SB200_OK;
}

// waits for the group's counts, appends its results to the context's result arrays
// An image with more candidates / keypoints than the context was sized for (exact ramps and the like: every pixel
// ties) does not fail the call: the per-candidate arrays of all slots are re-allocated larger and the detection
// stages are run again on the pyramids and extrema masks still resident in the slots.
int grow_capacity(sb200_ctx* ctx, uint32_t need) {
    for (auto& t : ctx->slot) {
        CU(cudaStreamSynchronize(t.stream));
        for (auto q : t.side) CU(cudaStreamSynchronize(q));
    }
    const uint64_t nc = std::max<uint64_t>((uint64_t)need + need / 4 + 1024, (uint64_t)ctx->cap * 2);
    if (nc > 0x3fffffffull) return fail(ctx, SB200_E_CAPACITY, "%u candidates per image exceed the supported maximum", need);
    ctx->cap = (uint32_t)nc;
    for (auto& t : ctx->slot) {
        for (auto& g : t.graphs) cudaGraphExecDestroy(g.exec);   // the graphs hold the old pointers and capacity
        t.graphs.clear();
        free_cand_arrays(t);
        int rc = alloc_cand_arrays(ctx, t);
        if (rc) return rc;
    }
    return SB200_OK;
}

int collect_group(sb200_ctx* ctx, Slot& s) {
    if (!s.busy) return SB200_OK;
    const uint32_t B = ctx->max_batch, n = s.n_imgs;
    const uint32_t* cand = s.h_counts;
    const uint32_t* kpc = s.h_counts + B;
    const uint32_t* outoff = s.h_counts + 3 * B;
    for (;;) {
        CU(cudaEventSynchronize(s.ev_counts));
        uint32_t need = 0;
        for (uint32_t i = 0; i < n; i++) need = std::max(need, std::max(cand[i], kpc[i]));
        if (need <= ctx->cap) break;
        int rc = grow_capacity(ctx, need);
        if (rc) return rc;
        for (auto& t : ctx->slot) {   // every group in flight lost its candidate arrays: detect again
            if (!t.busy) continue;
            rc = enqueue_detect(ctx, t, t.n_imgs, t.limit);
            if (rc) return rc;
            CU(cudaMemcpyAsync(t.h_counts, t.d_counts, (4 * (size_t)B + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, t.stream));
            CU(cudaEventRecord(t.ev_counts, t.stream));
        }
    }
    s.busy = false;
    const uint64_t total = outoff[n];
    int rc = ensure_result_capacity(ctx, ctx->res_n + total, s.first_img + n);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; i++) ctx->h_offsets[s.first_img + i] = ctx->res_n + outoff[i];
    ctx->h_offsets[s.first_img + n] = ctx->res_n + total;
    if (total) {
        CU(cudaMemcpyAsync(ctx->h_kps + ctx->res_n, s.d_out_kps, total * sizeof(sb200_keypoint),
                           cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(ctx->h_desc + ctx->res_n * SB200_DESC_SIZE, s.d_out_desc, total * SB200_DESC_SIZE,
                           cudaMemcpyDeviceToHost, s.stream));
    }
    ctx->res_n += total;
    return SB200_OK;
}

// OpenCV-style post-filters of the host result (cv::SIFT::detectAndCompute: KeyPointsFilter::removeDuplicatedSorted,
// then retainBest when nfeatures > 0; the comparison target of benches/sift.rs:99-113).  Host code on the pinned
// result arrays, per image, a few threads over the images of a batch:
//   * remove_duplicates: keypoints sorted by (x asc, y asc, size desc, angle asc, response desc, natural index) and
//     those equal to their predecessor in (x, y, size, angle) dropped -- the crate keeps such duplicates (two initial
//     extrema converging on one refined point), OpenCV does not;
//   * retain_best n: when more than n remain, every keypoint whose response is >= the n-th largest response is kept
//     (ties at the boundary survive, as in OpenCV), in the order the previous step left.
void apply_postfilter(sb200_ctx* ctx, uint32_t n_images) {
    if (!ctx->pf_dedup && ctx->pf_retain < 0) return;
    if (ctx->res_n == 0) return;
    std::vector<std::vector<uint32_t>> keep(n_images);
    auto one = [&](uint32_t im) {
        const uint64_t a = ctx->h_offsets[im], b = ctx->h_offsets[im + 1];
        const sb200_keypoint* k = ctx->h_kps + a;
        std::vector<uint32_t>& idx = keep[im];
        idx.resize(b - a);
        for (uint32_t i = 0; i < idx.size(); i++) idx[i] = i;
        if (ctx->pf_dedup && idx.size() > 1) {
            std::sort(idx.begin(), idx.end(), [k](uint32_t i, uint32_t j) {
                const sb200_keypoint &p = k[i], &q = k[j];
                if (p.x != q.x) return p.x < q.x;
                if (p.y != q.y) return p.y < q.y;
                if (p.size != q.size) return p.size > q.size;
                if (p.angle != q.angle) return p.angle < q.angle;
                if (p.response != q.response) return p.response > q.response;
                return i < j;
            });
            size_t m = 0;
            for (size_t j = 1; j < idx.size(); j++) {
                const sb200_keypoint &p = k[idx[m]], &q = k[idx[j]];
                if (p.x != q.x || p.y != q.y || p.size != q.size || p.angle != q.angle) idx[++m] = idx[j];
            }
            idx.resize(m + 1);
        }
        if (ctx->pf_retain >= 0 && idx.size() > (uint64_t)ctx->pf_retain) {
            if (ctx->pf_retain == 0) { idx.clear(); return; }
            std::vector<float> r(idx.size());
            for (size_t j = 0; j < idx.size(); j++) r[j] = k[idx[j]].response;
            std::nth_element(r.begin(), r.begin() + (ctx->pf_retain - 1), r.end(), std::greater<float>());
            const float thr = r[ctx->pf_retain - 1];
            size_t m = 0;
            for (size_t j = 0; j < idx.size(); j++) if (k[idx[j]].response >= thr) idx[m++] = idx[j];
            idx.resize(m);
        }
    };
    {
        std::atomic<uint32_t> next{0};
        auto work = [&]() { for (uint32_t im = next.fetch_add(1); im < n_images; im = next.fetch_add(1)) one(im); };
        const unsigned hw = std::thread::hardware_concurrency();
        const uint32_t nt = std::min<uint32_t>({n_images, std::max(1u, std::min(hw ? hw : 4u, 16u)), (uint32_t)(ctx->res_n / 4096 + 1)});
        std::vector<std::thread> th;
        for (uint32_t t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    // compaction towards the front, image by image (a destination never runs ahead of its source)
    std::vector<sb200_keypoint> tk;
    std::vector<uint8_t> td;
    uint64_t pos = 0;
    for (uint32_t im = 0; im < n_images; im++) {
        const uint64_t a = ctx->h_offsets[im], b = ctx->h_offsets[im + 1];
        tk.assign(ctx->h_kps + a, ctx->h_kps + b);
        td.assign(ctx->h_desc + a * SB200_DESC_SIZE, ctx->h_desc + b * SB200_DESC_SIZE);
        ctx->h_offsets[im] = pos;
        for (uint32_t i : keep[im]) {
            ctx->h_kps[pos] = tk[i];
            memcpy(ctx->h_desc + pos * SB200_DESC_SIZE, td.data() + (size_t)i * SB200_DESC_SIZE, SB200_DESC_SIZE);
            pos++;
        }
    }
    ctx->h_offsets[n_images] = pos;
    ctx->res_n = pos;
}

void fill_result(sb200_ctx* ctx, uint32_t n_images, sb200_result* out) {
    apply_postfilter(ctx, n_images);
    out->n = ctx->res_n;
    out->n_images = n_images;
    out->offsets = ctx->h_offsets;
    out->keypoints = ctx->h_kps;
    out->descriptors = ctx->h_desc;
}

int finish_all(sb200_ctx* ctx) {
    for (auto& s : ctx->slot) CU(cudaStreamSynchronize(s.stream));
    if (ctx->profiling) drain_stage_events(ctx);
    if (ctx->err_pending) {   // a device-pointer compute_descriptors ran since the last check
        ctx->err_pending = false;
        uint32_t e = 0;
        CU(cudaMemcpy(&e, ctx->d_err, sizeof e, cudaMemcpyDeviceToHost));
        if (e) {
            CU(cudaMemset(ctx->d_err, 0, sizeof e));
            return fail(ctx, SB200_E_INVALID, "compute_descriptors: a keypoint's scale is not in (0, %.1f] (descriptor window "
                        "radius above %d); its descriptor was zeroed", 12.0, DESC_MAX_RADIUS);
        }
    }
    return SB200_OK;
}

}  // namespace

// ============================================================================
// C ABI
// ============================================================================
extern "C" {

const char* sb200_status_string(int status) {
    switch (status) {
        case SB200_OK: return "ok";
        case SB200_E_INVALID: return "invalid argument";
        case SB200_E_CUDA: return "CUDA error";
        case SB200_E_CAPACITY: return "capacity exceeded";
        case SB200_E_STATE: return "invalid state";
        case SB200_E_UNSUPPORTED: return "unsupported on this machine";
        default: return "unknown status";
    }
}

const char* sb200_last_error(const sb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

const char* sb200_stage_name(uint32_t stage) { return stage < SB200_STAGE_COUNT ? kStageNames[stage] : "?"; }

int sb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return -SB200_E_CUDA;
    }
    return n;
}

int sb200_create(int device, uint32_t max_w, uint32_t max_h, uint32_t max_batch, uint32_t max_kp, sb200_ctx** out) {
    if (!out) return SB200_E_INVALID;
    *out = nullptr;
    if (max_w == 0 || max_h == 0 || max_batch == 0 || max_w > SB200_MAX_DIM || max_h > SB200_MAX_DIM ||
        max_batch > 65535)
        return SB200_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return SB200_E_CUDA;
    }
    sb200_ctx* ctx = new sb200_ctx();
    ctx->device = device;
    ctx->max_w = max_w; ctx->max_h = max_h; ctx->max_batch = max_batch;
    ctx->cap = max_kp ? max_kp : std::max<uint32_t>(16384u, (uint32_t)((uint64_t)max_w * max_h / 8));
    auto bail = [&](int rc) {
        std::string e = ctx->err;
        fprintf(stderr, "sb200_create: %s\n", e.c_str());
        sb200_destroy(ctx);
        return rc;
    };
    int rc = SB200_OK;
    auto body = [&]() -> int {
        CU(cudaSetDevice(device));
        cudaDeviceProp prop{};
        CU(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        // arenas sized for the largest image; a smaller image can have at most as many of everything
        PyrLayout L = make_layout(max_w, max_h);
        ctx->gauss_floats_cap = L.img_floats;
        ctx->mask_words_cap = L.img_mask_words;
        ctx->rows_cap = L.img_rows;
        // Gaussian taps -> constant memory
        float taps[N_LAYERS][32];
        memset(taps, 0, sizeof taps);
        for (int l = 0; l < N_LAYERS; l++) {
            int ks = 0;
            gaussian_taps(l == 0 ? seed_sigma() : octave_sigma(l), taps[l], ks);
            if (ks != 2 * blur_radius(l) + 1)
                return fail(ctx, SB200_E_INVALID, "internal: tap count %d of kernel %d", ks, l);
        }
        CU(cudaMemcpyToSymbol(c_taps, taps, sizeof taps));
        {
            float2 taps2[N_LAYERS][32];
            for (int l = 0; l < N_LAYERS; l++)
                for (int i = 0; i < 32; i++) taps2[l][i] = make_float2(taps[l][i], taps[l][i]);
            CU(cudaMemcpyToSymbol(c_taps2, taps2, sizeof taps2));
        }
        {   // the imageproc flavour's tap sets
            float tb[N_LAYERS][32];
            float2 tb2[N_LAYERS][32];
            memset(tb, 0, sizeof tb);
            for (int l = 0; l < N_LAYERS; l++) {
                int ks = 0;
                imageproc_taps(l == 0 ? seed_sigma() : octave_sigma(l), tb[l], ks);
                if (ks != 2 * blur_radius(l, FL_IMAGEPROC) + 1)
                    return fail(ctx, SB200_E_INVALID, "internal: tap count %d of imageproc kernel %d", ks, l);
                for (int i = 0; i < 32; i++) tb2[l][i] = make_float2(tb[l][i], tb[l][i]);
            }
            CU(cudaMemcpyToSymbol(c_taps_b, tb, sizeof tb));
            CU(cudaMemcpyToSymbol(c_taps2_b, tb2, sizeof tb2));
        }
        int r;
        if ((r = set_imageproc_attrs<0>(ctx)) || (r = set_imageproc_attrs<1>(ctx)) || (r = set_imageproc_attrs<2>(ctx)) ||
            (r = set_imageproc_attrs<3>(ctx)) || (r = set_imageproc_attrs<4>(ctx)) || (r = set_imageproc_attrs<5>(ctx)))
            return r;
        if ((r = set_blur_attr<0, true, false>(ctx))) return r;
        if ((r = set_blur_attr<1, false, false>(ctx))) return r;
        if ((r = set_blur_attr<2, false, false>(ctx))) return r;
        if ((r = set_blur_attr<3, false, false>(ctx))) return r;
        if ((r = set_blur_attr<3, false, true>(ctx))) return r;
        if ((r = set_blur_attr<4, false, false>(ctx))) return r;
        if ((r = set_blur_attr<5, false, false>(ctx))) return r;
        if ((r = set_tma_attr<0, false>(ctx)) || (r = set_tma_attr<1, false>(ctx)) || (r = set_tma_attr<2, false>(ctx)) || (r = set_tma_attr<3, false>(ctx)) ||
            (r = set_tma_attr<3, true>(ctx)) || (r = set_tma_attr<4, false>(ctx)) || (r = set_tma_attr<5, false>(ctx)))
            return r;
        if ((r = set_march_attr<0, false, FL_OPENCV, 1>(ctx)) || (r = set_march_attr<0, false, FL_OPENCV, 2>(ctx)) ||
            (r = set_march_attr<0, false>(ctx)) || (r = set_march_attr<1, false>(ctx)) || (r = set_march_attr<2, false>(ctx)) ||
            (r = set_march_attr<3, false>(ctx)) || (r = set_march_attr<3, true>(ctx)) || (r = set_march_attr<4, false>(ctx)) ||
            (r = set_march_attr<5, false>(ctx)))
            return r;
        {
            const char* e = getenv("SB200_BLUR");
            ctx->march = !(e && !strcmp(e, "tile"));
            const char* fk = getenv("SB200_FORK");
            ctx->fork_octaves = !(fk && !strcmp(fk, "0"));
            if (const char* sd = getenv("SB200_SIDES")) ctx->n_side = std::min(std::max(atoi(sd), 1), N_SIDE_MAX);
            const char* gr = getenv("SB200_GRAPHS");
            ctx->use_graphs = !(gr && !strcmp(gr, "0"));
            const char* jc = getenv("SB200_JPEG_CHUNK");
            if (jc && atoi(jc) >= 1) ctx->jpeg_chunk = (uint32_t)atoi(jc);
            const char* tl = getenv("SB200_TAIL");
            ctx->tail = !(tl && !strcmp(tl, "0"));
            CU(cudaFuncSetAttribute(k_tail<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_SMEM));
            CU(cudaFuncSetAttribute(k_tail<false, FL_IMAGEPROC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_SMEM));
            const char* sr = getenv("SB200_SEG_ROWS");
            if (sr && atoi(sr) >= 32) ctx->seg_rows_override = atoi(sr) / 32 * 32;
            const char* sd = getenv("SB200_SEED");
            ctx->seed_fused = !sd ? 2 : !strcmp(sd, "fused") ? 1 : !strcmp(sd, "split") ? 0 : 2;
            const char* pc = getenv("SB200_PIECES");
            if (pc && atoi(pc) >= 1) ctx->pieces_override = atoi(pc);
        }
        CU(cudaFuncSetAttribute(k_match_nn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
        CU(cudaFuncSetAttribute(k_extrema_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXT_SMEM));
        CU(cudaFuncSetAttribute(k_extrema_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXT_SMEM));
        CU(cudaFuncSetAttribute(k_extrema_tma<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CU(cudaFuncSetAttribute(k_descriptor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DESC_SMEM_BYTES));
        CU(cudaFuncSetAttribute(k_descriptor_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DESC_SMEM_BYTES));
        {
            int per_sm = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_descriptor, 32 * DESC_WARPS, DESC_SMEM_BYTES));
            // SB200_DESC_CTAS / SB200_ORI_CTAS (per SM) leave room on every SM for the other slot's kernels (experiments)
            const char* dc = getenv("SB200_DESC_CTAS");
            if (dc && atoi(dc) >= 1) per_sm = std::min(per_sm, atoi(dc));
            ctx->desc_ctas = std::max(1, per_sm) * ctx->sm_count;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_orient, 32 * ORI_WARPS, 0));
            const char* oc = getenv("SB200_ORI_CTAS");
            if (oc && atoi(oc) >= 1) per_sm = std::min(per_sm, atoi(oc));
            ctx->ori_ctas = std::max(1, per_sm) * ctx->sm_count;
        }
        {
            cudaDriverEntryPointQueryResult q;
            void* fn = nullptr;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
                q == cudaDriverEntryPointSuccess)
                ctx->encode_fn = fn;
            else
                return fail(ctx, SB200_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        }
        for (int i = 0; i < N_SLOTS; i++) {
            ctx->slot[i].index = i;
            if ((r = alloc_slot(ctx, ctx->slot[i]))) return r;
        }
        CU(cudaEventCreate(&ctx->t0));
        CU(cudaEventCreate(&ctx->t1));
        CU(dalloc(&ctx->d_err, 1));
        CU(cudaMemset(ctx->d_err, 0, sizeof(uint32_t)));
        return SB200_OK;
    };
    rc = body();
    if (rc) return bail(rc);
    *out = ctx;
    return SB200_OK;
}

void sb200_destroy(sb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto& s : ctx->slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        free_slot(s);
    }
    for (auto js : ctx->jstream) if (js) cudaStreamSynchronize(js);
    ctx->jpeg.release();   // before the stream and the buffers its states may still reference
    for (auto& g : ctx->jstage) {
        cudaFree(g.d);
        if (g.decoded) cudaEventDestroy(g.decoded);
        for (auto e : g.read) if (e) cudaEventDestroy(e);
    }
    for (auto js : ctx->jstream) if (js) cudaStreamDestroy(js);
    for (auto& p : ctx->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    cudaFreeHost(ctx->h_offsets); cudaFreeHost(ctx->h_kps); cudaFreeHost(ctx->h_desc);
    cudaFree(ctx->d_dimg); cudaFree(ctx->d_dkps); cudaFree(ctx->d_ddesc); cudaFree(ctx->d_flush); cudaFree(ctx->d_err);
    for (int i = 0; i < 2; i++) { cudaFree(ctx->d_mdesc[i]); cudaFree(ctx->d_mnorm[i]); cudaFree(ctx->d_mnbp[i]); cudaFree(ctx->d_mbest[i]); }
    cudaFree(ctx->d_mout); cudaFree(ctx->d_mcount);
    delete ctx;
}

int sb200_set_processing(sb200_ctx* ctx, int processing) {
    if (!ctx) return SB200_E_INVALID;
    if (processing != SB200_PROCESSING_OPENCV && processing != SB200_PROCESSING_IMAGEPROC)
        return fail(ctx, SB200_E_INVALID, "unknown processing flavour %d", processing);
    CU(cudaSetDevice(ctx->device));
    if (processing == ctx->flavour) return SB200_OK;
    for (auto& t : ctx->slot) {
        CU(cudaStreamSynchronize(t.stream));
        for (auto q : t.side) CU(cudaStreamSynchronize(q));
        for (auto& g : t.graphs) cudaGraphExecDestroy(g.exec);   // captured with the other flavour's kernels
        t.graphs.clear();
        t.busy = false;
    }
    ctx->flavour = processing == SB200_PROCESSING_IMAGEPROC ? FL_IMAGEPROC : FL_OPENCV;
    ctx->cur_w = ctx->cur_h = 0;   // the marching blur's TMA boxes depend on the tap radius: rebuilt at the next call
    ctx->have_pyramid = ctx->have_single = false;
    return SB200_OK;
}

int sb200_set_postfilter(sb200_ctx* ctx, int remove_duplicates, int64_t retain_best) {
    if (!ctx) return SB200_E_INVALID;
    ctx->pf_dedup = remove_duplicates != 0;
    ctx->pf_retain = retain_best < 0 ? -1 : retain_best;
    return SB200_OK;
}

int sb200_get_processing(const sb200_ctx* ctx) {
    return !ctx ? -SB200_E_INVALID : ctx->flavour == FL_IMAGEPROC ? SB200_PROCESSING_IMAGEPROC : SB200_PROCESSING_OPENCV;
}

static int extract_batch_impl(sb200_ctx* ctx, const Source& src, uint32_t n, uint32_t w, uint32_t h, int64_t features_limit,
                              sb200_result* out) {
    if (!ctx) return SB200_E_INVALID;
    if (!out || n == 0) return fail(ctx, SB200_E_INVALID, "bad arguments to extract_batch");
    if (!src.jpeg && (!src.img || (uint64_t)src.stride < (uint64_t)w * src.channels ||
                      (src.channels != 1 && src.channels != 3 && src.channels != 4)))
        return fail(ctx, SB200_E_INVALID, "bad arguments to extract_batch");
    CU(cudaSetDevice(ctx->device));
    int rc = set_image_size(ctx, w, h);
    if (rc) return rc;
    for (auto& s : ctx->slot) {
        CU(cudaStreamSynchronize(s.stream));  // results of the previous call are released
        s.busy = false;                        // groups a failed call left behind are dropped, not collected
        s.staged = false;
    }
    ctx->res_n = 0;
    ctx->have_pyramid = false;
    rc = ensure_result_capacity(ctx, 0, n);
    if (rc) return rc;
    const uint32_t B = ctx->max_batch;
    // group sizes: full groups of B images, but a long batch starts and ends with a quarter-size group so that the
    // upload of the first group and the download of the last one -- the two copies nothing can overlap -- are short
    std::vector<uint32_t> sizes;
    if (src.jpeg) {
        // bitstreams: groups never straddle the chunks they are decoded in
        for (uint64_t c0 = 0; c0 < n; c0 += src.chunk)
            for (uint64_t rem = std::min<uint64_t>(src.chunk, n - c0); rem;) {
                const uint32_t c = (uint32_t)std::min<uint64_t>(B, rem);
                sizes.push_back(c); rem -= c;
            }
    } else {
        uint64_t rem = n;
        const uint32_t q = std::max<uint32_t>(1, B / 4);
        if (n >= 2ull * B && B >= 4) {
            sizes.push_back(q); rem -= q;
            while (rem > (uint64_t)B + q) { sizes.push_back(B); rem -= B; }
            if (rem > q) { sizes.push_back((uint32_t)(rem - q)); rem = q; }
            sizes.push_back((uint32_t)rem);
        } else {
            while (rem) { const uint32_t c = (uint32_t)std::min<uint64_t>(B, rem); sizes.push_back(c); rem -= c; }
        }
    }
    uint32_t g = 0;
    uint64_t first = 0;
    for (; g < sizes.size(); first += sizes[g], g++) {
        Slot& s = ctx->slot[g % N_SLOTS];
        // groups complete in launch order; the slot's previous group was collected before this launch (below)
        rc = launch_group(ctx, s, src, first, sizes[g], w, h, features_limit);
        if (rc) return rc;
        // pageable input: pack the next group now, while the GPU works (its staging buffer is free long before its slot)
        if (g + 1 < sizes.size()) {
            rc = stage_group(ctx, ctx->slot[(g + 1) % N_SLOTS], src, first + sizes[g], sizes[g + 1], w, h);
            if (rc) return rc;
        }
        // the slot the NEXT group will use must be free again: collect the group that ran in it
        if (g + 1 >= (uint32_t)N_SLOTS) {
            rc = collect_group(ctx, ctx->slot[(g + 1) % N_SLOTS]);
            if (rc) return rc;
        }
    }
    // remaining groups in launch order (collect_group is a no-op for idle slots)
    for (uint32_t k = 0; k < (uint32_t)N_SLOTS; k++) {
        rc = collect_group(ctx, ctx->slot[(g + k) % N_SLOTS]);
        if (rc) return rc;
    }
    rc = finish_all(ctx);
    if (rc) return rc;
    ctx->have_single = (n == 1);
    ctx->have_pyramid = (n == 1);
    ctx->last_slot = 0;
    ctx->last_limit = features_limit;
    fill_result(ctx, n, out);
    return SB200_OK;
}

int sb200_extract_batch(sb200_ctx* ctx, const uint8_t* gray, uint32_t n, uint32_t w, uint32_t h, uint32_t stride,
                        uint64_t image_stride, int64_t features_limit, sb200_result* out) {
    Source src;
    src.img = gray; src.stride = stride; src.image_stride = image_stride; src.channels = 1;
    return extract_batch_impl(ctx, src, n, w, h, features_limit, out);
}

int sb200_extract_batch_rgb(sb200_ctx* ctx, const uint8_t* rgb, uint32_t n, uint32_t w, uint32_t h, uint32_t stride,
                            uint64_t image_stride, uint32_t channels, int64_t features_limit, sb200_result* out) {
    if (ctx && channels != 3 && channels != 4) return fail(ctx, SB200_E_INVALID, "channels must be 3 (RGB) or 4 (RGBA)");
    Source src;
    src.img = rgb; src.stride = stride; src.image_stride = image_stride; src.channels = channels;
    return extract_batch_impl(ctx, src, n, w, h, features_limit, out);
}

// ---- JPEG input ----
static_assert(sizeof(size_t) == sizeof(uint64_t), "lengths are passed to nvJPEG as size_t");

static int jpeg_ready(sb200_ctx* ctx) {
    std::string err;
    if (!ctx->jpeg.init(JPEG_STAGES, err)) return fail(ctx, SB200_E_UNSUPPORTED, "%s", err.c_str());
    return SB200_OK;
}

int sb200_jpeg_info(sb200_ctx* ctx, const uint8_t* jpeg, uint64_t length, uint32_t* w, uint32_t* h, uint32_t* components) {
    if (!ctx) return SB200_E_INVALID;
    if (!jpeg || length == 0) return fail(ctx, SB200_E_INVALID, "bad arguments to jpeg_info");
    CU(cudaSetDevice(ctx->device));
    int rc = jpeg_ready(ctx);
    if (rc) return rc;
    JpegDecoder::Info inf;
    std::string err;
    if (!ctx->jpeg.info(jpeg, (size_t)length, inf, err)) return fail(ctx, SB200_E_INVALID, "%s", err.c_str());
    if (w) *w = inf.w;
    if (h) *h = inf.h;
    if (components) *components = inf.components;
    return SB200_OK;
}

// headers of all streams: one frame size for the whole call; *rgb = some stream carries colour
static int jpeg_batch_info(sb200_ctx* ctx, const uint8_t* const* jpegs, const uint64_t* lengths, uint32_t n, uint32_t* w,
                           uint32_t* h, bool* rgb) {
    *rgb = false;
    for (uint32_t i = 0; i < n; i++) {
        if (!jpegs[i] || lengths[i] == 0) return fail(ctx, SB200_E_INVALID, "JPEG %u is empty", i);
        JpegDecoder::Info inf;
        std::string err;
        if (!ctx->jpeg.info(jpegs[i], (size_t)lengths[i], inf, err)) return fail(ctx, SB200_E_INVALID, "JPEG %u: %s", i, err.c_str());
        if (i == 0) { *w = inf.w; *h = inf.h; }
        else if (inf.w != *w || inf.h != *h)
            return fail(ctx, SB200_E_INVALID, "JPEG %u is %ux%u, the batch is %ux%u (one size per call)", i, inf.w, inf.h, *w, *h);
        if (inf.components == 3) *rgb = true;
    }
    return SB200_OK;
}

int sb200_extract_batch_jpeg(sb200_ctx* ctx, const uint8_t* const* jpegs, const uint64_t* lengths, uint32_t n,
                             int64_t features_limit, sb200_result* out) {
    if (!ctx) return SB200_E_INVALID;
    if (!jpegs || !lengths || !out || n == 0) return fail(ctx, SB200_E_INVALID, "bad arguments to extract_batch_jpeg");
    CU(cudaSetDevice(ctx->device));
    int rc = jpeg_ready(ctx);
    if (rc) return rc;
    uint32_t w = 0, h = 0;
    Source src;
    rc = jpeg_batch_info(ctx, jpegs, lengths, n, &w, &h, &src.jpeg_rgb);
    if (rc) return rc;
    if (w > ctx->max_w || h > ctx->max_h)
        return fail(ctx, SB200_E_INVALID, "image %ux%u larger than the context's %ux%u", w, h, ctx->max_w, ctx->max_h);
    src.jpeg = jpegs;
    src.jpeg_len = reinterpret_cast<const size_t*>(lengths);
    src.n = n;
    // chunks of about jpeg_chunk streams, equal in size (a short last chunk would fall back to Huffman decoding on the
    // host inside nvJPEG), a multiple of the group size
    const uint32_t B = ctx->max_batch;
    const uint32_t n_chunks = std::max<uint32_t>(1, n / std::max<uint32_t>(ctx->jpeg_chunk, 1));
    src.chunk = std::max<uint32_t>(1, ((n + n_chunks - 1) / n_chunks + B - 1) / B * B);
    rc = jpeg_stage_ready(ctx, (size_t)std::min(src.chunk, n) * w * h * (src.jpeg_rgb ? 3 : 1));
    if (rc) return rc;
    return extract_batch_impl(ctx, src, n, w, h, features_limit, out);
}

int sb200_decode_jpeg_luma(sb200_ctx* ctx, const uint8_t* jpeg, uint64_t length, uint8_t* gray, uint64_t capacity) {
    if (!ctx) return SB200_E_INVALID;
    if (!jpeg || length == 0 || !gray) return fail(ctx, SB200_E_INVALID, "bad arguments to decode_jpeg_luma");
    CU(cudaSetDevice(ctx->device));
    int rc = jpeg_ready(ctx);
    if (rc) return rc;
    uint32_t w = 0, h = 0;
    bool rgb = false;
    rc = jpeg_batch_info(ctx, &jpeg, &length, 1, &w, &h, &rgb);
    if (rc) return rc;
    if ((uint64_t)w * h > capacity) return fail(ctx, SB200_E_CAPACITY, "output buffer holds %llu bytes, the image has %llu",
                                                (unsigned long long)capacity, (unsigned long long)w * h);
    if (w > ctx->max_w || h > ctx->max_h)
        return fail(ctx, SB200_E_INVALID, "image %ux%u larger than the context's %ux%u", w, h, ctx->max_w, ctx->max_h);
    Slot& s = ctx->slot[0];
    CU(cudaStreamSynchronize(s.stream));
    Source src;
    const size_t len = (size_t)length;
    src.jpeg = &jpeg; src.jpeg_len = &len; src.jpeg_rgb = rgb; src.n = 1; src.chunk = 1;
    rc = jpeg_stage_ready(ctx, (size_t)w * h * (rgb ? 3 : 1));
    if (rc) return rc;
    rc = fetch_decoded(ctx, s, src, 0, 1, w, h);
    if (rc) return rc;
    CU(cudaMemcpyAsync(gray, s.d_in, (size_t)w * h, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    ctx->have_pyramid = ctx->have_single = false;   // the slot's input buffer no longer matches its pyramid
    return SB200_OK;
}

const char* sb200_jpeg_backend(const sb200_ctx* ctx) { return ctx ? ctx->jpeg.backend() : "none"; }

int sb200_rgb_to_luma(sb200_ctx* ctx, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t stride, uint32_t channels,
                      uint8_t* gray) {
    if (!ctx) return SB200_E_INVALID;
    if (!rgb || !gray || w == 0 || h == 0 || (channels != 3 && channels != 4) || (uint64_t)stride < (uint64_t)w * channels)
        return fail(ctx, SB200_E_INVALID, "bad arguments to rgb_to_luma");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    uint8_t *d_rgb = nullptr, *d_gray = nullptr;
    const size_t rowb = (size_t)w * channels, n_px = (size_t)w * h;
    cudaError_t e = cudaMalloc((void**)&d_rgb, rowb * h);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_gray, n_px);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_rgb, rowb, rgb, stride, rowb, h, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_luma<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>(d_rgb, d_gray, n_px, (int)channels);
        ctx->launches++;
        e = cudaMemcpyAsync(gray, d_gray, n_px, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_rgb); cudaFree(d_gray);
    if (e != cudaSuccess) return fail(ctx, SB200_E_CUDA, "rgb_to_luma: %s", cudaGetErrorString(e));
    return SB200_OK;
}

int sb200_extract(sb200_ctx* ctx, const uint8_t* gray, uint32_t w, uint32_t h, uint32_t stride, int64_t features_limit,
                  sb200_result* out) {
    return sb200_extract_batch(ctx, gray, 1, w, h, stride, (uint64_t)stride * h, features_limit, out);
}

int sb200_extract_batch_device(sb200_ctx* ctx, const uint8_t* d_gray, uint32_t n, uint32_t w, uint32_t h,
                               uint32_t stride, uint64_t image_stride, int64_t features_limit) {
    if (!ctx) return SB200_E_INVALID;
    if (!d_gray || n == 0 || n > ctx->max_batch || stride < w)
        return fail(ctx, SB200_E_INVALID, "bad arguments to extract_batch_device");
    CU(cudaSetDevice(ctx->device));
    int rc = set_image_size(ctx, w, h);
    if (rc) return rc;
    // consecutive calls alternate between the two slots so that the tail of one batch
    // (small octaves, keypoint kernels) overlaps the head of the next
    Slot& s = ctx->slot[ctx->dev_rr];
    ctx->last_slot = ctx->dev_rr;
    ctx->dev_rr = (ctx->dev_rr + 1) % N_SLOTS;
    if (ctx->use_graphs && !ctx->profiling && features_limit < 0 && d_gray != s.d_in) {
        // pack the group into the slot's own input buffer (1 B/px against ~370 B/px of pyramid traffic) so that the
        // captured graph does not depend on the caller's pointer and strides
        if (image_stride == (uint64_t)stride * h) {
            CU(cudaMemcpy2DAsync(s.d_in, w, d_gray, stride, w, (size_t)h * n, cudaMemcpyDeviceToDevice, s.stream));
        } else {
            for (uint32_t i = 0; i < n; i++)
                CU(cudaMemcpy2DAsync(s.d_in + (size_t)i * w * h, w, d_gray + i * image_stride, stride, w, h,
                                     cudaMemcpyDeviceToDevice, s.stream));
        }
        d_gray = s.d_in; stride = w; image_stride = (uint64_t)w * h;
    }
    rc = run_pipeline(ctx, s, n, w, h, stride, image_stride, d_gray, features_limit);
    if (rc) return rc;
    s.n_imgs = n;
    ctx->have_single = ctx->have_pyramid = (n == 1);
    ctx->last_limit = features_limit;
    return SB200_OK;
}

int sb200_pyramid_batch_device(sb200_ctx* ctx, const uint8_t* d_gray, uint32_t n, uint32_t w, uint32_t h, uint32_t stride,
                               uint64_t image_stride) {
    if (!ctx) return SB200_E_INVALID;
    if (!d_gray || n == 0 || n > ctx->max_batch || stride < w)
        return fail(ctx, SB200_E_INVALID, "bad arguments to pyramid_batch_device");
    CU(cudaSetDevice(ctx->device));
    int rc = set_image_size(ctx, w, h);
    if (rc) return rc;
    Slot& s = ctx->slot[ctx->dev_rr];
    ctx->last_slot = ctx->dev_rr;
    ctx->dev_rr = (ctx->dev_rr + 1) % N_SLOTS;
    if (ctx->use_graphs && !ctx->profiling && d_gray != s.d_in) {
        if (image_stride == (uint64_t)stride * h) {
            CU(cudaMemcpy2DAsync(s.d_in, w, d_gray, stride, w, (size_t)h * n, cudaMemcpyDeviceToDevice, s.stream));
        } else {
            for (uint32_t i = 0; i < n; i++)
                CU(cudaMemcpy2DAsync(s.d_in + (size_t)i * w * h, w, d_gray + i * image_stride, stride, w, h,
                                     cudaMemcpyDeviceToDevice, s.stream));
        }
        d_gray = s.d_in; stride = w; image_stride = (uint64_t)w * h;
    }
    rc = run_pipeline(ctx, s, n, w, h, stride, image_stride, d_gray, -1, false);
    if (rc) return rc;
    s.n_imgs = n;
    ctx->have_pyramid = (n == 1);
    ctx->have_single = false;
    return SB200_OK;
}

int sb200_device_result(sb200_ctx* ctx, uint32_t* counts, uint32_t n, const sb200_keypoint** d_keypoints,
                        const uint8_t** d_descriptors, uint32_t* capacity_per_image) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[ctx->last_slot];
    if (n > ctx->max_batch) return fail(ctx, SB200_E_INVALID, "n exceeds max_batch");
    CU(cudaStreamSynchronize(s.stream));
    if (counts && n)
        CU(cudaMemcpy(counts, s.d_counts + 2 * ctx->max_batch, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    // the kernels clamp at the per-image capacity: a denser image is truncated, and that has to be said (the host
    // entry points grow the capacity and run detection again; here the caller owns the schedule)
    const uint32_t m = std::min(n ? n : s.n_imgs, ctx->max_batch);
    if (m) {
        std::vector<uint32_t> cc(2 * (size_t)ctx->max_batch);
        CU(cudaMemcpy(cc.data(), s.d_counts, cc.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        uint32_t need = 0;
        for (uint32_t i = 0; i < m; i++) need = std::max(need, std::max(cc[i], cc[ctx->max_batch + i]));
        if (need > ctx->cap) {
            if (capacity_per_image) *capacity_per_image = ctx->cap;
            return fail(ctx, SB200_E_CAPACITY, "an image has %u candidates / keypoints, the context holds %u per image: the "
                        "device-resident result is truncated (create the context with a larger max_keypoints_per_image)",
                        need, ctx->cap);
        }
    }
    if (d_keypoints) *d_keypoints = reinterpret_cast<const sb200_keypoint*>(s.d_out_kps);
    if (d_descriptors) *d_descriptors = s.d_out_desc;
    if (capacity_per_image) *capacity_per_image = ctx->cap;
    return SB200_OK;
}

int sb200_sync(sb200_ctx* ctx) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    return finish_all(ctx);
}

int sb200_precompute(sb200_ctx* ctx, const uint8_t* gray, uint32_t w, uint32_t h, uint32_t stride) {
    if (!ctx) return SB200_E_INVALID;
    if (!gray || stride < w) return fail(ctx, SB200_E_INVALID, "bad arguments to precompute");
    CU(cudaSetDevice(ctx->device));
    int rc = set_image_size(ctx, w, h);
    if (rc) return rc;
    Slot& s = ctx->slot[0];
    for (auto& sl : ctx->slot) CU(cudaStreamSynchronize(sl.stream));
    for (uint32_t y = 0; y < h; y++) memcpy(s.h_in + (size_t)y * w, gray + (size_t)y * stride, w);
    CU(cudaMemcpyAsync(s.d_in, s.h_in, (size_t)w * h, cudaMemcpyHostToDevice, s.stream));
    rc = enqueue_pyramid(ctx, s, 1, w, h, w, (uint64_t)w * h, s.d_in);
    if (rc) return rc;
    rc = finish_all(ctx);
    if (rc) return rc;
    ctx->have_pyramid = true;
    ctx->have_single = false;
    ctx->last_slot = 0;
    return SB200_OK;
}

int sb200_extract_precomputed(sb200_ctx* ctx, int64_t features_limit, sb200_result* out) {
    if (!ctx) return SB200_E_INVALID;
    if (!out) return fail(ctx, SB200_E_INVALID, "null result");
    if (!ctx->have_pyramid) return fail(ctx, SB200_E_STATE, "no pyramid resident: call sb200_precompute first");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[ctx->last_slot];
    for (auto& sl : ctx->slot) {
        CU(cudaStreamSynchronize(sl.stream));
        sl.busy = false;
    }
    ctx->res_n = 0;
    int rc = ensure_result_capacity(ctx, 0, 1);
    if (rc) return rc;
    // the masks and row counters of the resident pyramid are intact: rerun detection only
    rc = enqueue_detect(ctx, s, 1, features_limit);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s.h_counts, s.d_counts, (4 * (size_t)ctx->max_batch + 1) * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, s.stream));
    CU(cudaEventRecord(s.ev_counts, s.stream));
    s.n_imgs = 1; s.first_img = 0; s.limit = features_limit; s.busy = true;
    rc = collect_group(ctx, s);
    if (rc) return rc;
    rc = finish_all(ctx);
    if (rc) return rc;
    ctx->have_single = true;
    ctx->last_limit = features_limit;
    fill_result(ctx, 1, out);
    return SB200_OK;
}

int sb200_pyramid_info(sb200_ctx* ctx, uint32_t* n_octaves, uint32_t* widths, uint32_t* heights, uint32_t cap) {
    if (!ctx) return SB200_E_INVALID;
    if (!ctx->have_pyramid) return fail(ctx, SB200_E_STATE, "no pyramid resident");
    if (n_octaves) *n_octaves = (uint32_t)ctx->L.n_oct;
    for (uint32_t o = 0; o < cap && o < (uint32_t)ctx->L.n_oct; o++) {
        if (widths) widths[o] = (uint32_t)ctx->L.o[o].w;
        if (heights) heights[o] = (uint32_t)ctx->L.o[o].h;
    }
    return SB200_OK;
}

int sb200_pyramid_layer(sb200_ctx* ctx, uint32_t octave, uint32_t layer, float* out) {
    if (!ctx) return SB200_E_INVALID;
    if (!ctx->have_pyramid) return fail(ctx, SB200_E_STATE, "no pyramid resident");
    if (!out || octave >= (uint32_t)ctx->L.n_oct || layer >= N_LAYERS) return fail(ctx, SB200_E_INVALID, "bad layer");
    CU(cudaSetDevice(ctx->device));
    const OctLayout& ol = ctx->L.o[octave];
    if (ol.w < 1 || ol.h < 1) return SB200_OK;
    Slot& s = ctx->slot[ctx->last_slot];
    CU(cudaMemcpy2DAsync(out, (size_t)ol.w * 4, s.d_gauss + ol.off + (long long)layer * ol.layer_stride,
                         (size_t)ol.pitch * 4, (size_t)ol.w * 4, ol.h, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return SB200_OK;
}

int sb200_pyramid_dog(sb200_ctx* ctx, uint32_t octave, uint32_t layer, float* out) {
    if (!ctx) return SB200_E_INVALID;
    if (!ctx->have_pyramid) return fail(ctx, SB200_E_STATE, "no pyramid resident");
    if (!out || octave >= (uint32_t)ctx->L.n_oct || layer >= N_DOG) return fail(ctx, SB200_E_INVALID, "bad layer");
    CU(cudaSetDevice(ctx->device));
    const OctLayout& ol = ctx->L.o[octave];
    if (ol.w < 1 || ol.h < 1) return SB200_OK;
    Slot& s = ctx->slot[ctx->last_slot];
    float* tmp = nullptr;
    CU(cudaMalloc((void**)&tmp, (size_t)ol.w * ol.h * 4));
    const float* a = s.d_gauss + ol.off + (long long)layer * ol.layer_stride;
    k_dog_layer<<<dim3((ol.w + 127) / 128, ol.h), 128, 0, s.stream>>>(a, a + ol.layer_stride, tmp, ol.w, ol.h, ol.pitch);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, tmp, (size_t)ol.w * ol.h * 4, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(ctx, SB200_E_CUDA, "dog download: %s", cudaGetErrorString(e));
    return SB200_OK;
}

int sb200_last_candidates(sb200_ctx* ctx, sb200_candidate* out, uint64_t cap, uint64_t* n) {
    // Parity view: the reference's complete candidate list (src/lib.rs:324-332), including the exactly-flat
    // neighbourhoods the pipeline drops early.  Recomputed from the resident pyramid into temporary buffers.
    if (!ctx) return SB200_E_INVALID;
    if (!ctx->have_single && !ctx->have_pyramid) return fail(ctx, SB200_E_STATE, "no single-image run resident");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[ctx->last_slot];
    const PyrLayout& L = ctx->L;
    cudaStream_t st = s.stream;
    CU(cudaStreamSynchronize(st));
    uint32_t *d_mask = nullptr, *d_rows = nullptr, *d_rowoff = nullptr, *d_cnt = nullptr;
    CandKey* d_keys = nullptr;
    int rc = SB200_OK;
    auto cleanup = [&]() { cudaFree(d_mask); cudaFree(d_rows); cudaFree(d_rowoff); cudaFree(d_cnt); cudaFree(d_keys); };
    auto body = [&]() -> int {
        CU(dalloc(&d_mask, (size_t)L.img_mask_words));
        CU(dalloc(&d_rows, (size_t)L.img_rows));
        CU(dalloc(&d_rowoff, (size_t)L.img_rows));
        CU(dalloc(&d_cnt, 1));
        CU(cudaMemsetAsync(d_rows, 0, (size_t)L.img_rows * 4, st));
        for (int o = 0; o < L.n_oct; o++) {
            const OctLayout& ol = L.o[o];
            if (!ol.scanned) continue;
            ExtremaParams e{};
            e.gauss = s.d_gauss + ol.off; e.img_stride = L.img_floats; e.layer_stride = ol.layer_stride;
            e.w = ol.w; e.h = ol.h; e.pitch = ol.pitch;
            e.mask = d_mask + ol.mask_off; e.mask_img_stride = L.img_mask_words; e.mask_pitch = ol.mask_pitch;
            e.rows = d_rows + ol.row_base; e.rows_img_stride = L.img_rows;
            launch_extrema<true>(ctx, st, s.index, o, e, 1);
            ctx->launches++;
        }
        k_rowscan<<<1, 1024, 0, st>>>(d_rows, d_rowoff, L.img_rows, d_cnt);
        ctx->launches++;
        uint32_t cnt = 0;
        CU(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (n) *n = cnt;
        const uint32_t m = (uint32_t)std::min<uint64_t>(cnt, cap);
        if (out && m) {
            CU(dalloc(&d_keys, (size_t)cnt));
            dim3 grid((L.img_rows + 7) / 8, 1);
            k_compact<<<grid, 256, 0, st>>>(L, d_mask, d_rows, d_rowoff, d_keys, cnt);
            ctx->launches++;
            std::vector<CandKey> keys(m);
            CU(cudaMemcpyAsync(keys.data(), d_keys, (size_t)m * sizeof(CandKey), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < m; i++) {
                int o, sc, y, x;
                unpack_key(keys[i], o, sc, y, x);
                out[i] = {o, sc, y, x};
            }
        }
        return SB200_OK;
    };
    rc = body();
    cleanup();
    return rc;
}

int sb200_last_sift_keypoints(sb200_ctx* ctx, sb200_sift_keypoint* out, uint64_t cap, uint64_t* n) {
    if (!ctx) return SB200_E_INVALID;
    if (!ctx->have_single) return fail(ctx, SB200_E_STATE, "no single-image run resident");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[ctx->last_slot];
    CU(cudaStreamSynchronize(s.stream));
    uint32_t cnt = 0;
    CU(cudaMemcpy(&cnt, s.d_counts + ctx->max_batch, 4, cudaMemcpyDeviceToHost));
    if (n) *n = cnt;
    const uint32_t m = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(cnt, ctx->cap), cap);
    if (out && m) {
        std::vector<DevKeyPoint> k(m);
        CU(cudaMemcpy(k.data(), s.d_kps, (size_t)m * sizeof(DevKeyPoint), cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < m; i++)
            out[i] = {k[i].x, k[i].y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].scale};
    }
    return SB200_OK;
}

int sb200_compute_descriptors_device(sb200_ctx* ctx, const float* d_img, uint32_t w, uint32_t h, uint32_t stride,
                                     const sb200_desc_in* d_kps, uint64_t n, uint8_t* d_out) {
    if (!ctx) return SB200_E_INVALID;
    if (!d_img || w < 3 || h < 3 || stride < w || (n && (!d_kps || !d_out)))
        return fail(ctx, SB200_E_INVALID, "bad arguments to compute_descriptors");
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return SB200_OK;
    cudaStream_t st = ctx->slot[0].stream;
    StageScope sc(ctx, st, SB200_STAGE_DESCRIPTOR);
    const int grid = (int)std::min<uint64_t>((n + DESC_WARPS - 1) / DESC_WARPS, (uint64_t)ctx->sm_count * 16);
    k_descriptor_list<<<grid, 32 * DESC_WARPS, DESC_SMEM_BYTES, st>>>(d_img, (int)w, (int)h, (int)stride,
                                                        reinterpret_cast<const DescIn*>(d_kps), n, d_out, ctx->d_err);
    ctx->err_pending = true;
    count_launch(ctx, SB200_STAGE_DESCRIPTOR);
    CU(cudaGetLastError());
    return SB200_OK;
}

int sb200_compute_descriptors(sb200_ctx* ctx, const float* img, uint32_t w, uint32_t h, uint32_t stride,
                              const sb200_desc_in* kps, uint64_t n, uint8_t* out) {
    if (!ctx) return SB200_E_INVALID;
    if (!img || w < 3 || h < 3 || stride < w || (n && (!kps || !out)))
        return fail(ctx, SB200_E_INVALID, "bad arguments to compute_descriptors");
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return SB200_OK;
    for (uint64_t i = 0; i < n; i++)   // the crate has no limit; the kernel's row table holds windows of radius <= 127
        if (!(kps[i].scale > 0.f) || descriptor_radius(kps[i].scale) > DESC_MAX_RADIUS)
            return fail(ctx, SB200_E_INVALID, "compute_descriptors: keypoint %llu has scale %g, supported range is (0, 12.0] "
                        "(descriptor window radius <= %d)", (unsigned long long)i, (double)kps[i].scale, DESC_MAX_RADIUS);
    cudaStream_t st = ctx->slot[0].stream;
    const size_t px = (size_t)w * h;
    if (px > ctx->dimg_cap) {
        cudaFree(ctx->d_dimg);
        ctx->d_dimg = nullptr;
        ctx->dimg_cap = 0;
        CU(dalloc(&ctx->d_dimg, px));
        ctx->dimg_cap = px;
    }
    if (n > ctx->dkps_cap) {
        cudaFree(ctx->d_dkps); cudaFree(ctx->d_ddesc);
        ctx->d_dkps = nullptr; ctx->d_ddesc = nullptr;
        ctx->dkps_cap = 0;
        CU(dalloc(&ctx->d_dkps, n));
        CU(dalloc(&ctx->d_ddesc, n * DESC_SIZE));
        ctx->dkps_cap = n;
    }
    CU(cudaMemcpy2DAsync(ctx->d_dimg, (size_t)w * 4, img, (size_t)stride * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_dkps, kps, n * sizeof(sb200_desc_in), cudaMemcpyHostToDevice, st));
    int rc = sb200_compute_descriptors_device(ctx, ctx->d_dimg, w, h, w, reinterpret_cast<const sb200_desc_in*>(ctx->d_dkps),
                                              n, ctx->d_ddesc);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, ctx->d_ddesc, n * DESC_SIZE, cudaMemcpyDeviceToHost, st));
    return finish_all(ctx);
}

// contiguous shards of ceil(n / n_ctx) images, one host thread per context; parts[d] = context d's own result
static int run_shards(sb200_ctx* const* ctxs, uint32_t n_ctx, const uint8_t* gray, uint32_t n, uint32_t w, uint32_t h,
                      uint32_t stride, uint64_t image_stride, int64_t features_limit, sb200_result* parts, uint64_t* first) {
    sb200_ctx* ctx = ctxs[0];
    const uint32_t per = (n + n_ctx - 1) / n_ctx;
    std::vector<int> rcs(n_ctx, SB200_OK);
    std::vector<std::thread> th;
    for (uint32_t d = 0; d < n_ctx; d++) parts[d] = sb200_result{0, 0, nullptr, nullptr, nullptr};
    for (uint32_t d = 0; d < n_ctx; d++) {
        const uint64_t f = std::min<uint64_t>((uint64_t)d * per, n);
        if (f >= n) break;
        const uint32_t cnt = (uint32_t)std::min<uint64_t>(per, n - f);
        auto shard = [=, &rcs]() {
            const auto t0 = std::chrono::steady_clock::now();
            rcs[d] = sb200_extract_batch(ctxs[d], gray + f * image_stride, cnt, w, h, stride, image_stride, features_limit,
                                         &parts[d]);
            ctxs[d]->last_shard_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        };
        if (d + 1 < n_ctx && f + cnt < n) th.emplace_back(shard);
        else { shard(); break; }   // the last non-empty shard runs on the calling thread
    }
    if (first) {
        for (uint32_t d = 0; d <= n_ctx; d++) first[d] = std::min<uint64_t>((uint64_t)d * per, n);
    }
    for (auto& t : th) t.join();
    for (uint32_t d = 0; d < n_ctx; d++)
        if (rcs[d]) return fail(ctx, rcs[d], "device %u: %s", d, sb200_last_error(ctxs[d]));
    return SB200_OK;
}

int sb200_extract_batch_multi_parts(sb200_ctx* const* ctxs, uint32_t n_ctx, const uint8_t* gray, uint32_t n, uint32_t w,
                                    uint32_t h, uint32_t stride, uint64_t image_stride, int64_t features_limit,
                                    sb200_result* parts, uint64_t* first_image) {
    if (!ctxs || n_ctx == 0 || !ctxs[0]) return SB200_E_INVALID;
    sb200_ctx* ctx = ctxs[0];
    for (uint32_t d = 1; d < n_ctx; d++)
        if (!ctxs[d]) return fail(ctx, SB200_E_INVALID, "context %u of %u is null", d, n_ctx);
    if (!gray || !parts || n == 0) return fail(ctx, SB200_E_INVALID, "bad arguments to extract_batch_multi_parts");
    return run_shards(ctxs, n_ctx, gray, n, w, h, stride, image_stride, features_limit, parts, first_image);
}

int sb200_extract_batch_multi(sb200_ctx* const* ctxs, uint32_t n_ctx, const uint8_t* gray, uint32_t n, uint32_t w,
                              uint32_t h, uint32_t stride, uint64_t image_stride, int64_t features_limit,
                              sb200_result* out) {
    if (!ctxs || n_ctx == 0 || !ctxs[0]) return SB200_E_INVALID;
    sb200_ctx* ctx = ctxs[0];
    for (uint32_t d = 1; d < n_ctx; d++)
        if (!ctxs[d]) return fail(ctx, SB200_E_INVALID, "context %u of %u is null", d, n_ctx);
    if (!gray || !out || n == 0) return fail(ctx, SB200_E_INVALID, "bad arguments to extract_batch_multi");
    ctx->last_gather_ms = 0.0;
    if (n_ctx == 1) return sb200_extract_batch(ctx, gray, n, w, h, stride, image_stride, features_limit, out);
    std::vector<sb200_result> parts(n_ctx);
    int rc = run_shards(ctxs, n_ctx, gray, n, w, h, stride, image_stride, features_limit, parts.data(), nullptr);
    if (rc) return rc;
    // Dense gather in image order into ctxs[0]'s result arrays (its own part already sits at their front and
    // ensure_result_capacity preserves it).  The other parts are copied by a few host threads in parallel; callers
    // that can take one part per device (sb200_extract_batch_multi_parts) skip this copy altogether.
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t total = 0;
    for (auto& p : parts) total += p.n;
    CU(cudaSetDevice(ctx->device));
    rc = ensure_result_capacity(ctx, total, n);
    if (rc) return rc;
    uint64_t pos = parts[0].n, img = parts[0].n_images;
    std::vector<CopyJob> jobs;
    for (uint32_t d = 1; d < n_ctx; d++) {
        const sb200_result& p = parts[d];
        if (p.n_images == 0) continue;
        jobs.push_back({ctx->h_kps + pos, p.keypoints, p.n * sizeof(sb200_keypoint)});
        jobs.push_back({ctx->h_desc + pos * SB200_DESC_SIZE, p.descriptors, p.n * SB200_DESC_SIZE});
        for (uint32_t i = 0; i < p.n_images; i++) ctx->h_offsets[img + i] = pos + p.offsets[i];
        pos += p.n;
        img += p.n_images;
    }
    parallel_copy(jobs);
    ctx->h_offsets[img] = pos;
    ctx->res_n = pos;
    fill_result(ctx, n, out);
    ctx->last_gather_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return SB200_OK;
}

double sb200_last_gather_ms(const sb200_ctx* ctx) { return ctx ? ctx->last_gather_ms : 0.0; }
double sb200_last_shard_ms(const sb200_ctx* ctx) { return ctx ? ctx->last_shard_ms : 0.0; }

// ---- descriptor matching (examples/sift-match.rs:30-35: BFMatcher(NORM_L2, crossCheck = true)) -----------------
namespace {
int encode_desc_map(sb200_ctx* ctx, CUtensorMap* tm, const uint8_t* d, uint64_t n, uint32_t box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)DESC_SIZE, (cuuint64_t)n};
    const cuuint64_t gstr[1] = {(cuuint64_t)DESC_SIZE};
    const cuuint32_t box[2] = {(cuuint32_t)DESC_SIZE, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)ctx->encode_fn)(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)d, gdim, gstr, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SB200_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for a descriptor matrix", (int)r);
    return SB200_OK;
}
}  // namespace

int sb200_match_descriptors_device(sb200_ctx* ctx, const uint8_t* d_query, uint64_t n_query, const uint8_t* d_train, uint64_t n_train,
                       sb200_dmatch* out, uint64_t cap, uint64_t* n_out) {
    if (!ctx) return SB200_E_INVALID;
    if (!n_out || (cap && !out) || (n_query && !d_query) || (n_train && !d_train) || n_query > 0x7fffffffull ||
        n_train > 0x7fffffffull)
        return fail(ctx, SB200_E_INVALID, "bad arguments to match");
    CU(cudaSetDevice(ctx->device));
    *n_out = 0;
    if (n_query == 0 || n_train == 0) return SB200_OK;
    if (((uintptr_t)d_query | (uintptr_t)d_train) & 15) return fail(ctx, SB200_E_INVALID, "descriptor matrices must be 16-byte aligned");
    cudaStream_t st = ctx->slot[0].stream;
    const uint8_t* d[2] = {d_query, d_train};
    const uint64_t n[2] = {n_query, n_train};
    for (int i = 0; i < 2; i++) {
        const size_t pad = (size_t)((n[i] + MT_N - 1) / MT_N) * MT_N;
        if (pad > ctx->m_cap[i]) {
            cudaFree(ctx->d_mnorm[i]); cudaFree(ctx->d_mnbp[i]); cudaFree(ctx->d_mbest[i]);
            ctx->d_mnorm[i] = nullptr; ctx->d_mnbp[i] = nullptr; ctx->d_mbest[i] = nullptr; ctx->m_cap[i] = 0;
            CU(dalloc(&ctx->d_mnorm[i], pad));
            CU(dalloc(&ctx->d_mnbp[i], pad));
            CU(dalloc(&ctx->d_mbest[i], pad));
            ctx->m_cap[i] = pad;
        }
        k_match_prep<<<(unsigned)((pad + 7) / 8), 256, 0, st>>>(d[i], (uint32_t)n[i], ctx->d_mnorm[i], ctx->d_mnbp[i], (uint32_t)pad);
        ctx->launches++;
    }
    if (n_query > ctx->mout_cap) {
        cudaFree(ctx->d_mout);
        ctx->d_mout = nullptr; ctx->mout_cap = 0;
        CU(dalloc(&ctx->d_mout, (size_t)n_query));
        ctx->mout_cap = n_query;
    }
    if (!ctx->d_mcount) CU(dalloc(&ctx->d_mcount, 1));
    // nearest neighbour in both directions: rows of `a` against all rows of `b`
    for (int dir = 0; dir < 2; dir++) {
        const int a = dir, b = 1 - dir;
        CUtensorMap tm_a, tm_b;
        int rc;
        if ((rc = encode_desc_map(ctx, &tm_a, d[a], n[a], MT_M)) || (rc = encode_desc_map(ctx, &tm_b, d[b], n[b], MT_N))) return rc;
        MatchParams mp{};
        mp.norm_a = ctx->d_mnorm[a]; mp.nbp = ctx->d_mnbp[b];
        mp.n_a = (uint32_t)n[a]; mp.n_b = (uint32_t)n[b];
        mp.best = ctx->d_mbest[a];
        // one CTA per 128 query rows and per range of train tiles: enough CTAs for two per SM-slot (one CTA holds
        // all of an SM's tensor memory), ranges of at least four tiles
        // one CTA per 128 query rows and per range of train tiles; one CTA holds all of an SM's tensor memory, so the
        // grid runs in rounds of sm_count CTAs: pick the split whose rounds x (tiles per CTA + start-up) is smallest
        const uint32_t row_ctas = (uint32_t)((n[a] + MT_M - 1) / MT_M), tiles = (uint32_t)((n[b] + MT_N - 1) / MT_N);
        uint32_t splits = 1;
        double best_cost = 1e30;
        for (uint32_t sp = 1; sp <= tiles; sp++) {
            const uint32_t per = (tiles + sp - 1) / sp, eff = (tiles + per - 1) / per;
            const uint64_t rounds = ((uint64_t)row_ctas * eff + ctx->sm_count - 1) / ctx->sm_count;
            const double cost = (double)rounds * (per + 1.5);
            if (cost < best_cost - 1e-9) { best_cost = cost; splits = sp; }
        }
        mp.tiles_per_cta = (tiles + splits - 1) / splits;
        splits = (tiles + mp.tiles_per_cta - 1) / mp.tiles_per_cta;
        CU(cudaMemsetAsync(ctx->d_mbest[a], 0xff, n[a] * sizeof(unsigned long long), st));
        k_match_nn<<<dim3(row_ctas, splits), MT_THREADS, MT_SMEM, st>>>(tm_a, tm_b, mp);
        ctx->launches++;
    }
    k_match_cross<<<1, 1024, 0, st>>>(ctx->d_mbest[0], ctx->d_mbest[1], (uint32_t)n_query, (uint32_t)n_train, ctx->d_mout,
                                      (uint32_t)std::min<uint64_t>(ctx->mout_cap, 0xffffffffull), ctx->d_mcount);
    ctx->launches++;
    CU(cudaGetLastError());
    // one round trip: the count and every row the caller's buffer could hold leave the device together (at most
    // n_query rows of 12 bytes), instead of count, synchronise, rows, synchronise
    uint32_t cnt = 0;
    const uint64_t room = std::min<uint64_t>(n_query, cap);
    CU(cudaMemcpyAsync(&cnt, ctx->d_mcount, 4, cudaMemcpyDeviceToHost, st));
    if (room) CU(cudaMemcpyAsync(out, ctx->d_mout, room * sizeof(sb200_dmatch), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_out = cnt;
    return cnt > cap ? fail(ctx, SB200_E_CAPACITY, "%u matches exceed the output capacity %llu", cnt, (unsigned long long)cap) : SB200_OK;
}

int sb200_match_descriptors(sb200_ctx* ctx, const uint8_t* query, uint64_t n_query, const uint8_t* train, uint64_t n_train,
                sb200_dmatch* out, uint64_t cap, uint64_t* n_out) {
    if (!ctx) return SB200_E_INVALID;
    if (!n_out || (n_query && !query) || (n_train && !train)) return fail(ctx, SB200_E_INVALID, "bad arguments to match");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    const uint8_t* h[2] = {query, train};
    const uint64_t n[2] = {n_query, n_train};
    for (int i = 0; i < 2; i++) {
        // staging copy of the host descriptors; the device entry point below sizes the other scratch arrays
        const size_t pad = std::max<size_t>((size_t)((n[i] + MT_N - 1) / MT_N) * MT_N, MT_N);
        if (pad > ctx->mdesc_cap[i]) {
            CU(cudaStreamSynchronize(st));
            cudaFree(ctx->d_mdesc[i]);
            ctx->d_mdesc[i] = nullptr; ctx->mdesc_cap[i] = 0;
            CU(dalloc(&ctx->d_mdesc[i], pad * DESC_SIZE));
            ctx->mdesc_cap[i] = pad;
        }
        if (n[i]) CU(cudaMemcpyAsync(ctx->d_mdesc[i], h[i], n[i] * DESC_SIZE, cudaMemcpyHostToDevice, st));
    }
    return sb200_match_descriptors_device(ctx, ctx->d_mdesc[0], n_query, ctx->d_mdesc[1], n_train, out, cap, n_out);
}

// ---- measurement ---------------------------------------------------------------
int sb200_set_profiling(sb200_ctx* ctx, int on) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = finish_all(ctx);
    ctx->profiling = on != 0;
    return rc;
}

int sb200_stage_stats(sb200_ctx* ctx, double* ms, uint64_t* launches, uint32_t cap) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = finish_all(ctx);
    if (rc) return rc;
    drain_stage_events(ctx);
    for (uint32_t i = 0; i < cap && i < SB200_STAGE_COUNT; i++) {
        if (ms) ms[i] = ctx->stage_ms[i];
        if (launches) launches[i] = ctx->stage_launches[i];
    }
    return SB200_OK;
}

int sb200_reset_stats(sb200_ctx* ctx) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = finish_all(ctx);
    drain_stage_events(ctx);
    for (int i = 0; i < SB200_STAGE_COUNT; i++) { ctx->stage_ms[i] = 0; ctx->stage_launches[i] = 0; }
    for (int i = 0; i < SB200_FINE_SLOTS; i++) { ctx->fine_ms[i] = 0; ctx->fine_launches[i] = 0; }
    return rc;
}

int sb200_launch_stats(sb200_ctx* ctx, double* ms, uint64_t* launches, uint32_t cap) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = finish_all(ctx);
    if (rc) return rc;
    drain_stage_events(ctx);
    for (uint32_t i = 0; i < cap && i < SB200_FINE_SLOTS; i++) {
        if (ms) ms[i] = ctx->fine_ms[i];
        if (launches) launches[i] = ctx->fine_launches[i];
    }
    return SB200_OK;
}

uint64_t sb200_launch_count(const sb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

uint64_t sb200_algorithmic_bytes(uint32_t w, uint32_t h, uint64_t* out, uint32_t cap) {
    // A(W,H) = W*H + 4*S  +  P*5*8 + P1*8  +  P_act*24
    PyrLayout L = make_layout(w, h);
    uint64_t S = (uint64_t)L.o[0].w * L.o[0].h, P = 0, Pact = 0;
    for (int o = 0; o < L.n_oct; o++) {
        const uint64_t px = (uint64_t)std::max(L.o[o].w, 0) * std::max(L.o[o].h, 0);
        P += px;
        if (L.o[o].scanned) Pact += px;
    }
    const uint64_t seed = (uint64_t)w * h + 4 * S;
    const uint64_t blur = P * 5 * 8 + (P - S) * 8;
    const uint64_t ext = Pact * 24;
    if (out && cap > 0) out[0] = seed;
    if (out && cap > 1) out[1] = blur;
    if (out && cap > 2) out[2] = ext;
    return seed + blur + ext;
}

int sb200_timer_start(sb200_ctx* ctx) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    // the timer lives on slot 0's stream; slot 1 is ordered behind it through a dependency event
    CU(cudaEventRecord(ctx->t0, ctx->slot[0].stream));
    for (int i = 1; i < N_SLOTS; i++) CU(cudaStreamWaitEvent(ctx->slot[i].stream, ctx->t0, 0));
    return SB200_OK;
}

int sb200_timer_stop(sb200_ctx* ctx) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    for (int i = 1; i < N_SLOTS; i++) {
        cudaEvent_t j = get_event(ctx);
        CU(cudaEventRecord(j, ctx->slot[i].stream));
        CU(cudaStreamWaitEvent(ctx->slot[0].stream, j, 0));
        ctx->ev_pool.push_back(j);
    }
    CU(cudaEventRecord(ctx->t1, ctx->slot[0].stream));
    return SB200_OK;
}

int sb200_timer_elapsed_ms(sb200_ctx* ctx, float* ms) {
    if (!ctx || !ms) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->t1));
    CU(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return SB200_OK;
}

int sb200_host_alloc(size_t bytes, void** out) {
    if (!out) return SB200_E_INVALID;
    return cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocDefault) == cudaSuccess ? SB200_OK : SB200_E_CUDA;
}
int sb200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? SB200_OK : SB200_E_CUDA; }

int sb200_device_alloc(sb200_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(out, std::max<size_t>(bytes, 1)));
    return SB200_OK;
}
int sb200_device_free(sb200_ctx* ctx, void* p) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(p));
    return SB200_OK;
}
int sb200_memcpy_h2d(sb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->slot[0].stream));
    CU(cudaStreamSynchronize(ctx->slot[0].stream));
    return SB200_OK;
}
int sb200_memcpy_d2h(sb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->slot[0].stream));
    CU(cudaStreamSynchronize(ctx->slot[0].stream));
    return SB200_OK;
}

int sb200_flush_l2(sb200_ctx* ctx) {
    if (!ctx) return SB200_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_flush) {
        ctx->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
        CU(cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
    }
    CU(cudaMemsetAsync(ctx->d_flush, 0x5a, ctx->flush_bytes, ctx->slot[0].stream));
    return SB200_OK;
}

}  // extern "C"
