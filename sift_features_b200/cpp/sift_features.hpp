// sift_features.hpp -- header-only C++ mirror of the crate's public interface over the C ABI
// (include/sift_b200.h).  Same names, argument meaning and result types as src/lib.rs:39-81,124-177,785;
// the crate panics on failure, this mirror throws sift_features::Error.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sift_b200.h"

namespace sift_features {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

// src/lib.rs:48-56
struct KeyPoint {
    float x, y, size, angle, response;
    bool operator==(const KeyPoint& o) const {
        return x == o.x && y == o.y && size == o.size && angle == o.angle && response == o.response;
    }
};
static_assert(sizeof(KeyPoint) == sizeof(sb200_keypoint), "layout");

// src/lib.rs:39-46: descriptors is (keypoints.size(), 128) row-major
struct SiftResult {
    std::vector<KeyPoint> keypoints;
    std::vector<uint8_t> descriptors;
};

// image::GrayImage (src/lib.rs:71): borrowed, row-major, `stride` bytes per row
struct GrayImageView {
    const uint8_t* data;
    uint32_t width, height, stride;
};

// src/lib.rs:86-90: the blur / resize implementation the pyramid is built with.  The crate's trait has three
// associated functions and static dispatch; here an implementation is a tag type whose FLAVOUR selects the device
// kernels (sb200_set_processing).
struct OpenCVProcessing {     // src/opencv_processing.rs:39-74 -- the flavour the crate's test and snapshots pin
    static constexpr int FLAVOUR = SB200_PROCESSING_OPENCV;
};
struct ImageprocProcessing {  // src/lib.rs:992-1007 -- the crate's default; restated, parity unpinned
    static constexpr int FLAVOUR = SB200_PROCESSING_IMAGEPROC;
};

// One context per (thread, device); owns the device arenas.  Not re-entrant.
class Extractor {
public:
    Extractor(uint32_t max_w, uint32_t max_h, uint32_t max_batch = 1, int device = 0,
              int processing = SB200_PROCESSING_OPENCV) {
        int st = sb200_create(device, max_w, max_h, max_batch, 0, &ctx_);
        if (st) throw Error(st, std::string("sb200_create: ") + sb200_status_string(st));
        check(sb200_set_processing(ctx_, processing));
    }
    ~Extractor() { sb200_destroy(ctx_); }
    Extractor(const Extractor&) = delete;
    Extractor& operator=(const Extractor&) = delete;

    // sift_with_processing::<P> with the context's flavour P, src/lib.rs:76-81
    SiftResult sift(const GrayImageView& img, std::optional<size_t> features_limit = std::nullopt) {
        sb200_result r{};
        check(sb200_extract(ctx_, img.data, img.width, img.height, img.stride,
                            features_limit ? (int64_t)*features_limit : -1, &r));
        return take(r, 0);
    }
    // n images of identical size -> one SiftResult per image
    std::vector<SiftResult> sift_batch(const uint8_t* data, uint32_t n, uint32_t w, uint32_t h, uint32_t stride,
                                       uint64_t image_stride, std::optional<size_t> features_limit = std::nullopt) {
        sb200_result r{};
        check(sb200_extract_batch(ctx_, data, n, w, h, stride, image_stride,
                                  features_limit ? (int64_t)*features_limit : -1, &r));
        std::vector<SiftResult> out;
        for (uint32_t i = 0; i < n; i++) out.push_back(take(r, i));
        return out;
    }
    // image::open(..).grayscale() + sift(), examples/run-sift.rs:8-19: JPEG bitstreams of one frame size, decoded
    // (nvJPEG) and converted to luma on the device
    std::vector<SiftResult> sift_jpeg(const std::vector<std::pair<const uint8_t*, uint64_t>>& jpegs,
                                      std::optional<size_t> features_limit = std::nullopt) {
        std::vector<const uint8_t*> ptr;
        std::vector<uint64_t> len;
        for (auto& j : jpegs) { ptr.push_back(j.first); len.push_back(j.second); }
        sb200_result r{};
        check(sb200_extract_batch_jpeg(ctx_, ptr.data(), len.data(), (uint32_t)ptr.size(),
                                       features_limit ? (int64_t)*features_limit : -1, &r));
        std::vector<SiftResult> out;
        for (uint32_t i = 0; i < ptr.size(); i++) out.push_back(take(r, i));
        return out;
    }
    // precompute_images, src/lib.rs:131-143 (the pyramid stays on the device)
    void precompute_images(const GrayImageView& img) {
        check(sb200_precompute(ctx_, img.data, img.width, img.height, img.stride));
    }
    // PrecomputedImages accessors, src/lib.rs:124-128: octave count / sizes and the (6,h,w) Gaussian or (5,h,w) DoG
    // stack of one octave, row-major
    uint32_t n_octaves() {
        uint32_t n = 0;
        check(sb200_pyramid_info(ctx_, &n, nullptr, nullptr, 0));
        return n;
    }
    std::pair<uint32_t, uint32_t> octave_size(uint32_t octave) {
        uint32_t n = 0, w[SB200_MAX_OCTAVES] = {0}, h[SB200_MAX_OCTAVES] = {0};
        check(sb200_pyramid_info(ctx_, &n, w, h, SB200_MAX_OCTAVES));
        if (octave >= n) throw Error(SB200_E_INVALID, "octave out of range");
        return {w[octave], h[octave]};
    }
    std::vector<float> scale_space(uint32_t octave) { return stack(octave, 6, sb200_pyramid_layer); }
    std::vector<float> dog(uint32_t octave) { return stack(octave, 5, sb200_pyramid_dog); }
    // sift_with_precomputed, src/lib.rs:147-177
    SiftResult sift_with_precomputed(std::optional<size_t> features_limit = std::nullopt) {
        sb200_result r{};
        check(sb200_extract_precomputed(ctx_, features_limit ? (int64_t)*features_limit : -1, &r));
        return take(r, 0);
    }
    // compute_descriptor, src/lib.rs:785-990
    std::vector<uint8_t> compute_descriptor(const float* img, uint32_t w, uint32_t h, float x, float y, float scale,
                                            float orientation) {
        sb200_desc_in k{x, y, scale, orientation};
        std::vector<uint8_t> out(SB200_DESC_SIZE);
        check(sb200_compute_descriptors(ctx_, img, w, h, w, &k, 1, out.data()));
        return out;
    }
    // BFMatcher(NORM_L2, crossCheck = true).match(query, train), examples/sift-match.rs:30-35: mutual nearest
    // neighbours of two (n,128) u8 descriptor matrices, ascending query order
    std::vector<sb200_dmatch> match(const uint8_t* query, uint64_t n_query, const uint8_t* train, uint64_t n_train) {
        std::vector<sb200_dmatch> out(n_query);
        uint64_t n = 0;
        check(sb200_match_descriptors(ctx_, query, n_query, train, n_train, out.data(), out.size(), &n));
        out.resize(n);
        return out;
    }
    sb200_ctx* handle() { return ctx_; }

private:
    void check(int st) {
        if (st) throw Error(st, sb200_last_error(ctx_));
    }
    std::vector<float> stack(uint32_t octave, uint32_t layers, int (*fn)(sb200_ctx*, uint32_t, uint32_t, float*)) {
        const auto wh = octave_size(octave);
        const size_t px = (size_t)wh.first * wh.second;
        std::vector<float> out(px * layers);
        for (uint32_t l = 0; l < layers; l++) check(fn(ctx_, octave, l, out.data() + px * l));
        return out;
    }
    static SiftResult take(const sb200_result& r, uint32_t i) {
        SiftResult s;
        const uint64_t a = r.offsets[i], b = r.offsets[i + 1];
        s.keypoints.resize(b - a);
        for (uint64_t k = a; k < b; k++) {
            const sb200_keypoint& q = r.keypoints[k];
            s.keypoints[k - a] = KeyPoint{q.x, q.y, q.size, q.angle, q.response};
        }
        s.descriptors.assign(r.descriptors + a * SB200_DESC_SIZE, r.descriptors + b * SB200_DESC_SIZE);
        return s;
    }
    sb200_ctx* ctx_ = nullptr;
};

// src/lib.rs:76 as a free function (creates a context sized for this image): sift_with_processing<P>(img, limit)
template <class P>
inline SiftResult sift_with_processing(const GrayImageView& img, std::optional<size_t> features_limit = std::nullopt) {
    Extractor ex(img.width, img.height, 1, 0, P::FLAVOUR);
    return ex.sift(img, features_limit);
}

// src/lib.rs:71-73: the crate's sift() is sift_with_processing::<ImageprocProcessing>
inline SiftResult sift(const GrayImageView& img, std::optional<size_t> features_limit = std::nullopt) {
    return sift_with_processing<ImageprocProcessing>(img, features_limit);
}

}  // namespace sift_features
