// abi_smoke_cpp.cpp -- the header-only C++ mirror (sift_features_b200/cpp/sift_features.hpp) compiled and run:
// sift_with_processing<OpenCVProcessing>, sift() (= ImageprocProcessing, as in the crate), Extractor::sift_batch,
// precompute_images + accessors + sift_with_precomputed, compute_descriptor, match.  Writes the same binary layout
// as abi_smoke.c for blocks 0 (OpenCV flavour), 3 (crate default) and the descriptor; prints a few facts.
//
//     abi_smoke_cpp IMAGE.raw WIDTH HEIGHT OUT.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sift_features_b200/cpp/sift_features.hpp"

namespace sfb = sift_features;

static void dump(FILE* f, const sfb::SiftResult& r) {
    const uint64_t n = r.keypoints.size();
    fwrite(&n, sizeof n, 1, f);
    fwrite(r.keypoints.data(), sizeof(sfb::KeyPoint), n, f);
    fwrite(r.descriptors.data(), 1, r.descriptors.size(), f);
}

int main(int argc, char** argv) {
    if (argc != 5) return 2;
    const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]);
    std::vector<uint8_t> img((size_t)w * h);
    FILE* in = fopen(argv[1], "rb");
    if (!in || fread(img.data(), 1, img.size(), in) != img.size()) return 2;
    fclose(in);
    FILE* out = fopen(argv[4], "wb");
    try {
        const sfb::GrayImageView view{img.data(), w, h, w};
        const sfb::SiftResult a = sfb::sift_with_processing<sfb::OpenCVProcessing>(view);
        dump(out, a);
        const sfb::SiftResult b = sfb::sift(view);
        dump(out, b);
        sfb::Extractor ex(w, h, 2);
        std::vector<uint8_t> two(img);
        two.insert(two.end(), img.begin(), img.end());
        const auto batch = ex.sift_batch(two.data(), 2, w, h, w, (uint64_t)w * h);
        if (batch.size() != 2 || !(batch[0].keypoints == a.keypoints) || batch[1].descriptors != a.descriptors) throw sfb::Error(-1, "batch differs from single");
        ex.precompute_images(view);
        const auto wh = ex.octave_size(0);
        const auto g0 = ex.scale_space(0);
        const auto d0 = ex.dog(0);
        if (wh.first != 2 * w || wh.second != 2 * h || g0.size() != (size_t)6 * wh.first * wh.second || d0.size() != (size_t)5 * wh.first * wh.second)
            throw sfb::Error(-1, "pyramid accessors");
        // DoG layer 0 is Gaussian layer 1 minus layer 0 (src/lib.rs:275)
        const size_t px = (size_t)wh.first * wh.second;
        for (size_t i = 0; i < px; i += 977)
            if (d0[i] != g0[px + i] - g0[i]) throw sfb::Error(-1, "dog != difference of Gaussians");
        const sfb::SiftResult staged = ex.sift_with_precomputed();
        if (!(staged.keypoints == a.keypoints)) throw sfb::Error(-1, "staged differs");
        std::vector<float> f(img.size());
        for (size_t i = 0; i < img.size(); i++) f[i] = (float)img[i] / 255.0f;
        const auto d = ex.compute_descriptor(f.data(), w, h, 100.0f, 100.0f, 2.1f, 123.0f);
        fwrite(d.data(), 1, d.size(), out);
        const auto m = ex.match(a.descriptors.data(), a.keypoints.size(), a.descriptors.data(), a.keypoints.size());
        printf("abi_smoke_cpp ok: %zu keypoints (opencv), %zu (crate default), %u octaves, %zu self-matches\n",
               a.keypoints.size(), b.keypoints.size(), ex.n_octaves(), m.size());
    } catch (const sfb::Error& e) {
        fprintf(stderr, "error %d: %s\n", e.status, e.what());
        return 1;
    }
    fclose(out);
    return 0;
}
