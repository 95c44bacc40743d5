"""Pins the oracle's Processing flavour A (src/opencv_processing.rs) against the real OpenCV
(cv2 4.x in this image) and its whole pipeline against cv2.SIFT_create on identical pixels
(the crate's stated goal: README.md / src/lib.rs:8-10)."""
import numpy as np
import pytest

from conftest import load_gray, noise_image

cv2 = pytest.importorskip("cv2")


def _host_has_fma() -> bool:
    try:
        return " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


def _sigmas(O):
    return [O.seed_sigma()] + [O.octave_sigma(s) for s in range(1, 6)]


def test_sigmas_match_reference_values(oracle):
    # SURVEY.md appendix A.2 (src/lib.rs:207, 220-229)
    assert oracle.seed_sigma() == pytest.approx(1.2489995996796799, abs=1e-15)
    exp = [1.2262734984654078, 1.5450077936447957, 1.9465878414647122, 2.452546996930815, 3.0900155872895905]
    for s, e in zip(range(1, 6), exp):
        assert oracle.octave_sigma(s) == pytest.approx(e, abs=1e-14)


def test_taps_equal_opencv(oracle):
    for s in _sigmas(oracle):
        ks = int(round(s * 8 + 1)) | 1
        assert np.array_equal(oracle.gaussian_taps(s), cv2.getGaussianKernel(ks, s, cv2.CV_32F).ravel())
    assert [len(oracle.gaussian_taps(s)) for s in _sigmas(oracle)] == [11, 11, 13, 17, 21, 27]


@pytest.mark.parametrize("shape", [(128, 256), (96, 160), (64, 1024)])
def test_blur_bit_exact_vs_opencv(oracle, shape):
    # widths are multiples of 16 so OpenCV's AVX2 / AVX-512 vector body covers every column
    # (its scalar tail does not fuse the multiply-add and differs by 1 ulp)
    img = noise_image(shape[1], shape[0], 3).astype(np.float32) / np.float32(255)
    for s in _sigmas(oracle):
        a, b = oracle.gaussian_blur(img, s), cv2.GaussianBlur(img, (0, 0), s)
        if not np.array_equal(a, b):
            # CPUs without FMA take OpenCV's non-fused path: still the same filter to 1 ulp
            assert np.abs(a - b).max() <= 2.4e-7
            pytest.skip("this host's OpenCV does not use the FMA path")


def test_blur_any_width_within_one_ulp(oracle):
    img = noise_image(131, 67, 4).astype(np.float32) / np.float32(255)
    for s in _sigmas(oracle):
        assert np.abs(oracle.gaussian_blur(img, s) - cv2.GaussianBlur(img, (0, 0), s)).max() <= 2.4e-7


@pytest.mark.parametrize("shape", [(53, 77), (213, 320), (300, 400), (7, 9), (2, 2)])
def test_resize_bit_exact_vs_opencv(oracle, shape):
    h, w = shape
    img = noise_image(w, h, 5).astype(np.float32) / np.float32(255)
    up = oracle.resize_linear_2x(img)
    ref = cv2.resize(img, (2 * w, 2 * h), interpolation=cv2.INTER_LINEAR)
    assert np.abs(up - ref).max() <= 1.2e-7
    if cv2.useOptimized():
        assert np.array_equal(up, ref)
    half = oracle.resize_nearest_half(up)
    assert np.array_equal(half, cv2.resize(up, (w, h), interpolation=cv2.INTER_NEAREST))
    odd = up[: 2 * h - 1, : 2 * w - 1].copy()
    if odd.shape[0] >= 2 and odd.shape[1] >= 2:
        assert np.array_equal(oracle.resize_nearest_half(odd),
                              cv2.resize(odd, (odd.shape[1] // 2, odd.shape[0] // 2), interpolation=cv2.INTER_NEAREST))


def test_pyramid_equals_opencv_chain(oracle):
    g = load_gray("bird_small")
    P = oracle.Pyramid(g)
    f = g.astype(np.float32) / np.float32(255)
    cur = cv2.GaussianBlur(cv2.resize(f, (g.shape[1] * 2, g.shape[0] * 2), interpolation=cv2.INTER_LINEAR), (0, 0),
                           oracle.seed_sigma())
    assert P.n_octaves == 8
    for o in range(P.n_octaves):
        layers = [cur]
        for s in range(1, 6):
            layers.append(cv2.GaussianBlur(layers[-1], (0, 0), oracle.octave_sigma(s)))
        for l in range(6):
            d = np.abs(P.gauss(o, l) - layers[l]).max()
            assert d <= 2.4e-7, (o, l, d)
            if layers[l].shape[1] % 16 == 0 and cv2.useOptimized() and _host_has_fma():
                assert d == 0.0, (o, l)
        cur = cv2.resize(layers[3], (layers[3].shape[1] // 2, layers[3].shape[0] // 2), interpolation=cv2.INTER_NEAREST)
    P.close()


@pytest.mark.parametrize("name", ["bird_small", "tree_small"])
def test_pipeline_matches_cv2_sift(oracle, name):
    """cv2.SIFT_create defaults are the crate's constants (3 / 0.04 / 10 / 1.6, src/lib.rs:92-94,185).
    Differences by design: size = cv/2, duplicates kept, order."""
    from scipy.spatial import cKDTree
    g = load_gray(name)
    kps, desc = oracle.sift(g)
    ckp, cdesc = cv2.SIFT_create().detectAndCompute(g, None)
    c = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in ckp], np.float32)
    assert abs(len(kps) - len(c)) <= 0.03 * len(c)
    A = np.stack([kps["x"], kps["y"], kps["size"] * 4, kps["angle"] / 20], 1)
    B = np.stack([c[:, 0], c[:, 1], c[:, 2] * 2, c[:, 3] / 20], 1)
    d, i = cKDTree(A).query(B)
    ok = d < 0.05
    assert ok.mean() >= 0.99
    j = i[ok]
    pos = np.hypot(kps["x"][j] - c[ok, 0], kps["y"][j] - c[ok, 1])
    assert np.median(pos) < 1e-4 and pos.max() < 5e-3
    assert np.median(np.abs(c[ok, 2] / kps["size"][j] - 2.0)) < 1e-4
    assert np.median(np.abs(c[ok, 4] / kps["response"][j] - 1.0)) < 1e-4
    dd = np.abs(desc[j].astype(int) - cdesc[ok].astype(int)).max(1)
    assert (dd <= 1).mean() >= 0.97 and (dd <= 2).mean() >= 0.99
