"""Processing flavour B = ImageprocProcessing (src/lib.rs:992-1007), the crate's default for sift() (src/lib.rs:71-73).

PARITY UNPINNED: imageproc 0.25 / image 0.25 are not part of the reference tree and no reference test, snapshot or
bench exercises this flavour, so neither the oracle's restatement of their published algorithms nor the kernels can
be checked against the crates themselves here.  What these tests pin:
  * (CPU) the oracle's literal restatement against independent numpy restatements of the same published algorithms
    and against the closed forms the kernels use (weights .25 / .75, odd-pixel decimation, clamp borders);
  * (GPU) the kernels against the oracle: pyramid, candidates, order and keypoint fields bit-exact, descriptors
    within +-1, exactly the bars of the pinned flavour.
"""
import numpy as np
import pytest

from conftest import load_gray, noise_image, smooth_image


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _np_taps(sigma64):
    f = np.float32
    sigma = f(sigma64)
    r = int(np.ceil(f(2.0) * sigma))
    norm = f(1.0) / (sigma * np.sqrt(f(2.0) * f(np.pi), dtype=f))
    x = np.arange(r + 1, dtype=f)
    half = (norm * np.exp(-(x * x) / (f(2.0) * (sigma * sigma)), dtype=f)).astype(f)
    return np.concatenate([half[:0:-1], half])


def _np_blur(img, sigma):
    k = _np_taps(sigma)
    r = len(k) // 2
    h, w = img.shape
    tmp = np.zeros_like(img)
    xs = np.arange(w)
    for i, ki in enumerate(k):
        tmp = (tmp + img[:, np.clip(xs + i - r, 0, w - 1)] * ki).astype(np.float32)
    out = np.zeros_like(img)
    ys = np.arange(h)
    for i, ki in enumerate(k):
        out = (out + tmp[np.clip(ys + i - r, 0, h - 1), :] * ki).astype(np.float32)
    return out


def test_oracle_imageproc_taps(oracle):
    for l, sigma in enumerate([oracle.seed_sigma()] + [oracle.octave_sigma(s) for s in range(1, 6)]):
        t = oracle.imageproc_taps(sigma)
        assert len(t) == [7, 7, 9, 9, 11, 15][l]                       # radius ceil(2 sigma)
        assert np.array_equal(t, t[::-1])
        ref = _np_taps(sigma)
        assert np.abs(t - ref).max() <= 2e-7 * ref.max()               # libm expf vs numpy exp: an ulp at most
        assert 0.97 < t.sum() < 1.0                                    # the pdf samples are NOT renormalised


@pytest.mark.parametrize("w,h", [(37, 23), (8, 5), (3, 70), (64, 64)])
def test_oracle_imageproc_blur_and_resizes(oracle, w, h):
    rng = np.random.default_rng(w * 100 + h)
    img = rng.random((h, w)).astype(np.float32)
    for sigma in (oracle.seed_sigma(), oracle.octave_sigma(3), oracle.octave_sigma(5)):
        got = oracle.gaussian_blur_imageproc(img, sigma)
        k, kr = oracle.imageproc_taps(sigma), _np_taps(sigma)
        if np.array_equal(k, kr):                                      # same taps: the accumulation must be bit-identical
            assert np.array_equal(_bits(got), _bits(_np_blur(img, sigma)))
        assert np.abs(got - _np_blur(img, sigma)).max() < 1e-6
    # Triangle 2x: vertical pass, then horizontal; weights (.25, .75) / (.75, .25), the clamped ends renormalise to 1
    def up1d(a):
        n = a.shape[0]
        out = np.empty((2 * n,) + a.shape[1:], np.float32)
        q, t = np.float32(0.25), np.float32(0.75)
        for k in range(n):
            out[2 * k] = a[0] if k == 0 else (a[k - 1] * q + a[k] * t)
            out[2 * k + 1] = a[k] if k == n - 1 else (a[k] * t + a[k + 1] * q)
        return out
    up = oracle.resize_triangle_2x(img)
    assert up.shape == (2 * h, 2 * w)
    assert np.array_equal(_bits(up), _bits(np.clip(up1d(up1d(img).T.copy()).T, 0, 1)))
    # Nearest 1/2: the box kernel with support 0 keeps source pixel floor((d + 0.5) * n / (n // 2)) = 2d + 1
    dn = oracle.resize_nearest_imageproc(img)
    assert np.array_equal(dn, img[1::2, 1::2][: h // 2, : w // 2])


def test_oracle_imageproc_pipeline_differs_from_opencv(oracle):
    g = load_gray("bird_small")
    a, _ = oracle.sift(g)
    b, db = oracle.sift(g, None, oracle.PROCESSING_IMAGEPROC)
    assert len(b) > 0 and len(a) != len(b) and db.shape == (len(b), 128)
    P = oracle.Pyramid(g, oracle.PROCESSING_IMAGEPROC)
    assert P.n_octaves == oracle.Pyramid(g).n_octaves
    # unnormalised taps: every blur dims the image a little, so the layers of an octave are ordered in brightness
    means = [float(P.gauss(0, l).mean()) for l in range(6)]
    assert all(means[l + 1] < means[l] for l in range(5))


# ---------------------------------------------------------------------------------------------------
def _check_imageproc(sf, O, gray, limit=None, pyramid=True):
    h, w = gray.shape
    P = O.Pyramid(gray, O.PROCESSING_IMAGEPROC)
    with sf.Extractor(w, h, 1, processing=sf.ImageprocProcessing) as ex:
        pre = ex.precompute_images(gray)
        assert pre.n_octaves == P.n_octaves and pre.dims == P.dims
        if pyramid:
            for o in range(P.n_octaves):
                g = pre.scale_space[o]
                for l in range(6):
                    assert np.array_equal(_bits(g[l]), _bits(P.gauss(o, l))), f"gaussian octave {o} layer {l}"
        res = ex.sift_with_precomputed(limit)
        assert np.array_equal(ex.last_candidates(), P.candidates())
        kg, ko = ex.last_sift_keypoints(), P.sift_keypoints()
        assert len(kg) == len(ko)
        for f in ("x", "y", "size", "response"):
            assert np.array_equal(_bits(kg[f]), _bits(ko[f])), f
        assert np.abs(kg["angle"] - ko["angle"]).max(initial=0) <= 1e-3 * 180 / np.pi
        okp, odesc = P.sift(limit)
        assert len(res) == len(okp)
        if len(res):
            dd = np.abs(res.descriptors.astype(int) - odesc.astype(int)).max(1)
            assert (dd <= 1).mean() >= 0.99 and dd.max() <= 2
        assert ex.sift(gray, limit) == res
    P.close()
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bird_small", "tree_small", "bird"])
def test_gpu_imageproc_reference_images(sf, oracle, name):
    res = _check_imageproc(sf, oracle, load_gray(name))
    assert len(res) > 100


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(640, 480), (129, 257), (33, 31), (32, 32), (31, 64), (75, 20), (10, 10), (1, 1),
                                 (257, 129), (511, 67), (1920, 1080)])
def test_gpu_imageproc_sizes(sf, oracle, w, h):
    img = smooth_image(w, h, w + h) if (w * h) % 2 else noise_image(w, h, w * 3 + h)
    _check_imageproc(sf, oracle, img, pyramid=(w * h < 400000))


@pytest.mark.gpu
def test_gpu_imageproc_seam(sf, oracle):
    """The Processing seam: the module-level sift() is the crate's sift() = ImageprocProcessing; the flavour is a
    property of the context and can be switched; batches and features_limit go through it unchanged."""
    g = load_gray("bird_small")
    ob, odb = oracle.sift(g, None, oracle.PROCESSING_IMAGEPROC)
    r = sf.sift(g)
    assert len(r) == len(ob) and np.array_equal(_bits(r.keypoint_array["x"]), _bits(ob["x"]))
    assert r == sf.sift_with_processing(g, None, sf.ImageprocProcessing)
    oa, _ = oracle.sift(g)
    assert len(sf.sift_with_processing(g, None, sf.OpenCVProcessing)) == len(oa)
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    with sf.Extractor(g.shape[1], g.shape[0], 2) as ex:
        a = ex.sift(g)
        assert lib.sb200_get_processing(ex.handle) == _ffi.PROCESSING_OPENCV
        ex._check(lib.sb200_set_processing(ex.handle, _ffi.PROCESSING_IMAGEPROC))
        b = ex.sift(g)
        offs, kp, de = ex.sift_batch(np.stack([g, g[::-1].copy(), g]))
        ex._check(lib.sb200_set_processing(ex.handle, _ffi.PROCESSING_OPENCV))
        assert ex.sift(g) == a and b == r and len(a) == len(oa)
        assert np.array_equal(kp[offs[0]:offs[1]], r.keypoint_array) and np.array_equal(kp[offs[2]:offs[3]], r.keypoint_array)
        assert lib.sb200_set_processing(ex.handle, 7) == _ffi.E_INVALID
    lim, _ = oracle.sift(g, 50, oracle.PROCESSING_IMAGEPROC)
    got = sf.sift(g, 50)
    assert len(got) == 50 and np.array_equal(_bits(got.keypoint_array["response"]), _bits(lim["response"]))
