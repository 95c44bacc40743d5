"""JPEG input (the step before the path, examples/run-sift.rs:8, src/lib.rs:1012) on the device.

What can be exact is: the extraction through the JPEG entry point equals the extraction of the pixels the device
decode produced, and those go through the oracle like any other image.  What cannot: two JPEG decoders do not produce
the same pixels (IDCT rounding, chroma upsampling) -- the reference's own zune-jpeg included -- so the decode is
compared with libjpeg-turbo (cv2.imdecode) + the oracle's integer luma at the tolerance written in each assert, and
the keypoints at the reference's snapshot criterion (count within a few percent, >= 90 % matched within 0.5 px)."""
import numpy as np
import pytest

from conftest import load_gray, smooth_image

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _encode(img, quality=92, sampling=None):
    params = [cv2.IMWRITE_JPEG_QUALITY, quality]
    if sampling is not None:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling]
    ok, buf = cv2.imencode(".jpg", img if img.ndim == 2 else img[..., ::-1], params)
    assert ok
    return buf.tobytes()


def _colour(gray, seed):
    """A colour image with real chroma structure derived from a gray fixture."""
    rng = np.random.default_rng(seed)
    h, w = gray.shape
    tint = smooth_image(w, h, seed).astype(np.int32) - 128
    rgb = np.stack([gray.astype(np.int32) + tint // 2, gray.astype(np.int32) - tint // 3,
                    np.roll(gray, 3, 1).astype(np.int32) + rng.integers(-6, 7, (h, w))], -1)
    return np.clip(rgb, 0, 255).astype(np.uint8)


def _cpu_luma(oracle, jpeg):
    """libjpeg-turbo decode + the oracle's integer luma (a one-component stream is its Y plane)."""
    img = cv2.imdecode(np.frombuffer(jpeg, np.uint8), cv2.IMREAD_UNCHANGED)
    return img if img.ndim == 2 else oracle.rgb_to_luma(np.ascontiguousarray(img[..., ::-1]))


def _matched_fraction(a, b, tol=0.5):
    if len(a) == 0 or len(b) == 0:
        return 1.0 if len(a) == len(b) else 0.0
    pa = np.stack([a["x"], a["y"]], 1).astype(np.float64)
    pb = np.stack([b["x"], b["y"]], 1).astype(np.float64)
    hit = 0
    for i in range(0, len(pa), 512):
        d = np.sqrt(((pa[i:i + 512, None, :] - pb[None]) ** 2).sum(-1))
        hit += int((d.min(1) < tol).sum())
    return hit / len(pa)


def test_gray_jpeg(sf, oracle):
    g = load_gray("bird_small")
    h, w = g.shape
    jpeg = _encode(g, 95)
    with sf.Extractor(w, h, 1) as ex:
        assert ex.jpeg_info(jpeg) == (w, h, 1)
        luma = ex.decode_jpeg_luma(jpeg)
        assert ex.jpeg_backend in ("hardware", "gpu", "default")
        # decoder against decoder: IDCT rounding only
        d = np.abs(luma.astype(int) - _cpu_luma(oracle, jpeg).astype(int))
        assert d.max() <= 2 and d.mean() < 0.25, (d.max(), d.mean())
        offs, kp, desc = ex.sift_jpeg([jpeg])
        assert sf.SiftResult(kp, desc) == ex.sift(luma)          # exact: same pixels, same path
    okp, odesc = oracle.sift(luma)                                # and the oracle on those pixels
    assert len(kp) == len(okp)
    assert np.array_equal(kp["x"].view(np.uint32), okp["x"].view(np.uint32))
    assert np.abs(desc.astype(int) - odesc.astype(int)).max(initial=0) <= 2


@pytest.mark.parametrize("sampling,tol_max,tol_mean", [("444", 3, 0.5), ("420", 24, 1.0)])
def test_colour_jpeg(sf, oracle, sampling, tol_max, tol_mean):
    g = load_gray("tree_small")
    h, w = g.shape
    rgb = _colour(g, 17)
    code = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}[sampling]
    jpeg = _encode(rgb, 92, code)
    with sf.Extractor(w, h, 1) as ex:
        assert ex.jpeg_info(jpeg) == (w, h, 3)
        luma = ex.decode_jpeg_luma(jpeg)
        ref = _cpu_luma(oracle, jpeg)
        d = np.abs(luma.astype(int) - ref.astype(int))
        # 4:4:4: IDCT + colour-conversion rounding; 4:2:0 adds the decoders' different chroma upsampling filters
        assert d.max() <= tol_max and d.mean() < tol_mean, (d.max(), d.mean())
        offs, kp, desc = ex.sift_jpeg([jpeg])
        assert sf.SiftResult(kp, desc) == ex.sift(luma)
    okp, _ = oracle.sift(ref)
    assert abs(len(kp) - len(okp)) <= 0.05 * len(okp)
    assert _matched_fraction(kp, okp) >= 0.9


@pytest.mark.parametrize("chunk", [None, "3", "1"])
def test_jpeg_batch(sf, monkeypatch, chunk):
    """Groups of bitstreams: mixed gray / colour streams in one call, more images than the context's batch, several
    decode chunks in flight (SB200_JPEG_CHUNK shrinks the 128-stream decode batches to test sizes); result identical to
    one image at a time."""
    if chunk:
        monkeypatch.setenv("SB200_JPEG_CHUNK", chunk)
    g = load_gray("bird_small")
    h, w = g.shape
    jpegs = []
    for i in range(7):
        img = np.roll(g, 5 * i, 1)
        jpegs.append(_encode(img, 90) if i % 3 == 0 else _encode(_colour(img, i), 88))
    with sf.Extractor(w, h, 2) as ex:
        offs, kp, desc = ex.sift_jpeg(jpegs)
        assert len(offs) == 8 and offs[-1] == len(kp)
        for i, j in enumerate(jpegs):
            one = ex.sift(ex.decode_jpeg_luma(j))
            assert sf.SiftResult(kp[offs[i]:offs[i + 1]], desc[offs[i]:offs[i + 1]]) == one
        lim = ex.sift_jpeg(jpegs[:3], features_limit=50)
        assert np.array_equal(np.diff(lim[0]), [50, 50, 50])


def test_jpeg_errors(sf):
    g = load_gray("bird_small")
    h, w = g.shape
    jpeg = _encode(g)
    with sf.Extractor(w, h, 2) as ex:
        with pytest.raises(sf.SiftError):
            ex.sift_jpeg([b"not a jpeg at all" * 10])
        with pytest.raises(sf.SiftError):
            ex.sift_jpeg([jpeg, _encode(g[:-8])])             # two frame sizes in one call
        assert len(ex.sift_jpeg([jpeg])[1]) > 0              # the context still works afterwards
    with sf.Extractor(w // 2, h // 2, 1) as ex:
        with pytest.raises(sf.SiftError):
            ex.sift_jpeg([jpeg])                             # larger than the context
