"""Descriptor matching (examples/sift-match.rs:30-35): oracle vs OpenCV's BFMatcher on
the CPU, GPU (tcgen05 Gram-matrix kernel, through the C ABI) vs oracle on the B200."""
import numpy as np
import pytest

from conftest import load_gray, load_snapshot


def _random_desc(n, seed, dup=0):
    rng = np.random.default_rng(seed)
    # SIFT-like rows: many small values, a few clamped ones
    d = np.minimum(rng.gamma(0.6, 30.0, (n, 128)), 255).astype(np.uint8)
    if dup and n > 4:
        src = rng.integers(0, n, dup)
        dst = rng.integers(0, n, dup)
        d[dst] = d[src]                      # exact duplicates: distance-0 ties
    return d


def test_oracle_matches_opencv_bfmatcher(oracle):
    cv2 = pytest.importorskip("cv2")
    _, d1 = load_snapshot("bird_small")
    _, d2 = load_snapshot("tree_small")
    for q, t in [(d1, d1[::-1].copy()), (d1, d2), (_random_desc(300, 1), _random_desc(400, 2))]:
        m = cv2.BFMatcher(cv2.NORM_L2, True).match(np.ascontiguousarray(q), np.ascontiguousarray(t))
        cv = np.array(sorted((x.queryIdx, x.trainIdx) for x in m), np.int64).reshape(-1, 2)
        o = oracle.match_cross_check(q, t)
        # identical up to exact-distance ties (OpenCV's tie order is unspecified)
        a = {tuple(r) for r in cv}
        b = {tuple(r[:2]) for r in o}
        assert len(a ^ b) <= max(2, len(a) // 200), (len(a), len(b), len(a ^ b))
        dist = {(x.queryIdx, x.trainIdx): x.distance for x in m}
        for qi, ti, d2v in o[:50]:
            if (qi, ti) in dist:
                assert abs(dist[(qi, ti)] - np.sqrt(d2v)) <= 1e-3 * max(1.0, np.sqrt(d2v))


@pytest.mark.gpu
@pytest.mark.parametrize("nq,nt", [(1, 1), (5, 3), (127, 129), (128, 256), (129, 257), (1000, 1500), (3000, 700),
                                   (8648, 8272)])
def test_gpu_match_random(sf, oracle, nq, nt):
    q, t = _random_desc(nq, 10 + nq, dup=nq // 50), _random_desc(nt, 20 + nt, dup=nt // 50)
    if nq > 10 and nt > 10:
        t[: min(nq, nt) // 3] = q[: min(nq, nt) // 3]        # a block of true correspondences
    with sf.Extractor(8, 8, 1) as ex:
        m = ex.match(q, t)
    o = oracle.match_cross_check(q, t)
    assert len(m) == len(o)
    assert np.array_equal(m["queryIdx"], o[:, 0]) and np.array_equal(m["trainIdx"], o[:, 1])
    assert np.array_equal(np.sqrt(o[:, 2].astype(np.float64)).astype(np.float32), m["distance"])


@pytest.mark.gpu
def test_gpu_match_real_descriptors(sf, oracle):
    """examples/sift-match.rs shape: extract two images on the GPU, match their descriptors."""
    a, b = load_gray("bird_small"), load_gray("bird_small")[:, ::-1].copy()
    with sf.Extractor(a.shape[1], a.shape[0], 1) as ex:
        ra, rb = ex.sift(a), ex.sift(b)
        m = ex.match(ra.descriptors, rb.descriptors)
        self_m = ex.match(ra.descriptors, ra.descriptors)
        empty = ex.match(ra.descriptors, np.zeros((0, 128), np.uint8))
    o = oracle.match_cross_check(ra.descriptors, rb.descriptors)
    assert np.array_equal(np.stack([m["queryIdx"], m["trainIdx"]], 1), o[:, :2])
    # an image against itself: every keypoint whose descriptor is unique matches itself at distance 0
    assert (self_m["distance"] == 0).all() and len(self_m) >= 0.95 * len(ra)
    assert len(empty) == 0


@pytest.mark.gpu
def test_gpu_match_scratch_capacities(sf, oracle):
    """Host and device entry points share scratch arrays with separate capacities: small host match, large device
    match, medium host match on one context (the sequence that used to overrun the host variant's staging buffer)."""
    import ctypes as C
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    small_q, small_t = _random_desc(40, 1), _random_desc(50, 2)
    big_q, big_t = _random_desc(3000, 3), _random_desc(2500, 4)
    mid_q, mid_t = _random_desc(900, 5), _random_desc(1100, 6)
    with sf.Extractor(8, 8, 1) as ex:
        a = ex.match(small_q, small_t)
        d_q, d_t = C.c_void_p(), C.c_void_p()
        assert lib.sb200_device_alloc(ex.handle, big_q.nbytes, C.byref(d_q)) == 0
        assert lib.sb200_device_alloc(ex.handle, big_t.nbytes, C.byref(d_t)) == 0
        lib.sb200_memcpy_h2d(ex.handle, d_q, big_q.ctypes.data, big_q.nbytes)
        lib.sb200_memcpy_h2d(ex.handle, d_t, big_t.ctypes.data, big_t.nbytes)
        raw = np.zeros(len(big_q), np.dtype([("query", np.uint32), ("train", np.uint32), ("dist2", np.uint32)]))
        cnt = C.c_uint64()
        assert lib.sb200_match_descriptors_device(ex.handle, d_q, len(big_q), d_t, len(big_t), raw.ctypes.data, len(raw),
                                                  C.byref(cnt)) == 0
        b = ex.match(mid_q, mid_t)
        lib.sb200_device_free(ex.handle, d_q); lib.sb200_device_free(ex.handle, d_t)
    for m, (q, t) in ((a, (small_q, small_t)), (b, (mid_q, mid_t))):
        o = oracle.match_cross_check(q, t)
        assert np.array_equal(m["queryIdx"], o[:, 0]) and np.array_equal(m["trainIdx"], o[:, 1])
    o = oracle.match_cross_check(big_q, big_t)
    assert cnt.value == len(o) and np.array_equal(raw["query"][: len(o)], o[:, 0])
