"""The oracle against the reference's own golden vectors: the four insta snapshots written by
#[test] sift_end2end (src/lib.rs:1009-1056, src/snapshots/*.snap).  The snapshots were produced from
pixels decoded by the Rust `image` crate (zune-jpeg); the fixtures here were decoded with libjpeg-turbo
(+-1 grey level on some pixels), so the comparison is a tolerance test -- see tests/golden/make_golden.py."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import assert_golden, load_gray, load_snapshot, noise_image


@pytest.mark.parametrize("name,count", [("bird_small", 225), ("tree_small", 1270)])
def test_snapshot_tolerance(oracle, name, count):
    skp, sdesc = load_snapshot(name)
    assert len(skp) == count and sdesc.shape == (count, 128)
    kps, desc = oracle.sift(load_gray(name))
    assert abs(len(kps) - count) <= 0.02 * count + 1
    d, i = cKDTree(np.stack([kps["x"], kps["y"]], 1)).query(skp[:, :2])
    assert (d < 0.5).mean() >= 0.90
    assert np.median(d) < 0.05
    # size / angle / response of the spatially matched keypoints
    A = np.stack([kps["x"], kps["y"], kps["size"] * 4, kps["angle"] / 20], 1)
    B = np.stack([skp[:, 0], skp[:, 1], skp[:, 2] * 4, skp[:, 3] / 20], 1)
    d4, i4 = cKDTree(A).query(B)
    ok = d4 < 0.5
    assert ok.mean() >= 0.85
    assert np.median(np.abs(kps["response"][i4[ok]] / skp[ok, 4] - 1)) < 0.02
    # descriptors of matched keypoints are close in L2 (512-norm vectors)
    l2 = np.linalg.norm(desc[i4[ok]].astype(float) - sdesc[ok].astype(float), axis=1)
    assert np.median(l2) < 25.0


@pytest.mark.parametrize("name", ["bird_small", "tree_small"])
def test_snapshot_conditional_descriptor_pin(oracle, name):
    """The tighter pin: golden keypoints that the oracle reproduces to within 0.02 px, 0.5 degrees and 1 % in size
    must also carry nearly the golden descriptor bytes -- this ties the descriptor arithmetic (not only the detector)
    to the crate's own golden vectors.  The residual is the +-1 grey level of the other JPEG decoder."""
    kps, desc = oracle.sift(load_gray(name))
    assert_golden(name, kps, desc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bird_small", "tree_small"])
def test_gpu_result_against_crate_snapshots(sf, name):
    """The same comparison with the CUDA path in the oracle's place: the reference's golden vectors checked against
    what the GPU returns, directly."""
    res = sf.sift_with_processing(load_gray(name), None, sf.OpenCVProcessing)
    assert_golden(name, res.keypoint_array, res.descriptors)


@pytest.mark.parametrize("name", ["bird_small", "tree_small"])
def test_snapshot_invariants_hold_for_oracle(oracle, name):
    """Properties of the golden data that do not depend on the decoder."""
    skp, sdesc = load_snapshot(name)
    kps, desc = oracle.sift(load_gray(name))
    for d in (sdesc, desc):
        n = np.linalg.norm(d.astype(float), axis=1)
        assert n.min() > 500 and n.max() < 520       # ~512, src/lib.rs:978
    for k in (skp[:, 3], kps["angle"]):
        assert k.min() > 0.0 and k.max() <= 360.0    # src/lib.rs:418
    assert (kps["response"] * 3 > 0.04).all()         # src/lib.rs:360
    # the golden keypoints are sorted by (x, y, size) (src/lib.rs:1020-1030)
    order = np.lexsort((skp[:, 2], skp[:, 1], skp[:, 0]))
    assert np.array_equal(order, np.arange(len(skp))) or np.allclose(skp[order], skp)


def test_duplicates_are_kept(oracle):
    # the crate keeps duplicate keypoints (snapshot -3 has two identical entries)
    skp, _ = load_snapshot("bird_small")
    assert len(np.unique(skp, axis=0)) < len(skp)
    kps, _ = oracle.sift(load_gray("tree_small"))
    a = np.stack([kps[f] for f in kps.dtype.names], 1)
    assert len(np.unique(a, axis=0)) < len(a)


def test_octave_count_rule(oracle):
    # src/lib.rs:133-134: round(log2(min(2W,2H)) - 2) + 1
    for (w, h), n in {(1920, 1080): 10, (3840, 2160): 11, (640, 480): 9, (320, 213): 8, (799, 533): 9}.items():
        P = oracle.Pyramid(np.zeros((h, w), np.uint8))
        assert P.n_octaves == n
        assert P.dims[0] == (2 * w, 2 * h) and P.dims[1] == (w, h)
        P.close()


def test_edge_cases(oracle):
    # constant image: every DoG value is 0 => no extrema (|v| <= 0 rejected, src/lib.rs:465)
    kps, desc = oracle.sift(np.full((64, 64), 128, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 128)
    # images too small to scan (seed < 10 px, src/lib.rs:315-317)
    for shape in [(1, 1), (2, 3), (4, 4)]:
        kps, _ = oracle.sift(noise_image(shape[1], shape[0], 1))
        assert len(kps) == 0
    # ragged sizes
    kps, desc = oracle.sift(noise_image(67, 131, 2))
    assert len(kps) > 0 and desc.shape == (len(kps), 128)


def test_features_limit(oracle):
    g = noise_image(160, 120, 7)
    kps, desc = oracle.sift(g)
    lk, ld = oracle.sift(g, 40)
    assert len(lk) == 40
    assert (np.diff(lk["response"]) <= 0).all()          # strongest first, src/lib.rs:158
    assert lk["response"][-1] >= np.sort(kps["response"])[::-1][39] - 1e-12
    # limit >= len keeps natural order untouched (src/lib.rs:157)
    k2, d2 = oracle.sift(g, len(kps) + 5)
    assert np.array_equal(k2, kps) and np.array_equal(d2, desc)
    # every limited keypoint (and its descriptor) is one of the full set
    full = {tuple(k.tolist()): i for i, k in enumerate(np.stack([kps[f] for f in kps.dtype.names], 1))}
    for k, d in zip(np.stack([lk[f] for f in lk.dtype.names], 1), ld):
        assert np.array_equal(desc[full[tuple(k.tolist())]], d)


def test_descriptor_bench_shape(oracle):
    # benches/descriptor.rs:18-32: raw image as f32, x=y=100, scale 2.1, 123 degrees
    img = load_gray("bird").astype(np.float32) / np.float32(255)
    d = oracle.compute_descriptor(img, 100.0, 100.0, 2.1, 123.0)
    assert d.shape == (128,) and 480 < np.linalg.norm(d.astype(float)) < 530
