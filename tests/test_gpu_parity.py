"""GPU-vs-oracle parity through the C ABI (run on the B200 box: pytest -m gpu).

Bars (BASELINE.json north_star): candidate set and keypoint order bit-exact; positions, scales, angles
within 1e-3; descriptor bytes within +-1 on >= 99 % of keypoints.  What the kernels actually achieve is
tighter -- the pyramid, the candidate list and every keypoint field are bit-identical to the oracle --
and the tests assert that tighter result so that a regression is caught."""
import numpy as np
import pytest

from conftest import load_gray, noise_image, smooth_image

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _check_image(sf, O, gray, limit=None, pyramid=True):
    h, w = gray.shape
    ex = sf.Extractor(w, h, 1)
    P = O.Pyramid(gray)
    try:
        pre = ex.precompute_images(gray)
        assert pre.n_octaves == P.n_octaves and pre.dims == P.dims
        if pyramid:
            for o in range(P.n_octaves):
                g = pre.scale_space[o]
                for l in range(6):
                    assert np.array_equal(_bits(g[l]), _bits(P.gauss(o, l))), f"gaussian octave {o} layer {l}"
            d0 = pre.dog[0]
            for l in range(5):
                assert np.array_equal(_bits(d0[l]), _bits(P.dog(0, l)))
        res = ex.sift_with_precomputed(limit)
        # integer extrema candidates: same set, same (natural) order
        assert np.array_equal(ex.last_candidates(), P.candidates())
        # SiftKeyPoints before any limit: same count, same order, 1e-3 bar (bit-exact in practice)
        kg, ko = ex.last_sift_keypoints(), P.sift_keypoints()
        assert len(kg) == len(ko)
        assert np.array_equal(kg["octave"], ko["octave"]) and np.array_equal(kg["scale"], ko["scale"])
        for f in ("x", "y", "size", "response"):
            assert np.array_equal(_bits(kg[f]), _bits(ko[f])), f
        assert np.abs(kg["angle"] - ko["angle"]).max(initial=0) <= 1e-3 * 180 / np.pi
        exact_angles = np.mean(_bits(kg["angle"]) == _bits(ko["angle"])) if len(kg) else 1.0
        assert exact_angles >= 0.999
        okp, odesc = P.sift(limit)
        assert len(res) == len(okp)
        ka = res.keypoint_array
        for f in ("x", "y", "size", "response"):
            assert np.array_equal(_bits(ka[f]), _bits(okp[f])), f
        assert np.abs(ka["angle"] - okp["angle"]).max(initial=0) <= 1e-3 * 180 / np.pi
        if len(res):
            dd = np.abs(res.descriptors.astype(int) - odesc.astype(int)).max(1)
            assert (dd <= 1).mean() >= 0.99, (dd <= 1).mean()
            assert dd.max() <= 2
        # the one-call path (sift(), src/lib.rs:71) gives the same answer as the staged path (:131 + :147)
        assert ex.sift(gray, limit) == res
        return res
    finally:
        ex.close()
        P.close()


@pytest.mark.parametrize("name", ["bird_small", "tree_small"])
def test_reference_test_images(sf, oracle, name):
    """The two images of the crate's only test (src/lib.rs:1038, 1047)."""
    res = _check_image(sf, oracle, load_gray(name))
    assert len(res) > 200


@pytest.mark.parametrize("w,h,seed", [(640, 480, 1234), (131, 67, 5), (96, 200, 6), (257, 129, 7)])
def test_noise_images(sf, oracle, w, h, seed):
    _check_image(sf, oracle, noise_image(w, h, seed))


@pytest.mark.parametrize("w,h", [(60, 34), (61, 35), (64, 64), (65, 51), (90, 68), (120, 17), (128, 96), (129, 33),
                                 (30, 128), (33, 257), (255, 16), (512, 9)])
def test_boundary_sizes(sf, oracle, w, h):
    """Sizes whose 2x seed image / octaves land on and next to the kernels' internal boundaries: 60-column extrema
    strips, 64- and 128-column blur strips, 32-row bands, 34-row extrema blocks, the 32-pixel TMA threshold and the
    64x36 tail threshold."""
    _check_image(sf, oracle, noise_image(w, h, 1000 + w * 7 + h))


@pytest.mark.parametrize("w,h,seed", [(400, 300, 11), (333, 222, 12)])
def test_smooth_images(sf, oracle, w, h, seed):
    _check_image(sf, oracle, smooth_image(w, h, seed))


def test_bench_image_bird(sf, oracle):
    """images/bird.jpg, the input of benches/sift.rs:79."""
    _check_image(sf, oracle, load_gray("bird"))


def test_1080p_noise_full_size(sf, oracle):
    """BASELINE.json configs[1]: one 1920x1080 synthetic image, default octaves/layers."""
    res = _check_image(sf, oracle, noise_image(1920, 1080, 1234), pyramid=False)
    assert 6000 < len(res) < 12000


def test_maximum_size(sf, oracle):
    """SB200_MAX_DIM: 4096 x 4096 (8192-pixel seed image, 12 octaves, 13-bit candidate coordinates at their limit)."""
    res = _check_image(sf, oracle, noise_image(4096, 4096, 1234), pyramid=False)
    assert len(res) > 50000


def test_8k_frame(sf, oracle):
    """Beyond the 4096-pixel limit of round 1 (64-bit candidate keys, SB200_MAX_DIM = 8192): a 7680 x 4320 frame (seed
    image 15360 x 8640, 12 octaves).  The CPU oracle needs minutes for the whole frame, so the check is structural: the
    right octave count, the keypoint density of the same noise at 4K, every keypoint inside the frame, and -- exactly --
    the oracle's keypoints in a strip of the frame whose pyramid does not depend on the rest of it (the top 64 rows of
    octave 0 see at most the first ~200 input rows)."""
    w, h = 7680, 4320
    img = noise_image(w, h, 4321)
    with sf.Extractor(w, h, 1) as ex:
        pre = ex.precompute_images(img)
        assert pre.n_octaves == 12 and pre.dims[0] == (2 * w, 2 * h)
        res = ex.sift_with_precomputed()
        ka = ex.last_sift_keypoints()
        cand = ex.last_candidates()
    assert 3.9e3 * 33.2 < len(res) < 4.4e3 * 33.2            # ~4.17 k keypoints per input megapixel on this noise
    kp = res.keypoint_array
    assert kp["x"].min() >= 0 and kp["x"].max() < w and kp["y"].min() >= 0 and kp["y"].max() < h
    assert kp["x"].max() > 4096 + 3000 and kp["y"].max() > 4096              # coordinates past the old 13-bit limit
    # octave-0 keypoints near the top edge depend on the top input rows only: compare with the oracle on a crop
    crop = np.ascontiguousarray(img[:256])
    okp, _ = oracle.sift(crop)
    top = lambda a: a[(a["y"] < 40)]
    got = top(kp)
    ref = top(okp)
    oct0 = lambda a: a[a["size"] < 3.5 / 2]                                   # octave 0: size = kp_scale / 2 < 1.8
    g0, r0 = oct0(got), oct0(ref)
    assert len(g0) == len(r0) and len(g0) > 1000
    assert np.array_equal(_bits(g0["x"]), _bits(r0["x"])) and np.array_equal(_bits(g0["y"]), _bits(r0["y"]))
    # white noise loses contrast with every octave: keypoints survive in octaves 0..2 only, extrema candidates much deeper
    assert ka["octave"].max() == 2 and cand["octave"].max() >= 7
    assert cand["x"].max() > 8192 and cand["y"].max() > 8192                 # 14-bit seed-image coordinates


@pytest.mark.parametrize("seg_rows", [512, 96])
def test_blur_segment_heights(sf, oracle, monkeypatch, seg_rows):
    """The marching blur cuts an octave into vertical segments whose height depends on the batch; a large batch
    of 1080p images uses 512-row segments.  Same bit-exact pyramid whatever the segmentation."""
    monkeypatch.setenv("SB200_SEG_ROWS", str(seg_rows))
    _check_image(sf, oracle, noise_image(700, 650, 21))


@pytest.mark.parametrize("pieces,size", [(1, (700, 650)), (3, (700, 650)), (7, (333, 517)), (64, (260, 200))])
def test_blur_aligned_pieces(sf, oracle, monkeypatch, pieces, size):
    """A launch of several waves (large batches) cuts every column of the marching blur into k equal pieces, one per
    CTA; SB200_PIECES forces that distribution on a single image, with k = 1, k not dividing the band count, and k
    larger than the band count of every octave (clamped to one band per CTA).  Same bit-exact pyramid."""
    monkeypatch.setenv("SB200_PIECES", str(pieces))
    _check_image(sf, oracle, noise_image(size[0], size[1], 33 + pieces))


@pytest.mark.parametrize("w,h,n", [(640, 480, 48), (1000, 700, 24)])
def test_blur_batch_distribution(sf, oracle, w, h, n):
    """A batch large enough for the multi-wave (aligned pieces) distribution on octave 0 and the single-wave one on
    the small octaves (the second shape: strips and bands that end inside the image, three pieces per column): every
    image of the batch equals the single-image result, and image 0 equals the oracle."""
    imgs = np.stack([noise_image(w, h, 900 + i) for i in range(n)])
    with sf.Extractor(w, h, n) as ex:
        off, kps, desc = ex.sift_batch(imgs)
    with sf.Extractor(w, h, 1) as ex1:
        for i in (0, n // 3 + 1, n - 1):
            one = ex1.sift(imgs[i])
            a, b = int(off[i]), int(off[i + 1])
            assert b - a == len(one) and b > a
            assert np.array_equal(desc[a:b], one.descriptors)
            assert kps[a:b].tobytes() == one.keypoint_array.tobytes()
    _check_image(sf, oracle, imgs[0])


@pytest.mark.parametrize("env", [{"SB200_BLUR": "tile"}, {"SB200_TAIL": "0"}, {"SB200_GRAPHS": "0", "SB200_FORK": "0"},
                                 {"SB200_SEED": "fused"}, {"SB200_SEED": "split"}, {"SB200_SIDES": "1"},
                                 {"SB200_SIDES": "4", "SB200_GRAPHS": "0"}, {"SB200_PDL": "0"},
                                 {"SB200_PDL": "0", "SB200_GRAPHS": "0"}])
def test_alternative_paths(sf, oracle, monkeypatch, env):
    """The debugging switches select older / simpler / alternative code paths (independent-tile TMA blur, per-layer
    launches for the small octaves, plain single-stream launches without graph capture, the seed blur that upsamples
    its own input bands in line instead of with producer warps, the separate upsample kernel, one / four side streams
    for the off-chain work of the octaves -- the latter as plain multi-stream launches --, launches without the
    programmatic-dependent-launch attribute): same bit-exact results."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _check_image(sf, oracle, noise_image(520, 390, 77))


@pytest.mark.parametrize("kind", ["ramp", "checker", "stripes", "impulses"])
def test_structured_images(sf, oracle, kind):
    """Synthetic structure that white noise and photographs never show: exact ramps (DoG = rounding noise, masses of
    ties), a checkerboard and 1-pixel stripes (periodic ties, saturated steps), isolated impulses on black."""
    h, w = 192, 256
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "ramp":
        g = ((xx + yy) // 2 % 256).astype(np.uint8)
    elif kind == "checker":
        g = (((xx // 8) + (yy // 8)) % 2 * 255).astype(np.uint8)
    elif kind == "stripes":
        g = ((xx % 2) * 200 + 20).astype(np.uint8)
    else:
        g = np.zeros((h, w), np.uint8)
        g[16::32, 16::32] = 255
        g[40, 70] = 90
    _check_image(sf, oracle, g)


def test_flat_regions(sf, oracle):
    """Saturated / constant blocks: every pixel of such a block passes the reference's extremum test with
    ties (src/lib.rs:437-506) and dies in interpolate_extremum; the candidate list must still be the reference's."""
    g = noise_image(300, 260, 31)
    g[20:120, 30:150] = 255
    g[150:240, 100:290] = 0
    g[125:140, :] = 77
    _check_image(sf, oracle, g)


def test_edge_cases(sf, oracle):
    # constant image, tiny images (never scanned, src/lib.rs:315-317), 1-pixel image
    for g in [np.full((64, 64), 128, np.uint8), noise_image(4, 4, 1), noise_image(1, 1, 1), noise_image(3, 2, 1),
              noise_image(7, 5, 2), noise_image(16, 16, 3)]:
        res = _check_image(sf, oracle, g)
        if g.shape[0] < 5:
            assert len(res) == 0 and res.descriptors.shape == (0, 128)


@pytest.mark.parametrize("limit", [0, 1, 37, 500, 100000])
def test_features_limit(sf, oracle, limit):
    """src/lib.rs:156-161: the `limit` strongest, strongest first; ties broken by natural order."""
    g = noise_image(320, 240, 21)
    res = _check_image(sf, oracle, g, limit=limit)
    total = len(oracle.sift(g)[0])
    assert len(res) == min(limit, total)
    if limit < total:   # sorted by response only when the limit actually truncates (src/lib.rs:157)
        assert (np.diff(res.keypoint_array["response"]) <= 0).all()


def test_strided_input(sf, oracle):
    big = noise_image(300, 120, 31)
    view = big[:, 20:220]                      # stride 300, width 200
    with sf.Extractor(200, 120) as ex:
        a = ex.sift(view)
        b = ex.sift(np.ascontiguousarray(view))
    assert a == b and len(a) > 0


def test_batch_equals_singles(sf, oracle):
    """Independent images of a batch (BASELINE.json configs[3] shape, scaled down) give exactly the
    per-image results, in image order, across group boundaries of the two pipelined slots."""
    n, w, h = 11, 160, 120
    imgs = np.stack([noise_image(w, h, 100 + i) for i in range(n)])
    imgs[3] = 77                                # an image without keypoints in the middle
    with sf.Extractor(w, h, 4) as ex:           # 11 images through groups of 4 -> 3 groups
        offs, kp, de = ex.sift_batch(imgs)
        assert len(offs) == n + 1 and offs[0] == 0 and offs[-1] == len(kp) == len(de)
        for i in range(n):
            single = ex.sift(imgs[i])
            assert np.array_equal(single.keypoint_array, kp[offs[i]:offs[i + 1]])
            assert np.array_equal(single.descriptors, de[offs[i]:offs[i + 1]])
        assert offs[4] == offs[3]
        # limit applies per image
        offs2, kp2, _ = ex.sift_batch(imgs, features_limit=10)
        assert all(offs2[i + 1] - offs2[i] == min(10, offs[i + 1] - offs[i]) for i in range(n))
    okp, odesc = oracle.sift(imgs[5])
    assert np.array_equal(_bits(kp[offs[5]:offs[6]]["x"]), _bits(okp["x"]))


def test_large_batch_equals_singles(sf):
    """A batch big enough for the batch-wide work queues of the orientation / descriptor kernels to run in their
    chunked, request-ahead mode (> 57k keypoints in one group) and for sift_batch to taper its first and last
    groups: still exactly the per-image results in image order."""
    n, w, h = 48, 640, 480
    imgs = np.stack([noise_image(w, h, 500 + i) for i in range(n)])
    with sf.Extractor(w, h, 1) as one:
        singles = [one.sift(imgs[i]) for i in range(n)]
    assert sum(len(s) for s in singles) > 57000
    for max_batch in (n, 8):        # one group of 48; groups of 2, 8, 8, 8, 8, 8, 4, 2
        with sf.Extractor(w, h, max_batch) as ex:
            offs, kp, de = ex.sift_batch(imgs)
        assert len(offs) == n + 1 and offs[-1] == len(kp)
        for i in range(n):
            assert np.array_equal(singles[i].keypoint_array, kp[offs[i]:offs[i + 1]]), (max_batch, i)
            assert np.array_equal(singles[i].descriptors, de[offs[i]:offs[i + 1]]), (max_batch, i)


def test_deterministic(sf):
    g = noise_image(640, 480, 5)
    with sf.Extractor(640, 480, 2) as ex:
        a, b = ex.sift(g), ex.sift(g)
        offs, kp, de = ex.sift_batch(np.stack([g, g]))
    assert np.array_equal(a.keypoint_array, b.keypoint_array)
    assert np.array_equal(kp[: offs[1]], kp[offs[1]:]) and np.array_equal(kp[: offs[1]], a.keypoint_array)
    # lane-private histogram copies summed in a fixed order, no atomics: descriptors repeat bit for bit, whichever
    # warp the work queue hands a keypoint to
    assert np.array_equal(a.descriptors, b.descriptors)
    assert np.array_equal(de[: offs[1]], a.descriptors) and np.array_equal(de[offs[1]:], a.descriptors)


def test_fuzz_sizes(sf, oracle):
    """Seeded sweep over odd sizes and content kinds: every tile / strip / segment remainder of the blur, extrema and
    tail kernels, each against the oracle bit for bit."""
    rng = np.random.default_rng(2024)
    for case in range(36):
        w, h = int(rng.integers(10, 330)), int(rng.integers(10, 250))
        kind = case % 3
        if kind == 0:
            g = noise_image(w, h, 900 + case)
        elif kind == 1:
            g = smooth_image(w, h, 900 + case)
        else:   # blocks: large flat areas, ties and strong edges
            g = np.kron(rng.integers(0, 4, (h // 16 + 1, w // 16 + 1)) * 80, np.ones((16, 16)))[:h, :w].astype(np.uint8)
        _check_image(sf, oracle, g, pyramid=(case % 6 == 0))


@pytest.mark.parametrize("w,h", [(4096, 10), (10, 4096), (2000, 16), (16, 2000), (4096, 1), (1, 4096), (3000, 33), (33, 3000),
                                 (8192, 12), (12, 8192), (5000, 40)])
def test_extreme_aspect(sf, oracle, w, h):
    """Strips and ribbons up to the maximum dimension: octaves that run out of rows or columns long before the other
    dimension does (TMA boxes wider / taller than the layer, single-row layers, the fused tail on 1-pixel-wide octaves)."""
    _check_image(sf, oracle, smooth_image(w, h, w + h), pyramid=False)
    _check_image(sf, oracle, noise_image(w, h, w * 7 + h), pyramid=(w * h < 50000))


def test_concurrent_contexts(sf):
    """The crate's functions are re-entrant: here one context per thread; four threads on one
    device at once give what each gives alone."""
    import threading
    imgs = [noise_image(320 + 32 * i, 240, 700 + i) for i in range(4)]
    alone = [sf.sift_with_processing(g) for g in imgs]
    got, errs = [None] * 4, []

    def work(i):
        try:
            with sf.Extractor(imgs[i].shape[1], imgs[i].shape[0], 2) as ex:
                for _ in range(6):
                    r = ex.sift(imgs[i])
                    offs, kp, de = ex.sift_batch(np.stack([imgs[i]] * 3))
                    assert np.array_equal(kp[offs[1]:offs[2]], r.keypoint_array)
                got[i] = r
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for a, b in zip(alone, got):
        assert a == b


def test_compute_descriptor_bench_shape(sf, oracle):
    """benches/descriptor.rs:18-32 and BASELINE.json configs[4] (scaled down for the oracle)."""
    img = load_gray("bird").astype(np.float32) / np.float32(255)
    d = sf.compute_descriptor(img, 100.0, 100.0, 2.1, 123.0)
    ref = oracle.compute_descriptor(img, 100.0, 100.0, 2.1, 123.0)
    assert np.abs(d.astype(int) - ref.astype(int)).max() <= 1
    rng = np.random.default_rng(99)
    n = 400
    k = np.stack([rng.uniform(0, img.shape[1], n), rng.uniform(0, img.shape[0], n),
                  np.exp(rng.uniform(np.log(1.8), np.log(3.6), n)), rng.uniform(0, 360, n)], 1).astype(np.float32)
    k[:4, :2] = [[0, 0], [img.shape[1] - 1, img.shape[0] - 1], [-3.2, 5], [5, 1e6]]   # borders / outside
    with sf.Extractor(img.shape[1], img.shape[0]) as ex:
        got = ex.compute_descriptors(img, k)
    exp = np.stack([oracle.compute_descriptor(img, *map(float, r)) for r in k])
    dd = np.abs(got.astype(int) - exp.astype(int)).max(1)
    assert (dd <= 1).mean() >= 0.99 and dd.max() <= 2


def test_errors(sf):
    from sift_features_b200 import _ffi
    with sf.Extractor(64, 64) as ex:
        with pytest.raises(sf.SiftError) as e:
            ex.sift(noise_image(65, 64, 1))                 # larger than the context
        assert e.value.status == _ffi.E_INVALID
        with pytest.raises(sf.SiftError) as e:
            ex.sift_with_precomputed()                      # no pyramid resident
        assert e.value.status == _ffi.E_STATE


def test_capacity_grows(sf, oracle):
    """More candidates / keypoints than the context was sized for is not an error (the reference never fails): the
    library re-allocates its per-candidate arrays and runs the detection stages again on the resident pyramids --
    for a single image, and in the middle of a pipelined batch."""
    img = noise_image(160, 120, 3)
    with sf.Extractor(160, 120, 1) as ex:
        ref = ex.sift(img)
    assert len(ref) > 64
    with sf.Extractor(160, 120, 1, max_keypoints_per_image=16) as ex:
        assert ex.sift(img) == ref
        assert ex.sift(img) == ref                          # and stays usable
    imgs = np.stack([noise_image(160, 120, 3 + (i % 3)) for i in range(9)])
    with sf.Extractor(160, 120, 2, max_keypoints_per_image=16) as ex:
        offs, kp, de = ex.sift_batch(imgs)
    with sf.Extractor(160, 120, 2) as ex:
        offs2, kp2, de2 = ex.sift_batch(imgs)
    assert np.array_equal(offs, offs2) and np.array_equal(kp, kp2) and np.array_equal(de, de2)
    assert np.array_equal(kp[offs[0]:offs[1]], ref.keypoint_array)


def test_opencv_cross_match(sf):
    """examples/opencv-cross-match.rs:34-43,63-73 with a numeric criterion: OpenCV SIFT on the image vs this
    library on the same image, BFMatcher(NORM_L2, crossCheck=true)."""
    cv2 = pytest.importorskip("cv2")
    g = load_gray("bird")
    ckp, cdesc = cv2.SIFT_create().detectAndCompute(g, None)
    res = sf.sift_with_processing(g, None, sf.OpenCVProcessing)
    m = cv2.BFMatcher(cv2.NORM_L2, True).match(res.descriptors.astype(np.float32), cdesc)
    assert len(m) >= 0.9 * min(len(res), len(ckp))
    ka = res.keypoint_array
    d = np.array([np.hypot(ka["x"][x.queryIdx] - ckp[x.trainIdx].pt[0], ka["y"][x.queryIdx] - ckp[x.trainIdx].pt[1])
                  for x in m])
    assert (d < 1.0).mean() >= 0.9


def test_multi_device_shards(sf, oracle):
    """sb200_extract_batch_multi_parts / _multi: contiguous shards over several contexts -- every device of the box, more
    than the module's context cache normally holds -- results in image order, the zero-copy
    parts and the dense gather identical.  With one device the same entry points run with one context."""
    from sift_features_b200 import _ffi
    ndev = _ffi.load().sb200_device_count()
    devices = list(range(ndev))
    n, w, h = 19, 160, 120
    imgs = np.stack([noise_image(w, h, 300 + i) for i in range(n)])
    res = sf.sift_batch(imgs, devices=devices, max_batch=2)
    dense = sf.sift_batch(imgs, devices=devices, max_batch=2, dense=True)
    assert len(res) == len(dense) == n
    for a, b in zip(res, dense):
        assert a == b
    for i in (0, 4, 8, 18):
        okp, odesc = oracle.sift(imgs[i])
        assert len(res[i]) == len(okp)
        assert np.array_equal(_bits(res[i].keypoint_array["x"]), _bits(okp["x"]))
        assert np.abs(res[i].descriptors.astype(int) - odesc.astype(int)).max(initial=0) <= 1
    if ndev < 2:
        pytest.skip("single device: multi-context gather exercised with one context only")


def test_multi_contexts_one_device(sf):
    """The sharding entry points with several contexts on ONE device (what a single-GPU lease can exercise): three
    contexts, shards of unequal size, an empty trailing shard, parts == dense == one context."""
    import ctypes as C
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    n, w, h = 7, 96, 80
    imgs = np.stack([noise_image(w, h, 500 + i) for i in range(n)])
    with sf.Extractor(w, h, 2) as ref:
        offs, kp, de = ref.sift_batch(imgs)
    for n_ctx in (2, 3, 4, 8):          # 8 contexts for 7 images: ceil(7/8) = 1 image each, the last context idle
        exs = [sf.Extractor(w, h, 2) for _ in range(n_ctx)]
        try:
            handles = (C.c_void_p * n_ctx)(*[e.handle for e in exs])
            parts = (_ffi.Result * n_ctx)()
            first = (C.c_uint64 * (n_ctx + 1))()
            assert lib.sb200_extract_batch_multi_parts(handles, n_ctx, imgs.ctypes.data, n, w, h, w, w * h, -1, parts, first) == 0
            assert list(first) == [min(d * -(-n // n_ctx), n) for d in range(n_ctx + 1)]
            got_kp, got_de, n_img = [], [], 0
            for d in range(n_ctx):
                if parts[d].n_images == 0:
                    continue
                assert first[d] == n_img
                o, k, dd = exs[d]._take(parts[d])
                assert np.array_equal(o + len(np.concatenate(got_kp)) if got_kp else o, offs[n_img:n_img + len(o)])
                got_kp.append(k); got_de.append(dd); n_img += int(parts[d].n_images)
            assert n_img == n
            assert np.array_equal(np.concatenate(got_kp), kp) and np.array_equal(np.concatenate(got_de), de)
            res = _ffi.Result()
            assert lib.sb200_extract_batch_multi(handles, n_ctx, imgs.ctypes.data, n, w, h, w, w * h, -1, C.byref(res)) == 0
            o2, k2, d2 = exs[0]._take(res)
            assert np.array_equal(o2, offs) and np.array_equal(k2, kp) and np.array_equal(d2, de)
            assert lib.sb200_last_gather_ms(exs[0].handle) >= 0.0
        finally:
            for e in exs:
                e.close()


def test_pageable_equals_pinned(sf):
    """Host input from pageable memory (what the crate's callers hold) goes through the pinned staging buffers, group
    after group, overlapped with the previous group's compute: same result as pinned input, strided rows included."""
    import ctypes as C
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    n, w, h = 11, 200, 150
    imgs = np.stack([noise_image(w, h, 700 + i) for i in range(n)])
    with sf.Extractor(w, h, 3) as ex:
        a = ex.sift_batch(imgs)                                   # pageable numpy memory
        p = C.c_void_p()
        assert lib.sb200_host_alloc(imgs.nbytes, C.byref(p)) == 0
        pinned = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=imgs.shape)
        pinned[...] = imgs
        b = ex.sift_batch(pinned)
        wide = np.zeros((n, h, w + 24), np.uint8)
        wide[:, :, :w] = imgs
        c = ex.sift_batch(wide[:, :, :w])                         # row stride != width
        lib.sb200_host_free(p)
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y) and np.array_equal(x, z)


def test_precomputed_images_are_owned(sf, oracle):
    """PrecomputedImages is an owned value in the crate (src/lib.rs:124-128): a later sift() on another image of the
    same shape must not change what sift_with_precomputed(pre) returns; a view of a shared context that went stale
    raises instead of returning the other image's results."""
    a, b = noise_image(128, 96, 1), noise_image(128, 96, 2)
    pre = sf.precompute_images(a)
    rb = sf.sift_with_processing(b)                               # same shape: the module-level cache's context
    ra = sf.sift_with_precomputed(pre)
    assert ra == sf.sift_with_processing(a) and not (ra == rb)
    assert np.array_equal(pre.scale_space[0][0].view(np.uint32), oracle.Pyramid(a).gauss(0, 0).view(np.uint32))
    pre.close()
    with sf.Extractor(128, 96, 1) as ex:
        view = ex.precompute_images(a)
        ex.sift(b)
        with pytest.raises(sf.SiftError):
            ex.sift_with_precomputed(pre=view)
        with pytest.raises(sf.SiftError):
            view.scale_space[0]


def test_device_result_reports_truncation(sf):
    """sb200_device_result: a device-resident result that hit the per-image capacity is reported (SB200_E_CAPACITY),
    not silently truncated."""
    import ctypes as C
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    img = noise_image(160, 120, 3)
    with sf.Extractor(160, 120, 1, max_keypoints_per_image=16) as ex:
        d = C.c_void_p()
        assert lib.sb200_device_alloc(ex.handle, img.nbytes, C.byref(d)) == 0
        assert lib.sb200_memcpy_h2d(ex.handle, d, img.ctypes.data, img.nbytes) == 0
        assert lib.sb200_extract_batch_device(ex.handle, d, 1, 160, 120, 160, 160 * 120, -1) == 0
        assert lib.sb200_sync(ex.handle) == 0
        counts = (C.c_uint32 * 1)()
        assert lib.sb200_device_result(ex.handle, counts, 1, None, None, None) == _ffi.E_CAPACITY
        lib.sb200_device_free(ex.handle, d)
    with sf.Extractor(160, 120, 1) as ex:
        d = C.c_void_p()
        assert lib.sb200_device_alloc(ex.handle, img.nbytes, C.byref(d)) == 0
        assert lib.sb200_memcpy_h2d(ex.handle, d, img.ctypes.data, img.nbytes) == 0
        assert lib.sb200_extract_batch_device(ex.handle, d, 1, 160, 120, 160, 160 * 120, -1) == 0
        counts = (C.c_uint32 * 1)()
        assert lib.sb200_device_result(ex.handle, counts, 1, None, None, None) == 0 and counts[0] > 64
        lib.sb200_device_free(ex.handle, d)


def test_descriptor_scale_limit(sf):
    """compute_descriptor with a window radius beyond the kernel's row table (scale > 12) is an invalid argument, on
    the host entry point before anything runs and on the device entry point at the next synchronising call."""
    import ctypes as C
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    img = (noise_image(256, 256, 9).astype(np.float32) / np.float32(255))
    with sf.Extractor(256, 256, 1) as ex:
        ok = ex.compute_descriptors(img, [[128, 128, 11.9, 30.0]])
        assert ok.shape == (1, 128) and ok.any()
        with pytest.raises(sf.SiftError) as e:
            ex.compute_descriptors(img, [[128, 128, 2.0, 0.0], [128, 128, 12.5, 0.0]])
        assert e.value.status == _ffi.E_INVALID
        k = np.array([[128, 128, 2.0, 0.0], [100, 90, 40.0, 10.0]], np.float32)
        d_img, d_k, d_out = C.c_void_p(), C.c_void_p(), C.c_void_p()
        for dptr, nbytes in ((d_img, img.nbytes), (d_k, k.nbytes), (d_out, 256)):
            assert lib.sb200_device_alloc(ex.handle, nbytes, C.byref(dptr)) == 0
        lib.sb200_memcpy_h2d(ex.handle, d_img, img.ctypes.data, img.nbytes)
        lib.sb200_memcpy_h2d(ex.handle, d_k, k.ctypes.data, k.nbytes)
        assert lib.sb200_compute_descriptors_device(ex.handle, d_img, 256, 256, 256, d_k, 2, d_out) == 0
        assert lib.sb200_sync(ex.handle) == _ffi.E_INVALID
        out = np.zeros((2, 128), np.uint8)
        lib.sb200_memcpy_d2h(ex.handle, out.ctypes.data, d_out, 256)
        assert out[0].any() and not out[1].any()             # the valid keypoint was computed, the other one zeroed
        assert lib.sb200_sync(ex.handle) == 0                 # reported once
        for dptr in (d_img, d_k, d_out):
            lib.sb200_device_free(ex.handle, dptr)


@pytest.mark.parametrize("channels", [3, 4])
def test_rgb_input(sf, oracle, channels):
    """Input prep of the reference's callers (image::grayscale, examples/run-sift.rs:8) on the device: same luma
    bytes as the oracle's integer formula, and sift_rgb(rgb) == sift(luma(rgb))."""
    rng = np.random.default_rng(41)
    base = smooth_image(210, 170, 5).astype(np.int32)
    rgb = np.clip(base[..., None] + rng.integers(-40, 41, (170, 210, channels)), 0, 255).astype(np.uint8)
    luma = oracle.rgb_to_luma(rgb)
    with sf.Extractor(210, 170, 1) as ex:
        assert np.array_equal(ex.rgb_to_luma(rgb), luma)
        assert ex.sift_rgb(rgb) == ex.sift(luma)
    # extremes of the integer formula
    edge = np.array([[[255, 255, 255], [0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 253], [3, 2, 200]]], np.uint8)
    edge = np.repeat(np.repeat(edge, 8, 0), 2, 1)
    with sf.Extractor(16, 8, 1) as ex:
        assert np.array_equal(ex.rgb_to_luma(edge), oracle.rgb_to_luma(edge))


def _np_postfilter(kp, dedup, retain):
    idx = list(range(len(kp)))
    if dedup and len(idx) > 1:
        idx.sort(key=lambda i: (kp["x"][i], kp["y"][i], -kp["size"][i], kp["angle"][i], -kp["response"][i], i))
        out = [idx[0]]
        for j in idx[1:]:
            p, q = kp[out[-1]], kp[j]
            if (p["x"], p["y"], p["size"], p["angle"]) != (q["x"], q["y"], q["size"], q["angle"]):
                out.append(j)
        idx = out
    if retain is not None and len(idx) > retain:
        if retain == 0:
            return []
        thr = np.sort(kp["response"][idx])[::-1][retain - 1]
        idx = [i for i in idx if kp["response"][i] >= thr]
    return idx


@pytest.mark.parametrize("name", ["bird", "tree_small"])
def test_opencv_post_filters(sf, name):
    """The optional OpenCV-style post-filters: duplicate removal and retainBest as
    cv::SIFT::detectAndCompute applies them -- against a numpy restatement on the unfiltered result (exact) and
    against cv2.SIFT_create's own keypoint counts (the crate differs from OpenCV by well under 1 % of keypoints)."""
    cv2 = pytest.importorskip("cv2")
    g = load_gray(name)
    with sf.Extractor(g.shape[1], g.shape[0], 2) as ex:
        plain = ex.sift(g)
        for dedup, retain in ((True, None), (False, 100), (True, 200), (True, 0), (True, 10 ** 6)):
            ex.set_postfilter(dedup, retain)
            got = ex.sift(g)
            idx = _np_postfilter(plain.keypoint_array, dedup, retain)
            assert np.array_equal(got.keypoint_array, plain.keypoint_array[idx])
            assert np.array_equal(got.descriptors, plain.descriptors[idx])
        ex.set_postfilter(True, None)
        dd = ex.sift(g)
        offs, kp, de = ex.sift_batch(np.stack([g, g, g]))         # batches go through the same filter, image by image
        assert all(np.array_equal(kp[offs[i]:offs[i + 1]], dd.keypoint_array) for i in range(3))
        ex.set_postfilter(True, 300)
        best = ex.sift(g)
        ex.set_postfilter(False, None)
        assert ex.sift(g) == plain                                # and off again
    assert len(dd) < len(plain) or name == "bird"                # the crate keeps duplicates, OpenCV does not
    ncv = len(cv2.SIFT_create().detect(g, None))
    assert abs(len(dd) - ncv) <= max(3, 0.01 * ncv), (len(dd), ncv)
    ncv300 = len(cv2.SIFT_create(nfeatures=300).detect(g, None))
    assert abs(len(best) - ncv300) <= 3, (len(best), ncv300)
