"""sift_features_b200 -- B200-native drop-in for the extraction path of the Rust crate
tnibler/sift-features (reference sources cited as src/lib.rs:LINE).

Host-side mirror of the crate's public interface (same names, argument meaning and
result types) over the C ABI of include/sift_b200.h:

    sift(img, features_limit=None)                      src/lib.rs:71
    sift_with_processing(img, features_limit, P)        src/lib.rs:76
    precompute_images(img) -> PrecomputedImages         src/lib.rs:131
    sift_with_precomputed(pre, features_limit)          src/lib.rs:147
    compute_descriptor(img_f32, x, y, scale, ori)       src/lib.rs:785
    SiftResult / KeyPoint                               src/lib.rs:39-56

plus the batched / multi-GPU entry points the B200 build adds (sift_batch).
All computation happens in hand-written sm_100a CUDA kernels inside
libsift_b200.so; this package only marshals buffers.  There is NO CPU fallback:
without the built library or without a CUDA device every call raises.

Processing flavour (the crate's `P: Processing`, src/lib.rs:76-90), both implemented on the GPU:
  * OpenCVProcessing (src/opencv_processing.rs) -- the flavour the crate's test, snapshots and benches use;
    pinned bit-for-bit against OpenCV 4.13.  Default of `Extractor` and of every batched entry point.
  * ImageprocProcessing (src/lib.rs:992-1007) -- what the crate's plain `sift()` means (src/lib.rs:71-73), and
    therefore what the module-level `sift()` here selects.  Its arithmetic lives in the imageproc / image crates,
    which are not part of the reference tree: restated from their published algorithms, PARITY UNPINNED
    (bit-exact against this repository's oracle only).  Use `sift_with_processing(img, limit, OpenCVProcessing)`
    for the pinned flavour, exactly as the crate's own test does.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _ffi

__all__ = [
    "KeyPoint", "SiftResult", "PrecomputedImages", "Processing", "OpenCVProcessing", "ImageprocProcessing", "Extractor",
    "SiftError", "sift", "sift_with_processing", "precompute_images", "sift_with_precomputed",
    "compute_descriptor", "compute_descriptors", "sift_batch", "match", "MATCH_DTYPE", "KEYPOINT_DTYPE",
]

#: layout of sb200_keypoint == the crate's KeyPoint (src/lib.rs:48-56)
KEYPOINT_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4")])
SIFT_KEYPOINT_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                                ("octave", "i4"), ("scale", "i4")])
CANDIDATE_DTYPE = np.dtype([("octave", "i4"), ("scale", "i4"), ("y", "i4"), ("x", "i4")])
DESC_IN_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("scale", "f4"), ("orientation", "f4")])
MATCH_DTYPE = np.dtype([("queryIdx", "i4"), ("trainIdx", "i4"), ("distance", "f4")])   # cv::DMatch fields
STAGE_NAMES = ("seed", "blur", "extrema", "refine", "orient", "descriptor", "top_blur")


class SiftError(RuntimeError):
    """A C-ABI call returned a non-zero status (the crate would have panicked)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"sift_b200 status {status}: {message}")
        self.status = status


@dataclass(frozen=True)
class KeyPoint:
    """src/lib.rs:48-56.  x, y, size in input-image pixels; size is sigma (half of OpenCV's)."""
    x: float
    y: float
    size: float
    angle: float
    response: float


class SiftResult:
    """src/lib.rs:39-46: keypoints and the (n, 128) u8 descriptors in the same order."""

    def __init__(self, keypoints: np.ndarray, descriptors: np.ndarray):
        self.keypoint_array = keypoints            # structured array, KEYPOINT_DTYPE
        self.descriptors = descriptors             # (n, 128) uint8

    @property
    def keypoints(self) -> List[KeyPoint]:
        a = self.keypoint_array
        return [KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]))
                for k in a]

    def __len__(self) -> int:
        return len(self.keypoint_array)

    def __eq__(self, other) -> bool:  # #[derive(PartialEq)]
        return (isinstance(other, SiftResult) and np.array_equal(self.keypoint_array, other.keypoint_array)
                and np.array_equal(self.descriptors, other.descriptors))


class Processing:
    """src/lib.rs:86-90: the blur / resize implementation the pyramid is built with.  The crate's trait has three
    associated functions and static dispatch; here an implementation is a marker class whose FLAVOUR selects the
    device kernels (sb200_set_processing)."""
    FLAVOUR = -1


class OpenCVProcessing(Processing):
    """src/opencv_processing.rs:39-74: GaussianBlur(ksize = 0, sigma), resize INTER_LINEAR / INTER_NEAREST.  Pinned."""
    FLAVOUR = _ffi.PROCESSING_OPENCV


class ImageprocProcessing(Processing):
    """src/lib.rs:992-1007: imageproc gaussian_blur_f32, image resize Triangle / Nearest -- the crate's default.
    Restated from the crates' published algorithms; parity unpinned."""
    FLAVOUR = _ffi.PROCESSING_IMAGEPROC


def _flavour(processing) -> int:
    f = getattr(processing, "FLAVOUR", None)
    if f not in (_ffi.PROCESSING_OPENCV, _ffi.PROCESSING_IMAGEPROC):
        raise ValueError("processing must be OpenCVProcessing or ImageprocProcessing")
    return f


def _as_gray(img) -> np.ndarray:
    a = np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("expected a 2-D uint8 (gray) image, like image::GrayImage")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


class Extractor:
    """One sb200 context: a device, its arenas and streams.  Not re-entrant (use one per thread)."""

    def __init__(self, max_width: int, max_height: int, max_batch: int = 1, device: int = 0,
                 max_keypoints_per_image: int = 0, processing=OpenCVProcessing):
        self._lib = _ffi.load()
        self.processing = processing
        flavour = _flavour(processing)
        self._h = C.c_void_p()
        st = self._lib.sb200_create(device, max_width, max_height, max_batch, max_keypoints_per_image,
                                    C.byref(self._h))
        if st != _ffi.OK:
            raise SiftError(st, f"sb200_create(device={device}, {max_width}x{max_height}, batch={max_batch}) failed: "
                            + self._lib.sb200_status_string(st).decode())
        self.device, self.max_width, self.max_height, self.max_batch = device, max_width, max_height, max_batch
        self.max_keypoints_per_image = max_keypoints_per_image or max(16384, max_width * max_height // 8)
        self.auto_grow = max_keypoints_per_image == 0   # an explicit capacity is a hard limit
        self._check(self._lib.sb200_set_processing(self._h, flavour))
        self._generation = 0   # bumped by every call that replaces the pyramid resident in the context

    def _grow(self) -> bool:
        """Rebuilds the context with 4x the per-image candidate/keypoint capacity (bounded by the number of
        (pixel, scale) positions the detector scans).  The crate has no such limit; this keeps the drop-in
        from failing on images that are far denser than the default sizing assumes."""
        worst = 16 * self.max_width * self.max_height
        if not self.auto_grow or self.max_keypoints_per_image >= worst:
            return False
        cap = min(self.max_keypoints_per_image * 4, worst)
        self._lib.sb200_destroy(self._h)
        self._h = C.c_void_p()
        st = self._lib.sb200_create(self.device, self.max_width, self.max_height, self.max_batch, cap, C.byref(self._h))
        if st != _ffi.OK:
            raise SiftError(st, "re-creating the context with a larger capacity failed")
        self._check(self._lib.sb200_set_processing(self._h, _flavour(self.processing)))
        if getattr(self, "_postfilter", None):
            self.set_postfilter(*self._postfilter)
        self.max_keypoints_per_image = cap
        return True

    def _retry_capacity(self, call):
        self._generation += 1
        while True:
            st = call()
            if st == _ffi.E_CAPACITY and self._grow():
                continue
            self._check(st)
            return

    def set_postfilter(self, remove_duplicates: bool = False, retain_best: Optional[int] = None):
        """OpenCV-style post-filters of the host results (sb200_set_postfilter): what cv::SIFT does after detection
        and the crate does not -- KeyPointsFilter::removeDuplicatedSorted and retainBest(n).  Off by default."""
        self._postfilter = (bool(remove_duplicates), retain_best)
        self._check(self._lib.sb200_set_postfilter(self._h, int(bool(remove_duplicates)),
                                                   -1 if retain_best is None else int(retain_best)))

    # -- plumbing ---------------------------------------------------------
    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def _check(self, st: int):
        if st != _ffi.OK:
            raise SiftError(st, self._lib.sb200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.sb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _take(self, res: _ffi.Result):
        n, ni = int(res.n), int(res.n_images)
        offs = np.ctypeslib.as_array(res.offsets, shape=(ni + 1,)).copy()
        if n:
            kp = np.frombuffer((C.c_char * (n * KEYPOINT_DTYPE.itemsize)).from_address(res.keypoints),
                               dtype=KEYPOINT_DTYPE).copy()
            de = np.frombuffer((C.c_char * (n * 128)).from_address(res.descriptors), dtype=np.uint8)
            de = de.reshape(n, 128).copy()
        else:
            kp = np.zeros(0, KEYPOINT_DTYPE)
            de = np.zeros((0, 128), np.uint8)
        return offs, kp, de

    # -- the crate's entry points ----------------------------------------
    def sift(self, img, features_limit: Optional[int] = None) -> SiftResult:
        """sift_with_processing::<P> (src/lib.rs:76-81) with the extractor's Processing P."""
        a = _as_gray(img)
        res = _ffi.Result()
        self._retry_capacity(lambda: self._lib.sb200_extract(
            self._h, a.ctypes.data, a.shape[1], a.shape[0], a.strides[0],
            -1 if features_limit is None else int(features_limit), C.byref(res)))
        _, kp, de = self._take(res)
        return SiftResult(kp, de)

    def sift_batch(self, images, features_limit: Optional[int] = None):
        """n same-sized images (n,H,W) u8 -> (offsets[n+1], keypoints, descriptors)."""
        a = np.asarray(images)
        if a.dtype != np.uint8 or a.ndim != 3:
            raise ValueError("expected an (n, H, W) uint8 array")
        if a.strides[2] != 1 or a.strides[1] < a.shape[2]:
            a = np.ascontiguousarray(a)
        res = _ffi.Result()
        self._retry_capacity(lambda: self._lib.sb200_extract_batch(
            self._h, a.ctypes.data, a.shape[0], a.shape[2], a.shape[1], a.strides[1], a.strides[0],
            -1 if features_limit is None else int(features_limit), C.byref(res)))
        return self._take(res)

    def sift_rgb(self, rgb, features_limit: Optional[int] = None) -> SiftResult:
        """sift() on an interleaved 8-bit RGB / RGBA image (H, W, 3|4): the `image` crate's grayscale() the reference's
        callers run first (examples/run-sift.rs:8) is done on the device."""
        a = np.ascontiguousarray(rgb)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] not in (3, 4):
            raise ValueError("expected an (H, W, 3|4) uint8 array")
        res = _ffi.Result()
        self._retry_capacity(lambda: self._lib.sb200_extract_batch_rgb(
            self._h, a.ctypes.data, 1, a.shape[1], a.shape[0], a.strides[0], a.strides[0] * a.shape[0], a.shape[2],
            -1 if features_limit is None else int(features_limit), C.byref(res)))
        _, kp, de = self._take(res)
        return SiftResult(kp, de)

    def rgb_to_luma(self, rgb) -> np.ndarray:
        """The device-side RGB(A) -> luma conversion alone."""
        a = np.ascontiguousarray(rgb)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] not in (3, 4):
            raise ValueError("expected an (H, W, 3|4) uint8 array")
        out = np.zeros(a.shape[:2], np.uint8)
        self._check(self._lib.sb200_rgb_to_luma(self._h, a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], a.shape[2],
                                                out.ctypes.data))
        return out

    def sift_jpeg(self, jpegs, features_limit: Optional[int] = None):
        """JPEG bitstreams (bytes-like objects, one frame size) -> (offsets[n+1], keypoints, descriptors): nvJPEG decode
        and the integer luma on the device, then sift() -- `image::open(..).grayscale()` + sift of examples/run-sift.rs:8-19
        without host pixels."""
        bufs = [np.frombuffer(j, np.uint8) for j in jpegs]
        if not bufs:
            raise ValueError("no images")
        ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        lens = (C.c_uint64 * len(bufs))(*[b.size for b in bufs])
        res = _ffi.Result()
        self._retry_capacity(lambda: self._lib.sb200_extract_batch_jpeg(
            self._h, ptrs, lens, len(bufs), -1 if features_limit is None else int(features_limit), C.byref(res)))
        return self._take(res)

    def jpeg_info(self, jpeg):
        """(width, height, components) from the JPEG header."""
        b = np.frombuffer(jpeg, np.uint8)
        w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._lib.sb200_jpeg_info(self._h, b.ctypes.data, b.size, C.byref(w), C.byref(h), C.byref(c)))
        return w.value, h.value, c.value

    def decode_jpeg_luma(self, jpeg) -> np.ndarray:
        """The device-side decode + luma step alone: the GrayImage sift_jpeg() runs on."""
        b = np.frombuffer(jpeg, np.uint8)
        w, h, _ = self.jpeg_info(jpeg)
        out = np.zeros((h, w), np.uint8)
        self._check(self._lib.sb200_decode_jpeg_luma(self._h, b.ctypes.data, b.size, out.ctypes.data, out.size))
        return out

    @property
    def jpeg_backend(self) -> str:
        return self._lib.sb200_jpeg_backend(self._h).decode()

    def precompute_images(self, img) -> "PrecomputedImages":
        """precompute_images::<P> (src/lib.rs:131-143) with the extractor's Processing; the pyramid stays on the device."""
        a = _as_gray(img)
        self._generation += 1
        self._check(self._lib.sb200_precompute(self._h, a.ctypes.data, a.shape[1], a.shape[0], a.strides[0]))
        return PrecomputedImages(self)

    def sift_with_precomputed(self, features_limit: Optional[int] = None, pre: "Optional[PrecomputedImages]" = None) -> SiftResult:
        """sift_with_precomputed (src/lib.rs:147-177) on the resident pyramid.  `pre`, when given, must still be the
        pyramid resident in this context (a later sift / precompute call on the same context replaces it)."""
        if pre is not None:
            pre._check_resident()
        res = _ffi.Result()
        self._check(self._lib.sb200_extract_precomputed(
            self._h, -1 if features_limit is None else int(features_limit), C.byref(res)))
        _, kp, de = self._take(res)
        return SiftResult(kp, de)

    def compute_descriptors(self, img_f32, keypoints) -> np.ndarray:
        """compute_descriptor (src/lib.rs:785-990) for many (x, y, scale, orientation_deg) rows."""
        img = np.ascontiguousarray(img_f32, np.float32)
        k = np.ascontiguousarray(np.asarray(keypoints, np.float32).reshape(-1, 4))
        out = np.zeros((len(k), 128), np.uint8)
        self._check(self._lib.sb200_compute_descriptors(self._h, img.ctypes.data, img.shape[1], img.shape[0],
                                                        img.shape[1], k.ctypes.data, len(k), out.ctypes.data))
        return out

    def match(self, query_descriptors, train_descriptors) -> np.ndarray:
        """Mutual nearest neighbours of two (N,128) u8 descriptor matrices -- what the reference's examples get
        from OpenCV's BFMatcher(NORM_L2, crossCheck=True).match(query, train) (examples/sift-match.rs:30-35).
        Returns MATCH_DTYPE rows (queryIdx, trainIdx, distance) in ascending query order."""
        q = np.ascontiguousarray(query_descriptors, np.uint8).reshape(-1, 128)
        t = np.ascontiguousarray(train_descriptors, np.uint8).reshape(-1, 128)
        raw = np.zeros(len(q), np.dtype([("query", np.uint32), ("train", np.uint32), ("dist2", np.uint32)]))
        n = C.c_uint64()
        self._check(self._lib.sb200_match_descriptors(self._h, q.ctypes.data, len(q), t.ctypes.data, len(t), raw.ctypes.data,
                                          len(raw), C.byref(n)))
        raw = raw[: n.value]
        out = np.zeros(len(raw), MATCH_DTYPE)
        out["queryIdx"], out["trainIdx"] = raw["query"], raw["train"]
        out["distance"] = np.sqrt(raw["dist2"].astype(np.float64)).astype(np.float32)
        return out

    # -- parity / debug views ------------------------------------------
    def last_candidates(self) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._lib.sb200_last_candidates(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, CANDIDATE_DTYPE)
        if n.value:
            self._check(self._lib.sb200_last_candidates(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def last_sift_keypoints(self) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._lib.sb200_last_sift_keypoints(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, SIFT_KEYPOINT_DTYPE)
        if n.value:
            self._check(self._lib.sb200_last_sift_keypoints(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    # -- measurement ----------------------------------------------------
    def set_profiling(self, on: bool):
        self._check(self._lib.sb200_set_profiling(self._h, int(on)))

    def stage_stats(self):
        ms = (C.c_double * _ffi.STAGE_COUNT)()
        ln = (C.c_uint64 * _ffi.STAGE_COUNT)()
        self._check(self._lib.sb200_stage_stats(self._h, ms, ln, _ffi.STAGE_COUNT))
        return {STAGE_NAMES[i]: {"ms": ms[i], "launches": int(ln[i])} for i in range(_ffi.STAGE_COUNT)}

    def reset_stats(self):
        self._check(self._lib.sb200_reset_stats(self._h))

    def launch_stats(self):
        """Per-launch view of the pyramid stages while profiling is on: {(octave, kind): (ms, launches)} with kind in
        "seed", "blur1".."blur5", "extrema", "tail" (sb200_launch_stats)."""
        n = _ffi.FINE_SLOTS
        ms = (C.c_double * n)()
        ln = (C.c_uint64 * n)()
        self._check(self._lib.sb200_launch_stats(self._h, ms, ln, n))
        kinds = ("seed", "blur1", "blur2", "blur3", "blur4", "blur5", "extrema", "tail")
        return {(i // 8, kinds[i % 8]): (ms[i], int(ln[i])) for i in range(n) if ln[i]}

    @property
    def launch_count(self) -> int:
        return int(self._lib.sb200_launch_count(self._h))


class PrecomputedImages:
    """src/lib.rs:124-128: `scale_space[o]` is the (6,h,w) Gaussian stack, `dog[o]` the (5,h,w) DoG stack of
    octave o.  The data lives on the device; indexing downloads the requested octave."""

    class _Stack:
        def __init__(self, pre, ex: Extractor, dims, layers, fn):
            self._pre, self._ex, self._dims, self._layers, self._fn = pre, ex, dims, layers, fn

        def __len__(self):
            return len(self._dims)

        def __getitem__(self, o: int) -> np.ndarray:
            self._pre._check_resident()
            w, h = self._dims[o]
            out = np.zeros((self._layers, h, w), np.float32)
            for l in range(self._layers):
                self._ex._check(self._fn(self._ex.handle, o, l, out[l].ctypes.data))
            return out

    def __init__(self, ex: Extractor, owns_extractor: bool = False):
        lib = ex._lib
        self._generation = ex._generation
        self._owns = owns_extractor
        n = C.c_uint32()
        ws = (C.c_uint32 * 16)()
        hs = (C.c_uint32 * 16)()
        ex._check(lib.sb200_pyramid_info(ex.handle, C.byref(n), ws, hs, 16))
        self.extractor = ex
        self.n_octaves = int(n.value)
        self.dims = [(int(ws[o]), int(hs[o])) for o in range(self.n_octaves)]
        self.scale_space = PrecomputedImages._Stack(self, ex, self.dims, 6, lib.sb200_pyramid_layer)
        self.dog = PrecomputedImages._Stack(self, ex, self.dims, 5, lib.sb200_pyramid_dog)

    def _check_resident(self):
        """In the crate PrecomputedImages is an owned value; here it is a view of the pyramid resident in a context.
        A later call on the same context replaces that pyramid -- using the stale view is an error, not another
        image's results."""
        if self.extractor._generation != self._generation:
            raise SiftError(_ffi.E_STATE, "this PrecomputedImages is stale: a later call on its Extractor replaced the "
                            "resident pyramid (precompute_images() at module level gives every result its own context)")

    def close(self):
        if self._owns:
            self.extractor.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# module-level functions with the crate's names; contexts are cached per shape
# ---------------------------------------------------------------------------
_cache: dict = {}
_CACHE_MAX = 8   # contexts hold device arenas: keep only a few alive


def _extractor(w: int, h: int, batch: int = 1, device: int = 0, keep=(), processing=OpenCVProcessing) -> Extractor:
    """Context for (device, shape, batch[, processing]) from a small LRU cache.  `keep`: keys the current call still
    uses -- they are never evicted (a call that shards over more devices than the cache normally holds grows it for
    its duration)."""
    key = (device, w, h, batch) if processing is OpenCVProcessing else (device, w, h, batch, _flavour(processing))
    ex = _cache.pop(key, None)
    if ex is None:
        while len(_cache) >= _CACHE_MAX:
            victim = next((k for k in _cache if k not in keep), None)
            if victim is None:
                break
            _cache.pop(victim).close()
        ex = Extractor(w, h, batch, device) if processing is OpenCVProcessing else Extractor(w, h, batch, device, 0, processing)
    _cache[key] = ex   # most recently used last
    return ex


def sift(img, features_limit: Optional[int] = None, device: int = 0) -> SiftResult:
    """src/lib.rs:71-73: `sift_with_processing::<ImageprocProcessing>(img, features_limit)`, like the crate.  (That
    flavour's arithmetic is restated from the imageproc / image crates and unpinned; the flavour the crate's own test
    pins is `sift_with_processing(img, limit, OpenCVProcessing)`.)"""
    return sift_with_processing(img, features_limit, ImageprocProcessing, device)


def sift_with_processing(img, features_limit: Optional[int] = None, processing=OpenCVProcessing,
                         device: int = 0) -> SiftResult:
    """src/lib.rs:76.  `processing` selects the blur/resize flavour (the crate's type parameter P)."""
    a = _as_gray(img)
    return _extractor(a.shape[1], a.shape[0], 1, device, processing=processing).sift(a, features_limit)


def precompute_images(img, processing=OpenCVProcessing, device: int = 0) -> PrecomputedImages:
    """src/lib.rs:131: precompute_images::<P>(img).  The crate returns an owned value; so does this: the result holds
    its own context (never shared with the module-level cache), released when the result is dropped."""
    a = _as_gray(img)
    ex = Extractor(a.shape[1], a.shape[0], 1, device, 0, processing)
    ex._generation += 1
    ex._check(ex._lib.sb200_precompute(ex.handle, a.ctypes.data, a.shape[1], a.shape[0], a.strides[0]))
    return PrecomputedImages(ex, owns_extractor=True)


def sift_with_precomputed(pre: PrecomputedImages, features_limit: Optional[int] = None) -> SiftResult:
    """src/lib.rs:147."""
    return pre.extractor.sift_with_precomputed(features_limit, pre)


def match(query_descriptors, train_descriptors, device: int = 0) -> np.ndarray:
    """BFMatcher(NORM_L2, crossCheck=True).match(query, train) of examples/sift-match.rs:30-35 on the GPU."""
    return _extractor(8, 8, 1, device).match(query_descriptors, train_descriptors)


def compute_descriptors(img_f32, keypoints, device: int = 0) -> np.ndarray:
    img = np.asarray(img_f32)
    return _extractor(max(img.shape[1], 8), max(img.shape[0], 8), 1, device).compute_descriptors(img, keypoints)


def compute_descriptor(img_f32, x: float, y: float, scale: float, orientation: float, device: int = 0) -> np.ndarray:
    """src/lib.rs:785: one keypoint; returns the 128 descriptor bytes."""
    return compute_descriptors(img_f32, [[x, y, scale, orientation]], device)[0]


def sift_batch(images, features_limit: Optional[int] = None, devices: Optional[Sequence[int]] = None,
               max_batch: int = 16, dense: bool = False, processing=OpenCVProcessing) -> List[SiftResult]:
    """n same-sized images -> one SiftResult per image.  With several devices the batch is split into
    contiguous shards, one host thread per device (no device-to-device traffic, no collective).  The per-image
    results are cut out of each device's own result arrays (sb200_extract_batch_multi_parts: nothing is gathered
    on the host); dense=True goes through sb200_extract_batch_multi, which first concatenates the parts into one
    dense array."""
    a = np.asarray(images)
    if a.dtype != np.uint8 or a.ndim != 3:
        raise ValueError("expected an (n, H, W) uint8 array")
    a = np.ascontiguousarray(a)
    devices = list(devices) if devices else [0]
    n, h, w = a.shape
    b = max(1, min(max_batch, n))
    keys = {(d, w, h, b) if processing is OpenCVProcessing else (d, w, h, b, _flavour(processing)) for d in devices}
    exs = [_extractor(w, h, b, d, keep=keys, processing=processing) for d in devices]
    lib = _ffi.load()
    lim = -1 if features_limit is None else int(features_limit)
    if len(exs) == 1:
        offs, kp, de = exs[0].sift_batch(a, features_limit)
        return [SiftResult(kp[offs[i]:offs[i + 1]], de[offs[i]:offs[i + 1]]) for i in range(n)]
    handles = (C.c_void_p * len(exs))(*[e.handle for e in exs])
    for e in exs:
        e._generation += 1
    if dense:
        res = _ffi.Result()
        exs[0]._check(lib.sb200_extract_batch_multi(handles, len(exs), a.ctypes.data, n, w, h, a.strides[1], a.strides[0],
                                                    lim, C.byref(res)))
        offs, kp, de = exs[0]._take(res)
        return [SiftResult(kp[offs[i]:offs[i + 1]], de[offs[i]:offs[i + 1]]) for i in range(n)]
    parts = (_ffi.Result * len(exs))()
    first = (C.c_uint64 * (len(exs) + 1))()
    exs[0]._check(lib.sb200_extract_batch_multi_parts(handles, len(exs), a.ctypes.data, n, w, h, a.strides[1],
                                                      a.strides[0], lim, parts, first))
    out: List[SiftResult] = []
    for d, e in enumerate(exs):
        if parts[d].n_images == 0:
            continue
        assert len(out) == first[d]
        offs, kp, de = e._take(parts[d])
        out += [SiftResult(kp[offs[i]:offs[i + 1]], de[offs[i]:offs[i + 1]]) for i in range(int(parts[d].n_images))]
    return out


def shard_ranges(n: int, parts: int) -> List[range]:
    """The partition sb200_extract_batch_multi uses: contiguous blocks of ceil(n/parts) images."""
    per = (n + parts - 1) // parts
    return [range(min(d * per, n), min((d + 1) * per, n)) for d in range(parts)]


def algorithmic_bytes(w: int, h: int):
    """A(W,H) of SURVEY.md section 8(d): (total, seed, blur+decimate, extrema) bytes for one image."""
    lib = _ffi.load()
    parts = (C.c_uint64 * 3)()
    total = lib.sb200_algorithmic_bytes(w, h, parts, 3)
    return int(total), int(parts[0]), int(parts[1]), int(parts[2])
