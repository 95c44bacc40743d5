"""ctypes loader for the CPU ORACLE (oracle/libsift_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never from the product
package (sift_features_b200/).  The library is a C restatement of
/root/reference/src/lib.rs (+ src/opencv_processing.rs); see sift_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsift_oracle.so")


class SiftKeyPoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("scale", C.c_int32)]


SIFT_KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"),
                          ("response", "f4"), ("octave", "i4"), ("scale", "i4")])
KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4")])
CAND_DTYPE = np.dtype([("octave", "i4"), ("scale", "i4"), ("y", "i4"), ("x", "i4")])


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sift_oracle.c")
    hdr = os.path.join(_HERE, "sift_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libsift_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    L.so_gaussian_ksize.restype = C.c_int
    L.so_gaussian_ksize.argtypes = [C.c_double]
    L.so_gaussian_taps.restype = C.c_int
    L.so_gaussian_taps.argtypes = [C.c_double, f32p, C.c_int]
    L.so_gaussian_blur.restype = None
    L.so_gaussian_blur.argtypes = [f32p, C.c_int, C.c_int, C.c_double, f32p]
    L.so_resize_linear_2x.restype = None
    L.so_resize_linear_2x.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.so_resize_nearest_half.restype = None
    L.so_resize_nearest_half.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.so_seed_sigma.restype = C.c_double
    L.so_octave_sigma.restype = C.c_double
    L.so_octave_sigma.argtypes = [C.c_int]
    L.so_precompute.restype = C.c_void_p
    L.so_precompute.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
    L.so_precompute_flavour.restype = C.c_void_p
    L.so_precompute_flavour.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.so_imageproc_taps.restype = C.c_int
    L.so_imageproc_taps.argtypes = [C.c_double, f32p, C.c_int]
    L.so_gaussian_blur_imageproc.restype = None
    L.so_gaussian_blur_imageproc.argtypes = [f32p, C.c_int, C.c_int, C.c_double, f32p]
    L.so_resize_triangle_2x.restype = None
    L.so_resize_triangle_2x.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.so_resize_nearest_imageproc.restype = None
    L.so_resize_nearest_imageproc.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.so_pyramid_free.restype = None
    L.so_pyramid_free.argtypes = [C.c_void_p]
    for name in ("so_pyramid_octaves",):
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = [C.c_void_p]
    for name in ("so_pyramid_width", "so_pyramid_height"):
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
    for name in ("so_pyramid_gauss", "so_pyramid_dog"):
        getattr(L, name).restype = C.POINTER(C.c_float)
        getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.so_find_candidates.restype = C.c_size_t
    L.so_find_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.so_find_keypoints.restype = C.c_size_t
    L.so_find_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.so_compute_descriptor.restype = None
    L.so_compute_descriptor.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_float, u8p]
    L.so_sift_with_precomputed.restype = C.c_size_t
    L.so_sift_with_precomputed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t]
    L.so_sift.restype = C.c_size_t
    L.so_sift.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t]
    _lib = L
    return L


# ---- Processing flavour A -------------------------------------------------
def gaussian_taps(sigma: float) -> np.ndarray:
    ks = lib().so_gaussian_ksize(sigma)
    t = np.zeros(ks, np.float32)
    lib().so_gaussian_taps(sigma, t, ks)
    return t


def gaussian_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty_like(img)
    lib().so_gaussian_blur(img, img.shape[1], img.shape[0], float(sigma), out)
    return out


def resize_linear_2x(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty((img.shape[0] * 2, img.shape[1] * 2), np.float32)
    lib().so_resize_linear_2x(img, img.shape[1], img.shape[0], out)
    return out


def resize_nearest_half(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty((img.shape[0] // 2, img.shape[1] // 2), np.float32)
    lib().so_resize_nearest_half(img, img.shape[1], img.shape[0], out)
    return out


def seed_sigma() -> float:
    return lib().so_seed_sigma()


def octave_sigma(s: int) -> float:
    return lib().so_octave_sigma(s)


# ---- Processing flavour B (ImageprocProcessing, src/lib.rs:992-1007; PARITY UNPINNED, see sift_oracle.h) ----
PROCESSING_OPENCV, PROCESSING_IMAGEPROC = 0, 1


def imageproc_taps(sigma: float) -> np.ndarray:
    t = np.zeros(64, np.float32)
    n = lib().so_imageproc_taps(sigma, t, 64)
    return t[:n].copy()


def gaussian_blur_imageproc(img: np.ndarray, sigma: float) -> np.ndarray:
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty_like(img)
    lib().so_gaussian_blur_imageproc(img, img.shape[1], img.shape[0], sigma, out)
    return out


def resize_triangle_2x(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty((img.shape[0] * 2, img.shape[1] * 2), np.float32)
    lib().so_resize_triangle_2x(img, img.shape[1], img.shape[0], out)
    return out


def resize_nearest_imageproc(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty((img.shape[0] // 2, img.shape[1] // 2), np.float32)
    lib().so_resize_nearest_imageproc(img, img.shape[1], img.shape[0], out)
    return out


# ---- pyramid ---------------------------------------------------------------
class Pyramid:
    """Mirror of PrecomputedImages (src/lib.rs:124-128)."""

    def __init__(self, gray: np.ndarray, processing: int = PROCESSING_OPENCV):
        gray = np.ascontiguousarray(gray, np.uint8)
        assert gray.ndim == 2
        self._h = lib().so_precompute_flavour(gray, gray.shape[1], gray.shape[0], gray.shape[1], processing)
        self.n_octaves = lib().so_pyramid_octaves(self._h)
        self.dims = [(lib().so_pyramid_width(self._h, o), lib().so_pyramid_height(self._h, o))
                     for o in range(self.n_octaves)]

    def _view(self, fn, o, l):
        w, h = self.dims[o]
        if w == 0 or h == 0:
            return np.zeros((h, w), np.float32)
        p = fn(self._h, o, l)
        return np.ctypeslib.as_array(p, shape=(h, w))

    def gauss(self, o: int, l: int) -> np.ndarray:
        return self._view(lib().so_pyramid_gauss, o, l)

    def dog(self, o: int, l: int) -> np.ndarray:
        return self._view(lib().so_pyramid_dog, o, l)

    def candidates(self) -> np.ndarray:
        n = lib().so_find_candidates(self._h, None, 0)
        out = np.zeros(n, CAND_DTYPE)
        if n:
            lib().so_find_candidates(self._h, out.ctypes.data, n)
        return out

    def sift_keypoints(self) -> np.ndarray:
        n = lib().so_find_keypoints(self._h, None, 0)
        out = np.zeros(n, SIFT_KP_DTYPE)
        if n:
            lib().so_find_keypoints(self._h, out.ctypes.data, n)
        return out

    def sift(self, features_limit: int | None = None):
        """sift_with_precomputed (src/lib.rs:147-177)."""
        lim = -1 if features_limit is None else int(features_limit)
        n = lib().so_sift_with_precomputed(self._h, lim, None, None, 0)
        kps = np.zeros(n, KP_DTYPE)
        desc = np.zeros((n, 128), np.uint8)
        if n:
            lib().so_sift_with_precomputed(self._h, lim, kps.ctypes.data, desc.ctypes.data, n)
        return kps, desc

    def close(self):
        if self._h:
            lib().so_pyramid_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sift(gray: np.ndarray, features_limit: int | None = None, processing: int = PROCESSING_OPENCV):
    """sift_with_processing::<P> (src/lib.rs:76-81); P = OpenCVProcessing unless processing says otherwise."""
    p = Pyramid(gray, processing)
    try:
        return p.sift(features_limit)
    finally:
        p.close()


def compute_descriptor(img: np.ndarray, x: float, y: float, scale: float, orientation: float) -> np.ndarray:
    """compute_descriptor (src/lib.rs:785-990)."""
    img = np.ascontiguousarray(img, np.float32)
    out = np.zeros(128, np.uint8)
    lib().so_compute_descriptor(img, img.shape[1], img.shape[0], x, y, scale, orientation, out)
    return out


def match_cross_check(query: np.ndarray, train: np.ndarray) -> np.ndarray:
    """What examples/sift-match.rs:30-35 and examples/opencv-cross-match.rs:34-43 ask of OpenCV:
    BFMatcher(NORM_L2, crossCheck=true).match(query, train) over (N,128) u8 descriptor matrices -- for every query
    row the nearest train row, kept when the query row is in turn the nearest of that train row.  Exact integer
    squared distances; ties resolve to the smallest index.  Returns rows (query, train, dist2), ascending query."""
    q = np.ascontiguousarray(query, np.uint8).reshape(-1, 128).astype(np.int64)
    t = np.ascontiguousarray(train, np.uint8).reshape(-1, 128).astype(np.int64)
    out = np.zeros((0, 3), np.int64)
    if len(q) == 0 or len(t) == 0:
        return out
    best_q = np.empty(len(q), np.int64)
    best_d = np.empty(len(q), np.int64)
    col_min = np.full(len(t), np.iinfo(np.int64).max)
    col_arg = np.zeros(len(t), np.int64)
    tn = (t * t).sum(1)
    for s in range(0, len(q), 2048):                      # blocked so that 100k x 100k stays in memory
        blk = q[s:s + 2048]
        d = (blk * blk).sum(1)[:, None] + tn[None, :] - 2 * (blk @ t.T)
        best_q[s:s + len(blk)] = d.argmin(1)
        best_d[s:s + len(blk)] = d.min(1)
        cm, ca = d.min(0), d.argmin(0) + s
        upd = cm < col_min                                # strict: earlier blocks (smaller query index) win ties
        col_min[upd], col_arg[upd] = cm[upd], ca[upd]
    keep = col_arg[best_q] == np.arange(len(q))
    idx = np.nonzero(keep)[0]
    return np.stack([idx, best_q[idx], best_d[idx]], 1)


def rgb_to_luma(rgb: np.ndarray) -> np.ndarray:
    """DynamicImage::grayscale() / to_luma8() of the `image` crate (0.25: color.rs rgb_to_luma, SRGB_LUMA = [2126, 7152,
    722], SRGB_LUMA_DIV = 10000, u32 arithmetic, truncating division), which the reference's callers apply before sift()
    (examples/run-sift.rs:8, src/lib.rs:1012).  The crate's source is not in the reference tree: restated from its
    published algorithm, unpinned."""
    a = np.asarray(rgb).astype(np.uint32)
    return ((2126 * a[..., 0] + 7152 * a[..., 1] + 722 * a[..., 2]) // 10000).astype(np.uint8)
