"""ctypes binding of the C ABI in include/sift_b200.h (libsift_b200.so).

This is the same stub a maintainer of the reference crate would write as an
`extern "C"` block; nothing here computes anything.  If the
library is missing or cannot be loaded the import fails loudly -- there is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SB200_LIB: an alternative build of the same library (development A/B runs); the default is the in-tree build
LIB_PATH = os.environ.get("SB200_LIB") or os.path.join(_HERE, "libsift_b200.so")

OK, E_INVALID, E_CUDA, E_CAPACITY, E_STATE = 0, 1, 2, 3, 4
DESC_SIZE = 128
STAGE_COUNT = 7
PROCESSING_OPENCV, PROCESSING_IMAGEPROC = 0, 1
FINE_SLOTS = 128

# every symbol include/sift_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "sb200_create", "sb200_destroy", "sb200_last_error", "sb200_status_string", "sb200_device_count",
    "sb200_set_processing", "sb200_get_processing", "sb200_set_postfilter", "sb200_extract", "sb200_extract_batch", "sb200_extract_batch_device", "sb200_pyramid_batch_device", "sb200_device_result", "sb200_sync",
    "sb200_precompute", "sb200_extract_precomputed", "sb200_pyramid_info", "sb200_pyramid_layer",
    "sb200_pyramid_dog", "sb200_last_candidates", "sb200_last_sift_keypoints", "sb200_compute_descriptors",
    "sb200_compute_descriptors_device", "sb200_extract_batch_multi", "sb200_extract_batch_multi_parts",
    "sb200_last_gather_ms", "sb200_last_shard_ms", "sb200_set_profiling", "sb200_stage_stats",
    "sb200_reset_stats", "sb200_launch_stats", "sb200_launch_count", "sb200_stage_name", "sb200_algorithmic_bytes", "sb200_timer_start",
    "sb200_timer_stop", "sb200_timer_elapsed_ms", "sb200_host_alloc", "sb200_host_free", "sb200_device_alloc",
    "sb200_device_free", "sb200_memcpy_h2d", "sb200_memcpy_d2h", "sb200_flush_l2", "sb200_match_descriptors", "sb200_match_descriptors_device",
    "sb200_extract_batch_rgb", "sb200_rgb_to_luma", "sb200_extract_batch_jpeg", "sb200_jpeg_info", "sb200_decode_jpeg_luma",
    "sb200_jpeg_backend",
]


class Result(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_images", C.c_uint32), ("offsets", C.POINTER(C.c_uint64)),
                ("keypoints", C.c_void_p), ("descriptors", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    """Loads libsift_b200.so; raises if it has not been built (python -m sift_features_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m sift_features_b200.build` "
                          "(needs nvcc; the B200 path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u8p, u32, u64, i64 = C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64
    sig = {
        "sb200_create": (C.c_int, [C.c_int, u32, u32, u32, u32, C.POINTER(vp)]),
        "sb200_destroy": (None, [vp]),
        "sb200_last_error": (C.c_char_p, [vp]),
        "sb200_status_string": (C.c_char_p, [C.c_int]),
        "sb200_device_count": (C.c_int, []),
        "sb200_set_processing": (C.c_int, [vp, C.c_int]),
        "sb200_get_processing": (C.c_int, [vp]),
        "sb200_set_postfilter": (C.c_int, [vp, C.c_int, i64]),
        "sb200_extract": (C.c_int, [vp, u8p, u32, u32, u32, i64, C.POINTER(Result)]),
        "sb200_extract_batch": (C.c_int, [vp, u8p, u32, u32, u32, u32, u64, i64, C.POINTER(Result)]),
        "sb200_extract_batch_device": (C.c_int, [vp, vp, u32, u32, u32, u32, u64, i64]),
        "sb200_pyramid_batch_device": (C.c_int, [vp, vp, u32, u32, u32, u32, u64]),
        "sb200_device_result": (C.c_int, [vp, vp, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(u32)]),
        "sb200_sync": (C.c_int, [vp]),
        "sb200_precompute": (C.c_int, [vp, u8p, u32, u32, u32]),
        "sb200_extract_precomputed": (C.c_int, [vp, i64, C.POINTER(Result)]),
        "sb200_pyramid_info": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), u32]),
        "sb200_pyramid_layer": (C.c_int, [vp, u32, u32, vp]),
        "sb200_pyramid_dog": (C.c_int, [vp, u32, u32, vp]),
        "sb200_last_candidates": (C.c_int, [vp, vp, u64, C.POINTER(u64)]),
        "sb200_last_sift_keypoints": (C.c_int, [vp, vp, u64, C.POINTER(u64)]),
        "sb200_compute_descriptors": (C.c_int, [vp, vp, u32, u32, u32, vp, u64, vp]),
        "sb200_compute_descriptors_device": (C.c_int, [vp, vp, u32, u32, u32, vp, u64, vp]),
        "sb200_extract_batch_rgb": (C.c_int, [vp, u8p, u32, u32, u32, u32, u64, u32, i64, C.POINTER(Result)]),
        "sb200_rgb_to_luma": (C.c_int, [vp, u8p, u32, u32, u32, u32, vp]),
        "sb200_extract_batch_jpeg": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64), u32, i64, C.POINTER(Result)]),
        "sb200_jpeg_info": (C.c_int, [vp, vp, u64, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "sb200_decode_jpeg_luma": (C.c_int, [vp, vp, u64, vp, u64]),
        "sb200_jpeg_backend": (C.c_char_p, [vp]),
        "sb200_match_descriptors": (C.c_int, [vp, vp, u64, vp, u64, vp, u64, C.POINTER(u64)]),
        "sb200_match_descriptors_device": (C.c_int, [vp, vp, u64, vp, u64, vp, u64, C.POINTER(u64)]),
        "sb200_extract_batch_multi": (C.c_int, [C.POINTER(vp), u32, u8p, u32, u32, u32, u32, u64, i64,
                                                 C.POINTER(Result)]),
        "sb200_extract_batch_multi_parts": (C.c_int, [C.POINTER(vp), u32, u8p, u32, u32, u32, u32, u64, i64,
                                                       C.POINTER(Result), C.POINTER(u64)]),
        "sb200_last_gather_ms": (C.c_double, [vp]),
        "sb200_last_shard_ms": (C.c_double, [vp]),
        "sb200_set_profiling": (C.c_int, [vp, C.c_int]),
        "sb200_stage_stats": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(u64), u32]),
        "sb200_reset_stats": (C.c_int, [vp]),
        "sb200_launch_stats": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(u64), u32]),
        "sb200_launch_count": (u64, [vp]),
        "sb200_stage_name": (C.c_char_p, [u32]),
        "sb200_algorithmic_bytes": (u64, [u32, u32, C.POINTER(u64), u32]),
        "sb200_timer_start": (C.c_int, [vp]),
        "sb200_timer_stop": (C.c_int, [vp]),
        "sb200_timer_elapsed_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "sb200_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
        "sb200_host_free": (C.c_int, [vp]),
        "sb200_device_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "sb200_device_free": (C.c_int, [vp, vp]),
        "sb200_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "sb200_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "sb200_flush_l2": (C.c_int, [vp]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
