"""Small fixed workload for ncu: `steps` device-resident batches of `batch` images (default 1080p)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf  # noqa: E402
from sift_features_b200 import _ffi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
w, h = map(int, a.size.split("x"))
lib = _ffi.load()
ex = sf.Extractor(w, h, a.batch)
H = ex.handle
imgs = np.stack([np.random.default_rng([1234, i]).integers(0, 256, (h, w), dtype=np.uint8) for i in range(a.batch)])
d = C.c_void_p()
assert lib.sb200_device_alloc(H, imgs.nbytes, C.byref(d)) == 0
assert lib.sb200_memcpy_h2d(H, d, imgs.ctypes.data, imgs.nbytes) == 0
for _ in range(a.steps):
    st = lib.sb200_extract_batch_device(H, d, a.batch, w, h, w, w * h, -1)
    assert st == 0, lib.sb200_last_error(H)
    assert lib.sb200_sync(H) == 0
counts = (C.c_uint32 * a.batch)()
assert lib.sb200_device_result(H, counts, a.batch, None, None, None) == 0
print("keypoints per image:", list(counts), "launches:", ex.launch_count)
ex.close()
