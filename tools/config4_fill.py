"""How much of a short per-device shard of config 4 is pipeline fill / drain: one device, n = 8192 / 2048 / 1024 / 512 images
of 640x480 through sb200_extract_batch_multi_parts (the per-device share at 1 / 4 / 8 / 16 GPUs), group sizes 128 / 64 / 32.
   python tools/config4_fill.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf
from sift_features_b200 import _ffi

lib = _ffi.load()
w, h, N = 640, 480, 8192
p = C.c_void_p()
assert lib.sb200_host_alloc(N * w * h, C.byref(p)) == 0
arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(N, h, w))
rng = np.random.default_rng(4321)
for i in range(0, N, 256):
    arr[i:i + 256] = rng.integers(0, 256, (256, h, w), dtype=np.uint8)
for B in [int(x) for x in sys.argv[1:]] or [128, 64, 32]:
    ex = sf.Extractor(w, h, B, device=0)
    handles = (C.c_void_p * 1)(ex.handle)
    parts = (_ffi.Result * 1)()
    first = (C.c_uint64 * 2)()
    base = None
    for n in (8192, 2048, 1024, 512):
        def run():
            assert lib.sb200_extract_batch_multi_parts(handles, 1, p, n, w, h, w, w * h, -1, parts, first) == 0
        run()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
        r = n / min(ts)
        base = base or r
        print(f"B={B:4d} n={n:5d}: {r:9.0f} images/s best ({1e3 * min(ts):7.2f} ms)  = {r / base:.3f} of the 8192-image rate", flush=True)
    ex.close()
