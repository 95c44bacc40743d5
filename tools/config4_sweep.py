"""Config 4 (8192 x 640x480 over all devices of the box) through sb200_extract_batch_multi_parts for several group sizes.
   python tools/config4_sweep.py [B ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf
from sift_features_b200 import _ffi

lib = _ffi.load()
ndev = lib.sb200_device_count()
w, h, n = 640, 480, 8192
p = C.c_void_p()
assert lib.sb200_host_alloc(n * w * h, C.byref(p)) == 0
arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n, h, w))
rng = np.random.default_rng(4321)
for i in range(0, n, 256):
    arr[i:i + 256] = rng.integers(0, 256, (256, h, w), dtype=np.uint8)
for B in [int(x) for x in sys.argv[1:]] or [128, 64, 32]:
    exs = [sf.Extractor(w, h, B, device=d) for d in range(ndev)]
    handles = (C.c_void_p * ndev)(*[e.handle for e in exs])
    parts = (_ffi.Result * ndev)()
    first = (C.c_uint64 * (ndev + 1))()
    for k in sorted({1, ndev}):
        def run():
            assert lib.sb200_extract_batch_multi_parts(handles, k, p, n, w, h, w, w * h, -1, parts, first) == 0
        run()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
        lib.sb200_last_shard_ms.restype = C.c_double
        shard = [round(lib.sb200_last_shard_ms(e.handle), 1) for e in exs[:k]]
        print(f"B={B:4d} devices={k}: {n / min(ts):9.0f} images/s best, {n / np.median(ts):9.0f} median ({1e3 * min(ts):.1f} ms); "
              f"last call: shard ms per device {shard}", flush=True)
    for e in exs:
        e.close()
