"""Where a single-image call spends its time: python tools/latency_breakdown.py [WxH] [opencv|imageproc]
   sift() from pageable / pinned host memory, the device-resident call (no copies), the pyramid alone."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf
from sift_features_b200 import _ffi
size = sys.argv[1] if len(sys.argv) > 1 else "1920x1080"
w, h = map(int, size.split("x"))
flavour = sys.argv[2] if len(sys.argv) > 2 else "opencv"
P = sf.ImageprocProcessing if flavour == "imageproc" else sf.OpenCVProcessing
lib = _ffi.load()
g = np.random.default_rng(1234).integers(0, 256, (h, w), dtype=np.uint8)


def med(f, n=40):
    for _ in range(5):
        f()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts))


with sf.Extractor(w, h, 1, processing=P) as ex:
    H = ex.handle
    t_sift = med(lambda: ex.sift(g))
    p = C.c_void_p(); assert lib.sb200_host_alloc(g.nbytes, C.byref(p)) == 0
    pin = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h, w)); pin[...] = g
    res = _ffi.Result()
    t_abi_pageable = med(lambda: lib.sb200_extract_batch(H, g.ctypes.data, 1, w, h, w, w * h, -1, C.byref(res)))
    t_abi_pinned = med(lambda: lib.sb200_extract_batch(H, p, 1, w, h, w, w * h, -1, C.byref(res)))
    d = C.c_void_p(); assert lib.sb200_device_alloc(H, g.nbytes, C.byref(d)) == 0
    assert lib.sb200_memcpy_h2d(H, d, g.ctypes.data, g.nbytes) == 0

    def dev():
        lib.sb200_extract_batch_device(H, d, 1, w, h, w, w * h, -1); lib.sb200_sync(H)

    def pyr():
        lib.sb200_pyramid_batch_device(H, d, 1, w, h, w, w * h); lib.sb200_sync(H)
    t_dev = med(dev)
    t_pyr = med(pyr)
    print(f"{size} {flavour} tail={os.environ.get('SB200_TAIL', '1')} fork={os.environ.get('SB200_FORK', '1')}: sift() {t_sift:.3f} ms | C ABI pageable {t_abi_pageable:.3f} pinned {t_abi_pinned:.3f} | "
          f"device-resident {t_dev:.3f} | pyramid only {t_pyr:.3f} | keypoints {int(res.n)}")
