"""e2e of sb200_extract_batch from pinned vs pageable input, 128 x 1080p per call (development aid).
   [SB200_COPY_THREADS=k] python tools/pageable_probe.py"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf
from sift_features_b200 import _ffi
lib = _ffi.load()
w, h, B, n = 1920, 1080, 32, 128
ex = sf.Extractor(w, h, B)
H = ex.handle
p = C.c_void_p()
assert lib.sb200_host_alloc(n * w * h, C.byref(p)) == 0
pin = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n, h, w))
rng = np.random.default_rng(1)
for i in range(0, n, 16):
    pin[i:i + 16] = rng.integers(0, 256, (16, h, w), dtype=np.uint8)
page = np.array(pin, copy=True)
res = _ffi.Result()
def run(ptr):
    assert lib.sb200_extract_batch(H, ptr, n, w, h, w, w * h, -1, C.byref(res)) == 0
for name, ptr in (("pinned", pin.ctypes.data), ("pageable", page.ctypes.data), ("pinned", pin.ctypes.data), ("pageable", page.ctypes.data)):
    run(ptr); run(ptr)
    t0 = time.perf_counter()
    for _ in range(8):
        run(ptr)
    dt = (time.perf_counter() - t0) / 8
    print(f"{name:9s} {n / dt:8.0f} images/s  {1e3 * dt:7.2f} ms per call", flush=True)
