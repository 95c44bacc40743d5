"""Per-launch device times of the pyramid stages (CUDA events, no profiler attached): which octave / layer costs what.

    python tools/fine_profile.py [1080p|vga|4k] [images per group] [repeats] [opencv|imageproc]

Prints, per (octave, kind), the average launch time, the pixels it covers, and the algorithmic GB/s
(8 B/px for a blur, 24 B/px for the extrema scan) against the measured HBM peak.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ctypes as C

import sift_features_b200 as sf
from sift_features_b200 import _ffi

SHAPES = {"1080p": (1920, 1080, 32), "vga": (640, 480, 128), "4k": (3840, 2160, 8)}
name = sys.argv[1] if len(sys.argv) > 1 else "1080p"
w, h, B = SHAPES[name]
if len(sys.argv) > 2:
    B = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
peak = 6541.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flavour = sys.argv[4] if len(sys.argv) > 4 else "opencv"
lib = _ffi.load()
ex = sf.Extractor(w, h, B, processing=sf.ImageprocProcessing if flavour == "imageproc" else sf.OpenCVProcessing)
H = ex.handle
imgs = np.stack([np.random.default_rng([1234, i]).integers(0, 256, (h, w), dtype=np.uint8) for i in range(B)])
d = C.c_void_p()
assert lib.sb200_device_alloc(H, imgs.nbytes, C.byref(d)) == 0
assert lib.sb200_memcpy_h2d(H, d, imgs.ctypes.data, imgs.nbytes) == 0
for _ in range(3):
    assert lib.sb200_extract_batch_device(H, d, B, w, h, w, w * h, -1) == 0
lib.sb200_sync(H)
ex.set_profiling(True)
ex.reset_stats()
for _ in range(reps):
    assert lib.sb200_extract_batch_device(H, d, B, w, h, w, w * h, -1) == 0
    lib.sb200_sync(H)
fine = ex.launch_stats()
stages = ex.stage_stats()
ex.set_profiling(False)
cw, ch = 2 * w, 2 * h
dims = []
while True:
    dims.append((cw, ch))
    if len(dims) >= 16 or min(cw, ch) <= 1:
        break
    cw, ch = cw // 2, ch // 2
print(f"# {name} ({flavour} flavour): {B} images per launch, {reps} repeats, HBM peak {peak} GB/s")
tot = 0.0
for (o, kind), (ms, n) in sorted(fine.items()):
    us = 1e3 * ms / n
    tot += us
    ow, oh = dims[o]
    px = ow * oh * B
    bpp = {"seed": 4 + 0.25, "extrema": 24, "tail": 0}.get(kind, 8)
    gbs = bpp * px / (us * 1e-6) / 1e9 if bpp else 0.0
    print(f"octave {o:2d} {ow:5d}x{oh:<5d} {kind:8s} {us:9.1f} us  {gbs:8.0f} GB/s  {gbs / peak:5.2f}")
print(f"sum of pyramid launches: {tot:.1f} us per group = {tot / B:.2f} us per image")
print("stages (ms per image):", {k: round(v['ms'] / reps / B, 5) for k, v in stages.items()})
ex.close()
