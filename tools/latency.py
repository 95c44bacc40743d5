"""Single-image latency of the public call (sift()) -- the way the reference crate is used.  python tools/latency.py [WxH]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf
size = sys.argv[1] if len(sys.argv) > 1 else "1920x1080"
w, h = map(int, size.split("x"))
g = np.random.default_rng(1234).integers(0, 256, (h, w), dtype=np.uint8)
with sf.Extractor(w, h, 1) as ex:
    for _ in range(5):
        r = ex.sift(g)
    ts = []
    for _ in range(30):
        t = time.perf_counter(); r = ex.sift(g); ts.append(time.perf_counter() - t)
    print(f"{size}: {len(r)} keypoints, latency median {1e3*np.median(ts):.3f} ms, min {1e3*min(ts):.3f} ms, launches/call {ex.launch_count // 35}")
