"""Dumps the descriptors / keypoints of two test images with the library SB200_LIB selects (development aid):
   SB200_LIB=a.so python tools/desc_ab.py out_a.npz ; SB200_LIB=b.so python tools/desc_ab.py out_b.npz ; then compare."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf
out = {}
for name, (w, h, seed) in {"noise1080": (1920, 1080, 1234), "noise_vga": (640, 480, 7)}.items():
    g = np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    with sf.Extractor(w, h, 1) as ex:
        r = ex.sift(g)
    out[name + "_d"] = r.descriptors
    out[name + "_k"] = r.keypoint_array.view(np.uint8)
t = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bird_gray.npy")) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bird_gray.npy")) else None
if t is not None:
    with sf.Extractor(t.shape[1], t.shape[0], 1) as ex:
        r = ex.sift(t)
    out["bird_d"] = r.descriptors
np.savez(sys.argv[1], **out)
if len(sys.argv) > 2:
    a = np.load(sys.argv[2])
    for k in out:
        same = np.array_equal(out[k], a[k])
        print(k, out[k].shape, "identical" if same else f"DIFFERENT: {np.count_nonzero(out[k] != a[k])} bytes")
