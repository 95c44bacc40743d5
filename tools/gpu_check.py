"""Verbose GPU-vs-oracle stage comparison (development aid; the assertions live in tests/).

    python tools/gpu_check.py [--size WxH] [--seed N] [--golden NAME]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)


def compare(gray, label, limit=None):
    h, w = gray.shape
    print(f"== {label}: {w}x{h}")
    ex = sf.Extractor(w, h, 1)
    t = time.time()
    pre = ex.precompute_images(gray)
    print(f"  gpu precompute {time.time()-t:.3f}s, octaves {pre.n_octaves} dims {pre.dims}")
    t = time.time()
    P = O.Pyramid(gray)
    print(f"  oracle pyramid {time.time()-t:.3f}s, octaves {P.n_octaves}")
    assert P.n_octaves == pre.n_octaves and P.dims == pre.dims
    bad_layers = 0
    for o in range(P.n_octaves):
        g = pre.scale_space[o]
        for l in range(6):
            ref = P.gauss(o, l)
            nbad = int(np.count_nonzero(g[l].view(np.uint32) != ref.view(np.uint32)))
            if nbad:
                bad_layers += 1
                d = np.abs(g[l] - ref)
                ys, xs = np.nonzero(g[l] != ref)
                print(f"  MISMATCH oct {o} layer {l}: {nbad} px, max {d.max():.3e}, first at y={ys[0]} x={xs[0]}"
                      f" (ylim {ys.min()}..{ys.max()} xlim {xs.min()}..{xs.max()})")
        if o < 2:
            d = pre.dog[o]
            for l in range(5):
                assert np.array_equal(d[l], P.dog(o, l)), ("dog", o, l)
    print(f"  pyramid: {bad_layers} mismatching layers")
    t = time.time()
    res = ex.sift_with_precomputed(limit)
    print(f"  gpu detect+describe {time.time()-t:.3f}s -> {len(res)} keypoints")
    cg = ex.last_candidates()
    co = P.candidates()
    same = len(cg) == len(co) and np.array_equal(cg, co)
    print(f"  candidates gpu {len(cg)} oracle {len(co)} identical(order too)={same}")
    if not same:
        sg = set(map(tuple, cg.tolist())); so = set(map(tuple, co.tolist()))
        print(f"    only gpu {len(sg-so)} only oracle {len(so-sg)}; first diffs {sorted(sg-so)[:3]} {sorted(so-sg)[:3]}")
    kg = ex.last_sift_keypoints()
    ko = P.sift_keypoints()
    print(f"  sift keypoints gpu {len(kg)} oracle {len(ko)}")
    if len(kg) == len(ko):
        for f in kg.dtype.names:
            a, b = kg[f], ko[f]
            if a.dtype.kind == "f":
                nb = int(np.count_nonzero(a.view(np.uint32) != b.view(np.uint32)))
                print(f"    {f}: bit-mismatches {nb}, max abs diff {np.abs(a-b).max() if len(a) else 0:.3e}")
            else:
                print(f"    {f}: mismatches {int(np.count_nonzero(a != b))}")
    t = time.time()
    okp, odesc = P.sift(limit)
    print(f"  oracle detect+describe {time.time()-t:.3f}s -> {len(okp)}")
    if len(okp) == len(res):
        ka = res.keypoint_array
        for f in ka.dtype.names:
            nb = int(np.count_nonzero(ka[f].view(np.uint32) != okp[f].view(np.uint32)))
            print(f"    out {f}: bit-mismatches {nb}, max abs diff {np.abs(ka[f]-okp[f]).max() if len(ka) else 0:.3e}")
        dd = np.abs(res.descriptors.astype(int) - odesc.astype(int))
        rowmax = dd.max(1) if len(dd) else np.zeros(0)
        print(f"    descriptors: rows exact {np.mean(rowmax == 0):.4f}, within +-1 {np.mean(rowmax <= 1):.4f}, "
              f"within +-2 {np.mean(rowmax <= 2):.4f}, max diff {int(rowmax.max()) if len(rowmax) else 0}, "
              f"bytes differing {np.mean(dd > 0):.5f}")
    # one-call path must agree with the staged path
    r2 = ex.sift(gray, limit)
    print(f"  one-call sift == staged: {r2 == res}")
    ex.close()
    P.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="640x480")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--golden", default="bird_small,tree_small")
    ap.add_argument("--limit", type=int, default=None)
    a = ap.parse_args()
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    for name in [n for n in a.golden.split(",") if n]:
        compare(np.load(os.path.join(G, name + "_gray.npy")), name, a.limit)
    w, h = map(int, a.size.split("x"))
    rng = np.random.default_rng(a.seed)
    compare(rng.integers(0, 256, (h, w), dtype=np.uint8), f"noise seed {a.seed}", a.limit)
    # odd sizes exercise ragged tiles and the reflect borders
    compare(rng.integers(0, 256, (67, 131), dtype=np.uint8), "noise 131x67", a.limit)


if __name__ == "__main__":
    main()
