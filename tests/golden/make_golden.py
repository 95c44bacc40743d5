"""Generates the committed golden fixtures from the read-only reference tree.

Run HERE (in the build container, where /root/reference exists):
    python tests/golden/make_golden.py

Outputs (small, committed):
  <name>_gray.npy       u8 (H,W): the reference's test images (images/*_small.jpg, src/lib.rs:1038,1047)
                        decoded with Pillow (libjpeg-turbo) and converted with the `image` crate's
                        integer Rec.709 luma (2126 R + 7152 G + 722 B) / 10000.  The crate itself decodes
                        with zune-jpeg, whose IDCT differs by +-1 grey level on some pixels -- the reason
                        the insta snapshots are a tolerance oracle here, not an equality oracle.
  <name>_snapshot.npz   keypoints (N,5) f32 [x,y,size,angle,response] and descriptors (N,128) u8 parsed
                        from src/snapshots/sift__sift_end2end{,-2,-3,-4}.snap (the reference's only
                        golden vectors; written by #[test] sift_end2end, src/lib.rs:1009-1056).
  bird_gray.npy         images/bird.jpg (799x533), the benches' input (benches/sift.rs:79, descriptor.rs:9)
  tree_gray.npy         images/tree.jpg (800x600): the densest of the reference's images (~9.6k keypoints / Mpx);
                        tiled to the bench shapes as the "natural image" workload
"""
import os
import re

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def gray_rec709(path):
    rgb = np.asarray(Image.open(path).convert("RGB")).astype(np.uint32)
    l = (2126 * rgb[..., 0] + 7152 * rgb[..., 1] + 722 * rgb[..., 2]) // 10000
    return l.astype(np.uint8)


def parse_keypoints(path):
    txt = open(path).read().split("---", 2)[2]
    vals = re.findall(r"^\s*-?\s*(x|y|size|angle|response): ([-+0-9.eE]+)$", txt, re.M)
    assert len(vals) % 5 == 0
    arr = np.array([float(v) for _, v in vals], np.float32).reshape(-1, 5)
    names = [k for k, _ in vals[:5]]
    assert names == ["x", "y", "size", "angle", "response"], names
    return arr


def parse_descriptors(path):
    txt = open(path).read().split("---", 2)[2]
    rows, cur = [], None
    for line in txt.splitlines():
        if line.startswith("- - "):
            if cur is not None:
                rows.append(cur)
            cur = [int(line[4:])]
        elif line.startswith("  - "):
            cur.append(int(line[4:]))
    if cur is not None:
        rows.append(cur)
    arr = np.array(rows, np.uint8)
    assert arr.shape[1] == 128, arr.shape
    return arr


def main():
    snaps = {"tree_small": ("sift__sift_end2end.snap", "sift__sift_end2end-2.snap"),
             "bird_small": ("sift__sift_end2end-3.snap", "sift__sift_end2end-4.snap")}
    for name, (ks, ds) in snaps.items():
        g = gray_rec709(f"{REF}/images/{name}.jpg")
        np.save(f"{HERE}/{name}_gray.npy", g)
        kp = parse_keypoints(f"{REF}/src/snapshots/{ks}")
        de = parse_descriptors(f"{REF}/src/snapshots/{ds}")
        assert len(kp) == len(de)
        np.savez_compressed(f"{HERE}/{name}_snapshot.npz", keypoints=kp, descriptors=de)
        print(name, g.shape, kp.shape, de.shape)
    g = gray_rec709(f"{REF}/images/bird.jpg")
    np.save(f"{HERE}/bird_gray.npy", g)
    print("bird", g.shape)
    g = gray_rec709(f"{REF}/images/tree.jpg")
    np.save(f"{HERE}/tree_gray.npy", g)
    print("tree", g.shape)


if __name__ == "__main__":
    main()
