"""OpenCV SIFT on IMAGE1 against this library's SIFT on IMAGE2: examples/opencv-cross-match.rs of the reference
(descriptors are interchangeable with OpenCV's; BFMatcher NORM_L2 + crossCheck, :34-43).

    python examples/opencv_cross_match.py IMAGE1 IMAGE2 [--processing=opencv|imageproc]      -> matches-b200-opencv.jpg
"""
import sys

import cv2
import numpy as np

from _common import draw_matches, load_and_sift, pop_processing, sf, to_cv_keypoints

P = pop_processing(sys.argv)
if len(sys.argv) != 3:
    raise SystemExit("Required args: IMAGE1 IMAGE2")
g1 = cv2.imread(sys.argv[1], cv2.IMREAD_GRAYSCALE)
if g1 is None:
    raise SystemExit(f"cannot decode {sys.argv[1]}")
ck1, cd1 = cv2.SIFT_create().detectAndCompute(g1, None)
g2, r2 = load_and_sift(sys.argv[2], None, P)
# OpenCV's descriptors are float32 holding integers 0..255: the same 128 bytes this library returns
train = np.clip(np.rint(cd1), 0, 255).astype(np.uint8)
pairs = sf.match(r2.descriptors, train)
print(f"{len(ck1)} OpenCV keypoints, {len(r2)} B200 keypoints, {len(pairs)} mutual matches")
draw_matches("matches-b200-opencv.jpg", g2, to_cv_keypoints(r2), g1, list(ck1), pairs)
