"""Match two images: examples/sift-match.rs of the reference.  Features of both images from the B200 library, mutual
nearest neighbours (BFMatcher NORM_L2 + crossCheck, :30-35) from its tcgen05 matcher, and -- like the reference --
OpenCV's own SIFT + BFMatcher beside it for comparison.

    python examples/sift_match.py IMAGE1 IMAGE2 [--processing=opencv|imageproc]      -> matches.jpg, cv_matches.jpg
"""
import sys

import cv2
import numpy as np

from _common import draw_matches, load_and_sift, pop_processing, sf, to_cv_keypoints

P = pop_processing(sys.argv)
if len(sys.argv) != 3:
    raise SystemExit("Required args: IMAGE1 IMAGE2")
g1, r1 = load_and_sift(sys.argv[1], None, P)
print(f"{len(r1)} keypoints")
g2, r2 = load_and_sift(sys.argv[2], None, P)
print(f"{len(r2)} keypoints")
pairs = sf.match(r2.descriptors, r1.descriptors)          # query = image 2, train = image 1, as in the reference
print(f"{len(pairs)} mutual matches")
draw_matches("matches.jpg", g2, to_cv_keypoints(r2), g1, to_cv_keypoints(r1), pairs)

cvsift = cv2.SIFT_create()
ck1, cd1 = cvsift.detectAndCompute(g1, None)
ck2, cd2 = cvsift.detectAndCompute(g2, None)
cm = cv2.BFMatcher(cv2.NORM_L2, True).match(cd2, cd1)
print(f"OpenCV: {len(ck1)} / {len(ck2)} keypoints, {len(cm)} mutual matches")
cv2.imwrite("cv_matches.jpg", cv2.drawMatches(g2, ck2, g1, ck1, cm, None))
