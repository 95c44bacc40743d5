"""Run sift and print the keypoint count: examples/run-sift.rs of the reference, on the B200 library.

    python examples/run_sift.py IMAGE [FEATURES_LIMIT] [--processing=opencv|imageproc]
"""
import sys

from _common import load_and_sift, pop_processing

P = pop_processing(sys.argv)
if len(sys.argv) < 2:
    raise SystemExit("Required args: IMAGE [FEATURES_LIMIT]")
_, res = load_and_sift(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None, P)
print(f"{len(res)} keypoints")
