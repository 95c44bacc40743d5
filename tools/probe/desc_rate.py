"""Fraction of descriptor rows identical to the oracle's / within +-1 (README claim check)."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sift_features_b200 as sf
from oracle import oracle as O
from conftest import load_gray, noise_image
O.build(); O.lib()
for name, g in [("bird", load_gray("bird")), ("tree_small", load_gray("tree_small")), ("noise640", noise_image(640, 480, 1234))]:
    r = sf.sift(g); okp, od = O.sift(g)
    d = np.abs(r.descriptors.astype(int) - od.astype(int))
    print(name, len(r), "rows identical %.4f" % (d.max(1) == 0).mean(), "rows within 1 %.4f" % (d.max(1) <= 1).mean(), "max", d.max(),
          "bytes differing %.6f" % (d > 0).mean())
