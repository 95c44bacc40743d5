import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import sift_features_b200 as sf
from oracle import oracle as O
from conftest import noise_image
g = noise_image(300, 260, 31)
g[20:120, 30:150] = 255
g[150:240, 100:290] = 0
g[125:140, :] = 77
ex = sf.Extractor(300, 260, 1)
res = ex.sift(g)
okp, odesc = O.sift(g)
dd = np.abs(res.descriptors.astype(int) - odesc.astype(int))
bad = np.nonzero(dd.max(1) > 1)[0]
print(len(res), "bad rows", len(bad))
ka = res.keypoint_array
for i in bad[:12]:
    print(i, ka[i], "maxdiff", dd[i].max(), "ndiff", (dd[i] > 0).sum(), "gpu sum", res.descriptors[i].astype(int).sum(), "ora sum", odesc[i].astype(int).sum())
    print("  gpu", res.descriptors[i][:32]); print("  ora", odesc[i][:32])
