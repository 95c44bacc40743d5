import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_gray(name: str) -> np.ndarray:
    return np.load(os.path.join(GOLDEN, f"{name}_gray.npy"))


def load_snapshot(name: str):
    z = np.load(os.path.join(GOLDEN, f"{name}_snapshot.npz"))
    return z["keypoints"], z["descriptors"]


def noise_image(w: int, h: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)


def smooth_image(w: int, h: int, seed: int) -> np.ndarray:
    """Band-limited texture: exercises flat regions, ties and low-contrast rejection that white noise never does."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float64)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(24):
        fx, fy = rng.uniform(-0.15, 0.15, 2)
        img += rng.uniform(0.2, 1.0) * np.cos(2 * np.pi * (fx * xx + fy * yy) + rng.uniform(0, 6.28))
    for _ in range(12):
        cx, cy, r = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(2, min(w, h) / 6 + 3)
        img += rng.uniform(-3, 3) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))
    img = (img - img.min()) / (img.max() - img.min() + 1e-9)
    return (img * 255).astype(np.uint8)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def sf():
    import sift_features_b200 as m
    from sift_features_b200 import _ffi
    _ffi.load()
    return m
