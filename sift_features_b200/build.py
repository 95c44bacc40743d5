"""Builds libsift_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m sift_features_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with
the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsift_b200.so")
SOURCES = ["sift_b200.cu"]
HEADERS = ["sb_common.cuh", "sb_math.cuh", "sb_pyramid.cuh", "sb_keypoints.cuh", "sb_match.cuh", "sb_jpeg.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",          # every FMA in the kernels is explicit (bit-exact arithmetic contract)
    "-Xcompiler", "-fPIC", "-shared",
    "-ldl",                  # nvJPEG is opened with dlopen on first use of the JPEG entry points (sb_jpeg.h)
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(HERE, "..", "include", "sift_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("SB200_NVCC_EXTRA", "").split()   # development knob, e.g. -DSB_MARCH_TW(l)=64
    out = os.environ.get("SB200_BUILD_OUT", LIB)              # development knob: build a variant next to the real one
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
