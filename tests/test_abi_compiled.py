"""The drop-in boundary without Python in the call path: a plain C program against include/sift_b200.h and a C++
program against the header-only mirror (sift_features_b200/cpp/sift_features.hpp), compiled with gcc / g++, linked
against libsift_b200.so, run on the GPU, their binary output diffed against what the ctypes binding returns."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_gray

HOST = os.path.join(ROOT, "tests", "host")
LIBDIR = os.path.join(ROOT, "sift_features_b200")


def _compile(tmp_path, src, exe, cxx):
    out = str(tmp_path / exe)
    cmd = ([("g++"), "-std=c++17"] if cxx else ["gcc", "-std=c11"]) + ["-O1", "-Wall", "-Wextra", "-o", out, os.path.join(HOST, src),
           "-L" + LIBDIR, "-l:libsift_b200.so", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "warning" not in r.stderr, r.stderr          # the header is clean C11 / C++17
    return out


def test_headers_compile_and_link(tmp_path):
    """Runs without a GPU: both programs compile warning-free and link against the built library."""
    _compile(tmp_path, "abi_smoke.c", "abi_smoke", False)
    _compile(tmp_path, "abi_smoke_cpp.cpp", "abi_smoke_cpp", True)


def _blocks(path, n_blocks):
    data = open(path, "rb").read()
    pos, out = 0, []
    for _ in range(n_blocks):
        (n,) = struct.unpack_from("<Q", data, pos); pos += 8
        kp = np.frombuffer(data, np.float32, n * 5, pos).reshape(n, 5); pos += n * 20
        de = np.frombuffer(data, np.uint8, n * 128, pos).reshape(n, 128); pos += n * 128
        out.append((kp, de))
    desc = np.frombuffer(data, np.uint8, 128, pos); pos += 128
    assert pos == len(data)
    return out, desc


@pytest.mark.gpu
def test_compiled_c_and_cpp_programs_match_ctypes(sf, tmp_path):
    g = load_gray("bird_small")
    h, w = g.shape
    raw = tmp_path / "img.raw"
    raw.write_bytes(g.tobytes())
    with sf.Extractor(w, h, 2) as ex:
        a = ex.sift(g)
        lim = ex.sift(g, 50)
        d = ex.compute_descriptors(g.astype(np.float32) / np.float32(255), [[100.0, 100.0, 2.1, 123.0]])[0]
    with sf.Extractor(w, h, 1, processing=sf.ImageprocProcessing) as ex:
        b = ex.sift(g)

    def as_rows(res):
        ka = res.keypoint_array
        return np.stack([ka[f] for f in ("x", "y", "size", "angle", "response")], 1)

    exe = _compile(tmp_path, "abi_smoke.c", "abi_smoke", False)
    r = subprocess.run([exe, str(raw), str(w), str(h), str(tmp_path / "c.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    blocks, desc = _blocks(tmp_path / "c.bin", 4)
    for (kp, de), ref in zip(blocks, (a, lim, a, b)):
        assert np.array_equal(kp, as_rows(ref)) and np.array_equal(de, ref.descriptors)
    assert np.array_equal(desc, d)

    exe = _compile(tmp_path, "abi_smoke_cpp.cpp", "abi_smoke_cpp", True)
    r = subprocess.run([exe, str(raw), str(w), str(h), str(tmp_path / "cpp.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert f"{len(a)} keypoints (opencv), {len(b)} (crate default)" in r.stdout
    blocks, desc = _blocks(tmp_path / "cpp.bin", 2)
    for (kp, de), ref in zip(blocks, (a, b)):
        assert np.array_equal(kp, as_rows(ref)) and np.array_equal(de, ref.descriptors)
    assert np.array_equal(desc, d)
