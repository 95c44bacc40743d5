"""sb_math.cuh (the device-side restatement of glibc expf / powf(2,.)) compiled for the host
must equal the host libm bit for bit -- the same header is what the CUDA kernels include."""
import os
import subprocess

from conftest import ROOT


def test_math_port_matches_libm(tmp_path):
    exe = str(tmp_path / "math_port_check")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-o", exe,
                    os.path.join(ROOT, "tests", "host", "math_port_check.cpp"), "-lm"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    rows = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    assert rows["expf"][0] > 10_000_000 and rows["expf"][1] == 0, rows
    assert rows["pow2f"][0] > 5_000_000 and rows["pow2f"][1] == 0, rows
