"""The example programs (examples/*.py = the reference's examples/run-sift.rs, sift-match.rs, opencv-cross-match.rs on this
library) run end to end on image files and report what the library calls report."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_gray

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _run(script, *args, cwd):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *map(str, args)], cwd=cwd,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_run_sift(sf, tmp_path):
    g = load_gray("bird_small")
    png, jpg = tmp_path / "a.png", tmp_path / "a.jpg"
    cv2.imwrite(str(png), g)
    cv2.imwrite(str(jpg), g, [cv2.IMWRITE_JPEG_QUALITY, 95])
    # like the reference's run-sift, the example calls the crate's plain sift() = ImageprocProcessing ...
    assert _run("run_sift.py", png, cwd=tmp_path) == f"{len(sf.sift(g))} keypoints\n"       # lossless file: exact
    # ... and --processing=opencv selects the flavour the crate's test pins
    n_cv = len(sf.sift_with_processing(g, None, sf.OpenCVProcessing))
    assert n_cv != len(sf.sift(g))
    assert _run("run_sift.py", png, "--processing=opencv", cwd=tmp_path) == f"{n_cv} keypoints\n"
    assert _run("run_sift.py", png, 100, cwd=tmp_path) == "100 keypoints\n"
    with sf.Extractor(g.shape[1], g.shape[0], 1) as ex:
        n_jpeg = len(ex.sift(ex.decode_jpeg_luma(jpg.read_bytes())))
    assert _run("run_sift.py", jpg, "--processing=opencv", cwd=tmp_path) == f"{n_jpeg} keypoints\n"
    colour = np.stack([g, np.roll(g, 2, 0), np.roll(g, 2, 1)], -1)
    cv2.imwrite(str(tmp_path / "c.png"), colour[..., ::-1])
    with sf.Extractor(g.shape[1], g.shape[0], 1) as ex:
        n_rgb = len(ex.sift_rgb(colour))
    assert _run("run_sift.py", tmp_path / "c.png", "--processing=opencv", cwd=tmp_path) == f"{n_rgb} keypoints\n"


def test_match_examples(sf, tmp_path):
    g = load_gray("tree_small")
    a, b = tmp_path / "a.png", tmp_path / "b.png"
    cv2.imwrite(str(a), g)
    cv2.imwrite(str(b), g[10:-6, 14:-10])                      # a crop: same content, shifted
    out = _run("sift_match.py", a, b, "--processing=opencv", cwd=tmp_path)
    n1, n2, m, cv_m = (int(x) for x in re.search(r"(\d+) keypoints\n(\d+) keypoints\n(\d+) mutual matches\n"
                                                  r"OpenCV: \d+ / \d+ keypoints, (\d+) mutual", out).groups())
    assert n1 == len(sf.sift_with_processing(g, None, sf.OpenCVProcessing)) and m >= 0.6 * n2
    assert abs(m - cv_m) <= 0.1 * cv_m                         # as many mutual matches as OpenCV finds on the pair
    assert (tmp_path / "matches.jpg").stat().st_size > 0 and (tmp_path / "cv_matches.jpg").stat().st_size > 0
    out = _run("opencv_cross_match.py", a, a, "--processing=opencv", cwd=tmp_path)    # same image both sides: nearly everything matches
    ncv, nb, m = (int(x) for x in re.search(r"(\d+) OpenCV keypoints, (\d+) B200 keypoints, (\d+) mutual", out).groups())
    assert m >= 0.9 * min(ncv, nb)
    assert (tmp_path / "matches-b200-opencv.jpg").stat().st_size > 0
    # the crate's default flavour runs through the same programs (its keypoints are not OpenCV-compatible: no threshold)
    out = _run("opencv_cross_match.py", a, a, cwd=tmp_path)
    assert re.search(r"(\d+) OpenCV keypoints, (\d+) B200 keypoints, (\d+) mutual", out)
