"""Shared helpers of the example programs (the callers either side of the extraction path)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sift_features_b200 as sf  # noqa: E402


def is_jpeg(data: bytes) -> bool:
    return data[:2] == b"\xff\xd8"


def pop_processing(argv):
    """`--processing=opencv|imageproc` anywhere on the command line.  The reference's examples call the crate's plain
    sift(), i.e. ImageprocProcessing (src/lib.rs:71-73) -- the default here too; `opencv` selects the flavour the crate's
    test pins (sift_with_processing::<OpenCVProcessing>)."""
    processing = sf.ImageprocProcessing
    for a in list(argv):
        if a.startswith("--processing="):
            argv.remove(a)
            name = a.split("=", 1)[1]
            if name not in ("opencv", "imageproc"):
                raise SystemExit("--processing must be opencv or imageproc")
            processing = sf.OpenCVProcessing if name == "opencv" else sf.ImageprocProcessing
    return processing


def load_and_sift(path: str, features_limit=None, processing=None):
    """`image::open(path).grayscale()` + `sift()` (examples/run-sift.rs:8-19).  A JPEG goes to the device as a bitstream
    (nvJPEG decode + integer luma there); anything else is decoded by OpenCV and converted on the device.
    Returns (gray image the features belong to, SiftResult)."""
    processing = processing or sf.ImageprocProcessing
    data = open(path, "rb").read()
    if is_jpeg(data):
        with sf.Extractor(8, 8, 1) as probe:
            w, h, _ = probe.jpeg_info(data)
        with sf.Extractor(w, h, 1, processing=processing) as ex:
            gray = ex.decode_jpeg_luma(data)
            _, kp, desc = ex.sift_jpeg([data], features_limit)
        return gray, sf.SiftResult(kp, desc)
    import cv2
    img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise SystemExit(f"cannot decode {path}")
    if img.dtype != np.uint8:
        raise SystemExit("wrong image type")      # the reference's examples accept 8-bit images only
    if img.ndim == 2:
        return img, sf.sift_with_processing(img, features_limit, processing)
    rgb = np.ascontiguousarray(img[..., 2::-1])    # BGR(A) -> RGB
    h, w = rgb.shape[:2]
    with sf.Extractor(w, h, 1, processing=processing) as ex:
        return ex.rgb_to_luma(rgb), ex.sift_rgb(rgb, features_limit)


def to_cv_keypoints(result):
    """KeyPoint -> cv2.KeyPoint the way examples/sift-match.rs:10-18 does it (size and angle copied, octave 1)."""
    import cv2
    ka = result.keypoint_array
    return [cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]), 1)
            for k in ka]


def draw_matches(path, img_query, kp_query, img_train, kp_train, pairs):
    """draw_matches_def + imwrite of examples/sift-match.rs:36-37 (host-side drawing, not part of the path)."""
    import cv2
    dm = [cv2.DMatch(int(p["queryIdx"]), int(p["trainIdx"]), float(p["distance"])) for p in pairs]
    out = cv2.drawMatches(img_query, kp_query, img_train, kp_train, dm, None)
    cv2.imwrite(path, out)
