"""The C-ABI library loads and exports every symbol include/sift_b200.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sift_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(sf):
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    assert sorted(_ffi.SYMBOLS) == declared           # the binding covers the whole header
    for name in declared:
        assert hasattr(lib, name), name


def test_header_cites_reference_lines():
    hdr = open(os.path.join(ROOT, "include", "sift_b200.h")).read()
    for cite in ["src/lib.rs:71-81", "src/lib.rs:131-143", ":147-177", "src/lib.rs:785-990", "src/lib.rs:48-56",
                 "src/lib.rs:39-46", "src/opencv_processing.rs"]:
        assert cite in hdr, cite


def test_status_strings_and_stage_names(sf):
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    assert lib.sb200_status_string(0) == b"ok"
    assert lib.sb200_status_string(3) == b"capacity exceeded"
    assert [lib.sb200_stage_name(i).decode() for i in range(7)] == list(sf.STAGE_NAMES)


def test_algorithmic_bytes_match_survey_table(sf):
    # 113.35 MB, 765.14 MB, 3060.62 MB
    assert sf.algorithmic_bytes(640, 480)[0] == 113_353_680
    total, seed, blur, ext = sf.algorithmic_bytes(1920, 1080)
    assert total == seed + blur + ext == 765_143_904
    assert seed == 1920 * 1080 + 4 * 3840 * 2160
    assert abs(sf.algorithmic_bytes(3840, 2160)[0] - 3_060_620_000) < 10_000


def test_no_cpu_fallback(sf):
    """Without a CUDA device context creation fails with SB200_E_CUDA -- never a silent CPU path."""
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    if lib.sb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(sf.SiftError) as e:
        sf.Extractor(64, 64)
    assert e.value.status == _ffi.E_CUDA
    with pytest.raises(sf.SiftError):
        sf.sift(np.zeros((32, 32), np.uint8))


def test_invalid_arguments_rejected_before_any_device_work(sf):
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    h = C.c_void_p()
    assert lib.sb200_create(0, 0, 10, 1, 0, C.byref(h)) == _ffi.E_INVALID
    assert lib.sb200_create(0, 9000, 10, 1, 0, C.byref(h)) == _ffi.E_INVALID   # > SB200_MAX_DIM
    assert lib.sb200_extract(None, None, 1, 1, 1, -1, None) == _ffi.E_INVALID
    with pytest.raises(ValueError):
        sf.sift(np.zeros((4, 4), np.float32))
    with pytest.raises(ValueError):                                              # not a Processing implementation
        sf.sift_with_processing(np.zeros((4, 4), np.uint8), None, object)
    assert lib.sb200_set_processing(None, 0) == _ffi.E_INVALID and lib.sb200_get_processing(None) < 0
    assert lib.sb200_set_postfilter(None, 1, 10) == _ffi.E_INVALID
    assert lib.sb200_extract_batch_multi_parts(None, 0, None, 1, 1, 1, 1, 1, -1, None, None) == _ffi.E_INVALID


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under sift_features_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sift_features_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs")):
                txt = open(os.path.join(dp, f)).read()
                assert "sift_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_shard_ranges(sf):
    assert [list(r) for r in sf.shard_ranges(10, 4)] == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9]]
    assert sum(len(r) for r in sf.shard_ranges(8192, 8)) == 8192
    assert [len(r) for r in sf.shard_ranges(3, 8)] == [1, 1, 1, 0, 0, 0, 0, 0]
