"""Host-side logic that needs no GPU: the module-level context cache (never evicts a context the current call still
uses, so sharding over more devices than it normally holds works), the shard partition, bench.py's byte model."""
import numpy as np

import sift_features_b200 as sf


class _FakeExtractor:
    alive = set()

    def __init__(self, w, h, batch, device):
        self.key = (device, w, h, batch)
        _FakeExtractor.alive.add(self.key)

    def close(self):
        _FakeExtractor.alive.discard(self.key)


def test_context_cache_never_evicts_contexts_in_use(monkeypatch):
    monkeypatch.setattr(sf, "Extractor", _FakeExtractor)
    monkeypatch.setattr(sf, "_cache", {})
    _FakeExtractor.alive = set()
    # sixteen devices, twice the cache's normal size, as sift_batch(devices=range(16)) asks for them
    keys = {(d, 64, 48, 4) for d in range(16)}
    exs = [sf._extractor(64, 48, 4, d, keep=keys) for d in range(16)]
    assert all(e.key in _FakeExtractor.alive for e in exs)           # none was closed while the call holds it
    again = [sf._extractor(64, 48, 4, d, keep=keys) for d in range(16)]
    assert all(a is b for a, b in zip(exs, again))                   # and they are reused, not rebuilt
    # a later, unrelated call shrinks the cache back: least recently used first
    other = sf._extractor(32, 32, 1, 0)
    assert other.key in _FakeExtractor.alive and len(sf._cache) <= sf._CACHE_MAX
    assert (0, 64, 48, 4) not in _FakeExtractor.alive               # oldest entry went first
    # plain LRU behaviour without a keep set
    for i in range(20):
        sf._extractor(100 + i, 10, 1, 0)
    assert len(sf._cache) == sf._CACHE_MAX and len(_FakeExtractor.alive) == sf._CACHE_MAX


def test_shard_ranges_match_the_c_partition():
    for n in (1, 5, 7, 8, 9, 8192):
        for parts in (1, 2, 3, 4, 8, 16):
            r = sf.shard_ranges(n, parts)
            assert len(r) == parts and sum(len(x) for x in r) == n
            flat = [i for x in r for i in x]
            assert flat == list(range(n))                             # contiguous, in image order
            per = -(-n // parts)
            assert all(len(x) <= per for x in r) and all(len(x) == per for x in r[: n // per])


def test_descriptor_byte_model():
    import bench
    # KeyPoint.size = kp_scale * 2^octave / 2 with kp_scale in (1.796, 3.592): scale 2.1 -> radius 22 at any octave
    kp = np.zeros(3, sf.KEYPOINT_DTYPE)
    kp["size"] = [2.1 / 2, 2.1, 2.1 * 4]
    assert bench.descriptor_bytes(kp) == 3 * (45 * 45 * 4 + 144)
    assert bench.descriptor_bytes(kp[:0]) == 0.0


def test_every_pipeline_kernel_waits_for_its_predecessor():
    """Programmatic dependent launch: the kernels of a group are launched through `klaunch` with
    the programmatic-stream-serialization attribute, so a kernel may start while its predecessor is still draining.
    That is only safe because every kernel's FIRST statement is `pdl_wait()` (griddepcontrol.wait) -- a kernel added
    without it would race silently.  Static check of the sources: every __global__ in the kernel headers opens with
    the wait, and the enqueue functions launch through the helper only."""
    import os
    import re
    csrc = os.path.join(os.path.dirname(os.path.abspath(sf.__file__)), "csrc")
    n = 0
    for name in ("sb_pyramid.cuh", "sb_keypoints.cuh"):
        src = open(os.path.join(csrc, name)).read()
        for m in re.finditer(r"__global__[^{;]*\{\s*", src):
            n += 1
            assert src[m.end():].startswith("pdl_wait();"), f"{name}: kernel at offset {m.start()} does not open with pdl_wait()"
    assert n >= 20
    host = open(os.path.join(csrc, "sift_b200.cu")).read()
    for fn in ("int enqueue_pyramid_imageproc(", "int enqueue_pyramid(", "int enqueue_detect("):
        a = host.index(fn)
        b = host.index("\n}\n", a)
        assert "<<<" not in host[a:b], f"{fn} launches a kernel without the helper"
        assert "klaunch(" in host[a:b] or "launch_blur" in host[a:b]
