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
