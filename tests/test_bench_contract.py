"""bench.py's JSON line: the reference arm (CPU oracle port; runs without a GPU) carries every key the driver reads, and
the b200 arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "vga")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                   # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "images/s"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                               # on the GPU box the arm is exercised by the bench itself
    r = _bench("--steps", "1", "--warmup", "3", "--no-cpu", timeout=300)
    assert r.returncode != 0                                 # loud failure, no CPU fallback, no JSON line
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
