#!/usr/bin/env python
"""bench.py -- throughput of the SIFT extraction hot path on B200 (BASELINE.json metric:
"1080p images/sec at 1/2/4/8 B200; descriptors/sec; % of HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k|vga|desc|match|jpeg] [--impl b200|reference]

One "step" = one pass of the whole hot path (seed, pyramid, DoG/extrema, refinement, orientation,
descriptors) over one batch of synthetic gray images per GPU.  Prints ONE JSON line on rank 0.

  value     images/s over all GPUs with the inputs already resident in HBM (results stay on the device),
            timed with CUDA events on the launching streams, max over ranks.
  e2e       the same metric through the public host API (sb200_extract_batch): pinned host inputs,
            host->device and device->host copies inside the timed region.
  roofline  the heaviest kernel of the step (by measured share) against the measured HBM copy peak
            (MEASURED_PEAKS.json), plus kernels[]: the top kernels with their share of the step, the bound that
            binds each and its fraction, and the stage-level view (blur stage, pyramid + DoG).
  workloads the other BASELINE.json configs in the same line (short sub-runs): 640x480 batches, 3840x2160,
            descriptor-only (200k keypoints), single-image 1080p latency, the reference's own bench shapes on
            bird.jpg (full pipeline / precomputed pyramid), natural-image 1080p (tiled tree.jpg), e2e from
            pageable host memory.  Under torchrun every rank runs them and rank 0 reports the aggregate.
  config4   BASELINE.json configs[3]: ONE batch of 8192 640x480 images sharded over all N GPUs by the product's
            own multi-device entry point (sb200_extract_batch_multi_parts / _multi), driven by rank 0 alone
            while the other ranks wait on the host.
  cpu_baseline  the in-repo oracle (a C port of the crate; the crate itself is Rust and cannot be built in
            this image) timed on a bounded sample on the host, N=1 only.

--impl reference times that same CPU port on all host cores (the reference's own CPU path stand-in).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, images per group (= context max_batch), groups per step)
    "1080p": (1920, 1080, 32, 4),  # BASELINE.json configs[1] shape, batched
    "4k": (3840, 2160, 8, 2),      # configs[2]
    "vga": (640, 480, 128, 4),     # configs[3] shape (8192 images = 16 such steps)
}
CONFIG4_IMAGES = 8192              # BASELINE.json configs[3]


def workload_text(name, w, h, natural=False):
    src = "tiled tree.jpg (the reference's densest image), a distinct shift per image" if natural else "i.i.d. uniform noise"
    return (f"{name}: {w}x{h} gray u8 {src}, full SIFT extraction "
            "(pyramid + DoG/extrema + refinement + orientation + descriptors)")


def headline_config(name, world):
    """config of the JSON line -- the same object for the b200 arm and for the reference arm."""
    w, h, B, G = WORKLOADS[name]
    return {"workload": workload_text(name, w, h), "images_per_step_per_gpu": B * G, "groups_per_step": G,
            "images_per_group": B, "parallelism": f"independent image shards x{world}, no collective",
            "l2": f"per-step working set ({B * G} pyramids) and rotating input sets exceed the 126 MB L2"}


def synth_images(n, w, h, seed):
    """i.i.d. uniform u8 noise, a distinct stream per image."""
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        out[i] = np.random.default_rng([seed, i]).integers(0, 256, (h, w), dtype=np.uint8)
    return out


def natural_images(n, w, h, seed):
    """tests/golden/tree_gray.npy (images/tree.jpg, 800x600) tiled to w x h, rolled by a distinct offset per image."""
    tree = np.load(os.path.join(ROOT, "tests", "golden", "tree_gray.npy"))
    tile = np.tile(tree, (h // tree.shape[0] + 1, w // tree.shape[1] + 1))[:h, :w]
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        out[i] = np.roll(tile, ((53 * (i + seed)) % h, (37 * (i + seed)) % w), (0, 1))
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.t = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class Dist:
    """torch.distributed plumbing for the N>1 contract (barrier + max over ranks); no data-path collective.
    host_barrier() waits on a gloo group: the waiting ranks leave their GPUs idle (an NCCL barrier would spin a
    kernel on them), which is what rank 0 needs while it drives every device for the config-4 measurement."""

    def __init__(self, rank, local_rank, world, cuda=True):
        self.rank, self.world, self.torch, self.cpu_group = rank, world, None, None
        if world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            self.cuda = cuda and torch.cuda.is_available()
            if self.cuda:
                torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl" if self.cuda else "gloo", rank=rank, world_size=world)
            self.torch, self.dist = torch, dist
            self.cpu_group = dist.new_group(backend="gloo") if self.cuda else None

    def barrier(self):
        if self.world > 1:
            if self.cuda:
                self.torch.cuda.synchronize()
            self.dist.barrier()

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group) if self.cpu_group is not None else self.dist.barrier()

    def reduce(self, value, op="max"):
        if self.world == 1:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device="cuda" if self.cuda else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def oracle_images_per_s(imgs, threads):
    """Times the CPU port (oracle) on `imgs` with `threads` host threads; returns (images/s, keypoints)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.lib()
    t0 = time.perf_counter()
    if threads <= 1:
        counts = [len(O.sift(im)[0]) for im in imgs]
    else:
        with ThreadPoolExecutor(threads) as ex:   # ctypes releases the GIL: real parallelism
            counts = list(ex.map(lambda im: len(O.sift(im)[0]), imgs))
    dt = time.perf_counter() - t0
    return len(imgs) / dt, int(sum(counts))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The crate is Rust
    and cannot be compiled here (no rustc/cargo), so this is the oracle port (kind "port").  Same config, metric,
    unit, steps and warm-up as the b200 arm; each step is a bounded sample of the step's images (one per host
    thread, at most 16) so that the run ends within minutes."""
    if rank != 0:
        return
    w, h, B, G = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    per_step = max(1, min(cores, 16 if args.workload != "4k" else 4))
    imgs = synth_images(per_step, w, h, 1234)
    for _ in range(args.warmup):
        oracle_images_per_s(imgs[: max(1, per_step // 4)], cores)
    t0 = time.perf_counter()
    kps = 0
    steps = max(1, args.steps)
    for _ in range(steps):
        _, k = oracle_images_per_s(imgs, cores)
        kps += k
    dt = time.perf_counter() - t0
    val = steps * per_step / dt
    line = {
        "impl": "reference", "metric": f"{args.workload} images/sec", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args.workload, world),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {per_step} of the step's {B * G} images on {cores} threads (oracle C "
                                   "port of src/lib.rs; the Rust crate cannot be built in this image)"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "keypoints_per_s": kps / dt,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def descriptor_bytes(kp):
    """Algorithmic bytes of compute_descriptor for the public keypoints `kp`:
    sum over keypoints of (2r+1)^2 * 4 (window pixels) + 128 (descriptor) + 16 (keypoint record), with
    r = round(3 * sigma_octave * sqrt(2) * 2.5) and sigma_octave in (1.796, 3.592) recovered from KeyPoint.size."""
    if len(kp) == 0:
        return 0.0
    size2 = 2.0 * kp["size"].astype(np.float64)          # seed-image size = kp_scale * 2^octave
    octave = np.floor(np.log2(size2 / 1.7959))
    sigma = size2 / np.exp2(octave)
    r = np.round(3.0 * sigma * np.sqrt(2.0) * 2.5)
    return float(((2 * r + 1) ** 2 * 4 + 144).sum())


class Measure:
    """One shape on one device: device-resident throughput (CUDA events), e2e through the host API, per-stage and
    per-launch times in a serialised repeat."""

    def __init__(self, lib, sf, dist, rank, local_rank, world):
        self.lib, self.sf, self.dist, self.rank, self.local_rank, self.world = lib, sf, dist, rank, local_rank, world

    def run(self, w, h, B, G, steps, warmup, natural=False, profile=True, pageable=False, keep_sets=False,
            processing=None):
        lib, sf, dist = self.lib, self.sf, self.dist
        per_step = B * G
        ex = sf.Extractor(w, h, B, device=self.local_rank, processing=processing or sf.OpenCVProcessing)
        H = ex.handle

        def chk(st):
            if st:
                raise RuntimeError(lib.sb200_last_error(H).decode())

        # distinct input sets, rotated so that consecutive steps never reuse L2-resident inputs
        set_bytes = per_step * w * h
        n_sets = int(min(max(2, (160 << 20) // set_bytes + 2), 24, max(2, steps + warmup)))
        sets_h, sets_d = [], []
        for s in range(n_sets):
            p = C.c_void_p()
            chk(lib.sb200_host_alloc(set_bytes, C.byref(p)))
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(per_step, h, w))
            arr[...] = (natural_images if natural else synth_images)(per_step, w, h, 1234 + 1000 * self.rank + s)
            d = C.c_void_p()
            chk(lib.sb200_device_alloc(H, set_bytes, C.byref(d)))
            chk(lib.sb200_memcpy_h2d(H, d, p, set_bytes))
            sets_h.append((p, arr)); sets_d.append(d)

        def device_step(i):
            d = sets_d[i % n_sets].value
            for g in range(G):
                chk(lib.sb200_extract_batch_device(H, d + g * B * w * h, B, w, h, w, w * h, -1))

        res = _ffi_result()

        def host_step(i, src=None):
            p = src if src is not None else sets_h[i % n_sets][0]
            chk(lib.sb200_extract_batch(H, p, per_step, w, h, w, w * h, -1, C.byref(res)))
            return int(res.n)

        for i in range(warmup):
            device_step(i)
        chk(lib.sb200_sync(H))
        counts = (C.c_uint32 * B)()
        chk(lib.sb200_device_result(H, counts, B, None, None, None))
        kp_per_group = int(sum(counts))
        # ---- timed region: device-resident inputs, CUDA events on the launching streams --------------
        chk(lib.sb200_set_profiling(H, 0))
        chk(lib.sb200_reset_stats(H))
        dist.barrier()
        l0 = ex.launch_count
        t_wall0 = time.perf_counter()
        chk(lib.sb200_timer_start(H))
        for i in range(steps):
            device_step(warmup + i)
        chk(lib.sb200_timer_stop(H))
        ms = C.c_float()
        chk(lib.sb200_timer_elapsed_ms(H, C.byref(ms)))
        chk(lib.sb200_sync(H))
        dist.barrier()
        launches = ex.launch_count - l0
        dev_ms = dist.reduce(ms.value, "max")
        # the pyramid stages alone under the production schedule (two slots in flight, CUDA graphs): what the
        # pyramid / DoG kernels achieve when the tail of one group overlaps the head of the next
        pyr_ms = None
        if profile:
            for i in range(2):
                d = sets_d[i % n_sets].value
                for g in range(G):
                    chk(lib.sb200_pyramid_batch_device(H, d + g * B * w * h, B, w, h, w, w * h))
            chk(lib.sb200_sync(H))
            chk(lib.sb200_timer_start(H))
            for i in range(steps):
                d = sets_d[(warmup + i) % n_sets].value
                for g in range(G):
                    chk(lib.sb200_pyramid_batch_device(H, d + g * B * w * h, B, w, h, w, w * h))
            chk(lib.sb200_timer_stop(H))
            chk(lib.sb200_timer_elapsed_ms(H, C.byref(ms)))
            chk(lib.sb200_sync(H))
            pyr_ms = ms.value
        # per-stage / per-launch device times: the same steps again with CUDA events bracketing every stage on the
        # launching stream and a sync after each group, so that no kernel of the other slot runs concurrently and
        # inflates a stage (the timed region above overlaps the two slots for throughput)
        stats = fine = None
        if profile and self.rank == 0:
            chk(lib.sb200_set_profiling(H, 1))
            chk(lib.sb200_reset_stats(H))
            for i in range(steps):
                d = sets_d[(warmup + i) % n_sets].value
                for g in range(G):
                    chk(lib.sb200_extract_batch_device(H, d + g * B * w * h, B, w, h, w, w * h, -1))
                    chk(lib.sb200_sync(H))
            stats = ex.stage_stats()
            fine = ex.launch_stats()
            chk(lib.sb200_set_profiling(H, 0))
        # ---- e2e: host buffers through the public API, wall clock around synchronous calls --------------
        for i in range(min(warmup, 2)):
            host_step(i)
        dist.barrier()
        t0 = time.perf_counter()
        kp_total = 0
        for i in range(steps):
            kp_total += host_step(warmup + i)
        t_e2e = dist.reduce(time.perf_counter() - t0, "max")
        kp_all = dist.reduce(kp_total, "sum")
        _, kp_last, _ = ex._take(res)
        desc_bytes_per_step = descriptor_bytes(kp_last)
        d2h = kp_total / max(1, steps) * (20 + 128) + (per_step + 1) * 8
        out = {
            "value": steps * per_step * self.world / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps, "dev_ms": dev_ms,
            "e2e": steps * per_step * self.world / t_e2e, "h2d": set_bytes, "d2h": int(d2h), "launches": int(launches),
            "kp_per_image": kp_per_group / B, "kp_per_s": kp_per_group * G * steps * self.world / (dev_ms * 1e-3),
            "e2e_kp_per_s": kp_all / t_e2e, "stats": stats, "fine": fine, "per_step": per_step, "n_sets": n_sets,
            "desc_bytes_per_image": desc_bytes_per_step / per_step, "t_wall0": t_wall0, "steps": steps,
            "pyramid_only_ms": pyr_ms,
        }
        if pageable:
            # the crate's callers hold pageable Vec<u8>: the library packs them into its pinned staging buffers
            pg = [np.array(sets_h[s][1], copy=True) for s in range(min(n_sets, 3))]
            for i in range(2):
                host_step(i, pg[i % len(pg)].ctypes.data)
            dist.barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                host_step(i, pg[i % len(pg)].ctypes.data)
            out["e2e_pageable"] = steps * per_step * self.world / dist.reduce(time.perf_counter() - t0, "max")
        if keep_sets:
            out["ctx"] = (ex, sets_h, sets_d, device_step, chk)
        else:
            for (p, _), d in zip(sets_h, sets_d):
                lib.sb200_device_free(H, d)
                lib.sb200_host_free(p)
            ex.close()
        return out


def _ffi_result():
    from sift_features_b200 import _ffi
    return _ffi.Result()


def stage_view(m, w, h, peak, sf):
    """pyramid / DoG stage fractions of the HBM roofline from the serialised stage times."""
    st = m["stats"]
    if st is None:
        return None
    tot, a_seed, a_blur, a_ext = sf.algorithmic_bytes(w, h)
    imgs = m["per_step"] * m["steps"]   # images the serialised repeat processed
    pyr_ms = st["seed"]["ms"] + st["blur"]["ms"] + st["extrema"]["ms"]

    def frac(bytes_img, ms):
        return bytes_img * imgs / (ms * 1e-3) / 1e9 / peak if ms > 0 else None
    return {
        "stages_ms_per_image": {k: v["ms"] / imgs for k, v in st.items()},
        "blur_stage_frac": frac(a_blur, st["blur"]["ms"]),
        "extrema_stage_frac": frac(a_ext, st["extrema"]["ms"]),
        "pyramid_dog_frac": frac(tot, pyr_ms),
        "pyramid_dog_pipelined_frac": frac(tot, m["pyramid_only_ms"]) if m.get("pyramid_only_ms") else None,
        "pyramid_dog_ms_per_image": pyr_ms / imgs,
        "algorithmic_bytes_per_image": tot,
        "descriptor_ns_per_keypoint": 1e6 * st["descriptor"]["ms"] / imgs / max(m["kp_per_image"], 1),
        "whole_path_frac_of_hbm_roofline": tot * m["value"] / m.get("world", 1) / 1e9 / peak,
    }


def kernel_table(m, w, h, B, peak, sf, sm_mhz):
    """Top kernels of the step: share of the serialised step, the bound that binds each, fraction of that bound."""
    st, fine = m["stats"], m["fine"]
    if st is None:
        return None, None
    tot, a_seed, a_blur, a_ext = sf.algorithmic_bytes(w, h)
    imgs = m["per_step"] * m["steps"]   # images the serialised repeat processed
    step_ms = sum(st[k]["ms"] for k in ("seed", "blur", "extrema", "refine", "orient", "descriptor"))
    dims, cw, ch = [], 2 * w, 2 * h
    while len(dims) < 16 and min(cw, ch) >= 1:
        dims.append((cw, ch)); cw //= 2; ch //= 2
    rows = []

    def add(name, ms, n_launch, bytes_img, bound, note=None, extra=None):
        """ms / n_launch: total device time and number of launches of this kernel in the serialised repeat"""
        if ms <= 0 or n_launch <= 0:
            return
        ach = bytes_img * imgs / (ms * 1e-3) / 1e9
        r = {"kernel": name, "ms_per_image": ms / imgs, "share_of_step": ms / step_ms, "launches": int(n_launch),
             "avg_launch_us": 1e3 * ms / n_launch, "bound": bound,
             "algorithmic_bytes_per_launch": bytes_img * imgs / n_launch, "achieved_gbs": ach, "hbm_frac": ach / peak}
        if note:
            r["note"] = note
        if extra:
            r.update(extra)
        rows.append(r)

    groups = max(1, imgs // B)
    desc_extra = {"ns_per_keypoint": 1e6 * st["descriptor"]["ms"] / imgs / max(m["kp_per_image"], 1)}
    try:   # the roofline this kernel is bound by: warp-instructions (ncu smsp__inst_executed.sum of the same kernel on the
        # 1080p noise workload, profiles/traffic.json) against the issue slots of the SMs, 4 per SM and clock
        wipk = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_descriptor_1080p"]["warp_instructions_per_keypoint"]
        if sm_mhz and st["descriptor"]["ms"] > 0:
            desc_extra["warp_instructions_per_keypoint"] = wipk
            desc_extra["issue_slot_frac"] = (wipk * m["kp_per_image"] * imgs / (st["descriptor"]["ms"] * 1e-3)
                                             / (148 * 4 * sm_mhz * 1e6))
    except Exception:
        pass
    add("k_descriptor", st["descriptor"]["ms"], groups, m["desc_bytes_per_image"],
        "issue", "instruction-issue bound on L2-resident patches (ncu: profiles/r02_*): the HBM fraction is low by nature; "
                 "issue_slot_frac = measured warp-instructions per second against 4 issue slots per SM and clock",
        desc_extra)
    add("k_orient (+ k_kpscan, k_emit)", st["orient"]["ms"], groups, 0.0, "issue / latency",
        "ordered per-bin accumulation on L2-resident patches")
    ex_ms = sum(v[0] for (o, k), v in fine.items() if k == "extrema")
    add("k_extrema_tma", ex_ms, sum(v[1] for (o, k), v in fine.items() if k == "extrema"), a_ext, "hbm",
        "24 B/px: six Gaussian layers read once")
    lane_ops = {1: 22, 2: 26, 3: 34, 4: 42, 5: 54}   # FP32 lane-operations per pixel the oracle's arithmetic fixes
    for l in (5, 4, 3, 2, 1):
        ms = sum(v[0] for (o, k), v in fine.items() if k == f"blur{l}")
        px = sum(dims[o][0] * dims[o][1] for (o, k) in fine if k == f"blur{l}")
        extra = None
        if sm_mhz and ms > 0:
            extra = {"fp32_pipe_frac": lane_ops[l] * px * imgs / (ms * 1e-3) / (148 * 128 * sm_mhz * 1e6)}
        add(f"k_blur_march<{l}>", ms, sum(v[1] for (o, k), v in fine.items() if k == f"blur{l}"), 8.0 * px,
            "hbm" if l <= 2 else "fp32 pipe / hbm", f"{[11, 13, 17, 21, 27][l - 1]}-tap separable Gaussian, 8 B/px, all octaves >= 32 px", extra)
    seed_ms = sum(v[0] for (o, k), v in fine.items() if k == "seed")
    if os.environ.get("SB200_SEED") == "split":
        add("k_upsample2x + k_blur_march<0>", seed_ms, 2 * groups, a_seed + 8.0 * dims[0][0] * dims[0][1], "hbm",
            "u8 -> f32 2x upsample, then the 11-tap seed blur (the upsampled image is written and read once more than "
            "the algorithmic model counts)")
    else:
        add("k_blur_march<0, seed>", seed_ms, groups, a_seed, "issue / shared memory",
            "u8 -> f32 2x upsample by producer warps straight into the stage buffers of the 11-tap seed blur: "
            "0.25 B in + 4 B out per seed pixel, the upsampled image never goes through HBM")
    tail_ms = sum(v[0] for (o, k), v in fine.items() if k == "tail")
    add("k_tail", tail_ms, groups, 0.0, "latency", "all octaves of at most 64x36 px in one launch")
    rows.sort(key=lambda r: -r["share_of_step"])
    return rows, step_ms / imgs


def run_config4(args, lib, sf, dist, rank, world):
    """BASELINE.json configs[3]: ONE batch of 8192 640x480 images sharded over the N devices of the box by the
    product's multi-device entry point, from one process (rank 0); the other ranks wait on the host."""
    from sift_features_b200 import _ffi
    out = None
    if rank == 0:
        w, h, B = 640, 480, 128
        n = CONFIG4_IMAGES
        ndev = world
        exs = [sf.Extractor(w, h, B, device=d) for d in range(ndev)]
        H0 = exs[0].handle

        def chk(st):
            if st:
                raise RuntimeError(lib.sb200_last_error(H0).decode())
        p = C.c_void_p()
        chk(lib.sb200_host_alloc(n * w * h, C.byref(p)))
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n, h, w))
        rng = np.random.default_rng(4321)
        for i in range(0, n, 256):   # i.i.d. uniform noise, one stream for the whole batch
            arr[i:i + 256] = rng.integers(0, 256, (min(256, n - i), h, w), dtype=np.uint8)
        handles = (C.c_void_p * ndev)(*[e.handle for e in exs])
        parts = (_ffi.Result * ndev)()
        first = (C.c_uint64 * (ndev + 1))()
        res = _ffi.Result()

        def run_parts(k):
            chk(lib.sb200_extract_batch_multi_parts(handles, k, p, n, w, h, w, w * h, -1, parts, first))
            return sum(int(parts[d].n) for d in range(k))

        def run_dense(k):
            chk(lib.sb200_extract_batch_multi(handles, k, p, n, w, h, w, w * h, -1, C.byref(res)))
            return int(res.n)

        def timed(fn, k, reps):
            fn(k)                                   # warm-up: graphs, pinned result arrays
            t0 = time.perf_counter()
            kp = 0
            for _ in range(reps):
                kp = fn(k)
            return (time.perf_counter() - t0) / reps, kp
        reps = 2
        t_parts, kp = timed(run_parts, ndev, reps)
        shard_ms = [round(lib.sb200_last_shard_ms(e.handle), 2) for e in exs]
        t_dense, _ = timed(run_dense, ndev, reps)
        gather_ms = lib.sb200_last_gather_ms(H0)
        out = {"workload": f"ONE batch of {n} 640x480 noise images sharded over {ndev} GPU(s) by "
                           "sb200_extract_batch_multi_parts (host buffers in, host results out; one host thread per device)",
               "n_gpus": ndev, "images": n, "keypoints": kp,
               "e2e_images_s": n / t_parts, "e2e_keypoints_s": kp / t_parts, "ms_per_batch": 1e3 * t_parts,
               "dense_gather": {"e2e_images_s": n / t_dense, "gather_ms": gather_ms,
                                "gather_frac_of_step": gather_ms / (1e3 * t_dense),
                                "note": "sb200_extract_batch_multi: the same plus a multi-threaded host concatenation "
                                        "into one dense result array"},
               "shard_ms_per_device": shard_ms,
               "h2d_bytes": n * w * h, "d2h_bytes": int(kp * 148 + (n + 1) * 8)}
        if ndev > 1:
            t1, _ = timed(run_parts, 1, 1)
            out["one_device_e2e_images_s"] = n / t1
            out["efficiency_vs_1dev"] = (n / t_parts) / (ndev * (n / t1))
        for e in exs:
            e.close()
        lib.sb200_host_free(p)
    dist.host_barrier()
    return out


def sub_workloads(args, M, lib, sf, dist, rank, local_rank, world, peak):
    """The other BASELINE.json configs as short sub-runs whose results ride in the same JSON line."""
    out = {}
    steps = max(3, min(args.steps, 6))

    def shape(name, natural=False, pageable=False, processing=None):
        w, h, B, G = WORKLOADS[name]
        m = M.run(w, h, B, G, steps, 3, natural=natural, pageable=pageable, processing=processing)
        m["world"] = world
        text = workload_text(name, w, h, natural)
        if processing is not None:
            text += " -- pyramid in the arithmetic of the crate's default Processing (ImageprocProcessing, src/lib.rs:992-1007)"
        r = {"workload": text, "images_per_step_per_gpu": B * G, "steps": steps,
             "value": m["value"], "unit": "images/s", "e2e": m["e2e"], "keypoints_per_image": m["kp_per_image"],
             "keypoints_per_s": m["kp_per_s"]}
        if "e2e_pageable" in m:
            r["e2e_pageable"] = m["e2e_pageable"]
        if rank == 0:
            sv = stage_view(m, w, h, peak, sf)
            if sv:
                r.update({k: sv[k] for k in ("pyramid_dog_frac", "pyramid_dog_pipelined_frac", "blur_stage_frac", "extrema_stage_frac",
                                             "descriptor_ns_per_keypoint", "whole_path_frac_of_hbm_roofline")})
                r["stages_ms_per_image"] = sv["stages_ms_per_image"]
        return r
    out["vga"] = shape("vga")
    out["4k"] = shape("4k")
    out["natural_1080p"] = shape("1080p", natural=True)
    out["imageproc_natural_1080p"] = shape("1080p", natural=True, processing=sf.ImageprocProcessing)
    out["desc"] = desc_measure(args, lib, sf, dist, rank, local_rank, world, steps, peak)
    if rank == 0:
        out["single_image"] = single_image(lib, sf, local_rank)
    return out


def single_image(lib, sf, device):
    """BASELINE.json configs[1] and [2] as literally written (ONE image per call through the public API), and the
    reference's own bench shapes on bird.jpg (benches/sift.rs:88-97 full pipeline, :115-121 precomputed pyramid)."""
    out = {}
    for name, (w, h) in {"1080p": (1920, 1080), "4k": (3840, 2160), "vga": (640, 480)}.items():
        img = synth_images(1, w, h, 77)[0]
        with sf.Extractor(w, h, 1, device=device) as ex:
            for _ in range(5):
                r = ex.sift(img)
            n = 30 if name != "4k" else 12
            t0 = time.perf_counter()
            for _ in range(n):
                r = ex.sift(img)
            dt = (time.perf_counter() - t0) / n
            # the same call at the C ABI (no numpy result copies), from pinned memory, and with the image resident
            H = ex.handle
            res = _ffi_result()
            pin = C.c_void_p()
            lib.sb200_host_alloc(img.nbytes, C.byref(pin))
            C.memmove(pin, img.ctypes.data, img.nbytes)
            dev = C.c_void_p()
            lib.sb200_device_alloc(H, img.nbytes, C.byref(dev))
            lib.sb200_memcpy_h2d(H, dev, img.ctypes.data, img.nbytes)

            def med(f):
                for _ in range(3):
                    f()
                ts = []
                for _ in range(n):
                    t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
                return 1e3 * float(np.median(ts))
            abi = {
                "pageable_host_ms": med(lambda: lib.sb200_extract_batch(H, img.ctypes.data, 1, w, h, w, w * h, -1, C.byref(res))),
                "pinned_host_ms": med(lambda: lib.sb200_extract_batch(H, pin, 1, w, h, w, w * h, -1, C.byref(res))),
                "device_resident_ms": med(lambda: (lib.sb200_extract_batch_device(H, dev, 1, w, h, w, w * h, -1), lib.sb200_sync(H))),
                "pyramid_only_ms": med(lambda: (lib.sb200_pyramid_batch_device(H, dev, 1, w, h, w, w * h), lib.sb200_sync(H))),
            }
            lib.sb200_device_free(H, dev)
            lib.sb200_host_free(pin)
        out[name] = {"ms_per_image": 1e3 * dt, "images_s": 1.0 / dt, "keypoints": len(r),
                     "call": "sift(img) on one pageable host image, result copied back into numpy arrays",
                     "c_abi": abi}
    bird = np.load(os.path.join(ROOT, "tests", "golden", "bird_gray.npy"))
    with sf.Extractor(bird.shape[1], bird.shape[0], 1, device=device) as ex:
        for _ in range(5):
            r = ex.sift(bird)
        t0 = time.perf_counter()
        for _ in range(30):
            r = ex.sift(bird)
        t_full = (time.perf_counter() - t0) / 30
        ex.precompute_images(bird)
        for _ in range(5):
            ex.sift_with_precomputed()
        t0 = time.perf_counter()
        for _ in range(30):
            r2 = ex.sift_with_precomputed()
        t_pre = (time.perf_counter() - t0) / 30
    out["bird_jpg_799x533"] = {
        "sift_with_opencv_preprocess_ms": 1e3 * t_full, "sift_no_preprocess_ms": 1e3 * t_pre, "keypoints": len(r),
        "shapes": "benches/sift.rs:88-97 (full pipeline) and :115-121 (sift_with_precomputed on a resident pyramid)",
        "same_result": bool(len(r) == len(r2))}
    try:
        import cv2
        cv2.setNumThreads(1)
        s = cv2.SIFT_create()
        t0 = time.perf_counter()
        k, _ = s.detectAndCompute(bird, None)
        out["bird_jpg_799x533"]["opencv_sift_cpu_ms"] = 1e3 * (time.perf_counter() - t0)   # benches/sift.rs:99-113
    except Exception:
        pass
    return out


def desc_measure(args, lib, sf, dist, rank, local_rank, world, steps, peak):
    """BASELINE.json configs[4]: descriptor-only, 200k precomputed keypoints, benches/descriptor.rs shape."""
    w, h, n = 1920, 1080, 200_000
    ex = sf.Extractor(w, h, 1, device=local_rank)
    H = ex.handle

    def chk(st):
        if st:
            raise RuntimeError(lib.sb200_last_error(H).decode())
    img = synth_images(1, w, h, 1234 + rank)[0].astype(np.float32) / np.float32(255)
    rng = np.random.default_rng(99 + rank)
    k = np.empty((n, 4), np.float32)
    k[:, 0] = rng.uniform(0, w, n); k[:, 1] = rng.uniform(0, h, n); k[:, 2] = 2.1; k[:, 3] = 123.0
    d_img, d_k, d_out = C.c_void_p(), C.c_void_p(), C.c_void_p()
    chk(lib.sb200_device_alloc(H, img.nbytes, C.byref(d_img)))
    chk(lib.sb200_device_alloc(H, k.nbytes, C.byref(d_k)))
    chk(lib.sb200_device_alloc(H, n * 128, C.byref(d_out)))
    chk(lib.sb200_memcpy_h2d(H, d_img, img.ctypes.data, img.nbytes))
    chk(lib.sb200_memcpy_h2d(H, d_k, k.ctypes.data, k.nbytes))
    for _ in range(3):
        chk(lib.sb200_compute_descriptors_device(H, d_img, w, h, w, d_k, n, d_out))
    chk(lib.sb200_sync(H))
    dist.barrier()
    l0 = ex.launch_count
    chk(lib.sb200_timer_start(H))
    for _ in range(steps):
        chk(lib.sb200_compute_descriptors_device(H, d_img, w, h, w, d_k, n, d_out))
    chk(lib.sb200_timer_stop(H))
    ms = C.c_float()
    chk(lib.sb200_timer_elapsed_ms(H, C.byref(ms)))
    chk(lib.sb200_sync(H))
    dev_ms = dist.reduce(ms.value, "max")
    launches = ex.launch_count - l0
    out = np.zeros((n, 128), np.uint8)
    t0 = time.perf_counter()
    for _ in range(steps):
        chk(lib.sb200_compute_descriptors(H, img.ctypes.data, w, h, w, k.ctypes.data, n, out.ctypes.data))
    t_e2e = dist.reduce(time.perf_counter() - t0, "max")
    bytes_kp = 45 * 45 * 4 + 128 + 16
    val = steps * n * world / (dev_ms * 1e-3)
    r = {"workload": "descriptor-only: 200k keypoints (scale 2.1, 123 deg, benches/descriptor.rs shape) on one "
                     "1920x1080 f32 noise image", "value": val, "unit": "keypoints/s", "ns_per_keypoint": 1e9 * world / val,
         "e2e": steps * n * world / t_e2e, "steps": steps, "gpu_launches": int(launches),
         "h2d_bytes_per_step": int(img.nbytes + k.nbytes), "d2h_bytes_per_step": n * 128,
         "hbm_frac": bytes_kp * val / world / 1e9 / peak, "algorithmic_bytes_per_keypoint": bytes_kp,
         "k": k, "img": img}
    for d in (d_img, d_k, d_out):
        lib.sb200_device_free(H, d)
    ex.close()
    return r


# ---------------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import sift_features_b200 as sf
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    dist = Dist(rank, local_rank, world)
    M = Measure(lib, sf, dist, rank, local_rank, world)
    w, h, B, G = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    if args.groups:
        G = args.groups
    peak, peak_src = peaks()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    m = M.run(w, h, B, G, args.steps, args.warmup, profile=args.profile_stages, pageable=True, keep_sets=True,
              natural=args.natural)
    m["world"] = world
    ex, sets_h, sets_d, device_step, chk = m.pop("ctx")
    H = ex.handle
    # clocks: the timed regions are short; keep the GPU under the same load until the sampler has >= 5 samples
    clocks = None
    if sampler:
        extra_t0 = time.perf_counter()
        i = 0
        while time.perf_counter() - extra_t0 < 1.0:
            device_step(i); i += 1
            if i % 8 == 0:
                chk(lib.sb200_sync(H))
        chk(lib.sb200_sync(H))
        sampler.stop()
        clocks = sampler.summary(m["t_wall0"], time.perf_counter())
        clocks["window"] = "timed region + e2e region + 1 s repeat of the timed loop"
    cpu_imgs = [np.array(a, copy=True) for a in sets_h[0][1][:24]] if rank == 0 else None
    for (p, _), d in zip(sets_h, sets_d):
        lib.sb200_device_free(H, d)
        lib.sb200_host_free(p)
    ex.close()
    dist.barrier()

    cfg = headline_config(args.workload, world)
    if args.natural:
        cfg["workload"] = workload_text(args.workload, w, h, True)
    if args.batch or args.groups:
        cfg.update({"images_per_step_per_gpu": B * G, "groups_per_step": G, "images_per_group": B})
    line = {
        "metric": f"{args.workload} images/sec", "value": m["value"], "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": m["e2e"], "unit": "images/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "from_pageable_host_memory": m.get("e2e_pageable")},
        "gpu_launches": m["launches"],
        "keypoints_per_s": m["kp_per_s"], "keypoints_per_image": m["kp_per_image"], "e2e_keypoints_per_s": m["e2e_kp_per_s"],
        "clocks": clocks,
    }
    if rank == 0 and m["stats"] is not None:
        sm_mhz = (clocks or {}).get("sm_mhz")
        rows, step_ms_img = kernel_table(m, w, h, B, peak, sf, sm_mhz)
        sv = stage_view(m, w, h, peak, sf)
        top = rows[0]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                t = json.load(open(tp)).get(f"{top['kernel'].split(' ')[0]}_{args.workload}")
                if t:
                    traffic = t["dram_bytes_per_image"] * B   # ncu --set full capture, scaled to this launch
            except Exception:
                pass
        line["roofline"] = {
            "bound": "hbm", "kernel": top["kernel"] + " -- the heaviest kernel of the step by measured time share",
            "achieved": top["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": top["hbm_frac"], "traffic": traffic,
            "peak_source": peak_src, "share_of_step": top["share_of_step"],
            "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"], "avg_launch_us": top["avg_launch_us"],
            "measured": "CUDA events around the launches on their launching stream, in a serialised repeat of the timed "
                        "steps (one group in flight)",
            "note": top.get("note"),
            "kernels": rows,
            "serialised_step_ms_per_image": step_ms_img,
            "blur_stage": {"frac": sv["blur_stage_frac"]},
            "extrema_stage": {"frac": sv["extrema_stage_frac"]},
            "pyramid_dog": {"algorithmic_bytes_per_image": sv["algorithmic_bytes_per_image"],
                            "ms_per_image": sv["pyramid_dog_ms_per_image"], "frac": sv["pyramid_dog_frac"],
                            "frac_pipelined": sv["pyramid_dog_pipelined_frac"],
                            "note": "frac: seed + blur + extrema stage times of the serialised repeat (one group in "
                                    "flight, a sync after every group); frac_pipelined: the same stages alone under the "
                                    "production schedule (CUDA graphs, two groups in flight)"},
        }
        line["stages_ms_per_image"] = sv["stages_ms_per_image"]
        line["whole_path_frac_of_hbm_roofline"] = sv["whole_path_frac_of_hbm_roofline"]
    # ---- the other configs of BASELINE.json, in the same line ------------------------------------------
    if args.workload == "1080p" and not args.no_extra:
        wl = sub_workloads(args, M, lib, sf, dist, rank, local_rank, world, peak)
        desc_k, desc_img = wl["desc"].pop("k"), wl["desc"].pop("img")
        line["workloads"] = wl
        line["config4"] = run_config4(args, lib, sf, dist, rank, world)
    if rank == 0:
        if world == 1 and not args.no_cpu:
            n_cpu = {"1080p": 3, "4k": 1, "vga": 24}[args.workload]
            v, _ = oracle_images_per_s(cpu_imgs[:n_cpu], 1)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": 1, "kind": "port",
                                    "sample": f"{n_cpu} of the step's images, single thread (the crate is "
                                              "single-threaded); oracle C port of src/lib.rs"}
            if "workloads" in line:
                from oracle import oracle as O
                nd = 1000
                t0 = time.perf_counter()
                for r in desc_k[:nd]:
                    O.compute_descriptor(desc_img, *map(float, r))
                line["workloads"]["desc"]["cpu_baseline"] = {
                    "value": nd / (time.perf_counter() - t0), "unit": "keypoints/s", "cores": 1, "kind": "port",
                    "sample": f"first {nd} of the 200k keypoints, single thread"}
            # the reference's own second bench (benches/sift.rs:99-113, `opencv_sift`): OpenCV's SIFT on one image
            try:
                import cv2
                cv2.setNumThreads(1)
                im = np.ascontiguousarray(cpu_imgs[0])
                sift = cv2.SIFT_create()
                t0 = time.perf_counter()
                kps, _ = sift.detectAndCompute(im, None)
                dt = time.perf_counter() - t0
                line["opencv_sift_cpu"] = {"value": 1.0 / dt, "unit": "images/s", "cores": 1, "keypoints": len(kps),
                                           "sample": "cv2.SIFT_create().detectAndCompute on 1 image, cv2.setNumThreads(1)"}
            except Exception as e:   # informational only
                line["opencv_sift_cpu"] = {"unavailable": str(e)[:80]}
        print(json.dumps(line), flush=True)
    dist.close()


def run_desc(args, rank, local_rank, world):
    """--workload desc: BASELINE.json configs[4] as its own line."""
    import sift_features_b200 as sf
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    dist = Dist(rank, local_rank, world)
    peak, peak_src = peaks()
    r = desc_measure(args, lib, sf, dist, rank, local_rank, world, args.steps, peak)
    k, img = r.pop("k"), r.pop("img")
    if rank == 0:
        line = {"metric": "descriptors/sec", "value": r["value"], "unit": "keypoints/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 200_000 * world / r["value"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": r["workload"]},
                "e2e": {"value": r["e2e"], "unit": "keypoints/s", "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": r["d2h_bytes_per_step"]},
                "gpu_launches": r["gpu_launches"], "ns_per_keypoint": r["ns_per_keypoint"],
                "roofline": {"bound": "hbm", "kernel": "k_descriptor_list", "achieved": r["hbm_frac"] * peak,
                             "peak": peak, "unit": "GB/s", "frac": r["hbm_frac"], "traffic": None, "peak_source": peak_src,
                             "note": "instruction-issue bound on L2-resident patches; HBM fraction is low by nature"}}
        if world == 1 and not args.no_cpu:
            from oracle import oracle as O
            nd = 2000
            t0 = time.perf_counter()
            for row in k[:nd]:
                O.compute_descriptor(img, *map(float, row))
            line["cpu_baseline"] = {"value": nd / (time.perf_counter() - t0), "unit": "keypoints/s", "cores": 1,
                                    "kind": "port", "sample": f"first {nd} of the 200k keypoints, single thread"}
        print(json.dumps(line), flush=True)
    dist.close()


def run_match(args, rank, local_rank, world):
    """Descriptor matching (examples/sift-match.rs:30-35): mutual nearest neighbours of
    two 1080p-sized descriptor sets (8648 x 128 u8 each, the keypoint count of the 1080p workload)."""
    import sift_features_b200 as sf
    from sift_features_b200 import _ffi
    lib = _ffi.load()
    dist = Dist(rank, local_rank, world)
    n = 8648
    ex = sf.Extractor(8, 8, 1, device=local_rank)
    H = ex.handle

    def chk(st):
        if st:
            raise RuntimeError(lib.sb200_last_error(H).decode())
    rng = np.random.default_rng(7 + rank)
    q = np.minimum(rng.gamma(0.6, 30.0, (n, 128)), 255).astype(np.uint8)
    t = np.minimum(rng.gamma(0.6, 30.0, (n, 128)), 255).astype(np.uint8)
    t[: n // 2] = q[rng.permutation(n)[: n // 2]]
    d_q, d_t = C.c_void_p(), C.c_void_p()
    chk(lib.sb200_device_alloc(H, q.nbytes, C.byref(d_q)))
    chk(lib.sb200_device_alloc(H, t.nbytes, C.byref(d_t)))
    chk(lib.sb200_memcpy_h2d(H, d_q, q.ctypes.data, q.nbytes))
    chk(lib.sb200_memcpy_h2d(H, d_t, t.ctypes.data, t.nbytes))
    out = np.zeros(n, np.dtype([("query", np.uint32), ("train", np.uint32), ("dist2", np.uint32)]))
    cnt = C.c_uint64()
    for _ in range(args.warmup):
        chk(lib.sb200_match_descriptors_device(H, d_q, n, d_t, n, out.ctypes.data, n, C.byref(cnt)))
    dist.barrier()
    l0 = ex.launch_count
    chk(lib.sb200_timer_start(H))
    for _ in range(args.steps):
        chk(lib.sb200_match_descriptors_device(H, d_q, n, d_t, n, out.ctypes.data, n, C.byref(cnt)))
    chk(lib.sb200_timer_stop(H))
    ms = C.c_float()
    chk(lib.sb200_timer_elapsed_ms(H, C.byref(ms)))
    dev_ms = dist.reduce(ms.value, "max")
    launches = ex.launch_count - l0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        chk(lib.sb200_match_descriptors(H, q.ctypes.data, n, t.ctypes.data, n, out.ctypes.data, n, C.byref(cnt)))
    t_e2e = dist.reduce(time.perf_counter() - t0, "max")
    if rank == 0:
        pairs = float(n) * n
        val = args.steps * pairs * world / (dev_ms * 1e-3)
        flops = 2.0 * 2 * pairs * 128          # both directions, multiply + add
        tf = None
        try:
            tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops"))
        except Exception:
            pass
        line = {"metric": "descriptor pairs/sec (mutual nearest neighbours, 8648 x 8648)", "value": val, "unit": "pairs/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> s32",
                "data": "synthetic", "config": {"workload": "match: two 8648 x 128 u8 descriptor sets, cross-check"},
                "e2e": {"value": args.steps * pairs * world / t_e2e, "unit": "pairs/s",
                        "h2d_bytes_per_step": int(q.nbytes + t.nbytes), "d2h_bytes_per_step": int(cnt.value) * 12},
                "gpu_launches": int(launches), "matches": int(cnt.value),
                "roofline": {"bound": "tensor", "kernel": "k_match_nn (tcgen05.mma kind::i8 + fused argmin epilogue)",
                             "achieved": flops * args.steps / (dev_ms * 1e-3) / 1e12, "peak": tf, "unit": "TOP/s (peak: dense bf16 TFLOP/s)",
                             "frac": (flops * args.steps / (dev_ms * 1e-3) / 1e12 / tf) if tf else None, "traffic": None,
                             "note": "timed through the synchronous device-pointer call (includes the count read-back); "
                                     "the step is epilogue / launch bound, not MMA bound, at this size"}}
        if world == 1 and not args.no_cpu:
            from oracle import oracle as O
            t0 = time.perf_counter()
            O.match_cross_check(q[:2048], t)
            line["cpu_baseline"] = {"value": 2048.0 * n / (time.perf_counter() - t0), "unit": "pairs/s", "cores": os.cpu_count(),
                                    "kind": "port", "sample": "2048 query rows against all train rows (numpy int64 GEMM)"}
        print(json.dumps(line), flush=True)
    ex.close()
    dist.close()


def run_jpeg(args, rank, local_rank, world):
    """JPEG input: 1080p bitstreams in host memory -> nvJPEG decode + luma on the device ->
    the extraction path -> keypoints and descriptors in host memory.  The CPU leg beside it is what the reference's
    callers do first: a libjpeg-turbo decode (cv2.imdecode) on one core."""
    import cv2
    import sift_features_b200 as sf
    dist = Dist(rank, local_rank, world)
    w, h, B, G = WORKLOADS["1080p"]
    B = args.batch or B
    G = args.groups or G
    bird = np.load(os.path.join(ROOT, "tests", "golden", "bird_gray.npy"))
    tile = np.tile(bird, (h // bird.shape[0] + 1, w // bird.shape[1] + 1))[:h, :w]
    jpegs = []
    for i in range(B * G):
        img = np.roll(tile, (37 * (i + 1 + rank * 1000)) % w, 1)
        if args.jpeg_colour:
            img = np.stack([img, np.roll(img, 5, 0), np.roll(img, 9, 1)], -1)
        jpegs.append(cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes())
    ex = sf.Extractor(w, h, B, device=local_rank)
    for _ in range(args.warmup):
        offs, kp, _d = ex.sift_jpeg(jpegs)
    dist.barrier()
    l0 = ex.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        offs, kp, _d = ex.sift_jpeg(jpegs)
    t = dist.reduce(time.perf_counter() - t0, "max")
    launches = ex.launch_count - l0
    if rank == 0:
        n = len(jpegs)
        line = {"metric": "1080p JPEG images/sec (decode + luma + extraction, host bitstreams -> host results)",
                "value": args.steps * n * world / t, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8 -> f32", "data": "synthetic",
                "config": {"workload": f"jpeg: {n} x 1920x1080 {'colour 4:2:0' if args.jpeg_colour else 'gray'} JPEGs "
                                       f"(quality 90, tiled bird fixture), groups of {B}",
                           "jpeg_backend": ex.jpeg_backend, "keypoints_per_image": float(len(kp)) / n},
                "e2e": {"value": args.steps * n * world / t, "unit": "images/s",
                        "h2d_bytes_per_step": int(sum(len(j) for j in jpegs)),
                        "d2h_bytes_per_step": int(len(kp)) * (20 + 128)},
                "gpu_launches": int(launches)}
        if world == 1 and not args.no_cpu:
            t0 = time.perf_counter()
            for j in jpegs[:16]:
                cv2.imdecode(np.frombuffer(j, np.uint8), cv2.IMREAD_GRAYSCALE)
            line["cpu_baseline"] = {"value": 16 / (time.perf_counter() - t0), "unit": "images/s", "cores": 1, "kind": "port",
                                    "sample": "decode only (cv2.imdecode, libjpeg-turbo) of 16 of the bitstreams"}
        print(json.dumps(line), flush=True)
    ex.close()
    dist.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=list(WORKLOADS) + ["desc", "match", "jpeg"])
    ap.add_argument("--batch", type=int, default=0, help="images per group (context max_batch)")
    ap.add_argument("--groups", type=int, default=0, help="groups per step")
    ap.add_argument("--natural", action="store_true",
                    help="tiled tree.jpg instead of noise as the headline workload's input (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (no sub-workloads, no config 4)")
    ap.add_argument("--jpeg-colour", action="store_true", help="jpeg workload: three-component 4:2:0 streams")
    ap.add_argument("--no-profile-stages", dest="profile_stages", action="store_false",
                    help="do not repeat the timed steps with per-stage CUDA events")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        if args.workload in ("desc", "match", "jpeg"):
            args.workload = "1080p"
        return run_reference(args, rank, world)
    if args.workload == "desc":
        return run_desc(args, rank, local_rank, world)
    if args.workload == "match":
        return run_match(args, rank, local_rank, world)
    if args.workload == "jpeg":
        return run_jpeg(args, rank, local_rank, world)
    return run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
