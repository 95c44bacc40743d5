"""N>1 host logic on CPU: world_size-2 gloo run of bench.py's distributed plumbing (barrier, max/sum over ranks)
and of the image sharding / host-side gather that sb200_extract_batch_multi performs, with the oracle standing in
for the per-device extractor (the product path itself needs a GPU and is covered by the -m gpu tests)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    import bench
    import sift_features_b200 as sf
    from oracle import oracle as O
    rank, local_rank, world = bench.dist_env()
    d = bench.Dist(rank, local_rank, world, cuda=False)      # gloo
    n, w, h = 5, 96, 64
    imgs = bench.synth_images(n, w, h, 77)                  # every rank generates the same batch
    mine = sf.shard_ranges(n, world)[rank]                  # contiguous shard of ceil(n/world) images
    counts = [len(O.sift(imgs[i])[0]) for i in mine]
    d.barrier()
    total = d.reduce(sum(counts), "sum")
    slowest = d.reduce(1.0 + rank, "max")
    out = dict(rank=rank, shard=list(mine), counts=counts, total=total, slowest=slowest)
    open(os.path.join({out!r}, f"rank{{rank}}.json"), "w").write(json.dumps(out))
    d.close()
""")


def test_two_rank_gloo_shard_and_gather(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    res = [json.load(open(tmp_path / f"rank{k}.json")) for k in range(2)]
    assert res[0]["shard"] == [0, 1, 2] and res[1]["shard"] == [3, 4]
    import bench
    imgs = bench.synth_images(5, 96, 64, 77)
    expect = [len(oracle.sift(im)[0]) for im in imgs]
    assert res[0]["counts"] + res[1]["counts"] == expect                 # gather in image order == single process
    assert res[0]["total"] == res[1]["total"] == float(sum(expect))      # all-reduce(sum) agrees on every rank
    assert res[0]["slowest"] == res[1]["slowest"] == 2.0                 # max over ranks, as the timing contract needs


def test_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under torchrun: rank 0 alone runs and prints; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--workload", "vga"], env=env, capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
