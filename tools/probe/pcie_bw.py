"""Host <-> device copy bandwidth per device, alone and with all devices copying at once, from ONE pinned host buffer
(the situation of config 4 through sb200_extract_batch_multi_parts).  python tools/probe/pcie_bw.py"""
import threading, time
import torch

n = torch.cuda.device_count()
MB = 256
host = torch.empty(n, MB << 20, dtype=torch.uint8).pin_memory()
host.fill_(1)
dev = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
streams = [torch.cuda.Stream(device=d) for d in range(n)]


def copy(d, h2d, reps, out):
    torch.cuda.set_device(d)
    with torch.cuda.stream(streams[d]):
        for _ in range(2):
            (dev[d].copy_(host[d], non_blocking=True) if h2d else host[d].copy_(dev[d], non_blocking=True))
        streams[d].synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            (dev[d].copy_(host[d], non_blocking=True) if h2d else host[d].copy_(dev[d], non_blocking=True))
        streams[d].synchronize()
        out[d] = reps * MB / 1024 / (time.perf_counter() - t0)


for h2d in (True, False):
    name = "H2D" if h2d else "D2H"
    alone = {}
    for d in range(n):
        copy(d, h2d, 8, alone)
    print(name, "alone      GB/s:", [round(alone[d], 1) for d in range(n)], flush=True)
    both = {}
    th = [threading.Thread(target=copy, args=(d, h2d, 8, both)) for d in range(n)]
    [t.start() for t in th]; [t.join() for t in th]
    print(name, "concurrent GB/s:", [round(both[d], 1) for d in range(n)], "sum", round(sum(both.values()), 1), flush=True)
# both directions on all devices at once
res = {}
def bidir(d):
    a, b = {}, {}
    t1 = threading.Thread(target=copy, args=(d, True, 8, a)); t1.start()
    # D2H on a second stream
    torch.cuda.set_device(d)
    s2 = torch.cuda.Stream(device=d)
    tmp = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
    src = torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{d}")
    with torch.cuda.stream(s2):
        t0 = time.perf_counter()
        for _ in range(8):
            tmp.copy_(src, non_blocking=True)
        s2.synchronize()
        b[d] = 8 * MB / 1024 / (time.perf_counter() - t0)
    t1.join()
    res[d] = (round(a[d], 1), round(b[d], 1))
th = [threading.Thread(target=bidir, args=(d,)) for d in range(n)]
[t.start() for t in th]; [t.join() for t in th]
print("bidirectional, all devices (H2D, D2H) GB/s:", [res[d] for d in range(n)], flush=True)
