"""PCIe copy rates with pinned host memory: one cudaMemcpyAsync against the same bytes split over 2 / 4 streams
(do several copy engines in one direction add up?), and both directions at once.  python tools/pcie_probe.py [GiB=4]"""
import json
import sys

import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30))
dev = torch.device("cuda", 0)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 4, dtype=torch.uint8, device=dev)
res = {}


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    step = n // parts

    def d2h():
        cur = torch.cuda.current_stream(dev)
        for i, s in enumerate(streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                h[i * step:(i + 1) * step].copy_(d[i * step:(i + 1) * step], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)

    def h2d():
        cur = torch.cuda.current_stream(dev)
        for i, s in enumerate(streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)

    ms = timed(d2h); res[f"d2h_{parts}"] = {"ms": ms, "GBps": n / ms / 1e6}
    ms = timed(h2d); res[f"h2d_{parts}"] = {"ms": ms, "GBps": n / ms / 1e6}

s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def both():
    cur = torch.cuda.current_stream(dev)
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


ms = timed(both); res["d2h_full_plus_h2d_quarter"] = {"ms": ms}
print(json.dumps(res))
