"""LCP array, BWT and inverse BWT of a resident 256 MiB repetitive text, device-timed (and the ncu target for
lcp.cu / bwt.cu).  Usage: python tools/extras_bench.py [MiB=256]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringsearch_b200 import _native as N  # noqa: E402
from stringsearch_b200 import synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = mib << 20
    dev = torch.device("cuda", 0)
    t = synth.repetitive(n, 3)
    d_t = torch.from_numpy(t).to(dev)
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    assert N.lib.gsa_build_device(d_t.data_ptr(), d_sa.data_ptr(), n, None, 0, st, None) == 0, N.last_error()
    res = {"text": f"rep_{mib}M"}
    d_lcp = torch.empty(n, dtype=torch.int32, device=dev)
    ws_bytes = N.lib.gsa_lcp_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ms = timed(lambda: N.lib.gsa_lcp_device(d_t.data_ptr(), d_sa.data_ptr(), d_lcp.data_ptr(), n, ws.data_ptr(), ws_bytes, st))
    res["lcp"] = {"ms": ms, "MB_per_s": n / ms / 1e3, "mean_lcp": float(d_lcp.double().mean().item()), "max_lcp": int(d_lcp.max().item())}
    del ws, d_lcp
    d_u = torch.empty(n, dtype=torch.uint8, device=dev)
    pidx = C.c_int32(0)
    ms = timed(lambda: N.lib.gsa_bwt_device(d_t.data_ptr(), d_sa.data_ptr(), n, d_u.data_ptr(), C.byref(pidx), st))
    res["bwt"] = {"ms": ms, "MB_per_s": n / ms / 1e3, "primary_index": pidx.value}
    d_back = torch.empty(n, dtype=torch.uint8, device=dev)
    ws_bytes = N.lib.gsa_inverse_bwt_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ms = timed(lambda: N.lib.gsa_inverse_bwt_device(d_u.data_ptr(), d_back.data_ptr(), n, pidx.value, ws.data_ptr(), ws_bytes, st))
    res["inverse_bwt"] = {"ms": ms, "MB_per_s": n / ms / 1e3, "round_trip_ok": bool(torch.equal(d_back, d_t))}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
