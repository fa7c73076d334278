"""BASELINE configs[4] shape in isolation: Q patterns of m bytes against the SA of an ACGT text, kernels only
(everything resident).  Used for A/B runs (GSA_NO_ACCEL=1, GSA_ACCEL_BITS=..) and as the ncu target for
k_lsm / k_search_all.   Usage: python tools/search_bench.py [text MiB=1024] [Q=10000000] [m=32] [check=0|1]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringsearch_b200 import _native as N  # noqa: E402
from stringsearch_b200 import synth  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    check = len(sys.argv) > 4 and sys.argv[4] == "1"
    n = mib << 20
    dev = torch.device("cuda", 0)
    t = synth.acgt(n, 5)
    h = C.c_void_p()
    assert N.lib.gsa_index_create(t.ctypes.data, n, 0, C.byref(h), None) == 0, N.last_error()
    flat, off = synth.patterns_from_text(t, Q, m, 6)
    d_p, d_o = torch.from_numpy(flat).to(dev), torch.from_numpy(off.astype(np.int64)).to(dev)
    d_s = torch.empty(Q, dtype=torch.int64, device=dev)
    d_l = torch.empty(Q, dtype=torch.int32, device=dev)
    d_left = torch.empty(Q, dtype=torch.int32, device=dev)
    d_cnt = torch.empty(Q, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    res = {"text": f"acgt_{mib}M", "patterns": Q, "pattern_len": m,
           "accel": os.environ.get("GSA_NO_ACCEL") is None, "accel_bits": os.environ.get("GSA_ACCEL_BITS", "default")}
    t0 = time.perf_counter()
    assert N.lib.gsa_lsm_device(h, d_p.data_ptr(), d_o.data_ptr(), 1, m, 0, 0, d_s.data_ptr(), d_l.data_ptr(), st) == 0
    torch.cuda.synchronize()
    res["first_query_ms (builds the prefix-bucket table)"] = (time.perf_counter() - t0) * 1e3
    for name, fn in (("longest_substring_match", lambda: N.lib.gsa_lsm_device(h, d_p.data_ptr(), d_o.data_ptr(), Q, m, 0, 0, d_s.data_ptr(), d_l.data_ptr(), st)),
                     ("search_all", lambda: N.lib.gsa_search_all_device(h, d_p.data_ptr(), d_o.data_ptr(), Q, m, d_left.data_ptr(), d_cnt.data_ptr(), st))):
        assert fn() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            assert fn() == 0
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[name] = {"queries_per_s": Q / (ms / 1e3), "ms": ms}
    res["checksum"] = [int(d_s.sum().item()), int(d_l.sum().item()), int(d_left.sum().item()), int(d_cnt.sum().item())]
    if check:
        from oracle import oracle

        port = oracle.port()
        sa = np.empty(n, dtype=np.int32)
        assert N.lib.gsa_index_sa(h, sa.ctypes.data) == 0
        cs, cl = port.lsm_batch(t, sa, (flat, off), threads=0)
        c_left, c_cnt = port.search_all_batch(t, sa, (flat, off), threads=0)
        res["equal_to_oracle"] = bool((d_s.cpu().numpy().astype(np.uint64) == cs).all() and (d_l.cpu().numpy().astype(np.uint32) == cl).all()
                                      and (d_left.cpu().numpy() == c_left).all() and (d_cnt.cpu().numpy() == c_cnt).all())
    N.lib.gsa_index_destroy(h)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
