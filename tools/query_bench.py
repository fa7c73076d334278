"""Long-needle queries on a repetitive text: effect of carrying the match lengths of the two
window ends (sa_search's lmatch / rmatch, utils.c:275-286) through the binary search.
Run twice: as is, and with GSA_NO_MATCH_CARRY=1.  Usage: python tools/query_bench.py [text MiB] [needle bytes]"""
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


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    n, Q = mib << 20, 200_000
    dev = torch.device("cuda", 0)
    t = synth.repetitive(n, 3)
    h = C.c_void_p()
    assert N.lib.gsa_index_create(t.ctypes.data, n, 0, C.byref(h), None) == 0, N.last_error()
    rng = np.random.default_rng(7)
    o = rng.integers(0, n - m, Q)
    pats = t[o[:, None] + np.arange(m)[None, :]]
    mut = rng.integers(0, m, Q)
    odd = np.arange(1, Q, 2)
    pats[odd, mut[odd]] ^= 1  # every second needle differs from the text in one byte
    flat = np.ascontiguousarray(pats.reshape(-1))
    off = np.arange(Q + 1, dtype=np.int64) * m
    d_p, d_o = torch.from_numpy(flat).to(dev), torch.from_numpy(off).to(dev)
    d_s = torch.empty(Q, dtype=torch.int64, device=dev)
    d_l = torch.empty(Q, dtype=torch.int32, device=dev)
    d_left = torch.empty(Q, dtype=torch.int32, device=dev)
    d_cnt = torch.empty(Q, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    res = {"text": f"rep_{mib}M", "patterns": Q, "pattern_len": m, "carry": os.environ.get("GSA_NO_MATCH_CARRY") is None}
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
    N.lib.gsa_index_destroy(h)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
