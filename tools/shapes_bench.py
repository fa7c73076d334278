"""Device-timed SA build of differently shaped 256 MiB inputs (catching pathological cases).
Usage: python tools/shapes_bench.py [MiB]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringsearch_b200 import _native as N  # noqa: E402
from stringsearch_b200 import synth  # noqa: E402


def fib(n):
    a, b = b"a", b"ab"
    while len(b) < n:
        a, b = b, b + a
    return np.frombuffer(b[:n], np.uint8).copy()


def zipf_words(n, rng, vocab=50_000):
    """Text-like: words of 2-10 letters drawn from a Zipf-distributed vocabulary, single spaces."""
    lens = rng.integers(2, 11, vocab)
    words = [bytes(rng.integers(97, 123, int(l), dtype=np.uint8)) + b" " for l in lens]
    ranks = np.minimum(rng.zipf(1.2, n // 4), vocab) - 1
    out = b"".join(words[r] for r in ranks)
    while len(out) < n:
        out += out
    return np.frombuffer(out[:n], np.uint8).copy()


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = mib << 20
    rng = np.random.default_rng(0)
    shapes = {
        "zeros": lambda: np.zeros(n, np.uint8),
        "ab": lambda: np.tile(np.frombuffer(b"ab", np.uint8), n // 2),
        "fibonacci": lambda: fib(n),
        "rep_period7": lambda: synth.repetitive(n, 5, period=7, mutation_rate=1e-4),
        "rep_period1000": lambda: synth.repetitive(n, 3),
        "rep_period100k_rare": lambda: synth.repetitive(n, 6, period=100_000, mutation_rate=1e-5),
        "square": lambda: np.tile(rng.integers(0, 256, n // 4, dtype=np.uint8), 4),
        "binary_random": lambda: rng.integers(0, 2, n, dtype=np.uint8),
        "zipf_words": lambda: zipf_words(n, rng),
        "dna_with_repeats": lambda: np.concatenate([synth.acgt(n // 2, 7), synth.acgt(n // 2, 7)[::-1].copy()[: n // 4], synth.acgt(n // 4, 8)]),
        "text_like": lambda: (rng.integers(0, 27, n, dtype=np.uint8) + 97).astype(np.uint8),
        "run_in_random": lambda: np.concatenate([rng.integers(0, 256, n // 4, dtype=np.uint8), np.full(n // 2, 65, np.uint8),
                                                 rng.integers(0, 256, n // 4, dtype=np.uint8)]),
    }
    dev = torch.device("cuda", 0)
    ws_bytes = N.lib.gsa_build_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    for name, mk in shapes.items():
        t = mk()
        d_t = torch.from_numpy(t).to(dev)
        stats = N.BuildStats()
        for _ in range(2):
            rc = N.lib.gsa_build_device(d_t.data_ptr(), d_sa.data_ptr(), n, ws.data_ptr(), ws_bytes, st, C.byref(stats))
            assert rc == 0, N.last_error()
        bad = C.c_int64(-1)
        rc = N.lib.gsa_sufcheck_device(d_t.data_ptr(), d_sa.data_ptr(), n, st, C.byref(bad))
        rl = stats.rounds_list()
        print(f"{name:22s} {stats.ms_total:9.2f} ms  {n / stats.ms_total / 1e3:9.1f} MB/s  rounds={stats.rounds:3d} "
              f"sorted={sum(r['sorted'] for r in rl[1:]) / n:6.2f}n walked={sum(r['live'] for r in rl[1:]) / n:6.2f}n "
              f"bag={sum(r['bag'] for r in rl) / n:5.2f}n sufcheck={'ok' if rc == 0 else 'FAIL'}", flush=True)


if __name__ == "__main__":
    main()
