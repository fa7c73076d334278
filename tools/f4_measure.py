"""SURVEY 8(f) rank 4 decision by measurement: what does sacapart's partitioning lose against one global
(un-partitioned) suffix array?  (crates/sacapart/src/lib.rs:5-25 "worse matches across boundaries".)
For every needle: (start, len) from PartitionedSuffixArray(P = 8) vs SuffixArray (one index over the whole text).
Reports the fraction of needles whose len differs, the mean / max length lost, and the fraction whose start
differs at equal len (a different occurrence of an equally long match: tie-break only).
Usage: python tools/f4_measure.py [acgt MiB=1024] [rep MiB=256]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringsearch_b200 import divsufsort, sacapart, synth  # noqa: E402


def compare(text, needles, P=8):
    full = divsufsort.sort(text, device=0)
    fs, fl = full.longest_substring_match_batch(needles)
    del full
    part = sacapart.PartitionedSuffixArray(text, P, devices=[0])
    ps, pl = part.longest_substring_match_batch(needles)
    part.close()
    q = fl.size
    lost = fl.astype(np.int64) - pl.astype(np.int64)
    assert (lost >= 0).all(), "a partition cannot find a longer match than the whole text holds"
    worse = lost > 0
    return {"needles": int(q), "partitions": P, "len_differs_fraction": float(worse.mean()), "len_differs": int(worse.sum()),
            "mean_len_lost_over_all": float(lost.mean()), "mean_len_lost_when_worse": float(lost[worse].mean()) if worse.any() else 0.0,
            "max_len_lost": int(lost.max()), "start_differs_at_equal_len_fraction": float(((fs != ps) & ~worse).mean()),
            "mean_len_full": float(fl.mean()), "mean_len_partitioned": float(pl.mean())}


def main():
    a_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    r_mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    out = {}
    t = synth.acgt(a_mib << 20, 5)
    needles = synth.patterns_from_text(t, 10_000_000, 32, 6)
    out[f"acgt_{a_mib}M, 10M x 32 B (BASELINE configs[4] needles: half cut from the text, half random)"] = compare(t, needles)
    # needles that straddle the partition boundaries on purpose: the worst case for partitioning
    n = t.size
    ps = n // 8 + 1
    rng = np.random.default_rng(8)
    k = 70_000
    starts = np.array([(1 + j % 7) * ps - 1 - (j // 7) % 31 for j in range(k)], dtype=np.int64)
    flat = t[starts[:, None] + np.arange(32)[None, :]].reshape(-1)
    off = np.arange(k + 1, dtype=np.uint64) * np.uint64(32)
    out[f"acgt_{a_mib}M, {k} needles cut across the 7 partition boundaries (1..31 bytes in front of the boundary)"] = compare(t, (flat, off))
    del t
    t = synth.repetitive(r_mib << 20, 3)
    n, m, q = t.size, 2048, 200_000
    o = rng.integers(0, n - m, q)
    pats = t[o[:, None] + np.arange(m)[None, :]]
    odd = np.arange(1, q, 2)
    pats[odd, rng.integers(0, m, q)[odd]] ^= 1
    out[f"rep_{r_mib}M (period 1000, 1e-3 mutations), 200k x 2 KiB needles cut from the text, every second one with one byte changed"] = compare(
        t, (np.ascontiguousarray(pats.reshape(-1)), np.arange(q + 1, dtype=np.uint64) * np.uint64(m)))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
