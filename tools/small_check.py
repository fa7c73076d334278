"""Small end-to-end checks through the C ABI with verbose diagnostics (no torch import, so it
is cheap to run under compute-sanitizer).  Usage: python tools/small_check.py [max_n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from stringsearch_b200 import _native as N  # noqa: E402
from stringsearch_b200 import divsufsort, synth  # noqa: E402


def main():
    max_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    port = oracle.port()
    rng = np.random.default_rng(0)
    cases = [("banana", np.frombuffer(b"banana", np.uint8)), ("mississippi", np.frombuffer(b"mississippi", np.uint8)),
             ("trap", np.frombuffer(b"a\0\0\0\0\0\0\0\0a\0", np.uint8)), ("aaa", np.frombuffer(b"aaa", np.uint8))]
    for n in (3, 8, 9, 100, 4096, 4097, 10_000, 100_000, 1 << 20, 4 << 20, 16 << 20):
        if n > max_n:
            break
        cases.append((f"rand256_{n}", rng.integers(0, 256, n, dtype=np.uint8)))
        cases.append((f"acgt_{n}", synth.acgt(n, 1)))
        cases.append((f"bin_{n}", rng.integers(0, 2, n, dtype=np.uint8)))
        cases.append((f"rep_{n}", synth.repetitive(n, 3, period=max(2, min(1000, n // 8)), mutation_rate=2e-3)))
        cases.append((f"zeros_{n}", np.zeros(n, np.uint8)))
        if n >= 100_000:  # huge groups: long period, few mutations; long runs of one byte inside random text
            cases.append((f"rep7_{n}", synth.repetitive(n, 5, period=7, mutation_rate=1e-4)))
            cases.append((f"rep1000_{n}", synth.repetitive(n, 6, period=1000, mutation_rate=1e-3)))
            x = rng.integers(0, 256, n, dtype=np.uint8)
            x[n // 3:n // 3 + n // 4] = 65
            cases.append((f"run_in_random_{n}", x))
            cases.append((f"abab_{n}", np.tile(np.frombuffer(b"ab", np.uint8), n // 2)))
    fails = 0
    for name, t in cases:
        stats = N.BuildStats()
        try:
            got = divsufsort.sort(t, device=0, stats=stats).sa
        except AssertionError as e:
            print(f"FAIL {name}: {e}")
            fails += 1
            continue
        exp = port.sa_build(t)
        rounds = [(r["depth"], r["live"], r["sorted"], r["groups"], r["passes"]) for r in stats.rounds_list()]
        if (got == exp).all():
            print(f"ok   {name}: sigma={stats.sigma} b={stats.bits_per_symbol} k={stats.symbols_per_key} "
                  f"ms={stats.ms_total:.3f} rounds={rounds}")
        else:
            fails += 1
            bad = np.flatnonzero(got != exp)
            print(f"FAIL {name}: {bad.size}/{exp.size} slots differ, first at {bad[0]}: gpu {got[bad[0]]} exp {exp[bad[0]]}; "
                  f"gpu[:12]={got[:12].tolist()} exp[:12]={exp[:12].tolist()} rounds={rounds}")
            srt = np.sort(got)
            if not (srt == np.arange(exp.size)).all():
                print("     not a permutation; min/max", got.min(), got.max(), "dups", exp.size - np.unique(got).size)
    print("FAILS:", fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
