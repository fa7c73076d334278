"""Time-bounded randomized parity stress: SA (and every 8th case LCP / BWT round trip / search) of random
structured texts against the oracle.  Usage: python tools/stress.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from stringsearch_b200 import divsufsort  # noqa: E402


def make(rng):
    n = int(rng.choice([rng.integers(1, 200), rng.integers(200, 5000), rng.integers(5000, 60000), rng.integers(60000, 400000)]))
    sigma = int(rng.choice([1, 2, 3, 4, 5, 16, 256]))
    kind = int(rng.integers(0, 8))
    r = lambda m: rng.integers(0, sigma, max(1, m), dtype=np.uint8)  # noqa: E731
    if kind == 0:
        t = r(n)
    elif kind == 1:
        t = np.tile(r(int(rng.integers(1, 2000))), n)[:n]
    elif kind == 2:
        t = np.tile(r(int(rng.integers(1, 2000))), n)[:n].copy()
        k = max(1, int(n * float(rng.choice([1e-4, 1e-3, 1e-2]))))
        t[rng.integers(0, n, k)] = r(k)
    elif kind == 3:
        t = np.repeat(r(n // 5 + 1), rng.integers(1, 12, n // 5 + 1))[:n]
    elif kind == 4:  # copies of one block
        b = r(max(1, n // int(rng.integers(2, 9))))
        t = np.tile(b, n // b.size + 1)[:n]
    elif kind == 5:  # long run inside random
        t = r(n)
        a, z = sorted(rng.integers(0, n, 2))
        t[a:z] = t[a] if z > a else 0
    elif kind == 6:  # two interleaved periods
        p, q = int(rng.integers(1, 50)), int(rng.integers(1, 50))
        t = ((np.tile(r(p), n)[:n].astype(np.int32) + np.tile(r(q), n)[:n]) % max(1, sigma)).astype(np.uint8)
    else:  # fibonacci-like substitution
        a, b = bytes([0]), bytes([0, 1 % max(1, sigma)])
        while len(b) < n:
            a, b = b, b + a
        t = np.frombuffer(b[:n], np.uint8).copy()
    return np.ascontiguousarray(t, dtype=np.uint8)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    port = oracle.port()
    ref = oracle.ref() if oracle.have_ref() else None
    t0, cases, fails = time.time(), 0, 0
    while time.time() - t0 < budget:
        t = make(rng)
        exp = (ref or port).sa_build(t)
        sa = divsufsort.sort(t, device=0)
        ok = (sa.sa == exp).all()
        if ok and cases % 8 == 0 and t.size > 1:
            ok = (divsufsort.lcp(t, exp, device=0) == port.lcp(t, exp)).all()
            u, pidx = divsufsort.bwt(t)
            ok = ok and (divsufsort.inverse_bwt(u, pidx) == t).all()
            pats = [t[o:o + m].tobytes() for o, m in zip(rng.integers(0, t.size, 50), rng.integers(0, 200, 50))]
            s, l = sa.longest_substring_match_batch(pats)
            es, el = port.lsm_batch(t, exp, pats)
            left, cnt = sa.search_all_batch(pats)
            eleft, ecnt = port.search_all_batch(t, exp, pats)
            ok = ok and (s == es).all() and (l == el).all() and (left == eleft).all() and (cnt == ecnt).all()
        if not ok:
            fails += 1
            np.save(os.path.join(ROOT, "gpurun_out", f"stress_fail_{seed}_{cases}.npy"), t)
            print(f"FAIL case {cases}: n={t.size}", flush=True)
        cases += 1
    print(f"stress: {cases} cases, {fails} failures, seed {seed}, {time.time() - t0:.0f} s")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
