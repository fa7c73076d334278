"""Small LCP / BWT / inverse-BWT / search round trips through the C ABI (no torch import: cheap under
compute-sanitizer).  Usage: python tools/small_extras.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from stringsearch_b200 import divsufsort, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
    port = oracle.port()
    rng = np.random.default_rng(1)
    fails = 0
    for name, t in (("acgt", synth.acgt(n, 1)), ("rep", synth.repetitive(n, 3, period=50, mutation_rate=2e-3)),
                    ("zeros", np.zeros(n // 4, np.uint8)), ("bin", rng.integers(0, 2, n, dtype=np.uint8))):
        sa, lcp = divsufsort.sort_with_lcp(t, device=0)
        exp_sa = port.sa_build(t)
        ok = (sa.sa == exp_sa).all() and (lcp == port.lcp(t, exp_sa)).all()
        u, pidx = divsufsort.bwt(t)
        ok = ok and (divsufsort.inverse_bwt(u, pidx) == t).all()
        pats = [t[o:o + m].tobytes() for o, m in zip(rng.integers(0, t.size - 1, 200), rng.integers(1, 300, 200))]
        s, l = sa.longest_substring_match_batch(pats)
        es, el = port.lsm_batch(t, exp_sa, pats)
        left, cnt = sa.search_all_batch(pats)
        eleft, ecnt = port.search_all_batch(t, exp_sa, pats)
        ok = ok and (s == es).all() and (l == el).all() and (left == eleft).all() and (cnt == ecnt).all()
        print(("ok   " if ok else "FAIL ") + name)
        fails += 0 if ok else 1
    print("FAILS:", fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
