"""numpy/python model of the GPU construction scheme (TEST INFRASTRUCTURE).

Mirrors stringsearch_b200/csrc/sa_build.cu step by step on the CPU so the *design* can be
checked against the oracle without a GPU:

* alphabet compaction, bit-packed round-0 keys, short-suffix ordering;
* rank[i] = "label" of i's group: any SA slot inside the group's slot range, +1 (0 = past the
  end).  A group keeps its label across rounds while the label stays inside its (shrinking)
  range, so most suffixes of a big group never need their rank rewritten; a new label is the
  middle of the new range;
* live suffixes are walked in (any) text-order list, key = (label(i) << 31) | label(i + h);
* sort, flag, head/tail scans, singleton finalisation (dead bit), slot compaction.

Not used by the product; far too slow for real inputs.
"""
from __future__ import annotations

import numpy as np

DEAD = 1 << 31


def bits_for(v: int) -> int:
    b = 1
    while b < 64 and (v >> b) != 0:
        b += 1
    return b


def build_sa_model(text: bytes, log=None, shuffle_seed=None, key_symbols=None, sparse=False) -> np.ndarray:
    t = np.frombuffer(bytes(text), dtype=np.uint8)
    n = t.size
    if n == 0:
        return np.zeros(0, np.int32)
    present = np.zeros(256, bool)
    present[t] = True
    code = np.cumsum(present) - present
    sigma = int(present.sum())
    b = bits_for(sigma - 1 if sigma > 1 else 1)
    k = min(64 // b, (bits_for(n) + 10 + b - 1) // b)  # adaptive round-0 depth, as sa_build.cu
    if key_symbols is not None:
        k = max(1, min(64 // b, key_symbols))
    ns = min(k - 1, n)
    codes = code[t].astype(object)
    keys = []
    for i in range(n):
        x = 0
        for s in range(k):
            x = (x << b) | (int(codes[i + s]) if i + s < n else 0)
        keys.append(x)
    init = [n - 1 - j if j < ns else j - ns for j in range(n)]
    order = sorted(range(n), key=lambda j: keys[init[j]])  # stable
    sufx = [init[j] for j in order]
    skey = [keys[i] for i in sufx]
    short_from = n - ns
    SA = np.full(n, -1, np.int64)
    rank = [0] * n
    rng = np.random.default_rng(shuffle_seed) if shuffle_seed is not None else None

    sa0 = list(sufx)  # complete round-0 order (sparse mode keeps it for lazy labels)

    def lazy_label(tpos):
        """sa_build.cu lazy_label(): slot + 1 of a suffix that was unique after round 0."""
        kt = keys[tpos]
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) // 2
            if keys[sa0[mid]] < kt:
                lo = mid + 1
            else:
                hi = mid
        extra = 0
        if kt & ((1 << b) - 1) == 0:
            first_short = n - ns
            frm = tpos + 1 if tpos >= first_short else first_short
            extra = sum(1 for j in range(frm, n) if keys[j] == kt)
        return lo + extra + 1

    def rebuild(skey, sufx, pos, round0):
        L = len(sufx)
        flag = [False] * (L + 1)
        flag[L] = True
        for l in range(L):
            f = l == 0 or skey[l] != skey[l - 1]
            if round0:
                f = f or sufx[l] >= short_from or (l > 0 and sufx[l - 1] >= short_from)
            flag[l] = f
        head = [0] * L
        tail = [0] * L
        cur = 0
        for l in range(L):
            if flag[l]:
                cur = pos[l]
            head[l] = cur
        for l in range(L - 1, -1, -1):
            if flag[l + 1]:
                cur = pos[l]
            tail[l] = cur
        out_pos = []
        writes = 0
        for l in range(L):
            s, e = head[l], tail[l]
            old = 0 if round0 else (skey[l] >> 31)
            keep = s + 1 <= old <= e + 1
            if s == e:
                assert SA[s] == -1
                SA[s] = sufx[l]
                if not (round0 and sparse):
                    rank[sufx[l]] = DEAD | (s + 1)
                writes += 1
            else:
                if not keep:
                    rank[sufx[l]] = s + (e - s) // 2 + 1
                    writes += 1
                out_pos.append(pos[l])
        return out_pos, writes

    pos, _ = rebuild(skey, sufx, list(range(n)), True)
    lst = list(range(n))  # candidates (identity list in round 1)
    h = k
    rounds = 1
    while pos:
        live = [i for i in lst if rank[i] != 0 and not (rank[i] & DEAD)]
        if rng is not None:
            rng.shuffle(live)  # the order of the sort input is irrelevant
        assert len(live) == len(pos)
        key = []
        for i in live:
            r2 = (rank[i + h] & (DEAD - 1)) if i + h < n else 0
            if sparse and i + h < n and r2 == 0:
                r2 = lazy_label(i + h)
                assert SA[r2 - 1] == i + h  # it really is the final slot of that suffix
            key.append(((rank[i] & (DEAD - 1)) << 31) | r2)
        order = sorted(range(len(live)), key=lambda l: key[l])
        skey = [key[l] for l in order]
        sufx = [live[l] for l in order]
        pos, writes = rebuild(skey, sufx, pos, False)
        if log is not None:
            log.append((h, len(live), writes))
        lst = live
        h *= 2
        rounds += 1
        assert rounds < 70
    assert (SA >= 0).all()
    return SA.astype(np.int32)
