"""numpy model of the GPU construction scheme (TEST INFRASTRUCTURE).

Mirrors stringsearch_b200/csrc/sa_build.cu step by step on the CPU -- alphabet
compaction, bit-packed round-0 keys with the short-suffix ordering, (group ordinal,
rank[i+h]) keys, flag/scan rank rebuild, singleton finalisation, live-set compaction --
so the *design* can be checked against the oracle without a GPU.  It is not used by the
product and is far too slow for real inputs.
"""
from __future__ import annotations

import numpy as np


def bits_for(v: int) -> int:
    b = 1
    while b < 64 and (v >> b) != 0:
        b += 1
    return b


def build_sa_model(text: bytes, log=None) -> np.ndarray:
    t = np.frombuffer(bytes(text), dtype=np.uint8)
    n = t.size
    if n == 0:
        return np.zeros(0, np.int32)
    present = np.zeros(256, bool)
    present[t] = True
    code = np.cumsum(present) - present  # exclusive count of present bytes below c
    sigma = int(present.sum())
    b = bits_for(sigma - 1 if sigma > 1 else 1)
    k = 64 // b
    key_bits = k * b
    ns = min(k - 1, n)
    codes = code[t].astype(object)
    # key(i) = next k symbols, zero padded (python ints: exact 64-bit arithmetic)
    keys = []
    for i in range(n):
        x = 0
        for s in range(k):
            x = (x << b) | (int(codes[i + s]) if i + s < n else 0)
        keys.append(x)
    init = [n - 1 - j if j < ns else j - ns for j in range(n)]
    order = sorted(range(n), key=lambda j: keys[init[j]])  # python sort is stable
    sufx = np.array([init[j] for j in order], dtype=np.int64)
    skey = [keys[i] for i in sufx]
    short_from = n - ns
    SA = np.full(n, -1, np.int64)
    rank = np.zeros(n, np.int64)

    def rebuild(skey, sufx, pos, round0):
        L = len(sufx)
        flag = np.zeros(L + 1, bool)
        flag[L] = True
        for l in range(L):
            f = l == 0 or skey[l] != skey[l - 1]
            if round0:
                f = f or sufx[l] >= short_from or (l > 0 and sufx[l - 1] >= short_from)
            flag[l] = f
        head = 0
        g = 0
        out_pos, out_sufx, out_ord = [], [], []
        for l in range(L):
            if flag[l]:
                head = pos[l] + 1
            rank[sufx[l]] = head
            if flag[l] and flag[l + 1]:
                assert SA[pos[l]] == -1
                SA[pos[l]] = sufx[l]
            else:
                if flag[l]:
                    g += 1
                out_pos.append(pos[l]); out_sufx.append(sufx[l]); out_ord.append(g - 1)
        return np.array(out_pos, np.int64), np.array(out_sufx, np.int64), np.array(out_ord, np.int64), g

    pos, sufx, ordv, G = rebuild(skey, sufx, np.arange(n), True)
    h = k
    rank_bits = bits_for(n)
    rounds = 1
    while len(sufx):
        L = len(sufx)
        key = []
        for l in range(L):
            tpos = int(sufx[l]) + h
            r2 = int(rank[tpos]) if tpos < n else 0
            key.append((int(ordv[l]) << rank_bits) | r2)
        kb = bits_for(G - 1 if G > 0 else 0) + rank_bits
        assert all(x < (1 << kb) for x in key) and kb <= 64
        order = sorted(range(L), key=lambda l: key[l])
        skey = [key[l] for l in order]
        sufx = sufx[order]
        if log is not None:
            log.append((h, L, G, kb))
        pos, sufx, ordv, G = rebuild(skey, sufx, pos, False)
        h *= 2
        rounds += 1
        assert rounds < 70
    assert (SA >= 0).all()
    return SA.astype(np.int32)
