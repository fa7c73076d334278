"""CPU model of the prefix-bucket search of stringsearch_b200/csrc/search.cu (k_lsm / k_search_all /
accel_build_device), statement by statement, so that its logic can be checked against the oracle
without a GPU (tests/test_host.py::test_accel_model_matches_oracle).  Test infrastructure only."""
from __future__ import annotations

import numpy as np


def bits_for(v: int) -> int:
    b = 1
    while b < 64 and (v >> b) != 0:
        b += 1
    return b


class Accel:
    def __init__(self, text: bytes, sa, bits: int):
        n = len(text)
        self.text, self.sa, self.n = text, [int(x) for x in sa], n
        self.present = [False] * 256
        for v in text:
            self.present[v] = True
        self.code, sigma = [0] * 256, 0
        for v in range(256):
            self.code[v] = min(sigma, 255)
            if self.present[v]:
                sigma += 1
        self.b = bits_for(sigma - 1 if sigma > 1 else 1)
        self.k = bits // self.b
        if self.k == 0:
            self.T = None
            return
        ln = (1 << (self.k * self.b)) + 1
        F = [0xFFFFFFFF] * ln
        prev = None
        for j in range(n):  # k_accel_sample / k_accel_mark: first SA index of every key
            kj = self.key(self.sa[j])
            if kj != prev:
                assert prev is None or kj > prev, "keys must be non-decreasing along the SA"
                F[kj] = j
                prev = kj
        carry = n  # suffix minimum, T[2^(kb)] = n
        for c in range(ln - 1, -1, -1):
            carry = min(carry, F[c])
            F[c] = carry
        self.T = F

    def key(self, s: int) -> int:
        key = 0
        for j in range(self.k):
            c = self.code[self.text[s + j]] if s + j < self.n else 0
            key = (key << self.b) | c
        return key

    def bounds(self, pat: bytes):
        n, m = self.n, len(pat)
        if self.k == 0:
            return 0, n, 0, n
        key, T, b, k = 0, self.T, self.b, self.k
        for j in range(k):
            if j >= m:
                sh = b * (k - j)
                klo, khi = key << sh, ((key + 1) << sh) - 1
                return T[klo], T[klo + 1], T[khi], T[khi + 1]
            v = pat[j]
            c = self.code[v]
            if not self.present[v]:
                kk = ((key << b) + c) << (b * (k - j - 1))
                hi_ = T[kk] if (c >> b) else T[kk + 1]
                return T[kk], hi_, T[kk], hi_
            key = (key << b) | c
        return T[key], T[key + 1], T[key], T[key + 1]

    def cmp(self, s: int, pat: bytes):
        """group_compare: (cpl, gt, lt) of pattern against the suffix at s."""
        suf = self.text[s:]
        lim = min(len(suf), len(pat))
        for i in range(lim):
            if suf[i] != pat[i]:
                return i, pat[i] > suf[i], pat[i] < suf[i]
        return lim, len(pat) > lim, False

    def lsm(self, pat: bytes):
        n, sa = self.n, self.sa
        lo, hi, _, _ = self.bounds(pat)
        lm = rm = 0
        lk = rk = False
        while lo < hi:
            mid = (lo + hi) >> 1
            cpl, gt, _ = self.cmp(sa[mid], pat)
            if gt:
                lo, lm, lk = mid + 1, cpl, True
            else:
                hi, rm, rk = mid, cpl, True
        ip = lo
        two = n >= 2
        a0 = 0 if ip == 0 else ip - 1
        if two and a0 > n - 2:
            a0 = n - 2
        x_known, y_known = lk and a0 + 1 == ip, two and rk and a0 + 1 == ip
        x_known2, y_known2 = rk and a0 == ip, two and lk and a0 + 2 == ip
        start = sa[a0]
        ln = lm if x_known else (rm if x_known2 else self.cmp(start, pat)[0])
        if two:
            s1 = sa[a0 + 1]
            y = rm if y_known else (lm if y_known2 else self.cmp(s1, pat)[0])
            if not ln > y:
                start, ln = s1, y
        return start, ln

    def search_all(self, pat: bytes):
        n, sa, m = self.n, self.sa, len(pat)
        if m == 0:
            return 0, n
        lo, hi, up_lo, up_hi = self.bounds(pat)
        rm, rk = 0, False
        while lo < hi:
            mid = (lo + hi) >> 1
            cpl, gt, _ = self.cmp(sa[mid], pat)
            if gt:
                lo = mid + 1
            else:
                hi, rm, rk = mid, cpl, True
        left = lo
        at_left = rm
        if not rk and left < n:
            at_left = self.cmp(sa[left], pat)[0]
        hit = left < n and at_left == m
        if not hit:
            return left, 0
        ulo, uhi = left + 1, max(up_hi, left + 1)
        ulo = max(ulo, up_lo)
        step, gallop = 1, True
        while ulo < uhi:
            probe = ulo + step - 1 if gallop else (ulo + uhi) >> 1
            probe = min(probe, uhi - 1)
            _, _, lt = self.cmp(sa[probe], pat)
            if not lt:
                ulo, step = probe + 1, step << 1
            else:
                uhi, gallop = probe, False
        return left, ulo - left
