"""sacabase on the GPU: SuffixArray, StringIndex, LongestCommonSubstring.

Mirrors crates/sacabase/src/lib.rs.  The search runs in batched CUDA kernels on a
device-resident copy of (text, sa); a single-needle call is a batch of one.
``contains`` / ``search_all`` follow libdivsufsort's sa_search
(crates/cdivsufsort/c-sources/utils.c:258-325), the in-tree semantics behind those names.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N


@dataclass
class LongestCommonSubstring:
    """lib.rs:4-21"""
    text: np.ndarray
    start: int
    len: int

    def as_bytes(self) -> bytes:
        return bytes(self.text[self.start:self.start + self.len])

    def __repr__(self) -> str:  # fmt::Debug, lib.rs:10-14
        return f"T[{self.start}..{self.start + self.len}]"


class NotSorted(Exception):
    """lib.rs:102-123"""

    def __init__(self, i: int, j: int):
        self.i, self.j = i, j
        super().__init__(f"invariant doesn't hold: suf(SA({i})) < suf(SA({j}))")


class StringIndex:
    """trait StringIndex (lib.rs:160-163)"""

    def longest_substring_match(self, needle) -> LongestCommonSubstring:  # pragma: no cover
        raise NotImplementedError


class _DeviceIndex:
    """Owner of a gsa_index handle (text + SA resident in HBM)."""

    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                N.lib.gsa_index_destroy(h)
            except Exception:
                pass


class SuffixArray(StringIndex):
    """sacabase::SuffixArray<'a, i32> (lib.rs:152-197)."""

    def __init__(self, text, sa: np.ndarray, device: int | None = None):
        """SuffixArray::new(text, sa) (lib.rs:170-172): takes ownership of `sa`.

        `device`: where the resident copy used by the queries and by verify() lives (None = the
        current CUDA device at the time of the first query).  The copy is made lazily; a suffix
        array whose length differs from the text's, or with an entry that is no text position, is
        refused there (ValueError / IndexError) instead of being read out of bounds."""
        self._text = N.as_u8(text)
        self._sa = np.ascontiguousarray(sa, dtype=np.int32)
        self._device = device
        self._dev = None

    # -- reference API ----------------------------------------------------------------
    def into_parts(self):
        """(text, sa), lib.rs:175-177"""
        return self._text, self._sa

    def text(self) -> np.ndarray:
        return self._text

    @property
    def sa(self) -> np.ndarray:
        return self._sa

    def verify(self) -> None:
        """lib.rs:127-149,180-182; raises NotSorted.  O(n) on the GPU (gsa_index_verify)."""
        if self._text.size == 0:
            raise OverflowError("attempt to subtract with overflow")  # input.len() - 1, lib.rs:143
        bad = C.c_int64(-1)
        rc = N.lib.gsa_index_verify(self._handle(), C.byref(bad))
        if rc == 1:
            raise NotSorted(bad.value, bad.value + 1)
        N.check(rc, "gsa_index_verify")

    def longest_substring_match(self, needle) -> LongestCommonSubstring:
        """lib.rs:39-99,190-197"""
        s, l = self.longest_substring_match_batch([needle])
        return LongestCommonSubstring(self._text, int(s[0]), int(l[0]))

    # -- batched forms (what the GPU actually runs) -----------------------------------------
    def longest_substring_match_batch(self, needles):
        flat, off = N.pack_patterns(needles)
        q = off.size - 1
        if self._sa.size == 0:
            raise IndexError("index out of bounds: the len is 0 but the index is 0")  # lib.rs:89-91
        start = np.empty(q, dtype=np.uint64)
        length = np.empty(q, dtype=np.uint32)
        rc = N.lib.gsa_lsm_batch(self._handle(), N.ptr(flat), N.ptr(off), q, N.ptr(start), N.ptr(length))
        N.check(rc, "gsa_lsm_batch")
        return start, length

    def search_all_batch(self, patterns):
        """-> (left, count): occurrences of pattern q are sa[left[q] : left[q]+count[q]] (sa_search)."""
        flat, off = N.pack_patterns(patterns)
        q = off.size - 1
        left = np.empty(q, dtype=np.int32)
        count = np.empty(q, dtype=np.int32)
        rc = N.lib.gsa_search_all_batch(self._handle(), N.ptr(flat), N.ptr(off), q, N.ptr(left), N.ptr(count))
        N.check(rc, "gsa_search_all_batch")
        return left, count

    def search_all(self, pattern) -> np.ndarray:
        left, count = self.search_all_batch([pattern])
        return self._sa[int(left[0]):int(left[0]) + int(count[0])] if count[0] > 0 else self._sa[:0]

    def contains_batch(self, patterns) -> np.ndarray:
        flat, off = N.pack_patterns(patterns)
        q = off.size - 1
        out = np.empty(q, dtype=np.uint8)
        rc = N.lib.gsa_contains_batch(self._handle(), N.ptr(flat), N.ptr(off), q, N.ptr(out))
        N.check(rc, "gsa_contains_batch")
        return out.astype(bool)

    def contains(self, pattern) -> bool:
        return bool(self.contains_batch([pattern])[0])

    # -- plumbing -------------------------------------------------------------------------
    def _handle(self):
        if self._dev is None:
            if self._sa.size != self._text.size:
                raise ValueError(f"text and suffix array should have same len ({self._text.size} != {self._sa.size})")
            h = C.c_void_p()
            dev = N.current_device() if self._device is None else int(self._device)
            rc = N.lib.gsa_index_from_parts(N.ptr(self._text), self._text.size, N.ptr(self._sa), self._sa.size, dev, C.byref(h))
            if rc == N.GSA_EPANIC:
                raise IndexError(f"index out of bounds: {N.last_error()}")  # lib.rs:53-57 slices text[sa[x]..]
            N.check(rc, "gsa_index_from_parts")
            self._device = dev
            self._dev = _DeviceIndex(h)
        return self._dev.h


def longest_substring_match(text, sa, needle) -> LongestCommonSubstring:
    """free function, lib.rs:39-99"""
    return SuffixArray(text, np.asarray(sa, dtype=np.int32)).longest_substring_match(needle)


def verify(text, sa) -> None:
    """free function, lib.rs:127-149"""
    SuffixArray(text, np.asarray(sa, dtype=np.int32)).verify()
