"""divsufsort::sort / sort_in_place on the GPU.

Mirrors crates/divsufsort/src/lib.rs:20-29 (and its twin crates/cdivsufsort/src/lib.rs:9-30):
same names, same argument meaning, and the reference's panics become Python exceptions
raised at the same preconditions.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .sacabase import SuffixArray

I32_MAX = 2**31 - 1


def sort_in_place(text, sa: np.ndarray, device: int | None = None, stats: N.BuildStats | None = None) -> None:
    """Sort suffixes of `text` and store their lexicographic order in `sa` (int32).

    Panics (AssertionError) like the reference when ``len(sa) != len(text)``
    (divsufsort.rs:4-8) or ``len(text) >= i32::MAX`` (divsufsort.rs:9-13).
    """
    t = N.as_u8(text)
    if not (isinstance(sa, np.ndarray) and sa.dtype == np.int32 and sa.flags.c_contiguous and sa.flags.writeable):
        raise TypeError("sa must be a writable C-contiguous int32 ndarray")
    # the reference's assert!s (they guard the buffers the library writes, so they are raised
    # explicitly: a bare `assert` disappears under python -O)
    if t.size != sa.size:
        raise AssertionError("text and suffix array should have same len")
    if t.size >= I32_MAX:
        raise AssertionError(f"text too large, should not exceed {I32_MAX - 1} bytes")
    if device is None and stats is None:
        rc = N.lib.gsa_divsufsort(N.ptr(t) if t.size else N.ptr(np.zeros(1, np.uint8)),
                                  N.ptr(sa) if sa.size else N.ptr(np.zeros(1, np.int32)), t.size)
    else:
        rc = N.lib.gsa_divsufsort_ex(N.ptr(t) if t.size else N.ptr(np.zeros(1, np.uint8)),
                                     N.ptr(sa) if sa.size else N.ptr(np.zeros(1, np.int32)), t.size,
                                     N.current_device() if device is None else device,
                                     C.byref(stats) if stats is not None else None)
    if rc != 0:  # cdivsufsort lib.rs:22 assert_eq!(0, ret)
        raise AssertionError(f"divsufsort returned {rc}: {N.last_error()}")


def sort(text, device: int | None = None, stats: N.BuildStats | None = None) -> SuffixArray:
    """-> sacabase.SuffixArray (text borrowed, sa owned), like divsufsort::sort (lib.rs:25-29)."""
    t = N.as_u8(text)
    sa = np.zeros(t.size, dtype=np.int32)
    dev = N.current_device() if device is None else int(device)  # one device for the build and the resident index
    sort_in_place(t, sa, device=dev, stats=stats)
    return SuffixArray(t, sa, device=dev)


def bwt(text):
    """divbwt (crates/cdivsufsort/c-sources/divsufsort.c:372-405): -> (U, primary_index)."""
    t = N.as_u8(text)
    if t.size >= I32_MAX:
        raise AssertionError(f"text too large, should not exceed {I32_MAX - 1} bytes")
    u = np.empty(t.size, dtype=np.uint8)
    rc = N.lib.gsa_divbwt(N.ptr(t) if t.size else N.ptr(np.zeros(1, np.uint8)),
                          N.ptr(u) if u.size else N.ptr(np.zeros(1, np.uint8)), None, t.size)
    if rc < 0:
        raise AssertionError(f"divbwt returned {rc}: {N.last_error()}")
    return u, rc


def inverse_bwt(u, primary_index: int) -> np.ndarray:
    """inverse_bw_transform (crates/cdivsufsort/c-sources/utils.c:111-156): (U, primary index) -> text."""
    b = N.as_u8(u)
    out = np.empty(b.size, dtype=np.uint8)
    one = np.zeros(1, np.uint8)
    rc = N.lib.gsa_inverse_bw_transform(N.ptr(b) if b.size else N.ptr(one), N.ptr(out) if out.size else N.ptr(one), None,
                                        b.size, int(primary_index))
    if rc != 0:
        raise N.GsaError(rc, "gsa_inverse_bw_transform", N.last_error())
    return out


def lcp(text, sa, device: int | None = None) -> np.ndarray:
    """LCP array of a suffix array: LCP[0] = 0, LCP[j] = lcp(suffix sa[j-1], suffix sa[j]).
    (No counterpart in the reference)"""
    t = N.as_u8(text)
    s = np.ascontiguousarray(sa, dtype=np.int32)
    if s.size != t.size:
        raise ValueError("sa and text must have the same length")
    out = np.zeros(t.size, dtype=np.int32)
    if t.size:
        rc = N.lib.gsa_lcp(N.ptr(t), N.ptr(s), N.ptr(out), t.size, N.current_device() if device is None else device)
        if rc != 0:
            raise N.GsaError(rc, "gsa_lcp", N.last_error())
    return out


def sort_with_lcp(text, device: int | None = None):
    """-> (SuffixArray, LCP) from one call: the text is uploaded once and the suffix array never
    leaves the device between the two steps."""
    t = N.as_u8(text)
    if t.size >= I32_MAX:
        raise AssertionError(f"text too large, should not exceed {I32_MAX - 1} bytes")
    sa = np.zeros(t.size, dtype=np.int32)
    out = np.zeros(t.size, dtype=np.int32)
    dev = N.current_device() if device is None else int(device)
    if t.size:
        rc = N.lib.gsa_divsufsort_lcp(N.ptr(t), N.ptr(sa), N.ptr(out), t.size, dev)
        if rc != 0:
            raise N.GsaError(rc, "gsa_divsufsort_lcp", N.last_error())
    return SuffixArray(t, sa, device=dev), out
