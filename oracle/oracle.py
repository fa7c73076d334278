"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs import this module (see oracle/oracle.h).  It exposes

* ``port``  -- liboracle.so, the plain-C restatement in oracle.c, and
* ``ref``   -- oracle/_ref/libdivsufsort_ref*.so, the reference's own vendored C
  libdivsufsort compiled from /root/reference by oracle/Makefile (prebuilt files
  are used as shipped when the reference checkout is absent, e.g. on the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> None:
    """Compile liboracle.so and (when /root/reference exists) oracle/_ref."""
    need = force or not os.path.exists(os.path.join(_HERE, "liboracle.so"))
    need = need or not os.path.exists(os.path.join(_HERE, "_ref", "libdivsufsort_ref.so"))
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


def _as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        assert b.dtype == np.uint8
        return np.ascontiguousarray(b)
    return np.frombuffer(bytes(b), dtype=np.uint8)


def _pack_patterns(pats):
    """list of bytes-likes -> (flat uint8 array, uint64 offsets[Q+1])."""
    if isinstance(pats, tuple) and len(pats) == 2:
        return np.ascontiguousarray(pats[0], dtype=np.uint8), np.ascontiguousarray(pats[1], dtype=np.uint64)
    lens = np.fromiter((len(p) for p in pats), dtype=np.uint64, count=len(pats))
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    flat = np.frombuffer(b"".join(bytes(p) for p in pats), dtype=np.uint8)
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(flat), off


class _Port:
    """liboracle.so (oracle.c)."""

    def __init__(self):
        build()
        self.lib = L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.oracle_sa_build.argtypes = [_u8p, _i32p, C.c_int32]
        L.oracle_sa_build.restype = C.c_int32
        L.oracle_longest_substring_match.argtypes = [_u8p, C.c_size_t, _i32p, C.c_size_t, _u8p, C.c_size_t, _u64p, _u64p]
        L.oracle_longest_substring_match.restype = C.c_int32
        L.oracle_verify.argtypes = [_u8p, C.c_size_t, _i32p, _u64p]
        L.oracle_verify.restype = C.c_int32
        L.oracle_sa_search.argtypes = [_u8p, C.c_int32, _u8p, C.c_int32, _i32p, C.c_int32, _i32p]
        L.oracle_sa_search.restype = C.c_int32
        L.oracle_sufcheck.argtypes = [_u8p, _i32p, C.c_int32]
        L.oracle_sufcheck.restype = C.c_int32
        L.oracle_lcp_kasai.argtypes = [_u8p, _i32p, C.c_int32, _i32p]
        L.oracle_lcp_kasai.restype = C.c_int32
        L.oracle_part_plan.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u64p]
        L.oracle_part_plan.restype = C.c_int32
        L.oracle_part_lsm.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_i32p), _u8p, C.c_size_t, _u64p, _u64p]
        L.oracle_part_lsm.restype = C.c_int32
        L.oracle_lsm_batch.argtypes = [_u8p, C.c_size_t, _i32p, C.c_size_t, _u8p, _u64p, C.c_uint64, _u64p, _u32p, C.c_int]
        L.oracle_lsm_batch.restype = C.c_int32
        L.oracle_search_all_batch.argtypes = [_u8p, C.c_int32, _i32p, _u8p, _u64p, C.c_uint64, _i32p, _i32p, C.c_int]
        L.oracle_search_all_batch.restype = C.c_int32
        L.oracle_part_lsm_batch.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_i32p), _u8p, _u64p, C.c_uint64, _u64p, _u32p, C.c_int]
        L.oracle_part_lsm_batch.restype = C.c_int32
        L.oracle_max_threads.restype = C.c_int

    # -- SA ----------------------------------------------------------------
    def sa_build(self, text) -> np.ndarray:
        t = _as_u8(text)
        sa = np.empty(t.size, dtype=np.int32)
        tp = _ptr(t, _u8p) if t.size else _ptr(np.zeros(1, np.uint8), _u8p)
        sp = _ptr(sa, _i32p) if t.size else _ptr(np.zeros(1, np.int32), _i32p)
        rc = self.lib.oracle_sa_build(tp, sp, t.size)
        if rc != 0:
            raise RuntimeError(f"oracle_sa_build rc={rc}")
        return sa

    def sufcheck(self, text, sa) -> int:
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        if t.size == 0:
            return 0
        return self.lib.oracle_sufcheck(_ptr(t, _u8p), _ptr(sa, _i32p), t.size)

    def lcp(self, text, sa) -> np.ndarray:
        """LCP[0] = 0, LCP[j] = lcp(suffix sa[j-1], suffix sa[j]) (Kasai)."""
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        out = np.zeros(t.size, np.int32)
        if t.size:
            rc = self.lib.oracle_lcp_kasai(_ptr(t, _u8p), _ptr(sa, _i32p), t.size, _ptr(out, _i32p))
            if rc != 0:
                raise RuntimeError(f"oracle_lcp_kasai rc={rc}")
        return out

    def verify(self, text, sa):
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        bad = C.c_uint64(0)
        rc = self.lib.oracle_verify(_ptr(t, _u8p) if t.size else None, t.size, _ptr(sa, _i32p) if t.size else None, C.byref(bad))
        return rc, bad.value

    # -- search ------------------------------------------------------------
    def longest_substring_match(self, text, sa, needle):
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        nd = _as_u8(needle)
        ndp = _ptr(nd, _u8p) if nd.size else _ptr(np.zeros(1, np.uint8), _u8p)
        s, l = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.oracle_longest_substring_match(_ptr(t, _u8p), t.size, _ptr(sa, _i32p), sa.size, ndp, nd.size, C.byref(s), C.byref(l))
        if rc != 0:
            raise IndexError("longest_substring_match on an empty suffix array (reference panics)")
        return s.value, l.value

    def sa_search(self, text, sa, pat):
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        p = _as_u8(pat)
        pp = _ptr(p, _u8p) if p.size else _ptr(np.zeros(1, np.uint8), _u8p)
        tp = _ptr(t, _u8p) if t.size else _ptr(np.zeros(1, np.uint8), _u8p)
        sp = _ptr(sa, _i32p) if sa.size else _ptr(np.zeros(1, np.int32), _i32p)
        idx = C.c_int32(-1)
        cnt = self.lib.oracle_sa_search(tp, t.size, pp, p.size, sp, sa.size, C.byref(idx))
        return cnt, idx.value

    def lsm_batch(self, text, sa, pats, threads: int = 0):
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        flat, off = _pack_patterns(pats)
        q = off.size - 1
        st = np.empty(q, dtype=np.uint64)
        ln = np.empty(q, dtype=np.uint32)
        rc = self.lib.oracle_lsm_batch(_ptr(t, _u8p), t.size, _ptr(sa, _i32p), sa.size, _ptr(flat, _u8p), _ptr(off, _u64p), q, _ptr(st, _u64p), _ptr(ln, _u32p), threads)
        if rc != 0:
            raise IndexError("empty suffix array")
        return st, ln

    def search_all_batch(self, text, sa, pats, threads: int = 0):
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        flat, off = _pack_patterns(pats)
        q = off.size - 1
        left = np.empty(q, dtype=np.int32)
        cnt = np.empty(q, dtype=np.int32)
        self.lib.oracle_search_all_batch(_ptr(t, _u8p), t.size, _ptr(sa, _i32p), _ptr(flat, _u8p), _ptr(off, _u64p), q, _ptr(left, _i32p), _ptr(cnt, _i32p), threads)
        return left, cnt

    # -- sacapart ----------------------------------------------------------
    def part_plan(self, n: int, num_partitions: int):
        ps, ap = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.oracle_part_plan(n, num_partitions, C.byref(ps), C.byref(ap))
        if rc != 0:
            raise ZeroDivisionError("num_partitions == 0")
        return ps.value, ap.value

    def part_build(self, text, num_partitions: int, builder=None):
        """-> (partition_size, [sa_0, sa_1, ...]) using `builder` (default: this oracle)."""
        t = _as_u8(text)
        ps, ap = self.part_plan(t.size, num_partitions)
        builder = builder or self.sa_build
        return ps, [builder(t[i * ps:min((i + 1) * ps, t.size)]) for i in range(ap)]

    def _sa_ptrs(self, sas):
        arr = (_i32p * max(1, len(sas)))()
        for i, s in enumerate(sas):
            arr[i] = _ptr(s, _i32p)
        return arr

    def part_lsm(self, text, partition_size, sas, needle):
        t = _as_u8(text)
        nd = _as_u8(needle)
        ndp = _ptr(nd, _u8p) if nd.size else _ptr(np.zeros(1, np.uint8), _u8p)
        s, l = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.oracle_part_lsm(_ptr(t, _u8p) if t.size else None, t.size, partition_size, len(sas), self._sa_ptrs(sas), ndp, nd.size, C.byref(s), C.byref(l))
        if rc != 0:
            raise RuntimeError("partitioned suffix arrays should always find at least one longest common substring")
        return s.value, l.value

    def part_lsm_batch(self, text, partition_size, sas, pats, threads: int = 0):
        t = _as_u8(text)
        flat, off = _pack_patterns(pats)
        q = off.size - 1
        st = np.empty(q, dtype=np.uint64)
        ln = np.empty(q, dtype=np.uint32)
        rc = self.lib.oracle_part_lsm_batch(_ptr(t, _u8p), t.size, partition_size, len(sas), self._sa_ptrs(sas), _ptr(flat, _u8p), _ptr(off, _u64p), q, _ptr(st, _u64p), _ptr(ln, _u32p), threads)
        if rc != 0:
            raise RuntimeError("zero partitions")
        return st, ln

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())


class _Ref:
    """oracle/_ref/libdivsufsort_ref*.so: the reference's own C code."""

    def __init__(self, ndebug: bool = False):
        build()
        name = "libdivsufsort_ref_ndebug.so" if ndebug else "libdivsufsort_ref.so"
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.divsufsort.argtypes = [_u8p, _i32p, C.c_int32]
        L.divsufsort.restype = C.c_int32
        L.sufcheck.argtypes = [_u8p, _i32p, C.c_int32, C.c_int32]
        L.sufcheck.restype = C.c_int32
        L.sa_search.argtypes = [_u8p, C.c_int32, _u8p, C.c_int32, _i32p, C.c_int32, _i32p]
        L.sa_search.restype = C.c_int32
        L.divbwt.argtypes = [_u8p, _u8p, _i32p, C.c_int32]
        L.divbwt.restype = C.c_int32

    def divbwt(self, text):
        """reference divbwt -> (U, primary index)"""
        t = _as_u8(text)
        u = np.empty(t.size, dtype=np.uint8)
        tp = _ptr(t, _u8p) if t.size else _ptr(np.zeros(1, np.uint8), _u8p)
        up = _ptr(u, _u8p) if u.size else _ptr(np.zeros(1, np.uint8), _u8p)
        return u, self.lib.divbwt(tp, up, None, t.size)

    def inverse_bwt(self, u, idx: int):
        """reference inverse_bw_transform -> (rc, text)"""
        b = _as_u8(u)
        out = np.zeros(b.size, dtype=np.uint8)
        one = np.zeros(1, np.uint8)
        self.lib.inverse_bw_transform.argtypes = [_u8p, _u8p, _i32p, C.c_int32, C.c_int32]
        self.lib.inverse_bw_transform.restype = C.c_int32
        rc = self.lib.inverse_bw_transform(_ptr(b, _u8p) if b.size else _ptr(one, _u8p),
                                           _ptr(out, _u8p) if out.size else _ptr(one, _u8p), None, b.size, int(idx))
        return rc, out

    def divsufsort_raw(self, tptr, saptr, n) -> int:
        return self.lib.divsufsort(tptr, saptr, n)

    def sa_build(self, text) -> np.ndarray:
        t = _as_u8(text)
        sa = np.empty(t.size, dtype=np.int32)
        tp = _ptr(t, _u8p) if t.size else _ptr(np.zeros(1, np.uint8), _u8p)
        sp = _ptr(sa, _i32p) if t.size else _ptr(np.zeros(1, np.int32), _i32p)
        rc = self.lib.divsufsort(tp, sp, t.size)
        if rc != 0:
            raise RuntimeError(f"divsufsort rc={rc}")
        return sa

    def sufcheck(self, text, sa) -> int:
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        if t.size == 0:
            return 0
        return self.lib.sufcheck(_ptr(t, _u8p), _ptr(sa, _i32p), t.size, 0)

    def sa_search(self, text, sa, pat):
        t = _as_u8(text)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        p = _as_u8(pat)
        pp = _ptr(p, _u8p) if p.size else _ptr(np.zeros(1, np.uint8), _u8p)
        tp = _ptr(t, _u8p) if t.size else _ptr(np.zeros(1, np.uint8), _u8p)
        sp = _ptr(sa, _i32p) if sa.size else _ptr(np.zeros(1, np.int32), _i32p)
        idx = C.c_int32(-1)
        cnt = self.lib.sa_search(tp, t.size, pp, p.size, sp, sa.size, C.byref(idx))
        return cnt, idx.value


_port = None
_refs = {}


def port() -> _Port:
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref(ndebug: bool = False) -> _Ref:
    if ndebug not in _refs:
        _refs[ndebug] = _Ref(ndebug)
    return _refs[ndebug]


def have_ref() -> bool:
    try:
        ref()
        return True
    except (OSError, FileNotFoundError, subprocess.CalledProcessError):
        return False
