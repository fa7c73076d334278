"""ctypes binding of libgsa.so (the C ABI declared in include/gsa.h).

The library is the product: if it is missing this module raises -- there is no
Python / numpy / CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsa.so")

GSA_OK, GSA_EINVAL, GSA_ENOMEM, GSA_ECUDA, GSA_EPANIC = 0, -1, -2, -3, -4
GSA_MAX_ROUNDS = 40

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)


class RoundStat(C.Structure):
    _fields_ = [
        ("depth", C.c_uint64),
        ("live", C.c_uint64),
        ("groups", C.c_uint32),
        ("key_bits", C.c_uint32),
        ("passes", C.c_uint32),
        ("sorted", C.c_uint32),
        ("ms_total", C.c_float),
        ("ms_sort", C.c_float),
        ("bag", C.c_uint32),
        ("reserved_", C.c_uint32),
    ]


class BuildStats(C.Structure):
    _fields_ = [
        ("rounds", C.c_uint32),
        ("sigma", C.c_uint32),
        ("bits_per_symbol", C.c_uint32),
        ("symbols_per_key", C.c_uint32),
        ("ms_total", C.c_float),
        ("ms_h2d", C.c_float),
        ("ms_d2h", C.c_float),
        ("radix_pass_launches", C.c_uint64),
        ("radix_pass_elements", C.c_uint64),
        ("ms_radix_passes", C.c_float),
        ("kernel_launches", C.c_uint64),
        ("radix_pass_bytes", C.c_uint64),
        ("round", RoundStat * GSA_MAX_ROUNDS),
    ]

    def rounds_list(self):
        return [
            dict(depth=int(r.depth), live=int(r.live), sorted=int(r.sorted), groups=int(r.groups), key_bits=int(r.key_bits),
                 passes=int(r.passes), ms_total=float(r.ms_total), ms_sort=float(r.ms_sort), bag=int(r.bag))
            for r in list(self.round)[: min(self.rounds, GSA_MAX_ROUNDS)]
        ]

    def algorithmic_bytes(self) -> int:
        """round 0 n*(41+24p); round k>=1 L*52 + S*24p + B*32, where S <= L is the
        number of suffixes actually sorted (inert members of huge groups are walked, not sorted) and
        B the suffixes of tiny groups refined in the bag (suffix, slot, label gather, key half out and
        in, suffix + slot out, SA: 8 words of 4 bytes)."""
        total = 0
        for i, r in enumerate(self.rounds_list()):
            if i == 0:
                total += r["live"] * (41 + 24 * r["passes"])
            else:
                total += r["live"] * 52 + r["sorted"] * 24 * r["passes"] + r["bag"] * 32
        return total


class GsaError(RuntimeError):
    def __init__(self, rc: int, where: str, detail: str = ""):
        self.rc = rc
        names = {-1: "GSA_EINVAL", -2: "GSA_ENOMEM", -3: "GSA_ECUDA", -4: "GSA_EPANIC"}
        super().__init__(f"{where}: {names.get(rc, rc)} {detail}".strip())


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C stringsearch_b200/csrc`). "
            "stringsearch_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "gsa_divsufsort": ([vp, vp, C.c_int32], C.c_int32),
        "gsa_divsufsort_ex": ([vp, vp, C.c_int32, C.c_int32, C.POINTER(BuildStats)], C.c_int32),
        "gsa_build_workspace_bytes": ([C.c_int32], C.c_size_t),
        "gsa_build_device": ([vp, vp, C.c_int32, vp, C.c_size_t, vp, C.POINTER(BuildStats)], C.c_int32),
        "gsa_divbwt": ([vp, vp, vp, C.c_int32], C.c_int32),
        "gsa_bwt_device": ([vp, vp, C.c_int32, vp, i32p, vp], C.c_int32),
        "gsa_inverse_bw_transform": ([vp, vp, vp, C.c_int32, C.c_int32], C.c_int32),
        "gsa_inverse_bwt_workspace_bytes": ([C.c_int32], C.c_size_t),
        "gsa_inverse_bwt_device": ([vp, vp, C.c_int32, C.c_int32, vp, C.c_size_t, vp], C.c_int32),
        "gsa_lcp_workspace_bytes": ([C.c_int32], C.c_size_t),
        "gsa_lcp_device": ([vp, vp, vp, C.c_int32, vp, C.c_size_t, vp], C.c_int32),
        "gsa_lcp": ([vp, vp, vp, C.c_int32, C.c_int32], C.c_int32),
        "gsa_divsufsort_lcp": ([vp, vp, vp, C.c_int32, C.c_int32], C.c_int32),
        "gsa_sufcheck_device": ([vp, vp, C.c_int32, vp, i64p], C.c_int32),
        "gsa_sufcheck": ([vp, vp, C.c_int32, C.c_int32, i64p], C.c_int32),
        "gsa_index_create": ([vp, C.c_int64, C.c_int32, C.POINTER(vp), C.POINTER(BuildStats)], C.c_int32),
        "gsa_index_from_parts": ([vp, C.c_int64, vp, C.c_int64, C.c_int32, C.POINTER(vp)], C.c_int32),
        "gsa_index_create_shard": ([vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(vp), C.POINTER(BuildStats)], C.c_int32),
        "gsa_index_len": ([vp], C.c_int64),
        "gsa_index_device": ([vp], C.c_int32),
        "gsa_index_sa": ([vp, vp], C.c_int32),
        "gsa_index_verify": ([vp, i64p], C.c_int32),
        "gsa_index_device_text": ([vp], vp),
        "gsa_index_device_sa": ([vp], vp),
        "gsa_index_destroy": ([vp], None),
        "gsa_lsm_batch": ([vp, vp, vp, C.c_uint64, vp, vp], C.c_int32),
        "gsa_search_all_batch": ([vp, vp, vp, C.c_uint64, vp, vp], C.c_int32),
        "gsa_contains_batch": ([vp, vp, vp, C.c_uint64, vp], C.c_int32),
        "gsa_lsm_device": ([vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, vp, vp, vp], C.c_int32),
        "gsa_search_all_device": ([vp, vp, vp, C.c_uint64, C.c_uint32, vp, vp, vp], C.c_int32),
        "gsa_lsm_reduce_device": ([vp, vp, C.c_uint64, C.c_uint32, vp], C.c_int32),
        "gsa_part_create": ([vp, C.c_uint64, C.c_uint64, i32p, C.c_int32, C.POINTER(vp)], C.c_int32),
        "gsa_part_num_partitions": ([vp], C.c_uint64),
        "gsa_part_partition_size": ([vp], C.c_uint64),
        "gsa_part_shard": ([vp, C.c_uint64], vp),
        "gsa_part_lsm_batch": ([vp, vp, vp, C.c_uint64, vp, vp], C.c_int32),
        "gsa_part_destroy": ([vp], None),
        "gsa_host_alloc": ([C.c_size_t], vp),
        "gsa_host_free": ([vp], None),
        "gsa_release_cached_memory": ([], None),
        "gsa_last_error": ([], C.c_char_p),
        "gsa_version": ([], C.c_char_p),
        "gsa_device_count": ([], C.c_int32),
        "gsa_current_device": ([], C.c_int32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.argtypes = args
        fn.restype = res
    L._gsa_symbols = tuple(sig)
    return L


lib = _load()
EXPORTED_SYMBOLS = lib._gsa_symbols


def last_error() -> str:
    s = lib.gsa_last_error()
    return s.decode(errors="replace") if s else ""


def current_device() -> int:
    """The calling thread's current CUDA device (the one gsa_divsufsort() builds on)."""
    return int(lib.gsa_current_device())


def check(rc: int, where: str) -> None:
    if rc != GSA_OK:
        raise GsaError(rc, where, last_error())


def as_u8(buf) -> np.ndarray:
    """bytes / bytearray / memoryview / uint8 ndarray -> contiguous uint8 ndarray (no copy when possible)."""
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8:
            raise TypeError("text must be uint8")
        return np.ascontiguousarray(buf)
    return np.frombuffer(buf, dtype=np.uint8)


def ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data if a.size else 0)


def pack_patterns(pats):
    """list of bytes-likes, or (flat uint8 array, uint64 offsets[Q+1]) -> (flat, offsets)."""
    if isinstance(pats, tuple) and len(pats) == 2:
        flat = np.ascontiguousarray(pats[0], dtype=np.uint8)
        off = np.ascontiguousarray(pats[1], dtype=np.uint64)
        # the library reads flat[off[q] : off[q+1]] for every q: refuse offsets that leave the buffer
        if off.ndim != 1 or off.size < 1:
            raise ValueError("pattern offsets must be a 1-d array of Q + 1 entries")
        if off.size > 1 and bool((off[1:] < off[:-1]).any()):
            raise ValueError("pattern offsets must be non-decreasing")
        if int(off[-1]) > flat.size:
            raise ValueError(f"pattern offsets end at {int(off[-1])} but only {flat.size} pattern bytes were given")
        return flat, off
    lens = np.fromiter((len(p) for p in pats), dtype=np.uint64, count=len(pats))
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    flat = np.frombuffer(b"".join(bytes(p) for p in pats), dtype=np.uint8)
    return np.ascontiguousarray(flat), off


class PinnedBuffer:
    """Page-locked host memory from gsa_host_alloc, exposed as a numpy array."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self._p = lib.gsa_host_alloc(max(1, self.nbytes))
        if not self._p:
            raise MemoryError(f"gsa_host_alloc({nbytes}) failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, self.nbytes)).from_address(self._p))[: self.nbytes]

    def view(self, dtype, count=None):
        a = self.array.view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if self._p:
            self.array = None
            lib.gsa_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
