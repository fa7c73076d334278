"""sacapart::PartitionedSuffixArray on one or more GPUs.

Mirrors crates/sacapart/src/lib.rs:26-98.  Two deployments:

* ``PartitionedSuffixArray``: one process driving `devices` (shard i lives on
  devices[i % len(devices)]), everything inside libgsa.so (gsa_part_*).
* ``DistributedPartitionedSuffixArray``: one process per GPU under torch.distributed
  (NCCL on GPUs, gloo in the CPU tests).  Rank r owns shards i with i % world == r; shards
  are built with no communication; a query broadcasts the pattern batch, every rank
  answers for its shards, and the per-rank (start, len) sets are all-gathered and merged
  with the reference's tie rule (longer wins; equal length -> lower partition).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .sacabase import LongestCommonSubstring, StringIndex


def partition_plan(n: int, num_partitions: int):
    """(partition_size, actual number of partitions): lib.rs:43 and par_chunks :45-49,60-62."""
    if num_partitions == 0:
        raise ZeroDivisionError("attempt to divide by zero")  # lib.rs:43
    ps = n // num_partitions + 1
    return ps, (n + ps - 1) // ps


class PartitionedSuffixArray(StringIndex):
    def __init__(self, text, num_partitions: int, devices=None):
        """PartitionedSuffixArray::new(text, num_partitions, divsufsort::sort) (lib.rs:39-58).

        The builder closure of the reference is fixed to the GPU divsufsort; `devices`
        (default: the current device) says where the shards go.
        """
        self._text = N.as_u8(text)
        if num_partitions == 0:
            raise ZeroDivisionError("attempt to divide by zero")
        devs = None
        nd = 0
        if devices is not None:
            devs = (C.c_int32 * len(devices))(*devices)
            nd = len(devices)
        h = C.c_void_p()
        rc = N.lib.gsa_part_create(N.ptr(self._text), self._text.size, num_partitions, devs, nd, C.byref(h))
        N.check(rc, "gsa_part_create")
        self._h = h

    def num_partitions(self) -> int:
        """lib.rs:60-62"""
        return int(N.lib.gsa_part_num_partitions(self._h))

    def partition_size(self) -> int:
        return int(N.lib.gsa_part_partition_size(self._h))

    def shard_sa(self, i: int) -> np.ndarray:
        ix = N.lib.gsa_part_shard(self._h, i)
        n = N.lib.gsa_index_len(ix)
        out = np.empty(n, dtype=np.int32)
        N.check(N.lib.gsa_index_sa(ix, N.ptr(out)), "gsa_index_sa")
        return out

    def longest_substring_match_batch(self, needles):
        flat, off = N.pack_patterns(needles)
        q = off.size - 1
        start = np.empty(q, dtype=np.uint64)
        length = np.empty(q, dtype=np.uint32)
        rc = N.lib.gsa_part_lsm_batch(self._h, N.ptr(flat), N.ptr(off), q, N.ptr(start), N.ptr(length))
        if rc == N.GSA_EPANIC:
            raise RuntimeError("partitioned suffix arrays should always find at least one longest common substring")
        N.check(rc, "gsa_part_lsm_batch")
        return start, length

    def longest_substring_match(self, needle) -> LongestCommonSubstring:
        """lib.rs:69-97"""
        s, l = self.longest_substring_match_batch([needle])
        return LongestCommonSubstring(self._text, int(s[0]), int(l[0]))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            N.lib.gsa_part_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DeviceBuffers:
    """Grow-only device buffers reused between query batches (a fresh torch allocation per collective and call
    means allocator traffic and cross-stream bookkeeping on every batch)."""

    def __init__(self):
        self._b = {}

    def get(self, name: str, numel: int, dtype, dev):
        import torch

        t = self._b.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype or t.device != dev:
            t = torch.empty(max(1, numel), dtype=dtype, device=dev)
            self._b[name] = t
        return t[:numel]


def merge_results(starts: np.ndarray, lens: np.ndarray):
    """Host statement of the cross-rank merge rule used by the distributed query
    (sacapart lib.rs:86-92): starts/lens are [nsets, Q]; longer wins, equal length -> the
    smaller start (= the lower partition, partitions being disjoint ascending ranges).
    Used by the gloo tests to check the device reduction; the product path reduces on the GPU."""
    best_s, best_l = starts[0].copy(), lens[0].copy()
    for s, l in zip(starts[1:], lens[1:]):
        take = (l > best_l) | ((l == best_l) & (s < best_s))
        best_s[take], best_l[take] = s[take], l[take]
    return best_s, best_l


class DistributedPartitionedSuffixArray(StringIndex):
    """One rank per GPU; see module docstring.  `text` is the full text on every rank (the
    reference's `&'a [u8]`); each rank uploads and indexes only its own shards (+ halo)."""

    def __init__(self, text, num_partitions: int, device: int, group=None, halo: int = 4096):
        import torch.distributed as dist

        self._dist = dist
        self._group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._text = N.as_u8(text)
        self._device = device
        self._ps, self._np = partition_plan(self._text.size, num_partitions)
        self._halo = halo
        self._bufs = _DeviceBuffers()
        self._shards = []  # (partition index, offset, handle)
        self.build_ms = []  # device time of every local shard build (gsa_build_stats.ms_total), in shard order
        for i in range(self.rank, self._np, self.world):
            off = i * self._ps
            ln = min(self._ps, self._text.size - off)
            self._shards.append((i, off, self._build_shard(off, ln)))

    # The two methods below are the only device-touching steps.  The gloo CPU tests override
    # them with the oracle to exercise the collective plumbing without a GPU; the product
    # implementation has no such path.
    def _build_shard(self, off: int, ln: int):
        h = C.c_void_p()
        st = N.BuildStats()
        rc = N.lib.gsa_index_create_shard(N.ptr(self._text), self._text.size, off, ln, self._halo, self._device,
                                          C.byref(h), C.byref(st))
        N.check(rc, "gsa_index_create_shard")
        self.build_ms.append(float(st.ms_total))
        return h

    def _answer_local(self, t_pat, t_off, q, t_start, t_len, dev, max_len: int = 0) -> None:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("stringsearch_b200 has no CPU path: shards need a CUDA device")
        stream = torch.cuda.current_stream(dev).cuda_stream
        for k, (_, offset, h) in enumerate(self._shards):
            rc = N.lib.gsa_lsm_device(h, t_pat.data_ptr(), t_off.data_ptr(), q, max_len, offset, 0 if k == 0 else 1,
                                      t_start.data_ptr(), t_len.data_ptr(), stream)
            N.check(rc, "gsa_lsm_device")

    def _destroy_shard(self, h) -> None:
        N.lib.gsa_index_destroy(h)

    def num_partitions(self) -> int:
        return self._np

    def local_partitions(self):
        return [i for i, _, _ in self._shards]

    def longest_substring_match_batch(self, needles, src: int = 0):
        """Collective: every rank calls it; `needles` is only read on rank `src`.
        Returns (start, len) numpy arrays on every rank."""
        import torch

        if self._np == 0:
            raise RuntimeError("partitioned suffix arrays should always find at least one longest common substring")
        dev = torch.device("cuda", self._device) if torch.cuda.is_available() else torch.device("cpu")
        t_pat = t_off = None
        if self.rank == src:
            flat, off = N.pack_patterns(needles)
            t_off = torch.from_numpy(off.astype(np.int64)).to(dev)
            t_pat = torch.from_numpy(flat if flat.size else np.zeros(1, np.uint8)).to(dev)
        t_start, t_len = self.longest_substring_match_device(t_pat, t_off, src)
        return t_start.cpu().numpy().astype(np.uint64), t_len.cpu().numpy().astype(np.uint32)

    def longest_substring_match_device(self, t_pat, t_off, src: int = 0):
        """The fan-out query on device tensors (collective).  On rank `src`: `t_pat` uint8 pattern
        bytes, `t_off` int64 offsets [Q + 1], both on this rank's GPU; ignored elsewhere.  Steps:
        broadcast of the batch, every rank answers for its shards in ascending partition order
        (offset / may_extend / strict-greater: gsa_lsm_device), all-gather of the per-rank
        (start, len) sets, merge on the device (gsa_lsm_reduce_device; lib.rs:86-92).
        Returns (start int64 [Q], len int32 [Q]) device tensors, identical on every rank."""
        import torch

        dist = self._dist
        on_gpu = torch.cuda.is_available()
        dev = torch.device("cuda", self._device) if on_gpu else torch.device("cpu")
        # ---- broadcast the pattern batch ------------------------------------------------
        if self.rank == src:
            hdr = torch.tensor([t_off.numel() - 1, t_pat.numel()], dtype=torch.int64, device=dev)
        else:
            hdr = torch.zeros(2, dtype=torch.int64, device=dev)
        if self.world > 1:
            dist.broadcast(hdr, src=src, group=self._group)
        h_host = hdr.tolist()
        q, nbytes = int(h_host[0]), int(h_host[1])
        if self.rank != src:
            t_off = self._bufs.get("off", q + 1, torch.int64, dev)
            t_pat = self._bufs.get("pat", max(1, nbytes), torch.uint8, dev)
        if self.world > 1:
            dist.broadcast(t_off, src=src, group=self._group)
            dist.broadcast(t_pat, src=src, group=self._group)
        max_len = int((t_off[1:] - t_off[:-1]).max()) if q else 0
        if max_len > self._halo + 1:
            raise ValueError(f"needle of {max_len} bytes exceeds the shard halo ({self._halo}); rebuild with a larger halo")
        # ---- local shards, ascending partition index --------------------------------------
        t_start = self._bufs.get("start", q, torch.int64, dev).zero_()
        t_len = self._bufs.get("len", q, torch.int32, dev).zero_()
        if self._shards:
            self._answer_local(t_pat, t_off, q, t_start, t_len, dev, max_len)
        else:
            t_start.fill_(2**62)  # a rank without shards never wins (len 0, huge start)
        # ---- gather + merge ---------------------------------------------------------------------
        if self.world > 1:
            g_start = self._bufs.get("g_start", self.world * q, torch.int64, dev)
            g_len = self._bufs.get("g_len", self.world * q, torch.int32, dev)
            dist.all_gather_into_tensor(g_start, t_start, group=self._group)
            dist.all_gather_into_tensor(g_len, t_len, group=self._group)
            if on_gpu:
                stream = torch.cuda.current_stream(dev).cuda_stream
                rc = N.lib.gsa_lsm_reduce_device(g_start.data_ptr(), g_len.data_ptr(), q, self.world, stream)
                N.check(rc, "gsa_lsm_reduce_device")
                t_start, t_len = g_start[:q], g_len[:q]
            else:  # gloo CPU tests of the plumbing only (no shards can exist without a GPU)
                s, l = merge_results(g_start.view(self.world, q).numpy().astype(np.uint64),
                                     g_len.view(self.world, q).numpy().astype(np.uint32))
                t_start, t_len = torch.from_numpy(s.astype(np.int64)), torch.from_numpy(l.astype(np.int32))
        return t_start, t_len

    def longest_substring_match(self, needle) -> LongestCommonSubstring:
        s, l = self.longest_substring_match_batch([needle])
        return LongestCommonSubstring(self._text, int(s[0]), int(l[0]))

    def close(self):
        for _, _, h in self._shards:
            self._destroy_shard(h)
        self._shards = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def slice_plan(q: int, world: int, chunks: int, chunk_min: int):
    """How a batch of `q` needles is cut for `world` ranks: rank r answers needles [r * per, (r + 1) * per) with
    per = ceil(q / world), and its slice travels in up to `chunks` pieces (one piece when a piece would hold fewer than
    `chunk_min` needles).  Returns (per, starts): starts[r * chunks + k] = first needle of piece k of rank r,
    starts[world * chunks] = q; empty pieces repeat the end of their rank's slice."""
    per = (q + world - 1) // world if q else 0
    ch = chunks if per >= chunk_min * chunks else 1
    cs = (per + ch - 1) // ch if per else 0
    starts = []
    for r in range(world):
        r_lo, r_hi = min(q, r * per), min(q, (r + 1) * per)
        starts += [min(r_hi, r_lo + k * cs) if k < ch else r_hi for k in range(chunks)]
    return per, starts + [q]


class ReplicatedSuffixArray(StringIndex):
    """The un-partitioned index on every GPU, queries split across the ranks (SURVEY.md 8(e):
    "1 GiB SA, 8 GPUs" -- pure data parallelism over the needles, answers identical to one GPU).

    Every rank builds its own copy (the construction is deterministic, so the copies are equal
    and no suffix array travels between GPUs); rank r answers needles [r * ceil(Q / W),
    (r + 1) * ceil(Q / W)) of a batch held by `src`, which sends every rank its slice, and one
    all-gather per result array puts the answers together on every rank."""

    CHUNKS = 4              # pieces a rank's slice of a batch travels in (transfer of piece k + 1 overlaps the search of piece k)
    CHUNK_MIN = 1 << 16     # ... when a piece holds at least this many needles

    def __init__(self, text, device: int, group=None):
        import torch.distributed as dist

        self._dist = dist
        self._group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._text = N.as_u8(text)
        self._device = device
        self._bufs = _DeviceBuffers()
        self._h = self._build()

    # device-touching steps (overridden with the oracle by the gloo CPU tests of the plumbing)
    def _build(self):
        h = C.c_void_p()
        N.check(N.lib.gsa_index_create(N.ptr(self._text), self._text.size, self._device, C.byref(h), None), "gsa_index_create")
        return h

    def _answer_local(self, t_pat, t_off, q, t_start, t_len, dev, max_len: int = 0) -> None:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("stringsearch_b200 has no CPU path: the index needs a CUDA device")
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = N.lib.gsa_lsm_device(self._h, t_pat.data_ptr(), t_off.data_ptr(), q, max_len, 0, 0,
                                  t_start.data_ptr(), t_len.data_ptr(), stream)
        N.check(rc, "gsa_lsm_device")

    def _destroy(self, h) -> None:
        N.lib.gsa_index_destroy(h)

    def _answer_local_search_all(self, t_pat, t_off, q, t_left, t_count, dev, max_len: int = 0) -> None:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("stringsearch_b200 has no CPU path: the index needs a CUDA device")
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = N.lib.gsa_search_all_device(self._h, t_pat.data_ptr(), t_off.data_ptr(), q, max_len,
                                         t_left.data_ptr(), t_count.data_ptr(), stream)
        N.check(rc, "gsa_search_all_device")

    def longest_substring_match_batch(self, needles, src: int = 0):
        """Collective: every rank calls it; `needles` is only read on rank `src`.
        Returns (start, len) numpy arrays for the whole batch on every rank."""
        if self._text.size == 0:
            raise IndexError("index out of bounds: the len is 0 but the index is 0")  # sacabase lib.rs:89-91
        a, b = self._split_and_gather(needles, src, "lsm")
        return a.astype(np.uint64), b.astype(np.uint32)

    def search_all_batch(self, needles, src: int = 0):
        """sa_search for every needle (utils.c:258-325): -> (left, count) int32 arrays, on every rank."""
        a, b = self._split_and_gather(needles, src, "search_all")
        return a.astype(np.int32), b.astype(np.int32)

    def _split_and_gather(self, needles, src: int, what: str):
        import torch

        dev = torch.device("cuda", self._device) if torch.cuda.is_available() else torch.device("cpu")
        t_pat = t_off = None
        if self.rank == src:
            flat, off = N.pack_patterns(needles)
            t_off = torch.from_numpy(off.astype(np.int64)).to(dev)
            t_pat = torch.from_numpy(flat if flat.size else np.zeros(1, np.uint8)).to(dev)
        a, b = self.query_device(t_pat, t_off, what, src)
        return a.cpu().numpy(), b.cpu().numpy()

    def query_device(self, t_pat, t_off, what: str = "lsm", src: int = 0):
        """One query batch on device tensors (collective).  On rank `src`: `t_pat` uint8 pattern
        bytes and `t_off` int64 offsets [Q + 1] on this rank's GPU; ignored elsewhere.  Rank r answers
        needles [r * ceil(Q / W), (r + 1) * ceil(Q / W)) against its copy of the index, so it is sent
        exactly those: a small header (batch size, longest needle, the byte range of every rank's slice)
        is broadcast, then `src` sends every rank its offsets and pattern bytes point to point (one
        grouped NCCL launch; over NVSwitch the W - 1 transfers run side by side, and 1 / W of the batch
        travels to each GPU instead of all of it), and one all-gather per result array puts the answers
        together on every rank.
        what = "lsm" -> (start int64 [Q], len int32 [Q]); "search_all" -> (left int32, count int32)."""
        import torch

        dist = self._dist
        W = self.world
        dev = torch.device("cuda", self._device) if torch.cuda.is_available() else torch.device("cpu")
        first = torch.int64 if what == "lsm" else torch.int32
        if W == 1:
            q = t_off.numel() - 1
            t_a = self._bufs.get("a_" + what, max(q, 1), first, dev).zero_()
            t_b = self._bufs.get("b_" + what, max(q, 1), torch.int32, dev).zero_()
            if q > 0:
                self._answer(what, t_pat, t_off, q, t_a, t_b, dev, int((t_off[1:] - t_off[:-1]).max()))
            return t_a[:q], t_b[:q]
        # ---- header: q, longest needle, byte offset of the first needle of every (rank, chunk) piece (+ the end) ----
        CH = self.CHUNKS

        def plan(q):
            return slice_plan(q, W, CH, self.CHUNK_MIN)

        if self.rank == src:
            q = t_off.numel() - 1
            _, starts = plan(q)
            cuts = torch.tensor(starts, dtype=torch.int64, device=dev)
            longest = (t_off[1:] - t_off[:-1]).max().view(1) if q else torch.zeros(1, dtype=torch.int64, device=dev)
            hdr = torch.cat([torch.tensor([q], dtype=torch.int64, device=dev), longest, t_off[cuts]])
        else:
            hdr = torch.zeros(W * CH + 3, dtype=torch.int64, device=dev)
        dist.broadcast(hdr, src=src, group=self._group)
        h_host = hdr.tolist()  # the one host round trip of a batch: buffer sizes depend on it
        q, max_len, cut_bytes = int(h_host[0]), int(h_host[1]), [int(x) for x in h_host[2:]]
        per, starts = plan(q)
        mine = range(self.rank * CH, (self.rank + 1) * CH)
        lo, hi = starts[mine[0]], starts[mine[-1] + 1]
        blo, bhi = cut_bytes[mine[0]], cut_bytes[mine[-1] + 1]
        # ---- every rank gets its slice of the offsets and of the pattern bytes, chunk after chunk: one grouped NCCL
        # launch per chunk, all issued now; the search of chunk k starts when chunk k has arrived, while the later
        # chunks are still on the wire ---------------------------------------------------------------------------
        if self.rank == src:
            my_off = t_off[lo:hi + 1]
            my_pat = t_pat[blo:bhi] if bhi > blo else t_pat[:1]
        else:
            my_off = self._bufs.get("off", max(1, hi - lo), torch.int64, dev)
            my_pat = self._bufs.get("pat", max(1, bhi - blo), torch.uint8, dev)
            peer_src = dist.get_global_rank(self._group, src) if self._group is not None else src
        arrivals = []
        for k in range(CH):
            ops = []
            if self.rank == src:
                for r in range(W):
                    p0, p1 = starts[r * CH + k], starts[r * CH + k + 1]
                    if r == src or p1 == p0:
                        continue
                    peer = dist.get_global_rank(self._group, r) if self._group is not None else r
                    # piece (r, k): the start offsets of its needles (the end offset of a piece is in the header)
                    ops.append(dist.P2POp(dist.isend, t_off[p0:p1], peer, self._group))
                    c0, c1 = cut_bytes[r * CH + k], cut_bytes[r * CH + k + 1]
                    if c1 > c0:
                        ops.append(dist.P2POp(dist.isend, t_pat[c0:c1], peer, self._group))
            else:
                p0, p1 = starts[mine[0] + k], starts[mine[0] + k + 1]
                if p1 > p0:
                    ops.append(dist.P2POp(dist.irecv, my_off[p0 - lo:p1 - lo], peer_src, self._group))
                    c0, c1 = cut_bytes[mine[0] + k], cut_bytes[mine[0] + k + 1]
                    if c1 > c0:
                        ops.append(dist.P2POp(dist.irecv, my_pat[c0 - blo:c1 - blo], peer_src, self._group))
            arrivals.append(dist.batch_isend_irecv(ops) if ops else [])
        t_a = self._bufs.get("a_" + what, max(per, 1), first, dev).zero_()
        t_b = self._bufs.get("b_" + what, max(per, 1), torch.int32, dev).zero_()
        rel = self._bufs.get("rel", hi - lo + 1, torch.int64, dev)
        for k in range(CH):
            for req in arrivals[k]:
                req.wait()
            p0, p1 = starts[mine[0] + k], starts[mine[0] + k + 1]
            if p1 == p0:
                continue
            # offsets into my slice of the pattern bytes; the end offset of the piece comes from the header
            piece = rel[p0 - lo:p1 - lo + 1]
            torch.sub(my_off[p0 - lo:p1 - lo], blo, out=piece[:-1])
            piece[-1:].fill_(cut_bytes[mine[0] + k + 1] - blo)
            self._answer(what, my_pat, piece, p1 - p0, t_a[p0 - lo:p1 - lo], t_b[p0 - lo:p1 - lo], dev, max_len)
        g_a = self._bufs.get("ga_" + what, W * max(per, 1), first, dev)
        g_b = self._bufs.get("gb_" + what, W * max(per, 1), torch.int32, dev)
        dist.all_gather_into_tensor(g_a, t_a, group=self._group)
        dist.all_gather_into_tensor(g_b, t_b, group=self._group)
        return g_a[:q], g_b[:q]  # views of buffers that the next batch of the same kind overwrites

    def _answer(self, what, t_pat, t_off, q, t_a, t_b, dev, max_len):
        if what == "lsm":
            self._answer_local(t_pat, t_off, q, t_a, t_b, dev, max_len)
        else:
            self._answer_local_search_all(t_pat, t_off, q, t_a, t_b, dev, max_len)

    def longest_substring_match(self, needle) -> LongestCommonSubstring:
        s, l = self.longest_substring_match_batch([needle])
        return LongestCommonSubstring(self._text, int(s[0]), int(l[0]))

    def close(self):
        if self._h is not None:
            self._destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
