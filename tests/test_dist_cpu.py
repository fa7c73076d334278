"""world_size-2 gloo test of the multi-GPU query plumbing (broadcast -> per-rank shards ->
all_gather -> merge) on CPU.  The two device-touching steps are replaced by the oracle so
that only the host-side collective logic is under test."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port_no, text_bytes, P, needles, expect, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from stringsearch_b200.sacapart import DistributedPartitionedSuffixArray

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = oracle.port()
    text = np.frombuffer(text_bytes, np.uint8)

    class OracleShards(DistributedPartitionedSuffixArray):
        def _build_shard(self, off, ln):
            return (off, ln, port.sa_build(self._text[off:off + ln]))

        def _destroy_shard(self, h):
            pass

        def _answer_local(self, t_pat, t_off, qn, t_start, t_len, dev, max_len=0):
            pats = t_pat.numpy()
            off = t_off.numpy()
            for k, (_, offset, (o, ln, sa)) in enumerate(self._shards):
                for j in range(qn):
                    nd = pats[off[j]:off[j + 1]].tobytes()
                    s, l = port.longest_substring_match(self._text[o:o + ln], sa, nd)
                    if s + l == ln:  # may_extend, sacapart lib.rs:77-84
                        l = int(port.lib.oracle_common_prefix_len(text_bytes[o + s:], len(text_bytes) - o - s, nd, len(nd)))
                    if k == 0 or l > int(t_len[j]):
                        t_start[j] = s + offset
                        t_len[j] = l

    try:
        psa = OracleShards(text, P, device=0)
        assert psa.local_partitions() == list(range(rank, psa.num_partitions(), world))
        s, l = psa.longest_substring_match_batch(needles if rank == 0 else None)
        got = list(zip(s.tolist(), l.tolist()))
        q.put((rank, got == expect, got))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [1, 2, 3, 5])
def test_distributed_query_plumbing_gloo(P, port):
    text = ("This is a rather long text. We can probably find matches that span two partitions. Oh yes. " * 3).encode()
    needles = [b"rather long", b"text. We can", b"We can probably find matches that span", b"zzz", b"Oh yes. This", b""]
    ps, sas = port.part_build(text, P)
    expect = [port.part_lsm(text, ps, sas, nd) for nd in needles]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 2000) + P
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, text, P, needles, expect, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, got in res:
        assert ok, (rank, got, expect)


def _worker_replicated(rank, world, port_no, text_bytes, needles, expect, q, chunk_min=1 << 16):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle
    from stringsearch_b200.sacapart import ReplicatedSuffixArray

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = oracle.port()
    text = np.frombuffer(text_bytes, np.uint8)
    answered = []

    class OracleReplica(ReplicatedSuffixArray):
        CHUNK_MIN = chunk_min  # 1: every rank's slice travels (and is answered) in up to CHUNKS pieces

        def _build(self):
            return port.sa_build(self._text)

        def _destroy(self, h):
            pass

        def _answer_local(self, t_pat, t_off, qn, t_start, t_len, dev, max_len=0):
            pats, off = t_pat.numpy(), t_off.numpy()
            answered.append(qn)
            for j in range(qn):
                s, l = port.longest_substring_match(self._text, self._h, pats[off[j]:off[j + 1]].tobytes())
                t_start[j], t_len[j] = s, l

    try:
        rsa = OracleReplica(text, device=0)
        s, l = rsa.longest_substring_match_batch(needles if rank == 0 else None)
        got = list(zip(s.tolist(), l.tolist()))
        q.put((rank, got == expect and sum(answered) <= (len(needles) + world - 1) // world, got))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nq,chunk_min,world", [(0, 1 << 16, 2), (1, 1 << 16, 2), (2, 1, 2), (5, 1 << 16, 2), (6, 1 << 16, 2), (6, 1, 2),
                                                (19, 1, 2), (23, 1, 3), (7, 1 << 16, 3)])
def test_replicated_query_plumbing_gloo(nq, chunk_min, world, port):
    """ReplicatedSuffixArray: the batch is split across the ranks (each answers only its slice) and
    the gathered answers are those of one un-partitioned index, in order."""
    text = ("This is a rather long text. We can probably find matches that span two partitions. Oh yes. " * 2).encode()
    needles = ([b"rather long", b"text. We can", b"zzz", b"Oh yes. This", b"", b"We can probably"] * 4)[:nq]
    port_base = 0
    sa = port.sa_build(text)
    expect = [tuple(port.longest_substring_match(text, sa, nd)) for nd in needles]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 31500 + (os.getpid() % 2000) + nq + (30 if chunk_min == 1 else 0) + 60 * (world - 2)
    procs = [ctx.Process(target=_worker_replicated, args=(r, world, port_no, text, needles, expect, q, chunk_min)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, got in res:
        assert ok, (rank, got, expect)
