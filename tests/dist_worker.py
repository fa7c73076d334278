"""torchrun worker for tests/test_gpu_dist.py: one rank per GPU, NCCL.
Builds a DistributedPartitionedSuffixArray, runs a collective query, and rank 0 compares the
result with the oracle's sacapart restatement."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def replicated(text, rank, world, local):
    """ReplicatedSuffixArray: every rank holds the whole index and answers its slice of the batch."""
    from stringsearch_b200 import sacapart

    rsa = sacapart.ReplicatedSuffixArray(text, device=local)
    rng = np.random.default_rng(6)
    needles = None
    if rank == 0:
        needles = []
        for i in range(5001):  # not a multiple of the world size
            o = int(rng.integers(0, len(text) - 1))
            b = bytearray(text[o:o + int(rng.integers(0, 300))].tobytes())
            if b and i % 3 == 0:
                b[-1] ^= 0x55
            needles.append(bytes(b))
    s, l = rsa.longest_substring_match_batch(needles)
    ok = True
    if rank == 0:
        from oracle import oracle

        port = oracle.port()
        sa = port.sa_build(text)
        es, el = port.lsm_batch(text, sa, needles)
        ok = bool((s == es).all() and (l == el).all())
    left, cnt = rsa.search_all_batch(needles)
    if rank == 0:
        eleft, ecnt = port.search_all_batch(text, sa, needles)
        ok = ok and bool((left == eleft).all() and (cnt == ecnt).all())
    # the same batch with every rank's slice travelling (and answered) in pieces, as large batches do
    rsa.CHUNK_MIN = 100
    s2, l2 = rsa.longest_substring_match_batch(needles)
    left2, cnt2 = rsa.search_all_batch(needles)
    if rank == 0:
        ok = ok and bool((s2 == es).all() and (l2 == el).all() and (left2 == eleft).all() and (cnt2 == ecnt).all())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OK" if ok else "DIST_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


def main():
    from stringsearch_b200 import sacapart, synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    text = synth.repetitive(600_000, 31, period=700, mutation_rate=3e-3)  # same bytes on every rank
    if len(sys.argv) > 1 and sys.argv[1] == "replicated":
        return replicated(text, rank, world, local)
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 2 * world
    psa = sacapart.DistributedPartitionedSuffixArray(text, P, device=local)
    assert psa.local_partitions() == list(range(rank, psa.num_partitions(), world))
    rng = np.random.default_rng(5)
    needles = None
    if rank == 0:
        needles = []
        ps = len(text) // P + 1
        for i in range(3000):
            o = int(rng.integers(0, len(text) - 1))
            m = int(rng.integers(0, 400))
            b = bytearray(text[o:o + m].tobytes())
            if b and i % 3 == 0:
                b[-1] ^= 0x55
            needles.append(bytes(b))
        for i in range(1, P):  # needles straddling partition boundaries (may_extend)
            needles.append(text[i * ps - 7:i * ps + 200].tobytes())
    s, l = psa.longest_substring_match_batch(needles)
    ok = True
    if rank == 0:
        from oracle import oracle

        port = oracle.port()
        ps, sas = port.part_build(text, P)
        es, el = port.part_lsm_batch(text, ps, sas, needles)
        ok = bool((s == es).all() and (l == el).all())
        if not ok:
            bad = np.flatnonzero((s != es) | (l != el))[:5]
            print("MISMATCH at", bad, s[bad], es[bad], l[bad], el[bad], flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OK" if ok else "DIST_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
