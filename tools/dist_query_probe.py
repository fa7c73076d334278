"""Phase timing of ReplicatedSuffixArray.query_device under torchrun (diagnostic):
broadcast of offsets / patterns, local kernel, the two all-gathers.  Usage (N ranks):
  python -m torch.distributed.run --nproc-per-node N tools/dist_query_probe.py [MiB=1024] [Q=10000000]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringsearch_b200 import sacapart, synth  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    t = synth.acgt(mib << 20, 5)
    rsa = sacapart.ReplicatedSuffixArray(t, lr)
    t_pat = t_off = None
    if rank == 0:
        flat, off = synth.patterns_from_text(t, Q, 32, 6)
        t_pat = torch.from_numpy(flat.copy()).to(dev)
        t_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    res = {}
    for what in ("lsm", "search_all", "lsm", "search_all"):
        rsa.query_device(t_pat, t_off, what)
        dist.barrier(); torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        for _ in range(3):
            rsa.query_device(t_pat, t_off, what)
        e[1].record()
        torch.cuda.synchronize()
        res.setdefault(what, []).append(e[0].elapsed_time(e[1]) / 3)
    # phases by hand
    q = Q
    per = (q + world - 1) // world
    b_off = t_off if rank == 0 else torch.empty(q + 1, dtype=torch.int64, device=dev)
    b_pat = t_pat if rank == 0 else torch.empty(q * 32, dtype=torch.uint8, device=dev)
    for name, fn in (("bcast_off_80MB", lambda: dist.broadcast(b_off, src=0)), ("bcast_pat_320MB", lambda: dist.broadcast(b_pat, src=0)),
                     ("allgather_i64", lambda: dist.all_gather_into_tensor(torch.empty(world * per, dtype=torch.int64, device=dev), torch.zeros(per, dtype=torch.int64, device=dev))),
                     ("allgather_i32", lambda: dist.all_gather_into_tensor(torch.empty(world * per, dtype=torch.int32, device=dev), torch.zeros(per, dtype=torch.int32, device=dev))),
                     ("maxlen", lambda: int((b_off[1:] - b_off[:-1]).max()))):
        fn(); dist.barrier(); torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        for _ in range(5):
            fn()
        e[1].record()
        torch.cuda.synchronize()
        res[name] = e[0].elapsed_time(e[1]) / 5
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
