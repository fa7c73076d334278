"""One process per GPU over NCCL (needs >= 2 GPUs; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("P", [2, 5, 8])
def test_distributed_partitioned_query_nccl(P):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + P), os.path.join(ROOT, "tests", "dist_worker.py"), str(P)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_replicated_index_queries_nccl():
    """ReplicatedSuffixArray over NCCL: needles split across the ranks, answers of one index."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29671", os.path.join(ROOT, "tests", "dist_worker.py"), "replicated"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_bench_two_ranks_small():
    """bench.py under torchrun on 2 GPUs (tiny workload): the JSON line carries the multi-GPU fields."""
    import json

    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3",
           "--workload", "rep_16M", "--no-queries", "--no-cpu-baseline"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["value"] > 0 and d["e2e"]["value"] > 0
