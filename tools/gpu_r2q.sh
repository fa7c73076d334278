#!/bin/bash
# round-2 multi-GPU batch q (gpurun --gpus N): replicated queries with scattered needle slices: NCCL tests + phase probe.
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}
(timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > $O/r2q_pytest_${N}gpu.log 2>&1; echo "rc=$?" >> $O/r2q_pytest_${N}gpu.log); tail -3 $O/r2q_pytest_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dist_query_probe.py 1024 10000000 > $O/r2q_probe_${N}gpu.json 2> $O/r2q_probe_${N}gpu.err; tail -2 $O/r2q_probe_${N}gpu.json; tail -3 $O/r2q_probe_${N}gpu.err
