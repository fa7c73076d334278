#!/bin/bash
# round-2 multi-GPU batch (run with gpurun --gpus N): bench.py under torchrun with the parity gate, the NCCL tests, the reference arm.
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/r2j_bench_${N}gpu.json 2> $O/r2j_bench_${N}gpu.err; echo "bench N=$N rc=$?") 2>&1 | tail -5
grep -c "NCCL INFO" $O/r2j_bench_${N}gpu.err; grep -m3 "nranks\|NVLS\|P2P" $O/r2j_bench_${N}gpu.err | cut -c1-200
python - "$O/r2j_bench_${N}gpu.json" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print("value", round(d['value']), "MB/s  e2e", round(d['e2e']['value']), " ms/step", round(d['ms_per_step'],1))
    for k in ('part_4G','queries','parity_gate'):
        v=d.get(k)
        if not v: print(k,'MISSING'); continue
        if 'error' in v: print(k,'ERROR',v['error'][:300]); continue
        if k=='part_4G': print(k,'build MB/s',round(v['build']['value']),'device',round(v['build']['device_only']['value']),'query q/s',round(v['query']['queries_per_s']/1e6),'M', v['query'].get('checked','')[:80])
        if k=='queries': print(k,'lsm',round(v['longest_substring_match']['queries_per_s']/1e6),'M/s search_all',round(v['search_all']['queries_per_s']/1e6),'M/s', v.get('checked','')[:80])
        if k=='parity_gate': print(k, v)
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
(timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_search.py -x -q -k "nccl or multi_device or two_ranks" > $O/r2j_pytest_${N}gpu.log 2>&1; echo "rc=$?" >> $O/r2j_pytest_${N}gpu.log); tail -3 $O/r2j_pytest_${N}gpu.log
