#!/bin/bash
# round-2 multi-GPU batch ab (gpurun --gpus N): bench.py under torchrun (parity gate, part_4G, replicated queries with sliced transfers).
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-4}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > $O/r2ab_bench_${N}gpu.json 2> $O/r2ab_bench_${N}gpu.err; echo "bench N=$N rc=$?" ) 2>&1 | tail -4
python - "$O/r2ab_bench_${N}gpu.json" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print("value", round(d['value']), "MB/s  e2e", round(d['e2e']['value']), " ms/step", round(d['ms_per_step'],1))
    for k in ('part_4G','queries','parity_gate'):
        v=d.get(k)
        if not v: print(k,'MISSING'); continue
        if 'error' in v: print(k,'ERROR',v['error'][:300]); continue
        if k=='part_4G': print(k,'build MB/s',round(v['build']['value']),'device',round(v['build']['device_only']['value']),'query q/s',round(v['query']['queries_per_s']/1e6),'M')
        if k=='queries': print(k,'lsm',round(v['longest_substring_match']['queries_per_s']/1e6),'M/s',round(v['longest_substring_match']['ms'],3),'ms  search_all',round(v['search_all']['queries_per_s']/1e6),'M/s', v.get('checked','')[:90])
        if k=='parity_gate': print(k, v.get('result'))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace('.json','.err')).read()[-2000:])
PY
