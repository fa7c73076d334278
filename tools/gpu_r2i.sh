#!/bin/bash
# round-2 GPU batch i: windowed look-back of the pass kernel (GSA_PASS_CFG 20..26): parity, then A/B timings.
cd "$(dirname "$0")/.."
O=gpurun_out
run_build_tests() {  # $1 = tag
  (timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2i_pytest_$1.log 2>&1; echo "rc=$?" >> $O/r2i_pytest_$1.log)
  echo "$1: $(tail -2 $O/r2i_pytest_$1.log | tr '\n' ' ')"
}
for c in ${PARITY_CFGS:-20 23 24}; do GSA_PASS_CFG=$c run_build_tests cfg$c; done
show() {
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print(sys.argv[2], "ms/step %.2f  pass frac %.3f (%.0f GB/s) share %.3f"%(d['ms_per_step'], r['frac'], r['achieved'], r['share_of_step']), " rounds ms:", [round(x['ms_total'],1) for x in d['rounds']])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for w in ${WORKLOADS:-rep_1G rand_256M acgt_512M}; do
  for cfg in ${BENCH_CFGS:-0 20 21 22 23 24 25 26}; do
    GSA_PASS_CFG=$cfg timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2i_bench_${w}_cfg$cfg.json 2> $O/r2i_bench_${w}_cfg$cfg.err
    show $O/r2i_bench_${w}_cfg$cfg.json "$w cfg$cfg"
  done
done
