#!/bin/bash
# round-2 GPU batch c: prefilter + persistent bulk-async pass, parity then A/B timings.
cd "$(dirname "$0")/.."
O=gpurun_out
run_build_tests() {  # $1 = tag
  (timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2c_pytest_$1.log 2>&1; echo "rc=$?" >> $O/r2c_pytest_$1.log)
  echo "$1: $(tail -2 $O/r2c_pytest_$1.log | tr '\n' ' ')"
}
run_build_tests default
GSA_PASS_CFG=10 run_build_tests cfg10
GSA_PASS_CFG=13 run_build_tests cfg13
b() {  # $1 = tag, rest = env assignments
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build > $O/r2c_bench_$tag.json 2> $O/r2c_bench_$tag.err
  python - "$O/r2c_bench_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print(sys.argv[2], "ms/step %.2f  pass frac %.3f (%.0f GB/s) share %.3f"%(d['ms_per_step'], r['frac'], r['achieved'], r['share_of_step']), " rounds ms:", [round(x['ms_total'],1) for x in d['rounds']])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b base GSA_NO_PREFILTER=1
b prefilter GSA_X=1
b pf_cfg10 GSA_PASS_CFG=10
b pf_cfg11 GSA_PASS_CFG=11
b pf_cfg12 GSA_PASS_CFG=12
b pf_cfg13 GSA_PASS_CFG=13
for w in rand_256M acgt_512M; do
  for cfg in 0 10 13; do
    GSA_PASS_CFG=$cfg timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2c_bench_${w}_cfg$cfg.json 2>/dev/null
    python - "$O/r2c_bench_${w}_cfg$cfg.json" "$w cfg$cfg" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print(sys.argv[2], "ms/step %.2f  pass frac %.3f (%.0f GB/s)"%(d['ms_per_step'], r['frac'], r['achieved']))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  done
done
timeout 600 python tools/shapes_bench.py 256 > $O/r2c_shapes.txt 2>&1; tail -15 $O/r2c_shapes.txt
