#!/bin/bash
# round-2 GPU batch aa: values loaded after the ranking (16 registers less during it), 3 and 4 CTAs/SM.
cd "$(dirname "$0")/.."
O=gpurun_out
show() {
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print(sys.argv[2], "ms/step %.2f  pass frac %.3f (%.0f GB/s) share %.3f"%(d['ms_per_step'], r['frac'], r['achieved'], r['share_of_step']))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b() {  # $1 = workload, $2 = tag, rest = env
  w=$1; tag=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2aa_bench_${w}_$tag.json 2> $O/r2aa_bench_${w}_$tag.err
  show $O/r2aa_bench_${w}_$tag.json "$w $tag"
}
cp stringsearch_b200/libgsa.so /tmp/libgsa_default.so
for v in default lv3 lv4; do
  [ $v != default ] && cp stringsearch_b200/libgsa_$v.so stringsearch_b200/libgsa.so
  if [ $v != default ]; then (timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2aa_pytest_$v.log 2>&1; echo "rc=$?" >> $O/r2aa_pytest_$v.log); echo "$v: $(tail -2 $O/r2aa_pytest_$v.log | tr '\n' ' ')"; fi
  for w in rep_1G rand_256M; do
    if [ $v = lv4 ]; then b $w $v GSA_PASS_PF=592; else b $w $v GSA_X=1; fi
  done
done
cp /tmp/libgsa_default.so stringsearch_b200/libgsa.so
