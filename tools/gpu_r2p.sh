#!/bin/bash
# round-2 GPU batch m: R2P ballot peers (default lib) and wide far steps of the look-back (libgsa_w8 / _w16).
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2p_pytest.log 2>&1; echo "rc=$?" >> $O/r2p_pytest.log); echo "default: $(tail -2 $O/r2p_pytest.log | tr '\n' ' ')"
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
b() {  # $1 = workload, $2 = tag, rest = env
  w=$1; tag=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2p_bench_${w}_$tag.json 2> $O/r2p_bench_${w}_$tag.err
  show $O/r2p_bench_${w}_$tag.json "$w $tag"
}
cp stringsearch_b200/libgsa.so /tmp/libgsa_default.so
for v in default q0 q2; do
  [ $v != default ] && cp stringsearch_b200/libgsa_$v.so stringsearch_b200/libgsa.so
  if [ $v != default ]; then (timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2p_pytest_$v.log 2>&1; echo "rc=$?" >> $O/r2p_pytest_$v.log); echo "$v: $(tail -2 $O/r2p_pytest_$v.log | tr '\n' ' ')"; fi
  for w in rep_1G rand_256M acgt_512M; do b $w $v GSA_X=1; done
done
cp /tmp/libgsa_default.so stringsearch_b200/libgsa.so
