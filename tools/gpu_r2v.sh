#!/bin/bash
# round-2 GPU batch v: sparse-mode survivor list + fewer host round trips below 8 MiB: parity, then timings.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not largest and not rep_1G" > $O/r2v_pytest.log 2>&1; echo "rc=$?" >> $O/r2v_pytest.log); echo "pytest: $(tail -2 $O/r2v_pytest.log | tr '\n' ' ')"
show() {
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print(sys.argv[2], "ms/step %.3f  pass frac %.3f  launches %s"%(d['ms_per_step'], r['frac'], d.get('gpu_launches')), " rounds ms:", [round(x['ms_total'],2) for x in d['rounds']])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b() {  # $1 = workload, $2 = tag, rest = env
  w=$1; tag=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2v_bench_${w}_$tag.json 2> $O/r2v_bench_${w}_$tag.err
  show $O/r2v_bench_${w}_$tag.json "$w $tag"
}
for w in acgt_4M rand_256M acgt_512M; do
  b $w default GSA_X=1
  b $w nolist GSA_NO_SURV_LIST=1
done
b rep_1G default GSA_X=1
