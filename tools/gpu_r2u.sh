#!/bin/bash
# round-2 GPU batch u: tile shapes of the pass kernel again, now with the L2 prefetch and the cheaper ballot ranking.
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
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2u_bench_${w}_$tag.json 2> $O/r2u_bench_${w}_$tag.err
  show $O/r2u_bench_${w}_$tag.json "$w $tag"
}
for w in rep_1G rand_256M; do
  b $w default GSA_X=1
  b $w cfg1_pf592 GSA_PASS_CFG=1 GSA_PASS_PF=592
  b $w cfg2_pf296 GSA_PASS_CFG=2 GSA_PASS_PF=296
  b $w cfg3_pf296 GSA_PASS_CFG=3 GSA_PASS_PF=296
  b $w pf296 GSA_PASS_PF=296
  b $w pf148 GSA_PASS_PF=148
done
