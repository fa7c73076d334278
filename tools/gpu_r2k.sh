#!/bin/bash
# round-2 GPU batch k: plain LDS/STS claim vs shared atomics in the ranking, L2 bulk prefetch, scanner-based prefixes.
cd "$(dirname "$0")/.."
O=gpurun_out
run_build_tests() {  # $1 = tag
  (timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2k_pytest_$1.log 2>&1; echo "rc=$?" >> $O/r2k_pytest_$1.log)
  echo "$1: $(tail -2 $O/r2k_pytest_$1.log | tr '\n' ' ')"
}
run_build_tests plain
GSA_PASS_CFG=30 run_build_tests cfg30
GSA_PASS_CFG=31 GSA_PASS_PF=444 run_build_tests cfg31_pf
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
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2k_bench_${w}_$tag.json 2> $O/r2k_bench_${w}_$tag.err
  show $O/r2k_bench_${w}_$tag.json "$w $tag"
}
for w in rep_1G rand_256M acgt_512M; do
  b $w plain GSA_X=1
  b $w pf444 GSA_PASS_PF=444
  b $w pf888 GSA_PASS_PF=888
  b $w pf222 GSA_PASS_PF=222
  b $w scan GSA_PASS_CFG=30
  b $w scan_pf444 GSA_PASS_CFG=30 GSA_PASS_PF=444
  b $w scan12_pf GSA_PASS_CFG=31 GSA_PASS_PF=592
done
cp stringsearch_b200/libgsa.so /tmp/libgsa_plain.so; cp stringsearch_b200/libgsa_atoms.so stringsearch_b200/libgsa.so
for w in rep_1G rand_256M acgt_512M; do
  b $w atoms GSA_X=1
  b $w atoms_scan GSA_PASS_CFG=30
done
cp /tmp/libgsa_plain.so stringsearch_b200/libgsa.so
