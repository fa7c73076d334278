#!/bin/bash
# round-2 GPU batch e: prefilter v3, narrow round-0 keys, chunked query API; ncu of both pass kernels.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_search.py -x -q -k "not full_size and not largest" > $O/r2e_pytest.log 2>&1; echo "rc=$?" >> $O/r2e_pytest.log)
echo "pytest: $(tail -2 $O/r2e_pytest.log | tr '\n' ' ')"
timeout 200 python tools/stress.py 60 12 > $O/r2e_stress.log 2>&1; tail -1 $O/r2e_stress.log
GSA_NO_SMALL_SORT=1 timeout 200 python tools/stress.py 30 13 > $O/r2e_stress_nosmall.log 2>&1; tail -1 $O/r2e_stress_nosmall.log
b() {  # $1 = tag, $2 = workload, rest = env assignments
  tag=$1; w=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2e_bench_$tag.json 2> $O/r2e_bench_$tag.err
  python - "$O/r2e_bench_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print(sys.argv[2], "ms/step %.3f  pass frac %.3f share %.3f launches/step %d"%(d['ms_per_step'], r['frac'], r['share_of_step'], d['gpu_launches']/d['steps']), " rounds ms:", [round(x['ms_total'],1) for x in d['rounds']][:14])
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
}
b rep1G_base rep_1G GSA_NO_PREFILTER=1 GSA_NO_NARROW=1
b rep1G_narrow rep_1G GSA_NO_PREFILTER=1
b rep1G_pf rep_1G GSA_NO_NARROW=1
b rep1G_default rep_1G GSA_X=1
b acgt4M acgt_4M GSA_X=1
b rand256M rand_256M GSA_X=1
b rep64M rep_64M GSA_X=1
timeout 600 python tools/shapes_bench.py 256 > $O/r2e_shapes.txt 2>&1; tail -13 $O/r2e_shapes.txt
(timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2e_bench_full.json 2> $O/r2e_bench_full.err; echo "full bench rc=$?")
GSA_NO_NARROW=1 GSA_PASS_CFG=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_radix_pass_p<512, 8, 0" -s 0 -c 1 -f -o $O/r2e_pass_p_cfg10 python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2e_ncu_passp.log 2>&1
GSA_NO_NARROW=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_radix_pass<256, 16, 0" -s 0 -c 1 -f -o $O/r2e_pass_classic python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2e_ncu_passc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_prefilter" -s 2 -c 1 -f -o $O/r2e_prefilter python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2e_ncu_pf.log 2>&1
ls -la $O/r2e*.ncu-rep
