#!/bin/bash
# round-2 GPU batch d: prefilter v2 (shared-memory group cache), merged sync points, small sort; f4 measurement; ncu of the bulk-async pass.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_search.py -x -q -k "not full_size and not largest" > $O/r2d_pytest.log 2>&1; echo "rc=$?" >> $O/r2d_pytest.log)
echo "pytest: $(tail -2 $O/r2d_pytest.log | tr '\n' ' ')"
timeout 200 python tools/stress.py 90 11 > $O/r2d_stress.log 2>&1; tail -2 $O/r2d_stress.log
b() {  # $1 = tag, $2 = workload, rest = env assignments
  tag=$1; w=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2d_bench_$tag.json 2> $O/r2d_bench_$tag.err
  python - "$O/r2d_bench_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print(sys.argv[2], "ms/step %.3f  pass frac %.3f share %.3f"%(d['ms_per_step'], r['frac'], r['share_of_step']), " rounds ms:", [round(x['ms_total'],1) for x in d['rounds']][:14])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b rep1G_nopf rep_1G GSA_NO_PREFILTER=1
b rep1G_pf rep_1G GSA_X=1
b acgt4M acgt_4M GSA_X=1
b acgt4M_nosmall acgt_4M GSA_NO_SMALL_SORT=1
b rand256M rand_256M GSA_X=1
b acgt512M acgt_512M GSA_X=1
b rep64M rep_64M GSA_X=1
timeout 600 python tools/shapes_bench.py 256 > $O/r2d_shapes.txt 2>&1; tail -13 $O/r2d_shapes.txt
timeout 900 python tools/f4_measure.py 1024 256 > $O/r2d_f4.json 2> $O/r2d_f4.err; tail -c 400 $O/r2d_f4.err; head -c 1500 $O/r2d_f4.json
# ncu: why does the bulk-async persistent pass lose?  one capture each, same pass of rep_256M's round 0
GSA_PASS_CFG=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_radix_pass_p" -s 9 -c 1 -f -o $O/r2d_pass_p_cfg10 python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2d_ncu_passp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_radix_pass<" -s 9 -c 1 -f -o $O/r2d_pass_classic python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2d_ncu_passc.log 2>&1
# ncu: the pre-filter and the gather behind it, a dense round of rep_256M
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_prefilter|k_gather" -s 20 -c 4 -f -o $O/r2d_prefilter python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2d_ncu_pf.log 2>&1
ls -la $O/*.ncu-rep | tail -5
