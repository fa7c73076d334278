#!/bin/bash
# round-2 GPU batch b: new search kernels (parity + A/B), bench, ncu captures.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_lcp.py tests/test_gpu_host_cpp.py -x -q > $O/r2b_pytest_search.log 2>&1; echo "rc=$?" >> $O/r2b_pytest_search.log)
tail -3 $O/r2b_pytest_search.log
(timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2b_pytest_build.log 2>&1; echo "rc=$?" >> $O/r2b_pytest_build.log)
tail -3 $O/r2b_pytest_build.log
for bits in default 20 22 26; do
  if [ $bits = default ]; then unset GSA_ACCEL_BITS; else export GSA_ACCEL_BITS=$bits; fi
  timeout 300 python tools/search_bench.py 1024 10000000 32 $([ $bits = default ] && echo 1 || echo 0) > $O/r2b_search_bits_$bits.json 2> $O/r2b_search_bits_$bits.err
done
unset GSA_ACCEL_BITS
GSA_NO_ACCEL=1 timeout 300 python tools/search_bench.py 1024 10000000 32 1 > $O/r2b_search_noaccel.json 2> $O/r2b_search_noaccel.err
cat $O/r2b_search_*.json
(timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?")
# ncu: search kernels, full sets
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lsm|k_search_all" -s 2 -c 2 -f -o $O/r2b_search_accel python tools/search_bench.py 1024 10000000 32 > $O/r2b_ncu_search.log 2>&1
# ncu: launch list of one build with DRAM bytes (measured traffic of the whole build)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2b_launches_rep1G.csv python bench.py --steps 1 --warmup 3 --only-build > $O/r2b_ncu_launch.log 2>&1
tail -2 $O/r2b_ncu_launch.log
ls -la $O | tail -15
