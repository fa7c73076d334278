#!/bin/bash
# round-2 GPU batch y: round 0 writes the sorted suffixes straight into the SA: parity, timings.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_build.py tests/test_gpu_search.py tests/test_gpu_host_cpp.py tests/test_gpu_lcp.py -x -q -k "not largest" > $O/r2y_pytest.log 2>&1; echo "rc=$?" >> $O/r2y_pytest.log); echo "pytest: $(tail -2 $O/r2y_pytest.log | tr '\n' ' ')"
timeout 300 python tools/stress.py 40 > $O/r2y_stress.log 2>&1; tail -2 $O/r2y_stress.log
for w in acgt_4M rand_256M acgt_512M rep_1G; do timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2y_bench_$w.json 2>/dev/null; python tools/show_bench.py $O/r2y_bench_$w.json 2>/dev/null | head -2; done
