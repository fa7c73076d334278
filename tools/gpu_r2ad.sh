#!/bin/bash
# round-2 GPU batch ad: the "too many to list" flag is set once per block (and only while it is zero): parity, rep_1G / rand_256M timings.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 600 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2ad_pytest.log 2>&1; echo "rc=$?" >> $O/r2ad_pytest.log); echo "pytest: $(tail -2 $O/r2ad_pytest.log | tr '\n' ' ')"
for w in rep_1G rand_256M; do timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2ad_bench_$w.json 2>/dev/null; python tools/show_bench.py $O/r2ad_bench_$w.json 2>/dev/null | head -2; done
