#!/bin/bash
# round-2 GPU batch x: vector loads in k_pack, one packed word per thread in k_hist0_windows: parity (build, BWT, LCP), timings.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_lcp.py -x -q -k "not largest and not rep_1G" > $O/r2z_pytest.log 2>&1; echo "rc=$?" >> $O/r2z_pytest.log); echo "pytest: $(tail -2 $O/r2z_pytest.log | tr '\n' ' ')"
for w in acgt_4M rand_256M acgt_512M rep_1G; do timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2z_bench_$w.json 2>/dev/null; python tools/show_bench.py $O/r2z_bench_$w.json 2>/dev/null | head -2; done
