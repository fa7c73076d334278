#!/bin/bash
# round-2 GPU batch x: vector loads in k_pack, one packed word per thread in k_hist0_windows: parity (build, BWT, LCP), timings.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_build.py -x -q -k "not largest and not rep_1G" > $O/r2ac_pytest.log 2>&1; echo "rc=$?" >> $O/r2ac_pytest.log); echo "pytest: $(tail -2 $O/r2ac_pytest.log | tr '\n' ' ')"
for w in acgt_4M rand_256M acgt_512M; do timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2ac_bench_$w.json 2>/dev/null; python tools/show_bench.py $O/r2ac_bench_$w.json 2>/dev/null | head -2; done
timeout 200 python tools/stress.py 30 > $O/r2ac_stress.log 2>&1; tail -1 $O/r2ac_stress.log
