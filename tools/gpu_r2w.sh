#!/bin/bash
# round-2 GPU batch w: full GPU test suite, default bench, ncu captures of the pass kernel (skewed and uniform digits), launch list with DRAM bytes.
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
(timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2w_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r2w_pytest_gpu.log); tail -3 $O/r2w_pytest_gpu.log
timeout 600 python bench.py > $O/r2w_bench.json 2> $O/r2w_bench.err; python tools/show_bench.py $O/r2w_bench.json 2>/dev/null | head -1
for w in acgt_4M rand_256M acgt_512M; do timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2w_bench_$w.json 2>/dev/null; python tools/show_bench.py $O/r2w_bench_$w.json 2>/dev/null | head -1; done
timeout 600 $NCU -k regex:'^k_radix_pass$' -s 1 -c 1 -o $O/r2w_pass_rep python bench.py --steps 1 --warmup 1 --only-build --workload rep_256M > $O/r2w_ncu1.log 2>&1
timeout 600 $NCU -k regex:'^k_radix_pass$' -s 1 -c 1 -o $O/r2w_pass_rand python bench.py --steps 1 --warmup 1 --only-build --workload rand_256M > $O/r2w_ncu2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2w_launches_rep1G.csv python bench.py --steps 1 --warmup 1 --only-build > $O/r2w_ncu3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2w_launches_rand256M.csv python bench.py --steps 1 --warmup 1 --only-build --workload rand_256M > $O/r2w_ncu4.log 2>&1
ls -la $O/r2w*
