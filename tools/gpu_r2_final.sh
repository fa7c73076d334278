#!/bin/bash
# round-2 final GPU batch: full GPU test suite, default bench, shapes, launch lists with DRAM bytes.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2f2_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r2f2_pytest_gpu.log); tail -3 $O/r2f2_pytest_gpu.log
( time timeout 600 python bench.py > $O/r2f2_bench.json 2> $O/r2f2_bench.err ) 2>&1 | grep real; python tools/show_bench.py $O/r2f2_bench.json 2>/dev/null | head -1
for w in acgt_4M rand_256M acgt_512M; do timeout 300 python bench.py --steps 5 --warmup 3 --only-build --workload $w > $O/r2f2_bench_$w.json 2>/dev/null; python tools/show_bench.py $O/r2f2_bench_$w.json 2>/dev/null | head -1; done
timeout 600 python tools/shapes_bench.py 256 > $O/r2f2_shapes.txt 2>&1; tail -22 $O/r2f2_shapes.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2f2_launches_rep1G.csv python bench.py --steps 1 --warmup 1 --only-build > $O/r2f2_ncu3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2f2_launches_rand256M.csv python bench.py --steps 1 --warmup 1 --only-build --workload rand_256M > $O/r2f2_ncu4.log 2>&1
ls $O/r2f2*
