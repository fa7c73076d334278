#!/bin/bash
# round-2 GPU batch l: default bench with the L2 prefetch, ncu of the pass kernel on uniform digits (rand_256M) and on rep_256M.
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 python bench.py > $O/r2l_bench.json 2> $O/r2l_bench.err; python tools/show_bench.py $O/r2l_bench.json | head -3
timeout 600 $NCU -k regex:'^k_radix_pass$' -s 1 -c 1 -o $O/r2l_pass_rand python bench.py --steps 1 --warmup 1 --only-build --workload rand_256M > $O/r2l_ncu1.log 2>&1
timeout 600 $NCU -k regex:'^k_radix_pass$' -s 1 -c 1 -o $O/r2l_pass_rep python bench.py --steps 1 --warmup 1 --only-build --workload rep_256M > $O/r2l_ncu2.log 2>&1
ls -la $O/r2l*.ncu-rep
