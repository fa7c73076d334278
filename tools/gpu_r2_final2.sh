#!/bin/bash
# round-2 final validation: full GPU test suite, default bench, smoke.
cd "$(dirname "$0")/.."
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2f3_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r2f3_pytest_gpu.log); tail -3 $O/r2f3_pytest_gpu.log
( time timeout 600 python bench.py > $O/r2f3_bench.json 2> $O/r2f3_bench.err ) 2>&1 | grep real; python tools/show_bench.py $O/r2f3_bench.json 2>/dev/null | head -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2f3_smoke.log
