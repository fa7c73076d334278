#!/bin/bash
# round-2 GPU batch g: ncu captures (pass kernels, search_all, rebuild / bag kernels, LCP, inverse BWT), final A/B of the pre-filter.
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
(timeout 600 python -m pytest tests/test_gpu_build.py -x -q -k "not full_size and not largest" > $O/r2g_pytest.log 2>&1; echo "rc=$?" >> $O/r2g_pytest.log); echo "pytest: $(tail -2 $O/r2g_pytest.log | tr '\n' ' ')"
for t in nopf:GSA_NO_PREFILTER=1 default:GSA_X=1; do
  tag=${t%%:*}; envs=${t#*:}
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --only-build > $O/r2g_bench_$tag.json 2> $O/r2g_bench_$tag.err
  python - "$O/r2g_bench_$tag.json" "$tag" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[2], "ms/step %.3f pass frac %.3f"%(d['ms_per_step'], d['roofline']['frac']), [round(x['ms_total'],1) for x in d['rounds']])
PY
done
timeout 300 python tools/extras_bench.py 256 > $O/r2g_extras.json 2> $O/r2g_extras.err; cat $O/r2g_extras.json
# pass kernels: the second launch of each = first non-generating pass of round 0 (2^28 elements, rep_256M)
GSA_PASS_CFG=10 timeout 600 $NCU -k regex:'^k_radix_pass_p$' -s 1 -c 1 -o $O/r2g_pass_p_cfg10 python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2g_ncu1.log 2>&1
timeout 600 $NCU -k regex:'^k_radix_pass$' -s 1 -c 1 -o $O/r2g_pass_classic python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2g_ncu2.log 2>&1
# the h = 2048 round of rep_256M: 10th launch of each of these kernels
timeout 600 $NCU -k regex:'^k_rebuild$|^k_bag_gather$|^k_bag_refine$|^k_slots$' -s 36 -c 4 -o $O/r2g_round2048 python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2g_ncu3.log 2>&1
timeout 600 $NCU -k regex:'^k_prefilter$|^k_gather$' -s 4 -c 2 -o $O/r2g_prefilter python bench.py --steps 1 --warmup 3 --only-build --workload rep_256M > $O/r2g_ncu4.log 2>&1
timeout 600 $NCU -k regex:'^k_search_all$' -s 1 -c 1 -o $O/r2g_search_all python tools/search_bench.py 1024 10000000 32 > $O/r2g_ncu5.log 2>&1
timeout 600 $NCU -k regex:'^k_irreducible$|^k_long$|^k_lcp_gather$|^k_reach_apply$|^k_ibwt_walk1$|^k_ibwt_walk2$|^k_bwt$' -c 7 -o $O/r2g_extras python tools/extras_bench.py 256 > $O/r2g_ncu6.log 2>&1
ls -la $O/r2g*.ncu-rep
