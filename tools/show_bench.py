"""Pretty-print bench.py JSON lines: python tools/show_bench.py gpurun_out/bench_*.json"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "ERR", e)
        continue
    e2e = d.get("e2e") or {}
    rf = d.get("roofline", {})
    print(f"{path}: {d['value']:.1f} {d['unit']}  {d['ms_per_step']:.1f} ms/step  e2e {e2e.get('value')}  "
          f"pass {rf.get('achieved', 0):.0f} GB/s frac {rf.get('frac', 0):.3f} share {rf.get('share_of_step', 0):.2f}  "
          f"whole-build frac {rf.get('whole_build', {}).get('frac_of_peak', 0):.3f}  cpu {d.get('cpu_baseline', {}).get('value')}")
    if "-v" in sys.argv or len(d.get("rounds", [])) > 1:
        for r in d.get("rounds", []):
            print(f"    h={r['depth']:<6} L={r['live']:<11} S={r.get('sorted', r['live']):<11} hugeG={r['groups']:<8} bits={r['key_bits']} p={r['passes']} "
                  f"ms={r['ms_total']:.2f} sort={r['ms_sort']:.2f} bag={r.get('bag', 0)}")
