"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU needed).
  python tools/ncu_summary.py launches <launches.csv>         per-kernel time shares
  python tools/ncu_summary.py raw <file.ncu-rep>              key raw metrics per captured launch
  python tools/ncu_summary.py stalls <file.ncu-rep> [kernel#] top stalled SASS lines
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    lines = open(path).read().splitlines()
    hi = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.OrderedDict()
    n = 0
    for r in csv.DictReader(lines[hi:]):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v}.get(r["Metric Unit"], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"{path}: {n} launches, {tot:.2f} ms of kernel time (cold-cache, serialised: compare shares)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {k[:44]:44s} n={a[0]:5d} ms={a[1]:10.3f} share={a[1] / tot:.3f}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    for r in rows:
        print(r[hdr.index("Kernel Name")][:80])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:60s} {r[i]} {units[i]}")


def stalls(path, which=0, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    b = blocks[which]
    ci = {h: i for i, h in enumerate(b["hdr"])}

    def f(r, k):
        try:
            return float(r[ci[k]])
        except (ValueError, IndexError, KeyError):
            return 0.0

    tot = sum(f(r, "# Samples") for r in b["rows"]) or 1.0
    st = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(f(r, s) for r in b["rows"]) for s in st}
    print(b["name"][:100], "samples", int(tot))
    print("  stall mix:", {k.replace("stall_", ""): round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
    for r in sorted(b["rows"], key=lambda r: -f(r, "# Samples"))[:top]:
        s2 = sorted([(s, f(r, s)) for s in st], key=lambda x: -x[1])[:2]
        print("  %5.2f%%  %-72s %s" % (100 * f(r, "# Samples") / tot, r[ci["Source"]][:72], [(a.replace("stall_", ""), int(v)) for a, v in s2]))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "raw":
        raw(sys.argv[2])
    else:
        stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
