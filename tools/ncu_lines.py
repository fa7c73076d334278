"""Per-source-line view of an `ncu --set full --import-source on` capture.

ncu's CSV source page is SASS-only; this joins it with `nvdisasm -g` of the in-tree libgsa.so
(same build as the one profiled) and sums executed instructions and stall samples per .cu line.

Usage: python tools/ncu_lines.py <capture.ncu-rep> <kernel-name-substring> [top] [launch index in the capture]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "stringsearch_b200", "libgsa.so")], cwd=tmp,
                   capture_output=True)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    table, cur_line, inside, fname = {}, None, False, None
    for ln in sass:
        if ln.startswith(".text."):
            inside = kernel_sub in ln and not table
            if inside:
                fname = ln.strip().rstrip(":")
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            table[int(m.group(1), 16)] = cur_line
    return table, fname


def main():
    rep, sub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    table, fname = line_table(sub)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the page holds one block per captured launch: "Kernel Name" row, header row, SASS rows
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    lo = starts[which]
    hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
    hdr = rows[lo + 1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
    base = int(body[0][0], 16)
    agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_i = tot_s = 0.0
    for r in body:
        off = int(r[0], 16) - base
        key = table.get(off, ("?", 0))
        ins = float(r[ci["Instructions Executed"]] or 0)
        smp = float(r[ci["# Samples"]] or 0)
        a = agg[key]
        a[0] += ins
        a[1] += smp
        for s in stalls:
            v = float(r[ci[s]] or 0)
            if v:
                a[2][s[6:]] += v
        tot_i += ins
        tot_s += smp
    src = {}
    for f in {k[0] for k in agg}:
        for p in glob.glob(os.path.join(ROOT, "stringsearch_b200", "csrc", f)):
            src[f] = open(p).read().splitlines()
    print(f"{fname}: {int(tot_i)} warp instructions, {int(tot_s)} samples, {len(body)} SASS rows ({len(table)} in nvdisasm)")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = src.get(key[0], [""] * (key[1] + 1))[key[1] - 1].strip()[:70] if key[1] else ""
        print("  smp %5.2f%%  ins %5.2f%%  %s:%-5d %-70s %s" % (100 * a[1] / max(tot_s, 1), 100 * a[0] / max(tot_i, 1), key[0], key[1],
                                                          text, dict(a[2].most_common(2))))


if __name__ == "__main__":
    main()
