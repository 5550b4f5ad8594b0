"""Per-source-line totals (instructions executed, stall samples) of one kernel from an .ncu-rep captured
with --import-source on.   usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hdr]
il, isrc, ie, iss = h.index("Line No"), h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
fpath = ""
agg = collections.defaultdict(lambda: [0.0, 0.0, ""])
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        fpath = r[1].split("/")[-1]
    if len(r) <= iss or not r[il].isdigit():
        continue            # rows with an empty line number are the SASS instructions of the line above
    key = (fpath, int(r[il]))
    try:   # source lines with quotes and commas (inline asm) can shift the CSV columns: skip those rows
        e, st = float(r[ie] or 0), float(r[iss] or 0)
    except ValueError:
        continue
    agg[key][0] += e
    agg[key][1] += st
    agg[key][2] = r[isrc].strip()[:90]
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print(f"total instructions {ti:.3e}, stall samples {ts:.0f}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:18s}:{l:4d} inst {100 * v[0] / ti:5.1f}%  stalls {100 * v[1] / ts:5.1f}%  {v[2]}")
