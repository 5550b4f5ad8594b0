"""Dynamic opcode mix and stall reasons of one kernel from an .ncu-rep (needs --import-source on).

usage: python tools/ncu_opmix.py report.ncu-rep warp_steps
  warp_steps = number of (warp, env-step) pairs the launch executed, e.g. 2**20/32*3000
"""
import collections
import csv
import subprocess
import sys

rep, per = sys.argv[1], float(eval(sys.argv[2]))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, data = rows[1], rows[2:]
ia, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
hist, sh = collections.Counter(), collections.Counter()
tot = ts = 0.0
for r in data:
    parts = r[ia].split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "")
    op = op.split(".")[0]
    e = float(r[ie] or 0)
    hist[op] += e
    sh[op] += float(r[iss] or 0)
    tot += e
    ts += float(r[iss] or 0)
fp64 = sum(hist[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"warp instructions per warp-step: {tot / per:.1f}   FP64-pipe (DFMA+DMUL+DADD+DSETP): {fp64 / per:.1f}")
for op, c in hist.most_common(24):
    print(f"  {op:10s} {c / per:8.1f}   stall samples {100 * sh[op] / ts:5.1f}%")
cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
agg = {h[i]: sum(float(r[i] or 0) for r in data) for i in cols}
s = sum(agg.values())
print("stall reasons:", {k: round(100 * v / s, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
