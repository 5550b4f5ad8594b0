"""Opcode histograms of the INNER loops (backward branches) of one kernel, largest first: a static check of a change
before spending GPU time.   usage: python tools/sass_loops.py KERNEL_SUBSTR [lib] [n_loops]"""
import collections
import re
import subprocess
import sys

pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "lqr_quadcopter_test_b200/libqtrack.so"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 4
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins, on = [], False
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, text in ins:
    if "BRA" in text:
        t = re.search(r"0x([0-9a-f]+)", text)
        if t and int(t.group(1), 16) < addr:
            loops.append((addr - int(t.group(1), 16), int(t.group(1), 16), addr))
for span, lo, hi in sorted(loops, reverse=True)[:top]:
    hist = collections.Counter()
    for addr, text in ins:
        if lo <= addr <= hi:
            parts = text.split()
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            hist[op.split(".")[0]] += 1
    tot = sum(hist.values())
    fp64 = sum(hist[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"loop 0x{lo:x}..0x{hi:x}: {tot} static instructions, FP64 {fp64} (DFMA {hist['DFMA']} DMUL {hist['DMUL']} DADD {hist['DADD']} DSETP {hist['DSETP']}), "
          f"LDCU {hist['LDCU']} LDC {hist['LDC']} LDL {hist['LDL']} STL {hist['STL']} BRA {hist['BRA']}")
