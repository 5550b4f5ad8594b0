"""Opcode histogram of the hot loop (the largest backward branch span) of one kernel.
usage: python tools/sass_loop.py KERNEL_SUBSTR [lib]"""
import collections
import re
import subprocess
import sys

pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "lqr_quadcopter_test_b200/libqtrack.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins, on = [], False
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = (0, 0, 0)
for addr, text in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)*`?\(?\.?L?_?x?_?(\w+)\)?", text)
    if "BRA" in text:
        t = re.search(r"0x([0-9a-f]+)", text)
        if t:
            tgt = int(t.group(1), 16)
            if tgt < addr and addr - tgt > best[0]:
                best = (addr - tgt, tgt, addr)
span, lo, hi = best
hist = collections.Counter()
for addr, text in ins:
    if lo <= addr <= hi:
        parts = text.split()
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        hist[op.split(".")[0]] += 1
tot = sum(hist.values())
fp64 = sum(hist[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"{pat}: loop 0x{lo:x}..0x{hi:x}: {tot} static instructions, FP64 {fp64}, LDL {hist['LDL']} STL {hist['STL']}")
print("  " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(20)))
