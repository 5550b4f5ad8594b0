"""Static opcode histogram of one kernel in libqtrack.so (cuobjdump -sass): a quick check before
spending GPU time.  usage: python tools/sass_count.py 'rollout_kernelILi1ELi0ELb0E' [lib]"""
import collections
import subprocess
import sys

pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "lqr_quadcopter_test_b200/libqtrack.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
hist, on = collections.Counter(), False
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if on and "/*" in line and ";" in line:
        body = line.split("*/", 1)[1].strip() if line.strip().startswith("/*") else line.strip()
        parts = body.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        hist[op.split(".")[0]] += 1
tot = sum(hist.values())
fp64 = sum(hist[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"{pat}: {tot} instructions, FP64 pipe {fp64}")
print("  " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(18)))
