"""Print the headline numbers of a bench.py JSON line (stdin)."""
import json
import sys

line = [l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
r = d.get("roofline", {})
print(f"value={d['value']:.4e} {d['unit']}  ms/step={d['ms_per_step']:.2f}  fp64_frac={r.get('frac')}"
      f"  e2e={d.get('e2e', {}).get('value', 0):.4e}  policy={(d.get("policy") or {}).get("value", 0):.3e}  dare={d.get('dare', {}).get('value', 0):.4e}"
      f"  clocks={d.get('clocks', {}).get('sm_mhz')}")
