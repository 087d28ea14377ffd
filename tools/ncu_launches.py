"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python tools/ncu_launches.py launches.csv [steps] > summary.md
Launches before the first kernel of this repository (parameter initialisation fills) are dropped."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    for r in csv.reader(f):
        if len(r) > 10 and r[0].isdigit():
            rows.append(r)
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
first = next((i for i, r in enumerate(rows) if "forde::" in r[4]), 0)
rows = rows[first:]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4]
    name = re.sub(r"\(.*", "", name)[:110]
    agg[name][0] += 1
    agg[name][1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if "forde" in k)
print(f"launches: {len(rows)}  total device time: {tot / 1e3:.2f} ms  (per step over {steps:g} steps: {tot / 1e3 / steps:.2f} ms)")
print(f"this repo's kernels (forde::*): {ours / 1e3:.2f} ms = {100 * ours / tot:.1f} % of device time\n")
print("| kernel | launches | total us | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k}` | {n} | {us:.0f} | {100 * us / tot:.1f}% | {us / n:.1f} |")
