"""Summarise the `ncu --page source --csv` dump of one kernel: stall totals and the hottest SASS lines.
Usage: ncu -i rep.ncu-rep --page source --csv --launch-skip N --launch-count 1 > src.csv; python tools/ncu_src.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) > 10:
        cur["data"].append(r)
for sec in sections:
    hdr, data = sec["hdr"], sec["data"]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print(sec["name"], "| samples", tot, "| SASS lines", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    print("  stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
    for r in top:
        st = sorted([(s[6:], int(r[ix[s]])) for s in stalls], key=lambda kv: -kv[1])[:2]
        print(r[ix["# Samples"]].rjust(7), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]].strip()[:72].ljust(72), st)
