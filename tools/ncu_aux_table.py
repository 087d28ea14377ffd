"""Markdown table from `ncu --page raw --csv` of tools/aux_once.py: per kernel duration, DRAM bytes and throughput,
tensor / XU (MUFU) pipe utilisation, registers, achieved occupancy.  python tools/ncu_aux_table.py raw.csv > table.md"""
import csv
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "byte/second": 1.0, "Kbyte/second": 1e3, "Mbyte/second": 1e6,
        "Gbyte/second": 1e9, "Tbyte/second": 1e12}
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hi], rows[hi + 1]
data = [r for r in rows[hi + 2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    if name not in ix or r[ix[name]] in ("", "n/a"):
        return None
    return float(r[ix[name]].replace(",", "")) * UNIT.get(units[ix[name]], 1.0)


def fmt(v, f="{:.1f}"):
    return "—" if v is None else f.format(v)


half = len(data) // 2          # aux_once.py launches everything twice; the second round is the warm one
print("| kernel | grid x block | us under ncu | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % of peak | tensor pipe % | XU (MUFU) pipe % | regs | achieved occupancy % |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in data[half:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")[:60]
    dur = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    bw = (rd + wr) / dur / 1e3 if dur and rd is not None and wr is not None else None
    print(f"| `{name}` | {r[ix['Grid Size']]} x {r[ix['Block Size']]} | {fmt(dur)} | {fmt(rd and rd / 1e6)} | {fmt(wr and wr / 1e6)} | "
          f"{fmt(bw, '{:.0f}')} | {fmt(val(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'))} | "
          f"{fmt(val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or val(r, 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active'))} | "
          f"{fmt(val(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'))} | "
          f"{fmt(val(r, 'launch__registers_per_thread'), '{:.0f}')} | {fmt(val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'))} |")
