"""Turn one `ncu --set full` capture of the six StatefulLayer GEMMs into profiles/ncu_traffic.json (+ a markdown table),
tied to the SHA-256 of the library that was profiled: bench.py only quotes `roofline.traffic` from this file when the
hash equals the hash of the library it loads.

On the GPU box (one GPU; never under torchrun):
  ncu --set full --clock-control none --import-source on -k regex:forde_gemm_kernel -c 6 -o gpurun_out/r2_layer \\
      python tools/layer_bench.py --once --assign zeros
  ncu -i gpurun_out/r2_layer.ncu-rep --page raw --csv > gpurun_out/r2_layer_raw.csv
Here:
  python tools/ncu_traffic.py gpurun_out/r2_layer_raw.csv --dims 50432 768 3072 --map zeros --out profiles/ncu_traffic.json \\
      --md profiles/r2_ncu_full_kernels.md
layer_bench --once launches the GEMMs in a fixed order (up_fwd, down_fwd, dz_bwd, dx_bwd, dw_up, dw_down); rows are named
by that order.  Algorithmic bytes are the formulas of DESIGN.md section 3 evaluated at --dims."""
import argparse
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ["sl_up_fwd", "sl_down_fwd", "sl_dz_bwd", "sl_dx_bwd", "sl_dw_up", "sl_dw_down"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}

ap = argparse.ArgumentParser()
ap.add_argument("raw_csv")
ap.add_argument("--dims", type=int, nargs=3, default=[50432, 768, 3072], metavar=("M", "d", "H"))
ap.add_argument("--map", default="zeros", help="brain map of the capture (layer_bench --assign)")
ap.add_argument("--lib", default=os.path.join(ROOT, "forde-vision-language_b200", "libforde_b200.so"))
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ncu_traffic.json"))
ap.add_argument("--md", default=None)
a = ap.parse_args()
M, d, H = a.dims

rows = [r for r in csv.reader(open(a.raw_csv)) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units, data = rows[hdr_i], rows[hdr_i + 1], [r for r in rows[hdr_i + 2:] if len(r) == len(rows[hdr_i])]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=None):
    if name not in ix or r[ix[name]] in ("", "n/a"):
        return default
    return float(r[ix[name]].replace(",", "")) * UNIT.get(units[ix[name]], 1.0)


alg = {  # algorithmic bytes per launch (all-generalist map: no z^T traffic; sign words M*H/8 bytes each way)
    "sl_up_fwd": 2 * (M * d + d * H + M * H) + M * H // 8,
    "sl_down_fwd": 2 * (M * H + H * d + M * d),
    "sl_dz_bwd": 2 * (M * d + d * H + M * H) + M * H // 8,
    "sl_dx_bwd": 2 * (M * H + d * H + M * d),
    "sl_dw_up": 2 * (M * d + M * H) + 4 * d * H,
    "sl_dw_down": 2 * (M * H + M * d) + 4 * d * H,
}
gemm_rows = [r for r in data if "forde_gemm_kernel" in r[ix["Kernel Name"]]]
kernels = {}
for name, r in zip(ORDER, gemm_rows):
    rd, wr = val(r, "dram__bytes_read.sum", 0.0), val(r, "dram__bytes_write.sum", 0.0)
    kernels[f"{name}[{M}x{d}x{H}]"] = {
        "kernel": r[ix["Kernel Name"]][:80],
        "dram_read_MB": round(rd / 1e6, 1), "dram_write_MB": round(wr / 1e6, 1), "traffic_bytes": rd + wr,
        "algorithmic_bytes": alg[name], "traffic_over_algorithmic": round((rd + wr) / alg[name], 3),
        "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        or val(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
        "duration_us_under_ncu": val(r, "gpu__time_duration.sum"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
        "grid_size": val(r, "launch__grid_size"),
        "inst_executed": val(r, "smsp__inst_executed.sum"),
    }
sha = hashlib.sha256(open(a.lib, "rb").read()).hexdigest()[:16]
try:   # hash of the sources the library was built from (written by __graft_entry__.build(); path-independent)
    src = open(a.lib + ".srchash").read().strip()[:16] or None
except OSError:
    src = None
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402
# hash of the GEMM engine's sources in THIS tree (run this right after the capture, on the tree that was profiled)
gemm_src = G.source_hash(G.GEMM_SOURCES)[:16] if src == G.source_hash()[:16] else None
out = {"lib_sha256_16": sha, "src_sha256_16": src, "gemm_src_sha256_16": gemm_src, "brain_map": a.map, "dims": {"M": M, "d": d, "H": H},
       "source": f"{os.path.relpath(a.raw_csv, ROOT)} (ncu --set full --clock-control none, one launch per kernel, "
                 f"tools/layer_bench.py --once --assign {a.map})",
       "kernels": kernels}
json.dump(out, open(a.out, "w"), indent=1)
print(f"wrote {a.out}: {len(kernels)} kernels, library {sha}")
if a.md:
    with open(a.md, "w") as f:
        f.write(f"# ncu --set full of the six StatefulLayer GEMMs (M={M}, d={d}, H={H}, brain map: {a.map}), library {sha}\n\n")
        f.write("One launch per kernel under ncu (cold caches, serialised): compare counters and shares, not absolute times; "
                "event-timed figures are in the layer-bench and bench records.\n\n")
        f.write("| kernel | DRAM read MB | DRAM write MB | traffic / algorithmic bytes | tensor pipe % | regs | grid (CTAs) | us under ncu |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for n, k in kernels.items():
            f.write(f"| `{n}` | {k['dram_read_MB']} | {k['dram_write_MB']} | {k['traffic_over_algorithmic']} | "
                    f"{k['tensor_pipe_pct']} | {k['registers_per_thread']} | {k['grid_size']} | {k['duration_us_under_ncu']} |\n")
