#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per launch with duration, DRAM bytes read / written, achieved
GB/s against the measured copy bandwidth, tensor-pipe and L1/shared data-pipe utilisation, issue-active.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x_raw.csv ; python profiles/ncu_table.py x_raw.csv [--json out.json]
"""
import csv
import json
import os
import re
import sys

raw = sys.argv[1]
rows = list(csv.reader(open(raw)))
while rows and not (rows[0] and rows[0][0] == "ID"):
    rows.pop(0)
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {n: i for i, n in enumerate(hdr)}
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
try:
    hbm = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0


def val(r, n, default=None):
    if n not in idx or r[idx[n]] in ("", "n/a"):
        return default
    return float(r[idx[n]].replace(",", "")) * mul.get(units[idx[n]], 1)


def first(r, names):
    for n in names:
        v = val(r, n)
        if v is not None:
            return v
    return None


out = []
print(f"{'kernel':58s} {'grid':>10s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s} {'%hbm':>6s} {'tensor%':>8s} {'l1data%':>8s} {'issue%':>7s}")
for r in data:
    name = re.sub(r"^void (sx::)?(tc::)?", "", r[idx["Kernel Name"]])
    name = re.sub(r"\(.*$", "", name)
    t = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum", 0.0), val(r, "dram__bytes_write.sum", 0.0)
    tensor = first(r, ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                       "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                       "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"])
    l1 = first(r, ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"])
    iss = first(r, ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active"])
    gbs = (rd + wr) / t / 1e9 if t else 0.0
    grid = r[idx["Grid Size"]].replace(" ", "")
    print(f"{name[:58]:58s} {grid:>10s} {t * 1e6:9.1f} {rd / 1e6:9.2f} {wr / 1e6:9.2f} {gbs:8.0f} {100 * gbs / hbm:6.1f} "
          f"{(tensor if tensor is not None else float('nan')):8.1f} {(l1 if l1 is not None else float('nan')):8.1f} "
          f"{(iss if iss is not None else float('nan')):7.1f}")
    out.append({"kernel": name, "grid": grid, "time_us": t * 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_gbs": gbs,
                "frac_of_measured_hbm": gbs / hbm, "tensor_pipe_pct": tensor, "l1_data_pipe_pct": l1, "issue_active_pct": iss})
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
