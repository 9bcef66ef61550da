"""profiles/conv_tc_traffic.json from one `ncu --set full` capture of the Conv2DMod kernels of ONE 256px forward at the
sweep's batch size:

    ncu --set full --clock-control none -k regex:conv_tc -s 28 -c 14 -o gpurun_out/fwd python profiles/exp_layers.py --batch 256 --iters 1
    ncu -i gpurun_out/fwd.ncu-rep --page raw --csv > gpurun_out/fwd_raw.csv
    python profiles/make_traffic_json.py gpurun_out/fwd_raw.csv 256

`dram_bytes_per_launch` = DRAM read + write bytes of a Conv2DMod launch averaged over the launches of one AttFind step:
conv c runs in every sweep batch whose perturbed coordinate lives in conv c' <= c, so its launch count is the number of
batches of convs 0..c (ceil(2 * coords / max_batch) each).  bench.py copies the number into roofline.traffic.
"""
import csv
import json
import math
import os
import re
import sys

raw, batch = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {n: i for i, n in enumerate(hdr)}
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, n):
    return float(r[idx[n]].replace(",", "")) * mul.get(units[idx[n]], 1)


pairs = [(512, 512), (512, 512), (512, 512), (512, 256), (256, 128), (128, 64), (64, 32)]      # 256px generator
coords = []
for ci, co in pairs:
    coords += [ci, co]
assert len(data) == 14, f"expected the 14 conv launches of one forward, got {len(data)}"
batches = [math.ceil(2 * c / batch) for c in coords]
launches = [sum(batches[: c + 1]) for c in range(14)]
layers, tot_b, tot_alg = [], 0.0, 0.0
for c, r in enumerate(data):
    ci, co = pairs[c // 2] if c % 2 == 0 else (pairs[c // 2][1], pairs[c // 2][1])
    hw = (4 << (c // 2)) ** 2
    ups = "1, 2, 8" in r[idx["Kernel Name"]] or re.search(r"true, \(int\)\d, \(int\)8", r[idx["Kernel Name"]]) is not None
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    t = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}[units[idx["gpu__time_duration.sum"]]]
    layers.append({"conv": c, "kernel": re.sub(r"^void sx::tc::", "", r[idx["Kernel Name"]]).split("(CUtensor")[0],
                   "batch": batch, "dram_read_bytes": rd, "dram_write_bytes": wr, "time_us": t * 1e6,
                   "launches_per_step": launches[c]})
    tot_b += (rd + wr) * launches[c]
json.dump({"source": "ncu --set full --clock-control none, the 14 Conv2DMod launches of one 256px forward at the sweep batch "
                     "size (profiles/make_traffic_json.py); weighted by each conv's launches per AttFind step",
           "batch": batch, "dram_bytes_per_launch": tot_b / sum(launches), "layers": layers},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "conv_tc_traffic.json"), "w"), indent=1)
print("dram_bytes_per_launch = %.1f MB over %d launches per step" % (tot_b / sum(launches) / 1e6, sum(launches)))
