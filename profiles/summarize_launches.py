#!/usr/bin/env python
"""Share table of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised times: SHARES, not
absolutes):   python profiles/summarize_launches.py gpurun_out/x/launches_step.csv "header comment" > profiles/rNN_launches.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
mul = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}
for r in rows:
    name = re.sub(r"^void ", "", r[4])
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^sx::", "", name)
    tot[name] += float(r[14].replace(",", "")) * mul.get(r[13], 1e-6)
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# total {total:.3f} ms over {len(rows)} launches.   total_ms %share launches kernel")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{t:9.3f} {100 * t / total:5.1f}% n={cnt[name]:4d}  {name[:150]}")
