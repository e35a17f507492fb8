#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, launches, summed device time and share of the step.
    python tools/launch_shares.py gpurun_out/.../launches.csv > profiles/<round>_launch_shares.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
    tot[k][0] += 1
    tot[k][1] += v
T = sum(v[1] for v in tot.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in tot.values())} launches, {T / 1e3:.1f} ms summed device time (serialised, cold cache: compare shares)")
for k, v in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k[:60]:60s} n={v[0]:4d} total_us={v[1]:12.1f} share={v[1] / T:.4f}")
