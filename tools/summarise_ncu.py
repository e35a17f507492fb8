#!/usr/bin/env python3
"""Summarise gpurun_out/prof_r1.ncu-rep (ncu --set full) into profiles/<name>.csv: per kernel launch the duration,
DRAM bytes read / written, DRAM throughput %, achieved occupancy, registers.  Run here (no GPU needed):
    python tools/summarise_ncu.py gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_full_summary.csv"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(w) if w in hdr else -1 for w in want]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(want)
    w.writerow([rows[1][i] if i >= 0 else "" for i in idx])      # units row
    for r in rows[2:]:
        w.writerow([(r[i][:60] if i >= 0 else "") for i in idx])
print("wrote", out, len(rows) - 2, "launches")
