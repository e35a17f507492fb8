#!/usr/bin/env python3
"""Per-source-line profile of one kernel from an `ncu --set full` report, without the GUI: joins the report's SASS page (samples / instructions per
address) with `nvdisasm -g` line info of the same kernel in the in-tree library.
    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin name, e.g. peac_pipeline> [top N]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, kre, cub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", cub, os.path.join(root, "planarslam_b200", "libpslam_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r)
hdr = rows[h]
name = rows[0][1]
short = re.match(r"(?:void )?(?:pslam::)?(\w+)", name).group(1)
ci = {k: hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed")}
stall_cols = [i for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
sass = []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        if sass and r and r[0] == "Kernel Name":
            break                                  # first matching launch only
        continue
    sass.append((int(r[ci["Address"]], 16), r[ci["Source"]], int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]), [int(r[i] or 0) for i in stall_cols]))
base = sass[0][0]
# line info of the kernel's section
line_of, cur, inside = {}, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        inside = short in m.group(1)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = {}
ts = sum(s[2] for s in sass) or 1
ti = sum(s[3] for s in sass) or 1
for a, src, smp, ins, st in sass:
    k = line_of.get(a - base)
    e = agg.setdefault(k, [0, 0, [0] * len(stall_cols)])
    e[0] += smp
    e[1] += ins
    e[2] = [x + y for x, y in zip(e[2], st)]
print(f"{name[:100]}  samples {ts}  warp instructions {ti}")
srcs = {}
for k, (smp, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        p = os.path.join(root, "planarslam_b200", "csrc", k[0])
        if os.path.exists(p):
            srcs.setdefault(p, open(p).read().splitlines())
            text = srcs[p][k[1] - 1].strip()[:110] if k[1] - 1 < len(srcs[p]) else ""
    tops = sorted(zip(st, [hdr[i][6:] for i in stall_cols]), reverse=True)[:2]
    print(f"{100 * smp / ts:5.1f}% smp {100 * ins / ti:5.1f}% ins  {k[0] if k else '?'}:{k[1] if k else 0:<5} {tops[0][1]}/{tops[1][1]:<12} {text}")
