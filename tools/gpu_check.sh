#!/bin/bash
# one gpurun call: GPU parity tests + default bench (+ the ablation without LSD); results under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 | cut -c1-1500
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_latest.log
if [ -n "$WITH_ABLATION" ]; then PSLAM_STAGES=orb,peac,pose timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_nolsd.log; fi
tail -5 gpurun_out/bench_err.log
python - <<PY
import json
import os
for f in [x for x in ("gpurun_out/bench_latest.log","gpurun_out/bench_nolsd.log") if os.path.exists(x)]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-2000:]); continue
    print(f, d["config"]["stages"], d["config"]["sub_batch"], "value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], d["clocks"], d.get("aux"), d.get("keylines_per_frame"))
    for k,v in d["roofline"]["per_kernel"].items(): print("   ", k, v["ms_total"], v["share"], v["achieved_gbs"])
PY
