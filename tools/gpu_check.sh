#!/bin/bash
# one gpurun call: GPU parity tests + default bench; results under gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -40 | cut -c1-1500
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_latest.log
python - <<PY
import json
for f in ("gpurun_out/bench_latest.log",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-2000:]); continue
    print(f, d["config"]["sub_batch"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
    for k,v in d["roofline"]["per_kernel"].items(): print("   ", k, v["ms_total"], v["share"], v["achieved_gbs"])
PY
