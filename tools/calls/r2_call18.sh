#!/bin/bash
# flood kernel with pipelined stage-A loads, rect_improve with the shared pixel pass (seed table / fast division reverted): parity + bench
set -u
OUT=gpurun_out/r2_call18
mkdir -p $OUT
( time timeout 900 python -m pytest -q -m gpu tests --durations=5 ) > $OUT/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $OUT/summary.txt
tail -12 $OUT/pytest_all.log
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -3 $OUT/bench_default.err
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call18/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"])
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
