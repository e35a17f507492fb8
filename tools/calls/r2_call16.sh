#!/bin/bash
# full GPU suite on the current tree (loop-closure path, adapters, map-plane update, rand jump, staged region2rect); LSD-only A/B of region2rect; default bench
set -u
OUT=gpurun_out/r2_call16
mkdir -p $OUT
( time timeout 900 python -m pytest -q -m gpu tests --durations=10 ) > $OUT/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $OUT/summary.txt
tail -22 $OUT/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
tail -2 $OUT/smoke.log
for v in staged shfl; do
  PSLAM_LSD_R2R=$v PSLAM_STAGES=lsd PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench_lsd_$v.json 2> $OUT/bench_lsd_$v.err; echo "bench lsd $v rc=$?" >> $OUT/summary.txt
done
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -3 $OUT/bench_default.err
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call16/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], d["metric"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"])
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
