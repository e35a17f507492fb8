#!/bin/bash
# frame-constructor call (parity + e2e), PEAC clustering occupancy variants, LSD at other sizes, full GPU suite
set -u
OUT=gpurun_out/r2_call10
mkdir -p $OUT
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 600 python -m pytest -q -m gpu -x tests/test_frame_construct_gpu.py tests/test_lsd_gpu.py tests/test_peac_gpu.py > $OUT/pytest_new.log 2>&1; echo "pytest new rc=$?" >> $OUT/summary.txt
tail -8 $OUT/pytest_new.log
for occ in 12 16 20; do
  PSLAM_PEAC_OCC=$occ PSLAM_STAGES=peac PSLAM_EXTRAS=0 timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench_peac_occ$occ.json 2> $OUT/bench_peac_occ$occ.err; echo "peac occ $occ rc=$?" >> $OUT/summary.txt
done
PSLAM_PEAC_OCC=16 timeout 300 python -m pytest -q -m gpu -x tests/test_peac_gpu.py > $OUT/pytest_peac16.log 2>&1; echo "pytest peac occ16 rc=$?" >> $OUT/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -5 $OUT/bench_default.err
timeout 1500 python -m pytest -q -m gpu tests > $OUT/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $OUT/summary.txt
tail -5 $OUT/pytest_all.log
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call10/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], d["metric"], "frames/step", d["config"]["frames_per_step"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"])
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
