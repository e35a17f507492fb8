#!/bin/bash
# e2e leg: frame-construct copy overlap + C batch pose call
set -u
OUT=gpurun_out/r2_call20
mkdir -p $OUT
timeout 300 python -m pytest -q -m gpu tests/test_frame_construct_gpu.py tests/test_pose_gpu.py > $OUT/pytest_fc.log 2>&1; echo "pytest frame construct rc=$?" >> $OUT/summary.txt
tail -3 $OUT/pytest_fc.log
( time PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 600 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -3 $OUT/bench_default.err
cat $OUT/summary.txt; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_call20/bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), d["e2e"])
PY
