#!/bin/bash
# device leg with two sub-batches of 3552 frames (the flood / seed kernels get 24 resident warps per SM instead of 12)
set -u
OUT=gpurun_out/r2_call21
mkdir -p $OUT
( time PSLAM_SUB_BATCH=3552 PSLAM_SUBS=2 PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 600 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_sub3552.json 2> $OUT/bench_sub3552.err; echo "bench sub3552 rc=$?" >> $OUT/summary.txt
tail -4 $OUT/bench_sub3552.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> $OUT/summary.txt
cat $OUT/summary.txt; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_call21/bench_sub3552.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), d["config"]["sub_batch"], d["config"]["frames_per_step"])
    print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
except Exception as e:
    print("failed", e)
PY
