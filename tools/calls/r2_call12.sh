#!/bin/bash
# warp-per-rectangle lsd_improve, division-free fast_cells, new exchange matcher, planes_post unroll: parity + bench; pose block size 32 vs 64
set -u
OUT=gpurun_out/r2_call12
mkdir -p $OUT
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 1500 python -m pytest -q -m gpu tests > $OUT/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $OUT/summary.txt
tail -6 $OUT/pytest_all.log
for nt in 32 64; do
  PSLAM_POSE_THREADS=$nt PSLAM_STAGES=pose PSLAM_EXTRAS=0 timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench_pose_$nt.json 2> $OUT/bench_pose_$nt.err; echo "pose $nt rc=$?" >> $OUT/summary.txt
done
timeout 900 python bench.py --steps 4 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -5 $OUT/bench_default.err
PSLAM_STAGES=orb PSLAM_EXTRAS=0 timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_orb_only.json 2> $OUT/bench_orb_only.err; echo "bench orb only rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call12/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], d["metric"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"], d["aux"].get("tracking_chain"), d["aux"].get("local_bundle_adjustments_per_sec"))
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
