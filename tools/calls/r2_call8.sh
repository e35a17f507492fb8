#!/bin/bash
# rewritten surface-normal / plane post-processing kernels, staged orient_describe, pose block size variants, bench with extras + exchange
set -u
OUT=gpurun_out/r2_call8
mkdir -p $OUT
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 1500 python -m pytest -q -m gpu -x tests/test_planepost_gpu.py tests/test_orb_gpu.py tests/test_pose_gpu.py tests/test_exchange_gpu.py tests/test_golden.py tests/test_vs_compiled_reference_gpu.py tests/test_cuda_vs_reference_functions_gpu.py tests/test_track_chain_gpu.py tests/test_zz_full_batch_gpu.py > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
tail -15 $OUT/pytest.log
for nt in 64 128 256; do
  PSLAM_POSE_THREADS=$nt PSLAM_STAGES=pose PSLAM_EXTRAS=0 timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench_pose_$nt.json 2> $OUT/bench_pose_$nt.err; echo "pose $nt rc=$?" >> $OUT/summary.txt
done
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_extras.json 2> $OUT/bench_extras.err; echo "bench extras rc=$?" >> $OUT/summary.txt
tail -5 $OUT/bench_extras.err
PSLAM_CPU_SECONDS=6 timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench reference rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call8/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl") == "reference":
            print(f.split("/")[-1], d["value"], d["cpu_baseline"]["cores"], d["ms_per_step"]); continue
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"], d.get("exchange"))
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
