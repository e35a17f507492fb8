#!/bin/bash
# new warp-per-frame lines3d / Manhattan kernels, ComputePlanes post-processing, and the bench step with the chained extras
set -u
OUT=gpurun_out/r2_call7
mkdir -p $OUT
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 1200 python -m pytest -q -m gpu -x tests/test_planepost_gpu.py tests/test_line3d_gpu.py tests/test_manhattan_gpu.py tests/test_cuda_vs_reference_functions_gpu.py > $OUT/pytest_new.log 2>&1; echo "pytest new rc=$?" >> $OUT/summary.txt
tail -30 $OUT/pytest_new.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_extras.json 2> $OUT/bench_extras.err; echo "bench extras rc=$?" >> $OUT/summary.txt
tail -5 $OUT/bench_extras.err
PSLAM_EXTRAS=0 timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_noextras.json 2> $OUT/bench_noextras.err; echo "bench noextras rc=$?" >> $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_extras.csv python bench.py --steps 1 --warmup 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call7/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"])
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>3})
    except Exception as e:
        print(f, "failed", e)
PY
