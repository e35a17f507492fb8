#!/bin/bash
# Round-2 GPU call 1: the full GPU suite (no xfails left, CUDA-vs-compiled-reference tests included), smoke, a short bench with the cv2-pinned LSD
# default, and an ncu launch list of the step.
set -u
OUT=gpurun_out/r2_call1
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1; echo "build rc=$?" >> $OUT/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/summary.txt
PSLAM_AUX_NEW=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 > $OUT/bench_under_ncu.log 2>&1
tail -n 8 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -3; cat $OUT/summary.txt; cat $OUT/bench.json
