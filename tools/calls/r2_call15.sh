#!/bin/bash
# loop-closure path + new adapters on the GPU; ncu --set full (with source) of the three dominant serial kernels for per-line profiles
set -u
OUT=gpurun_out/r2_call15
mkdir -p $OUT
( time timeout 600 python -m pytest -q -m gpu tests/test_loopclose_gpu.py tests/test_reference_adapter_gpu.py tests/test_bow_gpu.py tests/test_linesearch_gpu.py tests/test_mapplane_gpu.py --durations=8 ) > $OUT/pytest_new.log 2>&1; echo "pytest new rc=$?" >> $OUT/summary.txt
tail -25 $OUT/pytest_new.log
( time PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lsd_regions|k_peac_cluster|k_peac_flood" -c 3 -f -o $OUT/top3 python bench.py --steps 1 --warmup 1 ) > $OUT/ncu_top3.log 2>&1; echo "ncu top3 rc=$?" >> $OUT/summary.txt
tail -5 $OUT/ncu_top3.log
ls -la $OUT
cat $OUT/summary.txt
