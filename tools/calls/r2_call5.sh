#!/bin/bash
# TMA probe matrix + LBD / exchange tests
OUT=gpurun_out/r2_call5
mkdir -p $OUT
for v in "2 64 16 0 0 0" "2 64 16 0 0 1" "2 64 16 128 64 0" "2 64 16 124 61 0" "2 144 70 128 64 0" "2 144 70 124 61 0" "2 144 70 -4 -3 0" "2 256 32 0 0 0" \
         "3 64 16 0 0 0" "3 64 16 128 64 0" "3 144 70 128 64 0" "3 144 70 124 61 0" "3 144 70 124 61 1" "3 144 70 -4 -3 0" "3 128 64 124 61 0" "3 160 70 124 61 0"; do
  echo -n "tma_min $v: " >> $OUT/tma_matrix.log
  timeout 60 tools/dbg/tma_min $v >> $OUT/tma_matrix.log 2>&1 || true
done
cat $OUT/tma_matrix.log
PSLAM_NO_TMA=1 timeout 900 python -m pytest tests/test_lbd_gpu.py tests/test_exchange_gpu.py tests/test_match_gpu.py -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
tail -15 $OUT/pytest.log; cat $OUT/summary.txt
