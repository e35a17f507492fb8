#!/bin/bash
# First gpurun call of round 2: validates on a B200 everything that round 1 could only check on the host, in one go.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Outputs under gpurun_out/r2_first/: pytest logs (incl. the non-strict xfail tests run strictly), the aux throughput JSON, an A/B of the two
# LSD rectangle enumerations, and an ncu launch list of the new kernels.
set -u
OUT=gpurun_out/r2_first
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
# 1. the GPU suite as the driver runs it, then the xfail-marked tests strictly
python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $OUT/summary.txt
python -m pytest tests/test_line3d_gpu.py tests/test_manhattan_gpu.py tests/test_lsd_gpu.py tests/test_linefrustum_gpu.py tests/test_framefill.py tests/test_vs_compiled_reference_gpu.py -q -m gpu --runxfail > $OUT/pytest_runxfail.log 2>&1; echo "pytest runxfail rc=$?" >> $OUT/summary.txt
# 2. throughput + signatures of the new kernels
PYTHONPATH=. timeout 300 python tools/aux_new_kernels.py > $OUT/aux_new_kernels.json 2> $OUT/aux_new_kernels.err; echo "aux rc=$?" >> $OUT/summary.txt
# 3. LSD: published iterator vs OpenCV 4.x enumeration (same bench, LSD stage only)
PSLAM_AUX_NEW=0 PSLAM_STAGES=lsd python bench.py --steps 3 --warmup 3 > $OUT/bench_lsd_enum0.json 2> $OUT/bench_lsd_enum0.err
PSLAM_AUX_NEW=0 PSLAM_STAGES=lsd PSLAM_LSD_RECT_ENUM=cv4 python bench.py --steps 3 --warmup 3 > $OUT/bench_lsd_enum1.json 2> $OUT/bench_lsd_enum1.err
# 4. launch list of the new kernels (ncu serialises; times are cold-cache)
PYTHONPATH=. timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_lines3d|k_track_manhattan" -c 8 --csv --log-file $OUT/ncu_new_kernels.csv \
    python tools/aux_new_kernels.py > /dev/null 2> $OUT/ncu_new_kernels.err
tail -n 5 $OUT/pytest_gpu.log $OUT/pytest_runxfail.log; cat $OUT/summary.txt; cat $OUT/aux_new_kernels.json
