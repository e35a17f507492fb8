#!/bin/bash
# LSD v2 (bulk seed ordering kernel + 4x8-lane region growing): parity, LSD-only bench variants (occupancy target x launches per step), ncu --set full of
# k_lsd_regions at the bench launch size
set -u
OUT=gpurun_out/r2_call2
mkdir -p $OUT
timeout 900 python -m pytest tests/test_lsd_gpu.py tests/test_golden.py tests/test_cuda_vs_reference_functions_gpu.py tests/test_line3d_gpu.py -q -m gpu -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
export PSLAM_AUX_NEW=0 PSLAM_STAGES=lsd PSLAM_CPU_SECONDS=0.5
for v in "16 3" "24 2" "16 1" "24 1" "32 1"; do
  set -- $v
  PSLAM_LSD_OCC=$1 PSLAM_LSD_SUBS=$2 timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_lsd_occ$1_subs$2.json 2> $OUT/bench_lsd_occ$1_subs$2.err; echo "bench occ$1 subs$2 rc=$?" >> $OUT/summary.txt
done
PSLAM_LSD_OCC=16 PSLAM_LSD_SUBS=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lsd_regions -s 3 -c 1 -o $OUT/lsd_regions python bench.py --steps 1 --warmup 1 > $OUT/ncu.log 2>&1; echo "ncu rc=$?" >> $OUT/summary.txt
tail -n 15 $OUT/pytest.log; cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call2/bench_lsd_occ*.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"],1), {k:(v["ms_total"],v["launches"]) for k,v in d["roofline"]["per_kernel"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
