#!/bin/bash
# ncu --set full of the PEAC / pose / ORB kernels at the bench launch size + sub-batch layout experiments
set -u
OUT=gpurun_out/r2_call6
mkdir -p $OUT
for v in "3 160 70 -16 -3 0" "3 160 70 112 61 0" "3 160 70 512 445 1"; do echo -n "tma_min $v: " >> $OUT/tma_matrix.log; timeout 60 tools/dbg/tma_min $v >> $OUT/tma_matrix.log 2>&1 || true; done
cat $OUT/tma_matrix.log
timeout 600 python tools/tma_probe.py > $OUT/tma_probe.log 2>&1; cat $OUT/tma_probe.log
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
if grep -q "maps_global: rc=0 RESULT bit-exact" $OUT/tma_probe.log; then echo "TMA blur ok" >> $OUT/summary.txt; else export PSLAM_NO_TMA=1; echo "TMA blur FAILED: fallback" >> $OUT/summary.txt; fi
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $OUT/summary.txt
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/pytest_gpu.log | head
PSLAM_STAGES=peac timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_peac_cluster|k_peac_flood|k_peac_blocks" -s 6 -c 3 -o $OUT/peac_kernels python bench.py --steps 1 --warmup 1 > $OUT/ncu_peac.log 2>&1; echo "ncu peac rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=pose timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pose_optimization" -s 2 -c 1 -o $OUT/pose_kernel python bench.py --steps 1 --warmup 1 > $OUT/ncu_pose.log 2>&1; echo "ncu pose rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=orb timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fast_cells|k_orient_describe|k_blur_tma|k_blur_level" -s 4 -c 4 -o $OUT/orb_kernels python bench.py --steps 1 --warmup 1 > $OUT/ncu_orb.log 2>&1; echo "ncu orb rc=$?" >> $OUT/summary.txt
PSLAM_SUB_BATCH=3552 PSLAM_SUBS=2 PSLAM_LSD_SUBS=2 timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_sub3552.json 2> $OUT/bench_sub3552.err; echo "bench 3552 rc=$?" >> $OUT/summary.txt
PSLAM_SUB_BATCH=7104 PSLAM_SUBS=1 PSLAM_LSD_SUBS=1 timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_sub7104.json 2> $OUT/bench_sub7104.err; echo "bench 7104 rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call6/bench_*.json")):
    try:
        d=json.load(open(f))
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1))
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>5})
    except Exception as e:
        print(f, "failed", e)
PY
