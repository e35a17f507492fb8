#!/bin/bash
# Round-2 GPU call 3: whole GPU suite (LSD validate split, TMA blur, tracking chain, 4x8 region growing at 24 CTAs/SM), then the full step in three
# stream / occupancy layouts, then ncu --set full of the TMA blur and the LSD validation kernels.
set -u
OUT=gpurun_out/r2_call3
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $OUT/summary.txt
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_A_default.json 2> $OUT/bench_A.err; echo "bench A rc=$?" >> $OUT/summary.txt
PSLAM_LSD_STREAM=own PSLAM_LSD_OCC=32 PSLAM_LSD_SUBS=3 timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_B_own_occ32.json 2> $OUT/bench_B.err; echo "bench B rc=$?" >> $OUT/summary.txt
PSLAM_LSD_STREAM=own timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_C_own_occ24.json 2> $OUT/bench_C.err; echo "bench C rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=orb timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blur_tma -s 2 -c 1 -o $OUT/blur_tma python bench.py --steps 1 --warmup 1 > $OUT/ncu_blur.log 2>&1; echo "ncu blur rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=lsd timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lsd_validate|k_lsd_improve|k_lsd_order" -s 6 -c 3 -o $OUT/lsd_validate python bench.py --steps 1 --warmup 1 > $OUT/ncu_val.log 2>&1; echo "ncu validate rc=$?" >> $OUT/summary.txt
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -3; grep -E "^FAILED|^ERROR" $OUT/pytest_gpu.log | head -20; cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call3/bench_*.json")):
    try:
        d=json.load(open(f))
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1))
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>5})
    except Exception as e:
        print(f, "failed", e)
PY
