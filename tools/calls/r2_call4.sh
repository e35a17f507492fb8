#!/bin/bash
# Round-2 GPU call 4: TMA probe (maps in global memory vs parameter block, compute-sanitizer on failure), whole GPU suite (LSD v3 planes, validate split,
# tracking chain, reference-typed adapters), the full step in three stream / occupancy layouts, ncu of the TMA blur and LSD kernels.
set -u
OUT=gpurun_out/r2_call4
mkdir -p $OUT
timeout 900 python tools/tma_probe.py --sanitize > $OUT/tma_probe.log 2>&1; echo "tma probe rc=$?" >> $OUT/summary.txt
if grep -q "maps_global: rc=0 RESULT bit-exact" $OUT/tma_probe.log; then :; elif grep -q "maps_param: rc=0 RESULT bit-exact" $OUT/tma_probe.log; then export PSLAM_TMA_MAPS=param; else export PSLAM_NO_TMA=1; fi
echo "tma mode: ${PSLAM_TMA_MAPS:-global} no_tma=${PSLAM_NO_TMA:-0}" >> $OUT/summary.txt
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $OUT/summary.txt
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_A_default.json 2> $OUT/bench_A.err; echo "bench A rc=$?" >> $OUT/summary.txt
PSLAM_LSD_STREAM=own PSLAM_LSD_OCC=32 PSLAM_LSD_SUBS=3 timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_B_own_occ32.json 2> $OUT/bench_B.err; echo "bench B rc=$?" >> $OUT/summary.txt
PSLAM_LSD_STREAM=own timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_C_own_occ24.json 2> $OUT/bench_C.err; echo "bench C rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=orb timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blur_tma -s 2 -c 1 -o $OUT/blur_tma python bench.py --steps 1 --warmup 1 > $OUT/ncu_blur.log 2>&1; echo "ncu blur rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=lsd timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lsd_regions|k_lsd_validate|k_lsd_improve|k_lsd_order" -s 8 -c 4 -o $OUT/lsd_kernels python bench.py --steps 1 --warmup 1 > $OUT/ncu_lsd.log 2>&1; echo "ncu lsd rc=$?" >> $OUT/summary.txt
cat $OUT/tma_probe.log | head -60; grep -E "passed|failed" $OUT/pytest_gpu.log | tail -3; grep -E "^FAILED|^ERROR" $OUT/pytest_gpu.log | head -20; cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call4/bench_*.json")):
    try:
        d=json.load(open(f))
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1))
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>5})
    except Exception as e:
        print(f, "failed", e)
PY
