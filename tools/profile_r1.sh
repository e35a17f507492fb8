#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU): launch list with per-launch device time, then one `--set full`
# capture of the heaviest kernels.  Outputs land in gpurun_out/ and are summarised into profiles/ by tools/summarise_ncu.py.
set -x
export PSLAM_SUB_BATCH=${PSLAM_SUB_BATCH:-64}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_full.csv \
    python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_fast_cells|k_peac_cluster|k_peac_flood|k_peac_blocks|k_blur_level|k_resize_level|k_quadtree|k_orient_describe|k_pose_optimization|k_peac_final|k_peac_seed" \
    -s 40 -c 24 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
