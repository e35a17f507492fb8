#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU): launch list with per-launch device time of one bench run, then one
# `--set full` capture of one launch of every kernel family, exported to CSV on the box (the .ncu-rep itself is too large
# to travel back).  Outputs land in gpurun_out/ and are summarised into profiles/ by tools/summarise_ncu.py.
set -x
export PSLAM_SUB_BATCH=${PSLAM_SUB_BATCH:-296}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_full.csv \
    python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none \
    -k regex:"k_fast_cells|k_peac_cluster|k_peac_flood|k_peac_blocks|k_blur_level|k_resize_level|k_quadtree|k_orient_describe|k_pose_optimization|k_peac_seed|k_lsd_blur_scale|k_lsd_gradient|k_lsd_regions|k_lsd_validate|k_local_bundle" \
    -s 60 -c 36 -o /tmp/prof_r1 python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_r1.ncu-rep --page raw --csv > gpurun_out/prof_r1_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/prof_r1.ncu-rep | tail -12
