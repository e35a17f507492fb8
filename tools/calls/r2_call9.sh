#!/bin/bash
# packed FAST scorer + orient_describe without constant-memory serialisation: parity, bench, ncu of the per-pixel kernels; LSD at other image sizes
set -u
OUT=gpurun_out/r2_call9
mkdir -p $OUT
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 1500 python -m pytest -q -m gpu -x tests/test_orb_gpu.py tests/test_golden.py tests/test_vs_compiled_reference_gpu.py tests/test_lsd_gpu.py tests/test_zz_full_batch_gpu.py tests/test_track_chain_gpu.py > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/summary.txt
tail -15 $OUT/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_extras.json 2> $OUT/bench_extras.err; echo "bench extras rc=$?" >> $OUT/summary.txt
tail -5 $OUT/bench_extras.err
PSLAM_STAGES=orb PSLAM_EXTRAS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fast_cells|k_orient_describe|k_resize_level" -s 14 -c 4 -o $OUT/orb_kernels python bench.py --steps 1 --warmup 1 > $OUT/ncu_orb.log 2>&1; echo "ncu orb rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=lsd PSLAM_EXTRAS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lsd_gradient|k_lsd_blur_scale" -s 2 -c 2 -o $OUT/lsd_pixel_kernels python bench.py --steps 1 --warmup 1 > $OUT/ncu_lsd.log 2>&1; echo "ncu lsd rc=$?" >> $OUT/summary.txt
PSLAM_STAGES=peac PSLAM_EXTRAS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_planes_post|k_sn_integral|k_peac_blocks" -s 3 -c 3 -o $OUT/peac_extra_kernels python bench.py --steps 1 --warmup 1 > $OUT/ncu_peac.log 2>&1; echo "ncu peac rc=$?" >> $OUT/summary.txt
PSLAM_CONFIG=5 timeout 900 python bench.py --steps 2 --warmup 3 > $OUT/bench_config5.json 2> $OUT/bench_config5.err; echo "bench config5 rc=$?" >> $OUT/summary.txt
tail -5 $OUT/bench_config5.err
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call9/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], d["metric"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"], d.get("exchange"))
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
