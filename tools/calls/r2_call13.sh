#!/bin/bash
# lsd_gradient4, hamming_knn2 / exchange matcher with 4 queries per warp, lsd_blur_scale, blur_tma PRMT, lsd_improve reverted: parity + bench + launch list
set -u
OUT=gpurun_out/r2_call13
mkdir -p $OUT
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
timeout 1500 python -m pytest -q -m gpu tests > $OUT/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $OUT/summary.txt
tail -6 $OUT/pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 4 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -5 $OUT/bench_default.err
PSLAM_EXTRAS=0 timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/bench_noextras.json 2> $OUT/bench_noextras.err; echo "bench noextras rc=$?" >> $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_ncu.csv python bench.py --steps 1 --warmup 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call13/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], d["metric"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"], d["aux"].get("tracking_chain",{}).get("frames_per_sec_single_sequence"), d["aux"].get("local_bundle_adjustments_per_sec"))
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
