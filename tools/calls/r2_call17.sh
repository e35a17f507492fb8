#!/bin/bash
# seed LUT + exact fast division in k_lsd_regions (full GPU suite = the bit-exact LSD tests), LSD-only A/B of the division, e2e on 4 vs 2 contexts
set -u
OUT=gpurun_out/r2_call17
mkdir -p $OUT
( time timeout 900 python -m pytest -q -m gpu tests --durations=5 ) > $OUT/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $OUT/summary.txt
tail -12 $OUT/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
tail -2 $OUT/smoke.log
for v in 1 0; do
  PSLAM_LSD_FAST_DIV=$v PSLAM_STAGES=lsd PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench_lsd_fastdiv$v.json 2> $OUT/bench_lsd_fastdiv$v.err; echo "bench lsd fastdiv=$v rc=$?" >> $OUT/summary.txt
done
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -3 $OUT/bench_default.err
PSLAM_E2E_CONTEXTS=2 PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_e2e2.json 2> $OUT/bench_e2e2.err; echo "bench e2e 2 contexts rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call17/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"], "fastdiv", d.get("lsd_fast_division_verified_on_device"), d["e2e"]["call"][:90])
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
