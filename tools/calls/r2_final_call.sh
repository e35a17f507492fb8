#!/bin/bash
# round-2 final: full GPU suite, smoke, default bench (full CPU arm), launch list, config 5, ncu --set full of the two kernels rewritten last
set -u
OUT=gpurun_out/r2_final
mkdir -p $OUT
( time timeout 900 python -m pytest -q -m gpu tests --durations=5 ) > $OUT/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $OUT/summary.txt
tail -10 $OUT/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
tail -2 $OUT/smoke.log
( time timeout 600 python bench.py --steps 4 --warmup 3 ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?" >> $OUT/summary.txt
tail -3 $OUT/bench_default.err
PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_ncu.csv python bench.py --steps 1 --warmup 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/summary.txt
( time PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 PSLAM_CONFIG=5 timeout 420 python bench.py --steps 2 --warmup 3 ) > $OUT/bench_config5.json 2> $OUT/bench_config5.err; echo "bench config5 rc=$?" >> $OUT/summary.txt
( time PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5 timeout 420 ncu --set full --clock-control none -k regex:"k_lsd_regions|k_peac_flood" -c 2 -f -o $OUT/top2 python bench.py --steps 1 --warmup 1 ) > $OUT/ncu_top2.log 2>&1; echo "ncu top2 rc=$?" >> $OUT/summary.txt
tail -3 $OUT/ncu_top2.log
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl") == "reference":
            print(f.split("/")[-1], d["metric"], "value", d["value"], d.get("cpu_baseline"))
            continue
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], d["metric"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"])
        print("   ", {k:round(v["ms_total"],1) for k,v in pk.items() if v["ms_total"]>2})
    except Exception as e:
        print(f, "failed", e)
PY
