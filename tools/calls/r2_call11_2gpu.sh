#!/bin/bash
# 2 GPUs: peer-memory key-frame exchange check (fused matcher vs NCCL all_gather + single-GPU matcher), bench at N=2 for both arms, config 5 at N=2
set -u
OUT=gpurun_out/r2_call11
mkdir -p $OUT
export PSLAM_AUX_NEW=0 PSLAM_CPU_SECONDS=0.5
N=${PSLAM_NGPUS:-2}
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/exchange_check.py > $OUT/exchange_check.json 2> $OUT/exchange_check.err; echo "exchange check rc=$?" >> $OUT/summary.txt
cat $OUT/exchange_check.json; tail -3 $OUT/exchange_check.err
PSLAM_CONFIG=5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/exchange_check.py > $OUT/exchange_check_config5.json 2> $OUT/exchange_check_config5.err; echo "exchange check config5 rc=$?" >> $OUT/summary.txt
cat $OUT/exchange_check_config5.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$?" >> $OUT/summary.txt
tail -3 $OUT/bench_n$N.err
PSLAM_CONFIG=5 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2 --warmup 3 > $OUT/bench_config5_n$N.json 2> $OUT/bench_config5_n$N.err; echo "bench config5 N=$N rc=$?" >> $OUT/summary.txt
tail -3 $OUT/bench_config5_n$N.err
cat $OUT/summary.txt; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_call11/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pk=d["roofline"]["per_kernel"]
        print(f.split("/")[-1], d["metric"], "n_gpus", d["n_gpus"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), d.get("exchange"))
        print("   ", {k:round(v["ms_total"],2) for k,v in pk.items() if k.startswith("exchange")})
    except Exception as e:
        print(f, "failed", e)
PY
