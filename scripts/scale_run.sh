#!/bin/bash
# usage: bash scripts/scale_run.sh N   (bench.py exactly as the driver launches it for N GPUs)
N=$1; O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 50 --warmup 10 > $O/r02b_bench_ours_N$N.json 2> $O/bench_N$N.err
tail -3 $O/bench_N$N.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r02b_bench_ours_N$N.json").read().strip().splitlines()[-1])
    print("N=$N", j["ms_per_step"], j["value"], j.get("strong_scaling", {}).get("ms_per_step"), {k: (v.get("ms_per_step"), v.get("error")) for k, v in j["workloads"].items()})
    print(j["config"]["parallelism"][:200])
except Exception as e:
    print("unreadable", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 scripts/micro/exchange_probe.py > $O/exchange_probe_N$N.log 2>&1; tail -25 $O/exchange_probe_N$N.log
