O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 50 --warmup 10 --only --no-strong --grad-comm multimem > $O/b_mm_N2.json 2> $O/b_mm_N2.err
tail -4 $O/b_mm_N2.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/b_mm_N2.json").read().strip().splitlines()[-1])
print("N=2 multimem", j["ms_per_step"], j["ms_per_step_percentiles"], j["config"]["parallelism"][:160])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 scripts/micro/exchange_probe.py 2>&1 | grep -v "fused-barrier kernel only" | tail -22
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29574 scripts/check_pair_dist.py 2>&1 | tail -3
