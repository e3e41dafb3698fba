#!/bin/bash
# Collects the round's evidence on a B200 box into gpurun_out/ (copied to profiles/ afterwards by scripts/ncu_summary.py and by hand).
# usage: bash scripts/collect_profiles.sh [tag]     (tag defaults to r01)
T=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/${T}_bench_reference.json
timeout 400 python bench.py > gpurun_out/${T}_bench_ours.json 2> gpurun_out/bench_ours.err; cp gpurun_out/${T}_bench_ours.json gpurun_out/bench_ours.json; python scripts/_show_bench.py 2>/dev/null | cut -c1-600
# launch list of the bench command itself (graph nodes are profiled one by one: cold-cache, serialised -> compare SHARES)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/${T}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
# one full capture of each kernel of the step (eager launches of the same step; 19 sizing/warm-up steps skipped)
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_hash_field_bwd|k_hash_scatter|k_hash_field_fwd|k_composite_train_mse|k_march_count' -s 100 -c 5 -o gpurun_out/${T}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
