#!/bin/bash
# Collects the round's evidence on a B200 box into gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r01_bench_reference.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/r01_bench_reference.json
timeout 400 python bench.py > gpurun_out/r01_bench_ours.json 2> gpurun_out/bench_ours.err; python scripts/_show_bench.py 2>/dev/null | cut -c1-300
cp gpurun_out/r01_bench_ours.json gpurun_out/bench_ours.json
timeout 300 python bench.py --no-pipeline --no-cpu-baseline > gpurun_out/r01_bench_ours_serial.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_l.log 2>&1
PVD_TRACE=1 timeout 200 python scripts/trace_kernels.py > gpurun_out/r01_trace_timeline.txt 2>&1; tail -3 gpurun_out/r01_trace_timeline.txt
