#!/bin/bash
# Short end-of-round refresh (after the vm two-kernel backward): tests, the three workloads whose kernels changed, one ncu capture.
T=${1:-r01f}
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; tail -2 $O/${T}_pytest_gpu.log
timeout 150 python bench.py > $O/${T}_bench_ours.json 2> $O/bench_ours.err; tail -c 400 $O/${T}_bench_ours.json
for W in vm hash-vm; do
  timeout 120 python bench.py --workload $W --steps 100 --cpu-budget 6 > $O/${T}_bench_ours_${W}.json 2> $O/bench_ours_${W}.err; tail -c 250 $O/${T}_bench_ours_${W}.json
done
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:k_vm_field_fwd|k_vm_field_bwd|k_vm_scatter' -s 57 -c 3 -o $O/${T}_vm_full python bench.py --workload vm --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_fv.log 2>&1
ncu -i $O/${T}_vm_full.ncu-rep --page raw --csv > $O/${T}_vm_ncu_full_raw.csv 2>/dev/null
ls $O | grep ${T}
