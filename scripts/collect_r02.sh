#!/bin/bash
# Round-2 evidence on ONE B200 (writes gpurun_out/; summarised into profiles/ afterwards by scripts/ncu_summary.py).
#   bash scripts/collect_r02.sh [tag]
T=${1:-r02}
O=gpurun_out
mkdir -p $O
# launch list of the bench command itself (graph nodes are profiled one by one: cold-cache, serialised -> compare SHARES, not times)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/${T}_ncu_launches.csv \
    python bench.py --steps 3 --warmup 3 --only --no-cpu-baseline > $O/ncu_l.log 2>&1
# one full capture of each kernel of the hash step + the optimizer (eager launches of the same step)
timeout 600 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_hash_field_bwd|k_hash_field_fwd|k_composite_train_mse|k_march_count|k_adamw_multi|k_grad_nonfinite_multi' -s 80 -c 8 -o $O/${T}_full \
    python bench.py --steps 2 --warmup 3 --only --no-cpu-baseline --no-graph > $O/ncu_f.log 2>&1
ncu -i $O/${T}_full.ncu-rep --page raw --csv > $O/${T}_ncu_full_raw.csv 2>/dev/null
# the TMA-staged and the per-thread-load variant of the forward, same capture settings (N2 evidence)
PVD_FWD_TMA=0 timeout 300 ncu --set full --clock-control none -k 'regex:k_hash_field_fwd' -s 20 -c 1 -o $O/${T}_fwd_notma \
    python bench.py --steps 2 --warmup 3 --only --no-cpu-baseline --no-graph > $O/ncu_fn.log 2>&1
ncu -i $O/${T}_fwd_notma.ncu-rep --page raw --csv > $O/${T}_fwd_notma_ncu_full_raw.csv 2>/dev/null
# vm + pair kernels
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:k_vm_field_fwd|k_vm_field_bwd|k_vm_scatter|k_pair_sample_sq|k_pair_composite|k_pair_combine' \
    -s 120 -c 6 -o $O/${T}_pair_full python bench.py --workload hash-vm --steps 2 --warmup 3 --only --no-cpu-baseline --no-graph > $O/ncu_fp.log 2>&1
ncu -i $O/${T}_pair_full.ncu-rep --page raw --csv > $O/${T}_pair_ncu_full_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_mlp_field_fwd' -s 20 -c 1 -o $O/${T}_mlp_full \
    python bench.py --workload mlp-hash --steps 2 --warmup 3 --only --no-cpu-baseline --no-graph > $O/ncu_fm.log 2>&1
ncu -i $O/${T}_mlp_full.ncu-rep --page raw --csv > $O/${T}_mlp_ncu_full_raw.csv 2>/dev/null
# the NeRF-MLP training kernels: forward that saves its tiles, trunk data gradients, tensor-core weight gradients
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:k_mlp_field_fwd|k_mlp_trunk_bwd|k_mlp_wgrad' -s 30 -c 3 -o $O/${T}_mlptrain_full \
    python bench.py --workload mlp --steps 2 --warmup 3 --only --no-cpu-baseline --no-graph > $O/ncu_fmt.log 2>&1
ncu -i $O/${T}_mlptrain_full.ncu-rep --page raw --csv > $O/${T}_mlptrain_ncu_full_raw.csv 2>/dev/null
ls -la $O | grep ${T}_ | tail -20
