#!/bin/bash
# Collects the round's evidence on a B200 box into gpurun_out/ (summarised into profiles/ afterwards by scripts/ncu_summary.py).
# usage: bash scripts/collect_profiles.sh [tag] [quick]     (tag defaults to r01)
T=${1:-r01}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; tail -3 $O/${T}_pytest_gpu.log
# headline (BASELINE configs[1]): reference arm, then ours
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2> $O/bench_ref.err; tail -c 300 $O/${T}_bench_reference.json
timeout 400 python bench.py > $O/${T}_bench_ours.json 2> $O/bench_ours.err; cp $O/${T}_bench_ours.json $O/bench_ours.json; python scripts/_show_bench.py 2>/dev/null | cut -c1-700
# the other BASELINE configurations through the same harness
for W in vm hash-vm mlp-hash; do
  timeout 300 python bench.py --workload $W --impl reference --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_reference_${W}.json 2> $O/bench_ref_${W}.err; tail -c 200 $O/bench_ref_${W}.err
  timeout 400 python bench.py --workload $W --steps 100 --cpu-budget 8 > $O/${T}_bench_ours_${W}.json 2> $O/bench_ours_${W}.err; tail -c 200 $O/bench_ours_${W}.err
  python - <<PY
import json
for arm in ("reference", "ours"):
    try:
        d = json.loads(open("$O/${T}_bench_%s_${W}.json" % arm).read().strip().splitlines()[-1])
        print("${W}", arm, round(d["value"]), "rays/s", round(d["ms_per_step"], 4), "ms", d.get("kernel_ms"))
    except Exception as e:
        print("${W}", arm, "FAILED", e)
PY
done
[ "$2" = "quick" ] && exit 0
# launch list of the bench command itself (graph nodes are profiled one by one: cold-cache, serialised -> compare SHARES)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file $O/${T}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_l.log 2>&1
# one full capture of each kernel of the step (eager launches of the same step; 19 sizing/warm-up steps x 4 kernels skipped)
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_hash_field_bwd|k_hash_field_fwd|k_composite_train_mse|k_march_count' -s 76 -c 4 -o $O/${T}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_f.log 2>&1
ncu -i $O/${T}_full.ncu-rep --page raw --csv > $O/${T}_ncu_full_raw.csv 2>/dev/null
# the two-launch variant of the backward (PVD_SPLIT_SCATTER=1): the MLP backward and the stand-alone scatter kernel
PVD_SPLIT_SCATTER=1 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_hash_field_bwd|k_hash_scatter' -s 38 -c 2 -o $O/${T}_split_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_fs.log 2>&1
ncu -i $O/${T}_split_full.ncu-rep --page raw --csv > $O/${T}_split_ncu_full_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_vm_field_fwd|k_vm_field_bwd|k_pair_sample_sq|k_pair_composite|k_pair_combine' -s 100 -c 5 -o $O/${T}_pair_full python bench.py --workload hash-vm --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_fp.log 2>&1
ncu -i $O/${T}_pair_full.ncu-rep --page raw --csv > $O/${T}_pair_ncu_full_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_mlp_field_fwd' -s 20 -c 1 -o $O/${T}_mlp_full python bench.py --workload mlp-hash --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_fm.log 2>&1
ncu -i $O/${T}_mlp_full.ncu-rep --page raw --csv > $O/${T}_mlp_ncu_full_raw.csv 2>/dev/null
ls -la $O | tail -30
