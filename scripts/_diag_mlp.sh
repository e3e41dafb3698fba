timeout 600 python -m pytest tests/test_gpu_fused_mlp.py -m gpu -q > gpurun_out/t_mlp.log 2>&1; tail -5 gpurun_out/t_mlp.log
for d in 0 7 16 23 8; do
PVD_MLP_DIAG=$d python bench.py --workload mlp-hash --only --steps 30 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('diag $d', j['ms_per_step'], j['kernel_ms']['teacher_fwd'])"
done
