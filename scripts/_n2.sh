timeout 600 python -m pytest tests/test_gpu_pair_engine.py tests/test_gpu_fp32_path.py -m gpu -q > gpurun_out/t_pair.log 2>&1; tail -5 gpurun_out/t_pair.log; grep -n "AssertionError\|Error:" gpurun_out/t_pair.log | head -5
bash scripts/_scale.sh 2 2>&1 | grep -v "fused-barrier kernel only"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02b_bench_ours_N2.json").read().strip().splitlines()[-1])
print("percentiles", j["ms_per_step_percentiles"], "hash-fp32" in j["workloads"], "mlp" in j["workloads"])
PY
