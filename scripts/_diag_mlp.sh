timeout 900 python -m pytest tests/test_gpu_fp32_path.py tests/test_gpu_engine.py tests/test_gpu_optim.py -m gpu -q > gpurun_out/t_fp32.log 2>&1; tail -6 gpurun_out/t_fp32.log; grep -n "AssertionError\|Error:" gpurun_out/t_fp32.log | head
python bench.py --workload hash-fp32 --only --steps 50 --no-cpu-baseline 2>gpurun_out/b_f32_err.log | tail -1 > gpurun_out/b_f32.json
python -c "
import json; j=json.load(open('gpurun_out/b_f32.json')); print('hash-fp32 ours', j['ms_per_step'], j['kernel_ms'], j.get('iteration',{}).get('ms'), j.get('iteration_error'))"
tail -3 gpurun_out/b_f32_err.log
python bench.py --impl reference --workload hash-fp32 --only --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/b_f32_ref_err.log | tail -1 > gpurun_out/b_f32_ref.json
python -c "
import json; j=json.load(open('gpurun_out/b_f32_ref.json')); print('hash-fp32 ref', j.get('ms_per_step'), j.get('iteration'))"
tail -3 gpurun_out/b_f32_ref_err.log
