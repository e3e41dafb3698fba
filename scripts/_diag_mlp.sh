timeout 900 python -m pytest tests/test_gpu_fp32_path.py tests/test_gpu_network_golden.py tests/test_gpu_fused_mlp.py -m gpu -q > gpurun_out/t_fp32.log 2>&1; tail -30 gpurun_out/t_fp32.log
python - <<'PY'
import sys, time, torch
sys.path[:0] = ["aaai2023-pvd_b200", "."]
from pvd_b200.fused import HashNeRFField
torch.manual_seed(0)
M = 73000
x = (torch.rand(M, 3, device="cuda") * 2 - 1) * 0.99
d = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda"), dim=-1)
for fp32 in (True, False):
    net = HashNeRFField(num_levels=14, desired_resolution=2048, fp32=fp32, table_fp16=False).cuda()
    net.train()
    def step():
        net.zero_grad(set_to_none=True)
        s, c = net(x, d)
        (s.sum() * 1e-3 + c.sum()).backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print("hash field fwd+bwd through autograd, 73k samples, fp32 =", fp32, ":", e0.elapsed_time(e1) / 10, "ms")
PY
