import json,sys
d=json.loads(open("gpurun_out/bench_ours.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","serial_ms_per_step","e2e","kernel_ms","roofline")})
