"""bench.py's contract on a box without a GPU: the reference arm falls back to the CPU oracle port and still prints ONE JSON line
with the keys the driver reads; our arm refuses to run (there is no CPU path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="CPU-fallback contract; on a GPU box the reference arm runs the reference extensions")
@pytest.mark.timeout(240)
def test_reference_arm_falls_back_to_the_cpu_port():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3", "--rays", "512",
                        "--cpu-budget", "1"], capture_output=True, text=True, timeout=220, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "training rays/sec (fwd+bwd)" and d["unit"] == "rays/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None


@pytest.mark.skipif(_has_cuda(), reason="needs a box without a GPU")
def test_our_arm_has_no_cpu_path():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_roofline_tables_cover_every_workload():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.WORKLOADS) == {"hash", "vm", "hash-vm", "mlp-hash", "mlp", "hash-fp32"} == set(bench.DEFAULT_RAYS) == set(bench.ALL_WORKLOADS)
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
    for k in ("k_hash_field_fwd", "k_hash_field_bwd", "k_vm_field_fwd", "k_vm_field_bwd", "k_vm_scatter", "k_mlp_field_fwd", "k_pair_composite"):
        assert k in t and t[k]["dram_bytes"] > 0 and t[k]["source"].startswith("profiles/") and os.path.exists(os.path.join(ROOT, t[k]["source"])), k
    tr, src = bench.ncu_traffic("k_vm_field_bwd+k_vm_scatter")
    assert tr == t["k_vm_field_bwd"]["dram_bytes"] + t["k_vm_scatter"]["dram_bytes"] and src
    assert bench.ncu_traffic("k_not_a_kernel") == (None, None)
