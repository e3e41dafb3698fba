"""Diagnostic: run the tcgen05 probe kernel (scripts/micro/tc_probe.cu -- NOT part of libpvd_b200.so) in all modes and report
errors / the TMEM lane mapping.  Builds its own small shared object next to this file.
Usage on the GPU box: python scripts/micro/tc_probe.py"""
import ctypes as C
import os
import sys

import torch

import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "aaai2023-pvd_b200"))
from pvd_b200 import _native as nv  # noqa: E402

_SO = os.path.join(HERE, "libtc_probe.so")
if not os.path.exists(_SO):
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                           f"-I{os.path.join(ROOT, 'aaai2023-pvd_b200', 'csrc')}", f"-I{os.path.join(ROOT, 'include')}", "-DPVD_BUILDING",
                           os.path.join(HERE, "tc_probe.cu"), "-o", _SO])
_probe = C.CDLL(_SO)


def probe(mode, A, B, N):
    out = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = _probe.pvd_tc_probe(C.c_int(mode), nv.ptr(A), C.c_uint32(A.shape[0]), C.c_uint32(A.shape[1]), nv.ptr(B),
                               C.c_uint32(B.shape[0]), C.c_uint32(B.shape[1]), nv.ptr(out), C.c_uint32(N), nv.ptr(status),
                               nv.stream_of(A))
    nv.check(rc)
    torch.cuda.synchronize()
    return out, int(status.item())


def lane_map(out, ref):
    """for each expected row find the dumped lane that matches"""
    m = []
    for r in range(ref.shape[0]):
        d = (out - ref[r][None]).abs().amax(dim=1)
        j = int(torch.argmin(torch.nan_to_num(d, nan=1e9)))
        m.append((j, float(d[j])))
    return m


def main():
    torch.manual_seed(0)
    res = {}
    for K, N in ((32, 64), (64, 16), (64, 64), (32, 16)):
        A = (torch.randn(128, K, device="cuda") * 0.5).half()
        B = (torch.randn(N, K, device="cuda") * 0.5).half()
        out, st = probe(0, A, B, N)
        ref = A.float() @ B.float().T
        err = (out - ref).abs().max().item()
        print(f"mode0 K={K} N={N}: status={st} max_err={err:.4e} ref_max={ref.abs().max().item():.3f}")
        res[f"m0_{K}_{N}"] = err
    for K, N in ((64, 32), (16, 64), (64, 64)):
        A = (torch.randn(128, K, device="cuda") * 0.5).half()
        W = (torch.randn(K, N, device="cuda") * 0.5).half()
        out, st = probe(1, A, W, N)
        ref = A.float() @ W.float()
        err = (out - ref).abs().max().item()
        print(f"mode1 K={K} N={N}: status={st} max_err={err:.4e}")
        res[f"m1_{K}_{N}"] = err
    for C1, N in ((64, 32), (64, 16), (64, 64)):
        T1 = (torch.randn(128, C1, device="cuda") * 0.5).half()
        T2 = (torch.randn(128, N, device="cuda") * 0.5).half()
        out, st = probe(2, T1, T2, N)
        ref = T1.float().T @ T2.float()
        lm = lane_map(out, ref)
        worst = max(e for _, e in lm)
        print(f"mode2 C1={C1} N={N}: status={st} lane-matched max_err={worst:.4e}")
        print("   row->lane:", [j for j, _ in lm])
    T1 = (torch.randn(128, 128, device="cuda") * 0.5).half()
    T2 = (torch.randn(128, 32, device="cuda") * 0.5).half()
    out, st = probe(3, T1, T2, 32)
    ref = T1.float().T @ T2.float()
    print(f"mode3: status={st} max_err={(out - ref).abs().max().item():.4e}")


if __name__ == "__main__":
    main()
