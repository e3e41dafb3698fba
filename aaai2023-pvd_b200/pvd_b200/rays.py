"""On-device ray generation: `get_rays` of the reference (distill_mutual/utils.py:324-404, same signature and result dict) with the
meshgrid / gather / normalise / matmul chain (~12 torch launches and two [B, H*W] temporaries per call) replaced by one kernel
(csrc/field_tensors.cu::k_get_rays).  The random pixel choice is the reference's own torch call (torch.randint / torch.multinomial),
so a seeded run draws the same pixels."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nv


def get_rays(poses, intrinsics, H, W, N=-1, error_map=None):
    """poses [B, 4, 4] cam2world (CUDA), intrinsics (fx, fy, cx, cy) -> {"rays_o", "rays_d" [B, N, 3], "inds" [B, N] (when N > 0),
    "inds_coarse" (with an error_map)}."""
    device = poses.device
    if not poses.is_cuda:
        raise RuntimeError("pvd_b200.rays.get_rays: poses must be a CUDA tensor (there is no CPU path)")
    B = poses.shape[0]
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    results = {}
    inds = None
    stride = 0
    if N > 0:
        N = min(N, H * W)
        if error_map is None:
            row = torch.randint(0, H * W, size=[N], device=device)  # may duplicate (utils.py:354)
            inds = row.expand([B, N])
            inds_k, stride = row.contiguous(), 0                     # one row shared by every pose
        else:
            inds_coarse = torch.multinomial(error_map.to(device), N, replacement=False)  # [B, N] in [0, 128*128)
            inds_x, inds_y = inds_coarse // 128, inds_coarse % 128
            sx, sy = H / 128, W / 128
            inds_x = (inds_x * sx + torch.rand(B, N, device=device) * sx).long().clamp(max=H - 1)
            inds_y = (inds_y * sy + torch.rand(B, N, device=device) * sy).long().clamp(max=W - 1)
            inds = inds_x * W + inds_y
            results["inds_coarse"] = inds_coarse
            inds_k, stride = inds.contiguous(), N
        results["inds"] = inds
        n = N
    else:
        inds_k, n = None, H * W
    P = poses.detach().float().contiguous()
    rays_o = torch.empty(B, n, 3, dtype=torch.float32, device=device)
    rays_d = torch.empty(B, n, 3, dtype=torch.float32, device=device)
    with nv.on_device(P):
        nv.check(nv.lib().pvd_get_rays(nv.ptr(P), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_uint32(H), C.c_uint32(W),
                                       nv.ptr(inds_k), C.c_uint32(stride), C.c_uint32(B), C.c_uint32(n), nv.ptr(rays_o), nv.ptr(rays_d),
                                       nv.stream_of(P)))
    results["rays_o"] = rays_o
    results["rays_d"] = rays_d
    return results
