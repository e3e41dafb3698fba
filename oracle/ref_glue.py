"""Restatement of the reference's Python glue around its native extensions (oracle/_ref).  TEST INFRASTRUCTURE ONLY.

The reference's wrappers (raymarching/raymarching.py, gridencoder/grid.py, shencoder/sphere_harmonics.py) cannot travel to
the GPU box (no /root/reference there) and must not be copied, so the little they do around each native call --
allocate outputs, zero-fill, permute -- is restated here, each function citing the lines it follows.  `ext` is the dict
of imported reference modules from tests/conftest.py::ref_ext.
"""
from __future__ import annotations

import numpy as np
import torch


def near_far_from_aabb(ext, rays_o, rays_d, aabb, min_near=0.2):
    # raymarching.py:40-50
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    ext["raymarching"].near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
    return nears, fars


def march_rays_train_raw(ext, rays_o, rays_d, bound, bitfield, C, H, nears, fars, M, perturb, dt_gamma=0.0, max_steps=1024):
    # raymarching.py:240-271: zero-filled outputs, counter starts at zero
    N = rays_o.shape[0]
    dev = rays_o.device
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    ext["raymarching"].march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs,
                                        dirs, deltas, rays, counter, int(perturb))
    return xyzs, dirs, deltas, rays, counter


def canonicalize(xyzs, dirs, deltas, rays, M_out=None):
    """Re-order a reference march result (atomic-arrival offsets) into canonical ray-id order.

    Returns (xyzs, dirs, deltas, rays) where rays[n] = (n, exclusive prefix sum of counts, count) and the sample rows of
    every ray are moved accordingly.  Rays the reference dropped for overflow must not occur (use a large M).
    """
    order = torch.argsort(rays[:, 0].long())
    r = rays[order].long()
    cnt = r[:, 2]
    can_off = torch.cumsum(cnt, 0) - cnt
    total = int(cnt.sum().item())
    ray_of = torch.repeat_interleave(torch.arange(r.shape[0], device=rays.device), cnt)
    within = torch.arange(total, device=rays.device) - can_off[ray_of]
    src = r[ray_of, 1] + within
    M_out = M_out or xyzs.shape[0]

    def move(a):
        out = torch.zeros((M_out,) + a.shape[1:], dtype=a.dtype, device=a.device)
        out[:total] = a[src]
        return out

    can_rays = torch.stack([r[:, 0], can_off, cnt], dim=1).int()
    return move(xyzs), move(dirs), move(deltas), can_rays


def composite_rays_train_forward(ext, sigmas, rgbs, deltas, rays):
    # raymarching.py:311-320
    M, N = sigmas.shape[0], rays.shape[0]
    ws = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
    depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
    image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
    ext["raymarching"].composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, ws, depth, image)
    return ws, depth, image


def composite_rays_train_backward(ext, grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image):
    # raymarching.py:339-355
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
    ext["raymarching"].composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, M, N, gs, gc)
    return gs, gc


def grid_encode_forward(ext, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                        align_corners=False):
    # grid.py:39-84 (embeddings already in the dtype to test)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    S = np.log2(per_level_scale)
    outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
    dy_dx = (torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs
             else torch.empty(1, device=inputs.device, dtype=embeddings.dtype))
    ext["gridencoder"].grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, base_resolution,
                                           calc_grad_inputs, dy_dx, gridtype, align_corners)
    return outputs.permute(1, 0, 2).reshape(B, L * C), dy_dx


def grid_encode_backward(ext, grad, inputs, embeddings, offsets, per_level_scale, base_resolution, dy_dx=None, gridtype=0,
                         align_corners=False):
    # grid.py:98-130
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    S = np.log2(per_level_scale)
    cgi = dy_dx is not None
    g = grad.view(B, L, C).permute(1, 0, 2).contiguous()
    ge = torch.zeros_like(embeddings)
    gi = torch.zeros_like(inputs, dtype=embeddings.dtype) if cgi else torch.zeros(1, device=inputs.device, dtype=embeddings.dtype)
    if not cgi:
        dy_dx = torch.empty(1, device=inputs.device, dtype=embeddings.dtype)
    ext["gridencoder"].grid_encode_backward(g, inputs, embeddings, offsets, ge, B, D, C, L, S, base_resolution, cgi, dy_dx, gi,
                                            gridtype, align_corners)
    return ge, (gi if cgi else None)


def sh_encode_forward(ext, inputs, degree, calc_grad_inputs=False):
    # sphere_harmonics.py:22-37
    B, D = inputs.shape
    out = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
    dy_dx = (torch.empty(B, D * degree ** 2, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs
             else torch.empty(1, dtype=inputs.dtype, device=inputs.device))
    ext["shencoder"].sh_encode_forward(inputs, out, B, D, degree, calc_grad_inputs, dy_dx)
    return out, (dy_dx if calc_grad_inputs else None)
