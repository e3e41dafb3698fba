"""The reference's training step, driven through ITS OWN compiled CUDA extensions (oracle/_ref).  TEST / BENCH INFRASTRUCTURE ONLY.

`/root/reference` cannot travel to the GPU box and its sources must not be copied, so the thin Python layer the reference
wraps around its native kernels is restated here, each piece citing the lines it follows:

    _RefGridEncode     gridencoder/grid.py:20-139        (half cast of the whole table per call, [L,B,C] + permute, zeros_like)
    _RefSH             shencoder/sphere_harmonics.py:15-64
    _RefTruncExp       tools/activation.py:6-21
    _RefComposite      raymarching/raymarching.py:292-360
    ref_march          raymarching/raymarching.py:176-289  (zero-filled M-row buffers, mean_count sizing)
    RefHashNetwork     distill_mutual/network.py:103-152,335-343,413-437  (hash branch: nn.Linear stacks under autocast)
    RefTrainer.step    distill_mutual/renderer.py:359-448 + just_train_tea/utils.py:588-606,841-846 (autocast, MSE, GradScaler)

Everything numerical (ray marching, hash gather/scatter, SH, compositing) runs in the reference's kernels, the GEMMs in
cuBLAS through F.linear, exactly as in the reference.  This is the "reference extensions rebuilt for sm_100a" baseline of
BASELINE.md section 2a.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.amp import custom_bwd, custom_fwd

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_ext():
    d = os.path.join(_HERE, "_ref")
    if d not in sys.path:
        sys.path.insert(0, d)
    import _gridencoder
    import _raymarching
    import _shencoder
    return {"raymarching": _raymarching, "gridencoder": _gridencoder, "shencoder": _shencoder}


def make_ops(ext):
    rm, ge, sh = ext["raymarching"], ext["gridencoder"], ext["shencoder"]

    class _RefGridEncode(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda")
        def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution):
            inputs = inputs.contiguous()
            B, D = inputs.shape
            L, C = offsets.shape[0] - 1, embeddings.shape[1]
            S = np.log2(per_level_scale)
            if torch.is_autocast_enabled("cuda") and C % 2 == 0:
                embeddings = embeddings.to(torch.half)  # grid.py:51-52: whole table, every forward
            outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
            dy_dx = torch.empty(1, device=inputs.device, dtype=embeddings.dtype)
            ge.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, base_resolution, False, dy_dx, 0, False)
            outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
            ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
            ctx.dims = [B, D, C, L, S, base_resolution]
            return outputs

        @staticmethod
        @custom_bwd(device_type="cuda")
        def backward(ctx, grad):
            inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
            B, D, C, L, S, H = ctx.dims
            grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
            grad_embeddings = torch.zeros_like(embeddings)  # grid.py:106
            grad_inputs = torch.zeros(1, device=inputs.device, dtype=embeddings.dtype)
            ge.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, False, dy_dx, grad_inputs,
                                    0, False)
            return None, grad_embeddings, None, None, None

    class _RefSH(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
        def forward(ctx, inputs, degree):
            inputs = inputs.contiguous()
            B, D = inputs.shape
            outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
            dy_dx = torch.empty(1, dtype=inputs.dtype, device=inputs.device)
            sh.sh_encode_forward(inputs, outputs, B, D, degree, False, dy_dx)
            return outputs

        @staticmethod
        def backward(ctx, grad):
            return None, None

    class _RefTruncExp(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
        def forward(ctx, x):
            ctx.save_for_backward(x)
            return torch.exp(x)

        @staticmethod
        @custom_bwd(device_type="cuda")
        def backward(ctx, g):
            return g * torch.exp(ctx.saved_tensors[0].clamp(-12, 12))

    class _RefComposite(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
        def forward(ctx, sigmas, rgbs, deltas, rays):
            sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
            M, N = sigmas.shape[0], rays.shape[0]
            ws = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
            depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
            image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
            rm.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, ws, depth, image)
            ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, depth, image)
            ctx.dims = [M, N]
            return ws, depth, image

        @staticmethod
        @custom_bwd(device_type="cuda")
        def backward(ctx, gws, gdepth, gimage):
            gws, gimage = gws.contiguous(), gimage.contiguous()
            sigmas, rgbs, deltas, rays, ws, depth, image = ctx.saved_tensors
            M, N = ctx.dims
            gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
            rm.composite_rays_train_backward(gws, gimage, sigmas, rgbs, deltas, rays, ws, image, M, N, gs, gc)
            return gs, gc, None, None

    return _RefGridEncode, _RefSH, _RefTruncExp, _RefComposite


class RefHashNetwork(nn.Module):
    """The hash branch of the reference's NeRFNetwork (network.py:47-51,103-152,335-343,413-437)."""

    def __init__(self, ext, offsets, per_level_scale, base_resolution=16, bound=1.0, clip_min=-2.0, clip_max=7.0):
        super().__init__()
        self.ops = make_ops(ext)
        self.register_buffer("offsets", torch.as_tensor(offsets, dtype=torch.int32))
        self.per_level_scale, self.base_resolution, self.bound = per_level_scale, base_resolution, bound
        self.clip_min, self.clip_max = clip_min, clip_max
        L = len(offsets) - 1
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), 2).uniform_(-1e-4, 1e-4))
        self.sigma_net = nn.ModuleList([nn.Linear(2 * L, 64, bias=False), nn.Linear(64, 16, bias=False)])
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 3, bias=False)])

    def forward(self, x, d):
        GridEncode, SH, TruncExp, _ = self.ops
        x = (x + self.bound) / (2 * self.bound)
        h = GridEncode.apply(x.view(-1, 3), self.embeddings, self.offsets, self.per_level_scale, self.base_resolution)
        for l in range(2):
            h = self.sigma_net[l](h)
            if l != 1:
                h = F.relu(h, inplace=True)
        h[..., 0] = torch.clamp(h[..., 0].clone(), self.clip_min, self.clip_max)
        sigma = TruncExp.apply(h[..., 0])
        geo_feat = h[..., 1:]
        d = SH.apply(d.reshape(-1, 3), 4)
        h = torch.cat([d, geo_feat], dim=-1)
        for l in range(3):
            h = self.color_net[l](h)
            if l != 2:
                h = F.relu(h, inplace=True)
        return sigma, torch.sigmoid(h)


class RefTrainer:
    """run_cuda (training branch) + MSE + scaled backward, as the reference's train_one_epoch does per iteration."""

    def __init__(self, ext, net: RefHashNetwork, bitfield, bound=1.0, cascade=1, grid_size=128, min_near=0.2, max_steps=1024,
                 loss_scale=65536.0):
        self.ext, self.net = ext, net
        self.bitfield = bitfield
        self.bound, self.cascade, self.grid_size, self.min_near, self.max_steps = bound, cascade, grid_size, min_near, max_steps
        dev = bitfield.device
        self.aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32, device=dev)  # renderer.py:110-113
        self.mean_count, self.local_step = 0, 0
        self.loss_scale = loss_scale

    def march(self, rays_o, rays_d, nears, fars, counter, perturb=True):
        rm = self.ext["raymarching"]
        N = rays_o.shape[0]
        M = N * self.max_steps
        if self.mean_count > 0:  # raymarching.py:235-238
            M = self.mean_count + (128 - self.mean_count % 128)
        dev = rays_o.device
        xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        rm.march_rays_train(rays_o, rays_d, self.bitfield, self.bound, 0.0, self.max_steps, N, self.cascade, self.grid_size, M, nears,
                            fars, xyzs, dirs, deltas, rays, counter, int(perturb))
        if self.mean_count <= 0:  # raymarching.py:276-284
            m = counter[0].item()
            m += 128 - m % 128
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
            torch.cuda.empty_cache()
        return xyzs, dirs, deltas, rays

    def step(self, rays_o, rays_d, gt, bg_color=1.0):
        rm = self.ext["raymarching"]
        Composite = self.net.ops[3]
        for p in self.net.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16):  # fp16=True is forced by both CLIs (main_distill_mutual.py:251-254)
            N = rays_o.shape[0]
            nears = torch.empty(N, device=rays_o.device)
            fars = torch.empty(N, device=rays_o.device)
            rm.near_far_from_aabb(rays_o, rays_d, self.aabb, N, self.min_near, nears, fars)
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = self.march(rays_o, rays_d, nears, fars, counter)
            sigmas, rgbs = self.net(xyzs, dirs)
            ws, depth, image = Composite.apply(sigmas, rgbs, deltas, rays)
            image = image + (1 - ws).unsqueeze(-1) * bg_color
            depth = torch.clamp(depth - nears, min=0) / (fars - nears + 1e-6)
            loss = torch.mean((image - gt) ** 2)
        (loss * self.loss_scale).backward()
        return loss

    def update_mean_count(self):
        total_step = min(16, self.local_step)  # renderer.py:768-773
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0
