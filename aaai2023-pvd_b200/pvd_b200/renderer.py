"""`NeRFRenderer` with the reference's contract (distill_mutual/renderer.py:66-814) on top of the B200 operators.

Kept identical for callers (`Trainer.train_step`, distill_mutual/utils.py:998-1018):
  * `render(rays_o, rays_d, staged, bg_color, perturb, force_all_rays, inherited_params, **opt)` -> dict with keys
    `depth, image, inherited_params, sigmas, rays` (+ `stage1` / `stage2` markers), renderer.py:421-438,546-559;
  * the student marches and the teacher re-uses the SAME samples through `inherited_params` (renderer.py:374-394);
  * buffers `aabb_train, aabb_infer, density_grid, density_bitfield, step_counter` live in the state_dict (renderer.py:92-113);
  * `update_extra_state` / `mark_untrained_grid` / `reset_extra_state` (renderer.py:127-137,561-773).
The dead pure-PyTorch `run` path of the reference (it calls `self.color`, which asserts False, network.py:516) is not rebuilt:
`render` requires `cuda_ray=True`, which both CLIs force (main_distill_mutual.py:251-254).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn as nn

import raymarching


def _meshgrid_ij(*xs):
    return torch.meshgrid(*xs, indexing="ij")


class NeRFRenderer(nn.Module):
    def __init__(self, bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1, grid_size=128):
        super().__init__()
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = grid_size
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        if bg_radius > 0:
            # distill_mutual/renderer.py:350-355 (polar_from_ray + self.background) is not built: the reference's own background
            # model asserts (`assert 1 == 2`), both CLIs keep bg_radius = -1.  Refuse rather than silently render on white.
            raise NotImplementedError("bg_radius > 0 (spherical background model) is not supported by the B200 renderer")
        self.bg_radius = bg_radius
        aabb = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer("aabb_train", aabb)
        self.register_buffer("aabb_infer", aabb.clone())
        self.cuda_ray = cuda_ray
        if cuda_ray:
            self.register_buffer("density_grid", torch.zeros([self.cascade, grid_size ** 3]))
            self.register_buffer("density_bitfield", torch.zeros(self.cascade * grid_size ** 3 // 8, dtype=torch.uint8))
            self.mean_density = 0
            self.iter_density = 0
            self.register_buffer("step_counter", torch.zeros(16, 2, dtype=torch.int32))
            self.mean_count = 0
            self.local_step = 0

    # subclasses provide the field
    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    # ------------------------------------------------------------------------------------------ rendering
    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 inherited_params=[], **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer,
                                                     self.min_near)
        if bg_color is None:
            bg_color = 1
        args = getattr(self, "args", None)
        stage_iters = getattr(args, "stage_iters", {"stage1": -1, "stage2": -1})
        global_step = getattr(args, "global_step", 10 ** 9)
        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            stu_first = getattr(args, "render_stu_first", True)
            is_teacher = getattr(self, "is_teacher", False)
            marches = (not is_teacher) if stu_first else is_teacher
            if marches or len(inherited_params) == 0:
                xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield,
                                                                        self.cascade, self.grid_size, nears, fars, counter,
                                                                        self.mean_count, perturb, 128, force_all_rays, dt_gamma,
                                                                        max_steps)
                inherited_params = [xyzs, dirs, deltas, rays]
            else:
                xyzs, dirs, deltas, rays = inherited_params
            sigmas, rgbs = self(xyzs, dirs)
            if global_step < stage_iters["stage1"] or global_step < stage_iters["stage2"]:
                key = "stage1" if global_step < stage_iters["stage1"] else "stage2"
                return {key: global_step, "depth": None, "image": None, "inherited_params": inherited_params, "sigmas": sigmas,
                        "rays": rays}
            sigmas = self.density_scale * sigmas
            weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays)
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
            depth = torch.clamp(depth - nears, min=0) / (fars - nears + 1e-6)
            return {"depth": depth.view(*prefix), "image": image.view(*prefix, 3), "inherited_params": inherited_params,
                    "sigmas": sigmas, "rays": rays}

        # inference.  A hash field renders through ONE persistent kernel (march + field + composite fused, rays pulled from a device
        # queue: csrc/field_hash.cu::k_hash_render_persistent) -- same per-ray results as the host loop below, no host round trips.
        # PVD_PERSISTENT_INFER=0, a jittered march (perturb) or another field type take the reference's loop (renderer.py:450-543).
        if (not perturb and os.environ.get("PVD_PERSISTENT_INFER", "1") != "0" and getattr(self, "model_type", None) == "hash"
                and hasattr(self, "render_persistent")):
            weights_sum, depth, image = self.render_persistent(rays_o, rays_d, nears, fars, dt_gamma, max_steps)
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
            depth = torch.clamp(depth - nears, min=0) / (fars - nears)
            return {"depth": depth.view(*prefix), "image": image.view(*prefix, 3), "inherited_params": inherited_params}
        # march a few steps per alive ray, composite in place, compact (renderer.py:450-543)
        weights_sum = torch.zeros(N, dtype=torch.float32, device=device)
        depth = torch.zeros(N, dtype=torch.float32, device=device)
        image = torch.zeros(N, 3, dtype=torch.float32, device=device)
        n_alive = N
        alive_counter = torch.zeros([1], dtype=torch.int32, device=device)
        rays_alive = torch.zeros(2, n_alive, dtype=torch.int32, device=device)
        rays_t = torch.zeros(2, n_alive, dtype=torch.float32, device=device)
        step, i = 0, 0
        while step < max_steps:
            if step == 0:
                torch.arange(n_alive, out=rays_alive[0])
                rays_t[0] = nears
            else:
                alive_counter.zero_()
                raymarching.compact_rays(n_alive, rays_alive[i % 2], rays_alive[(i + 1) % 2], rays_t[i % 2], rays_t[(i + 1) % 2],
                                         alive_counter)
                n_alive = alive_counter.item()
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], rays_o, rays_d, self.bound,
                                                        self.density_bitfield, self.cascade, self.grid_size, nears, fars, 128,
                                                        perturb, dt_gamma, max_steps)
            sigmas, rgbs = self(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            raymarching.composite_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], sigmas, rgbs, deltas, weights_sum, depth,
                                       image)
            step += n_step
            i += 1
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {"depth": depth.view(*prefix), "image": image.view(*prefix, 3), "inherited_params": inherited_params}

    def _update_extra_state_fused(self, decay):
        """update_extra_state through pvd_density_grid_points / _update / pvd_packbits_mean (include/pvd_b200.h): per cascade one
        kernel builds the jittered query points of all cells, the field's density kernel evaluates them, one kernel applies the EMA
        rule and accumulates the mean, one packs the bitfield with the threshold formed on the device.  The random numbers are the
        reference's (torch.rand_like / torch.randint, same shapes, same order), so a run with the same seed visits the same points."""
        from . import _native as nv
        l = nv.lib()
        dev = self.density_grid.device
        H = self.grid_size
        n_cells = H ** 3
        grid = self.density_grid
        if not hasattr(self, "_upkeep_sum") or self._upkeep_sum.device != dev:
            self._upkeep_sum = torch.zeros(1, dtype=torch.float64, device=dev)
            self._upkeep_tmp = torch.empty(n_cells, dtype=torch.float32, device=dev)
            self._mean_density_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self._upkeep_sum.zero_()
        full = self.iter_density < 16
        with nv.on_device(grid):
            st = nv.stream_of(grid)
            for cas in range(self.cascade):
                bound = float(min(2 ** cas, self.bound))
                if full:
                    n, indices = n_cells, None
                    # the reference draws rand_like over the meshgrid-ordered points; here element j belongs to Morton cell j --
                    # the same distribution, one draw per cell
                    noise = torch.rand(n, 3, device=dev)
                else:
                    n = n_cells // 4
                    coords = torch.randint(0, H, (n, 3), device=dev)
                    indices = raymarching.morton3D(coords)
                    occ = torch.nonzero(grid[cas] > 0).squeeze(-1)
                    if occ.numel() > 0:   # the one size the host must know (the reference reads it the same way, renderer.py:716-727)
                        occ = occ[torch.randint(0, occ.shape[0], [n], dtype=torch.long, device=dev)]
                        indices = torch.cat([indices, occ.to(indices.dtype)], dim=0)
                    indices = indices.to(torch.int32).contiguous()
                    n = indices.shape[0]
                    noise = torch.rand(n, 3, device=dev)
                xyzs = torch.empty(n, 3, device=dev)
                nv.check(l.pvd_density_grid_points(nv.ptr(indices), nv.ptr(noise), C.c_uint32(n), C.c_uint32(H), C.c_float(bound),
                                                   nv.ptr(xyzs), st))
                sig = self.density(xyzs)["sigma"].reshape(-1).detach().float().contiguous()
                g = grid[cas]
                nv.check(l.pvd_density_grid_update(nv.ptr(g), nv.ptr(self._upkeep_tmp), nv.ptr(indices), nv.ptr(sig), C.c_uint32(n),
                                                   C.c_uint32(n_cells), C.c_float(self.density_scale), C.c_float(decay),
                                                   nv.ptr(self._upkeep_sum), st))
            nv.check(l.pvd_packbits_mean(nv.ptr(grid), C.c_uint32(self.cascade * n_cells // 8), nv.ptr(self._upkeep_sum),
                                         C.c_uint32(self.cascade * n_cells), C.c_float(self.density_thresh),
                                         nv.ptr(self.density_bitfield), nv.ptr(self._mean_density_dev), st))
        self._mean_density = None      # read lazily from the device (see the property)
        self.iter_density += 1
        total = min(16, self.local_step)
        if total > 0:
            self.mean_count = int(self.step_counter[:total, 0].sum().item() / total)
        self.local_step = 0

    @property
    def mean_density(self):
        """renderer.py:750-752.  After a fused update the value lives on the device and is fetched only when somebody asks."""
        if getattr(self, "_mean_density", 0) is None:
            self._mean_density = float(self._mean_density_dev.item())
        return getattr(self, "_mean_density", 0)

    @mean_density.setter
    def mean_density(self, v):
        self._mean_density = v

    def render(self, rays_o, rays_d, staged=False, max_ray_batch=4096, **kwargs):
        if not self.cuda_ray:
            raise RuntimeError("only the cuda_ray path exists (the reference's pure-PyTorch `run` is dead code: network.py:516)")
        return self.run_cuda(rays_o, rays_d, **kwargs)

    # ------------------------------------------------------------------------------------------ density-grid upkeep
    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """Cells that no training camera sees get density -1 (renderer.py:561-643)."""
        if not self.cuda_ray:
            return
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        B = poses.shape[0]
        fx, fy, cx, cy = intrinsic
        dev = self.density_grid.device
        axis = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        count = torch.zeros_like(self.density_grid)
        poses = poses.to(dev)
        for xs in axis:
            for ys in axis:
                for zs in axis:
                    xx, yy, zz = _meshgrid_ij(xs, ys, zs)
                    coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
                    indices = raymarching.morton3D(coords).long()
                    world = (2 * coords.float() / (self.grid_size - 1) - 1).unsqueeze(0)
                    for cas in range(self.cascade):
                        bound = min(2 ** cas, self.bound)
                        half = bound / self.grid_size
                        cas_world = world * (bound - half)
                        head = 0
                        while head < B:
                            tail = min(head + S, B)
                            cam = cas_world - poses[head:tail, :3, 3].unsqueeze(1)
                            cam = cam @ poses[head:tail, :3, :3]
                            m = (cam[:, :, 2] > 0) & (cam[:, :, 0].abs() < cx / fx * cam[:, :, 2] + half * 2) & \
                                (cam[:, :, 1].abs() < cy / fy * cam[:, :, 2] + half * 2)
                            count[cas, indices] += m.sum(0).reshape(-1)
                            head += S
        self.density_grid[count == 0] = -1

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128, fused=None):
        """EMA update of the density grid from the field, repack the bitfield, refresh mean_count (renderer.py:647-773).
        `fused` (default: PVD_FUSED_UPKEEP != 0) takes the three-kernel path of csrc/density_grid.cu without reading anything
        back.  It consumes the same torch random numbers in the same order as the torch flow below (`fused=False`), and
        tests/test_density_grid.py checks that 19 updates leave the same grid / bitfield / mean as that flow.  Against the
        REFERENCE the full sweep is statistically, not bitwise, equivalent: the reference draws its noise per 128^3-cell chunk of
        the meshgrid sweep (renderer.py:690-693), here one triple is drawn per Morton cell."""
        if not self.cuda_ray:
            return
        if fused is None:
            fused = os.environ.get("PVD_FUSED_UPKEEP", "1") != "0"
        if fused and self.density_grid.is_cuda:
            return self._update_extra_state_fused(decay)
        dev = self.density_grid.device
        tmp = -torch.ones_like(self.density_grid)
        H = self.grid_size
        if self.iter_density < 16:  # full sweep
            axis = torch.arange(H, dtype=torch.int32, device=dev).split(S)
            for xs in axis:
                for ys in axis:
                    for zs in axis:
                        xx, yy, zz = _meshgrid_ij(xs, ys, zs)
                        coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
                        indices = raymarching.morton3D(coords).long()
                        xyzs = 2 * coords.float() / (H - 1) - 1
                        for cas in range(self.cascade):
                            bound = min(2 ** cas, self.bound)
                            half = bound / H
                            p = xyzs * (bound - half)
                            p = p + (torch.rand_like(p) * 2 - 1) * half
                            sig = self.density(p)["sigma"].reshape(-1).detach().float() * self.density_scale
                            tmp[cas, indices] = sig
        else:  # random quarter of the cells + as many currently occupied ones
            n = H ** 3 // 4
            for cas in range(self.cascade):
                coords = torch.randint(0, H, (n, 3), device=dev)
                indices = raymarching.morton3D(coords).long()
                occ = torch.nonzero(self.density_grid[cas] > 0).squeeze(-1)
                if occ.numel() > 0:
                    occ = occ[torch.randint(0, occ.shape[0], [n], dtype=torch.long, device=dev)]
                    indices = torch.cat([indices, occ], dim=0)
                    coords = torch.cat([coords, raymarching.morton3D_invert(occ).long()], dim=0)
                xyzs = 2 * coords.float() / (H - 1) - 1
                bound = min(2 ** cas, self.bound)
                half = bound / H
                p = xyzs * (bound - half)
                p = p + (torch.rand_like(p) * 2 - 1) * half
                sig = self.density(p)["sigma"].reshape(-1).detach().float() * self.density_scale
                tmp[cas, indices] = sig
        valid = (self.density_grid >= 0) & (tmp >= 0)
        self.density_grid[valid] = torch.maximum(self.density_grid[valid] * decay, tmp[valid])
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        self.iter_density += 1
        thresh = min(self.mean_density, self.density_thresh)
        self.density_bitfield = raymarching.packbits(self.density_grid, thresh, self.density_bitfield)
        total = min(16, self.local_step)
        if total > 0:
            self.mean_count = int(self.step_counter[:total, 0].sum().item() / total)
        self.local_step = 0
