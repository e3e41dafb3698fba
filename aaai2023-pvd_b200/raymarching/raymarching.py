"""Drop-in `raymarching` operators backed by libpvd_b200.so (sm_100a).

Same callables, positional signatures, return values and dtypes as the reference module
(`raymarching/raymarching.py`: near_far_from_aabb :53, polar_from_ray :87, morton3D :113, morton3D_invert :138,
packbits :169, march_rays_train :289, composite_rays_train :360, march_rays :454, composite_rays :502,
compact_rays :527), so `renderer.py` / `network.py` of the reference import and call them unchanged.

Differences that a caller can observe:
  * `march_rays_train` hands out sample offsets deterministically (ray-id order) instead of in atomicAdd
    arrival order, and in the warm-up path (mean_count <= 0 or force_all_rays) allocates exactly the rows it
    needs after the counting phase instead of zero-filling N*max_steps rows;
  * `compact_rays` keeps survivors in their previous relative order;
  * launches go to the current stream of the tensors' device; native errors raise RuntimeError.
There is no CPU path: CPU tensors are moved to the GPU exactly where the reference does (`.cuda()`).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from pvd_b200 import _native as nv

__all__ = [
    "near_far_from_aabb", "polar_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
    "composite_rays_train", "march_rays", "composite_rays", "compact_rays",
]

_u32, _f32 = C.c_uint32, C.c_float


def _cuda(t):
    return t if t.is_cuda else t.cuda()


def _rays(t):
    return _cuda(t).contiguous().view(-1, 3)


# ------------------------------------------------------------------------------------------ utils
class _near_far_from_aabb(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """rays_o/rays_d [N,3], aabb [6] -> nears, fars [N] (reference: raymarching.py:20-53)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        aabb = aabb.to(rays_o.device, torch.float32).contiguous()
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        with nv.on_device(rays_o):
            nv.check(nv.lib().pvd_near_far_from_aabb(nv.ptr(rays_o), nv.ptr(rays_d), nv.ptr(aabb), _u32(N),
                                                     _f32(min_near), nv.ptr(nears), nv.ptr(fars), nv.stream_of(rays_o)))
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _polar_from_ray(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, radius):
        """Polar coordinates on the background sphere, [N,2] in [-1,1] (reference: raymarching.py:56-87)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        with nv.on_device(rays_o):
            nv.check(nv.lib().pvd_polar_from_ray(nv.ptr(rays_o), nv.ptr(rays_d), _f32(radius), _u32(N), nv.ptr(coords),
                                                 nv.stream_of(rays_o)))
        return coords


polar_from_ray = _polar_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """coords [N,3] int -> Morton indices [N] int32 (reference: raymarching.py:90-113)."""
        coords = _cuda(coords).int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        with nv.on_device(coords):
            nv.check(nv.lib().pvd_morton3D(nv.ptr(coords), _u32(N), nv.ptr(indices), nv.stream_of(coords)))
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """indices [N] int -> coords [N,3] int32 (reference: raymarching.py:116-138)."""
        indices = _cuda(indices).int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        with nv.on_device(indices):
            nv.check(nv.lib().pvd_morton3D_invert(nv.ptr(indices), _u32(N), nv.ptr(coords), nv.stream_of(indices)))
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, grid, thresh, bitfield=None):
        """grid [C, H^3] float -> bitfield [C*H^3/8] uint8 (reference: raymarching.py:141-169)."""
        grid = _cuda(grid).contiguous()
        N = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        with nv.on_device(grid):
            nv.check(nv.lib().pvd_packbits(nv.ptr(grid), _u32(N), _f32(thresh), nv.ptr(bitfield), nv.stream_of(grid)))
        return bitfield


packbits = _packbits.apply


# ------------------------------------------------------------------------------------------ train
def _align_up_strict(m: int, align: int) -> int:
    # the reference always adds: m += align - m % align (raymarching.py:236-237,278-279)
    return m + (align - m % align) if align > 0 else m


class _march_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C_, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        """Sample points along rays through the occupancy grid (reference: raymarching.py:176-289).

        Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3] = (ray id, point offset, point count).
        """
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        density_bitfield = _cuda(density_bitfield).contiguous()
        dev = rays_o.device
        N = rays_o.shape[0]
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        l = nv.lib()
        ws = torch.empty(int(l.pvd_march_rays_train_workspace_words(N, max_steps)), dtype=torch.int32, device=dev)
        st = nv.stream_of(rays_o)
        with nv.on_device(rays_o):
            nv.check(l.pvd_march_rays_train_count(nv.ptr(rays_o), nv.ptr(rays_d), nv.ptr(density_bitfield), _f32(bound),
                                                  _f32(dt_gamma), _u32(max_steps), _u32(N), _u32(C_), _u32(H), nv.ptr(nears),
                                                  nv.ptr(fars), nv.ptr(rays), nv.ptr(step_counter), _u32(int(bool(perturb))),
                                                  nv.ptr(ws), st))
            if not force_all_rays and mean_count > 0:
                M = _align_up_strict(int(mean_count), align)
                M_drop = M
            else:
                # warm-up: one D2H read, as raymarching.py:277, but before allocating instead of after
                M = _align_up_strict(int(step_counter[0].item()), align)
                # the reference applies its drop rule (offset + n >= M, raymarching.cu:419) against N*max_steps here
                M_drop = N * max_steps
            xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
            dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
            deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
            nv.check(l.pvd_march_rays_train_write(nv.ptr(rays_o), nv.ptr(rays_d), _f32(bound), _u32(max_steps), _u32(N),
                                                  _u32(M_drop), nv.ptr(rays), nv.ptr(ws), nv.ptr(xyzs), nv.ptr(dirs),
                                                  nv.ptr(deltas), st))
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays):
        """Alpha-composite per-sample (sigma, rgb) into weights_sum [N], depth [N], image [N,3]
        (reference: raymarching.py:292-325)."""
        sigmas, rgbs, deltas = sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=dev)
        depth = torch.empty(N, dtype=sigmas.dtype, device=dev)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=dev)
        with nv.on_device(sigmas):
            nv.check(nv.lib().pvd_composite_rays_train_forward(nv.ptr(sigmas), nv.ptr(rgbs), nv.ptr(deltas), nv.ptr(rays),
                                                               _u32(M), _u32(N), nv.ptr(weights_sum), nv.ptr(depth),
                                                               nv.ptr(image), nv.stream_of(sigmas)))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N]
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is ignored, as in the reference (raymarching.py:331)
        grad_weights_sum, grad_image = grad_weights_sum.contiguous(), grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        with nv.on_device(sigmas):
            nv.check(nv.lib().pvd_composite_rays_train_backward(
                nv.ptr(grad_weights_sum), nv.ptr(grad_image), nv.ptr(sigmas), nv.ptr(rgbs), nv.ptr(deltas), nv.ptr(rays),
                nv.ptr(weights_sum), nv.ptr(image), _u32(M), _u32(N), nv.ptr(grad_sigmas), nv.ptr(grad_rgbs),
                nv.stream_of(sigmas)))
        return grad_sigmas, grad_rgbs, None, None


composite_rays_train = _composite_rays_train.apply


# ------------------------------------------------------------------------------------------ infer
class _march_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C_, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        """March every alive ray up to n_step samples (reference: raymarching.py:367-454)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        dev = rays_o.device
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
        with nv.on_device(rays_o):
            nv.check(nv.lib().pvd_march_rays(_u32(n_alive), _u32(n_step), nv.ptr(rays_alive), nv.ptr(rays_t), nv.ptr(rays_o),
                                             nv.ptr(rays_d), _f32(bound), _f32(dt_gamma), _u32(max_steps), _u32(C_), _u32(H),
                                             nv.ptr(density_bitfield), nv.ptr(near), nv.ptr(far), nv.ptr(xyzs), nv.ptr(dirs),
                                             nv.ptr(deltas), _u32(int(perturb)), nv.stream_of(rays_o)))
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
        """In-place compositing for inference (reference: raymarching.py:457-502)."""
        sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
        with nv.on_device(sigmas):
            nv.check(nv.lib().pvd_composite_rays(_u32(n_alive), _u32(n_step), nv.ptr(rays_alive), nv.ptr(rays_t),
                                                 nv.ptr(sigmas), nv.ptr(rgbs), nv.ptr(deltas), nv.ptr(weights_sum),
                                                 nv.ptr(depth), nv.ptr(image), nv.stream_of(sigmas)))
        return tuple()


composite_rays = _composite_rays.apply


class _compact_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter):
        """Drop rays whose rays_t_old < 0 (reference: raymarching.py:505-527)."""
        with nv.on_device(rays_t):
            nv.check(nv.lib().pvd_compact_rays(_u32(n_alive), nv.ptr(rays_alive), nv.ptr(rays_alive_old), nv.ptr(rays_t),
                                               nv.ptr(rays_t_old), nv.ptr(alive_counter), nv.stream_of(rays_t)))
        return tuple()


compact_rays = _compact_rays.apply
