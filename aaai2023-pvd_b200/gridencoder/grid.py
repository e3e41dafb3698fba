"""Drop-in `gridencoder` (multiresolution hash / tiled grid encoder) backed by libpvd_b200.so.

Mirrors gridencoder/grid.py of the reference: `grid_encode(inputs, embeddings, offsets, per_level_scale,
base_resolution, calc_grad_inputs=False, gridtype=0, align_corners=False)` (:20-139) and the `GridEncoder` module
(:142-232) with the same constructor arguments, parameter/buffer names (`embeddings`, `offsets`) and shapes, so
checkpoints load unchanged.

B200-side differences: the kernels write / read activations sample-major ([B, L*C]) so the reference's two permute
copies (grid.py:84,104) disappear.  Under autocast the fp16 copy of the table is re-cast on every forward like the reference
(grid.py:52); the dense cast is taken out of the hot path by the training engines (explicit stage() / fused optimizer), not by a
cache here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from pvd_b200 import _native as nv

_gridtype_to_id = {"hash": 0, "tiled": 1}
_u32, _f32, _int = C.c_uint32, C.c_float, C.c_int

def _half_table(emb: torch.Tensor) -> torch.Tensor:
    """fp16 copy of the table for an autocast forward -- cast on EVERY call, exactly as the reference does (grid.py:51-52).
    A copy cached per parameter version is not safe here: in-place writes through `.data` (torch_ema's copy_to() / restore(),
    distill_mutual/utils.py:1210-1212, and reset_parameters) leave `_version` untouched, so evaluation after an EMA swap would
    silently gather stale weights.  The training engines do not come through here: they own an explicit `stage()` (or let the fused
    optimizer write the fp16 shadow), pvd_b200/engine.py."""
    return emb.detach().to(torch.half)


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False):
        # inputs [B, D] float in [0,1]; embeddings [sO, C]; offsets [L+1] int32 -> [B, L*C]
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        Cc = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        # half tables under autocast when C is even, float otherwise (grid.py:49-52)
        if torch.is_autocast_enabled("cuda") and Cc % 2 == 0:
            table = _half_table(embeddings)
        else:
            table = embeddings.detach().contiguous()
        if table.dtype not in (torch.float32, torch.float16):
            raise RuntimeError("embeddings must be a float32 or float16 tensor")  # CHECK_IS_FLOATING, gridencoder.cu:433
        if not (inputs.is_cuda and table.is_cuda and offsets.is_cuda):
            raise RuntimeError("inputs, embeddings and offsets must be CUDA tensors")  # CHECK_CUDA, gridencoder.cu:420-424
        if offsets.dtype != torch.int32:
            raise RuntimeError("offsets must be an int tensor")  # CHECK_IS_INT, gridencoder.cu:434
        inputs = inputs.float()
        dt = nv.F16 if table.dtype == torch.float16 else nv.F32
        outputs = torch.empty(B, L * Cc, device=inputs.device, dtype=table.dtype)
        dy_dx = torch.empty(B, L * D * Cc, device=inputs.device, dtype=table.dtype) if calc_grad_inputs else None
        with nv.on_device(inputs):
            nv.check(nv.lib().pvd_grid_encode_forward(nv.ptr(inputs), nv.ptr(table), nv.ptr(offsets), nv.ptr(outputs), _u32(B),
                                                      _u32(D), _u32(Cc), _u32(L), _f32(S), _u32(H), _int(int(calc_grad_inputs)),
                                                      nv.ptr(dy_dx), _u32(gridtype), _int(int(align_corners)), _int(dt),
                                                      _int(1), nv.stream_of(inputs)))
        ctx.save_for_backward(inputs, table, offsets, dy_dx if dy_dx is not None else torch.empty(0, device=inputs.device))
        ctx.dims = [B, D, Cc, L, S, H, gridtype]
        ctx.calc_grad_inputs = calc_grad_inputs
        ctx.align_corners = align_corners
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, table, offsets, dy_dx = ctx.saved_tensors
        B, D, Cc, L, S, H, gridtype = ctx.dims
        cgi = ctx.calc_grad_inputs
        grad = grad.contiguous().to(table.dtype)  # [B, L*C], sample-major: no permute needed
        grad_embeddings = torch.zeros_like(table)
        grad_inputs = torch.zeros(B, D, device=inputs.device, dtype=table.dtype) if cgi else None
        dt = nv.F16 if table.dtype == torch.float16 else nv.F32
        with nv.on_device(inputs):
            nv.check(nv.lib().pvd_grid_encode_backward(nv.ptr(grad), nv.ptr(inputs), nv.ptr(table), nv.ptr(offsets),
                                                       nv.ptr(grad_embeddings), _u32(B), _u32(D), _u32(Cc), _u32(L), _f32(S),
                                                       _u32(H), _int(int(cgi)), nv.ptr(dy_dx if cgi else None),
                                                       nv.ptr(grad_inputs), _u32(gridtype), _int(int(ctx.align_corners)),
                                                       _int(dt), _int(1), nv.stream_of(inputs)))
        if cgi:
            return grad_inputs.to(inputs.dtype), grad_embeddings, None, None, None, None, None, None
        return None, grad_embeddings, None, None, None, None, None, None


grid_encode = _grid_encode.apply


def level_table(offsets: torch.Tensor, per_level_scale, base_resolution):
    """(scales float32[L], resolutions int32[L]) exactly as the kernels compute them on the device."""
    L = offsets.shape[0] - 1
    scales = torch.empty(L, dtype=torch.float32, device=offsets.device)
    res = torch.empty(L, dtype=torch.int32, device=offsets.device)
    with nv.on_device(offsets):
        nv.check(nv.lib().pvd_grid_level_table(nv.ptr(offsets), _u32(L), _f32(float(np.log2(per_level_scale))),
                                               _u32(int(base_resolution)), nv.ptr(scales), nv.ptr(res), nv.stream_of(offsets)))
    return scales, res


def level_offsets(input_dim, num_levels, base_resolution, per_level_scale, log2_hashmap_size, align_corners):
    """Entry offset of every level (grid.py:177-190): dense (res+1)^D grids until they exceed 2^log2_hashmap_size."""
    cap = 2 ** log2_hashmap_size
    offs, total = [], 0
    for lvl in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** lvl))
        n = min(cap, (res if align_corners else res + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offs.append(total)
        total += n
    offs.append(total)
    return np.array(offs, dtype=np.int32)


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False):
        super().__init__()
        if desired_resolution is not None:  # overrides per_level_scale (grid.py:157-161)
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size
        offsets = level_offsets(input_dim, num_levels, base_resolution, per_level_scale, log2_hashmap_size, align_corners)
        self.register_buffer("offsets", torch.from_numpy(offsets))
        self.n_params = int(offsets[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)  # grid.py:200-202

    def __repr__(self):
        top = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {top} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners}")

    def forward(self, inputs, bound=1):
        # inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)
        lead = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        out = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                          inputs.requires_grad, self.gridtype_id, self.align_corners)
        return out.view(lead + [self.output_dim])
