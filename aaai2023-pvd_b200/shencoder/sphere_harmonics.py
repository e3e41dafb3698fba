"""Drop-in `shencoder` (real spherical-harmonics direction encoding) backed by libpvd_b200.so.

Mirrors shencoder/sphere_harmonics.py of the reference: `sh_encode(inputs, degree, calc_grad_inputs=False)` (:15-64,
forced to float32) and the `SHEncoder(input_dim=3, degree=4)` module (:67-95).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from pvd_b200 import _native as nv

_u32, _int = C.c_uint32, C.c_int


class _sh_encoder(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        # inputs [B, 3] in [-1,1] -> [B, degree^2]
        inputs = inputs.contiguous()
        if not inputs.is_cuda:
            raise RuntimeError("inputs must be a CUDA tensor")  # CHECK_CUDA, shencoder.cu:403
        B, input_dim = inputs.shape
        out_dim = degree ** 2
        outputs = torch.empty(B, out_dim, dtype=inputs.dtype, device=inputs.device)
        dy_dx = torch.empty(B, input_dim * out_dim, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else None
        with nv.on_device(inputs):
            nv.check(nv.lib().pvd_sh_encode_forward(nv.ptr(inputs), nv.ptr(outputs), _u32(B), _u32(input_dim), _u32(degree),
                                                    _int(int(calc_grad_inputs)), nv.ptr(dy_dx), nv.stream_of(inputs)))
        ctx.save_for_backward(inputs, dy_dx if dy_dx is not None else torch.empty(0, device=inputs.device))
        ctx.dims = [B, input_dim, degree]
        ctx.calc_grad_inputs = calc_grad_inputs
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        if not ctx.calc_grad_inputs:
            return None, None, None
        grad = grad.contiguous()
        inputs, dy_dx = ctx.saved_tensors
        B, input_dim, degree = ctx.dims
        grad_inputs = torch.zeros_like(inputs)
        with nv.on_device(inputs):
            nv.check(nv.lib().pvd_sh_encode_backward(nv.ptr(grad), nv.ptr(inputs), _u32(B), _u32(input_dim), _u32(degree),
                                                     nv.ptr(dy_dx), nv.ptr(grad_inputs), nv.stream_of(inputs)))
        return grad_inputs, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert 0 < self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        # inputs [..., 3] in [-size, size] -> [..., degree^2]
        inputs = inputs / size
        lead = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        out = sh_encode(inputs, self.degree, inputs.requires_grad)
        return out.reshape(lead + [self.output_dim])
