"""Host side of the fused "tensors" (Plenoxels-style dense voxel) field: autograd op + network module.

`TensorsNeRFField` mirrors `NeRFNetwork(model_type="tensors")` of the reference (distill_mutual/network.py:91-96 volume,
:184-191 init_plenoxel_volume, :311-322 compute_plenoxel_fea, :383-409 forward, :461-476 density, :677-681 get_params):
one parameter `tensor_volume.0` of shape [1, 3 * degree^2 + 1, D, H, W] -- kept in torch.channels_last_3d memory, so that the 8
corners of a trilinear tap are 8 contiguous 112-byte reads instead of 8 x 28 strided ones; the shape in the state_dict is the
reference's -- and an SH encoder of `plenoxel_degree`.  No MLP: sigma = trunc_exp(clamp(h0)), rgb = sigmoid(<h_rgb, SH(d)>).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _native as nv
from .fused import _Args
from .renderer import NeRFRenderer


class PvdTensorsField(C.Structure):
    _fields_ = [("volume", C.c_void_p), ("res", C.c_uint32 * 3), ("degree", C.c_uint32), ("aabb", C.c_float * 6),
                ("sigma_clip_min", C.c_float), ("sigma_clip_max", C.c_float), ("density_scale", C.c_float)]


def tensors_struct(volume, degree, aabb, clip_min, clip_max, density_scale=1.0):
    _, Cc, D, H, W = volume.shape
    assert Cc == 3 * degree * degree + 1, "volume channels must be 3 * degree^2 + 1"
    assert volume.is_contiguous(memory_format=torch.channels_last_3d) and volume.dtype == torch.float32, \
        "the plenoxel volume must be fp32 in torch.channels_last_3d memory"
    return PvdTensorsField(volume=volume.data_ptr(), res=(C.c_uint32 * 3)(D, H, W), degree=degree, aabb=(C.c_float * 6)(*aabb),
                           sigma_clip_min=clip_min, sigma_clip_max=clip_max, density_scale=density_scale)


class _FusedTensorsField(Function):
    @staticmethod
    def forward(ctx, xyzs, dirs, volume, meta):
        degree, aabb, clip_min, clip_max = meta
        xyzs = xyzs.detach().float().contiguous()
        dirs = dirs.detach().float().contiguous()
        vol = volume.detach()
        if not vol.is_contiguous(memory_format=torch.channels_last_3d):
            vol = vol.contiguous(memory_format=torch.channels_last_3d)
        M, dev = xyzs.shape[0], xyzs.device
        sigmas = torch.empty(M, dtype=torch.float32, device=dev)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        f = tensors_struct(vol, degree, aabb, clip_min, clip_max)
        with nv.on_device(xyzs):
            nv.check(nv.lib().pvd_tensors_field_forward(C.byref(f), nv.ptr(xyzs), nv.ptr(dirs), C.c_uint32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                        nv.stream_of(xyzs)))
        ctx.save_for_backward(xyzs, dirs, volume)
        ctx.meta = meta
        return sigmas, rgbs

    @staticmethod
    def backward(ctx, grad_sigmas, grad_rgbs):
        xyzs, dirs, volume = ctx.saved_tensors
        degree, aabb, clip_min, clip_max = ctx.meta
        M, dev = xyzs.shape[0], xyzs.device
        gs = (grad_sigmas if grad_sigmas is not None else torch.zeros(M, device=dev)).float().contiguous()
        gc = (grad_rgbs if grad_rgbs is not None else torch.zeros(M, 3, device=dev)).float().contiguous()
        vol = volume.detach()
        if not vol.is_contiguous(memory_format=torch.channels_last_3d):
            vol = vol.contiguous(memory_format=torch.channels_last_3d)
        grad = torch.zeros_like(vol, memory_format=torch.preserve_format)
        f = tensors_struct(vol, degree, aabb, clip_min, clip_max)
        with nv.on_device(xyzs):
            nv.check(nv.lib().pvd_tensors_field_backward(C.byref(f), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(gs), nv.ptr(gc), C.c_uint32(M), None,
                                                         nv.ptr(grad), nv.stream_of(xyzs)))
        return None, None, grad, None


class TensorsNeRFField(NeRFRenderer):
    def __init__(self, plenoxel_degree=3, plenoxel_res=(128, 128, 128), bound=1, args=None, density_scale=1.0, is_teacher=False, scale=0.02,
                 **renderer_kwargs):
        super().__init__(bound=bound, density_scale=density_scale, **renderer_kwargs)
        from shencoder import SHEncoder
        self.is_teacher = is_teacher
        self.model_type = "tensors"
        self.args = args or _Args()
        self.plenoxel_degree = int(plenoxel_degree)
        self.plenoxel_res = list(plenoxel_res)
        assert len(self.plenoxel_res) == 3
        fea_dim = self.plenoxel_degree ** 2 * 3 + 1
        vol = scale * torch.randn((1, fea_dim, *self.plenoxel_res))                           # network.py:184-191 (s = 0.02, :93)
        self.tensor_volume = nn.ParameterList([nn.Parameter(vol.contiguous(memory_format=torch.channels_last_3d))])
        self.encoder_dir = SHEncoder(degree=self.plenoxel_degree)
        self.feature_sigma_color = None
        self.sigma_l = None
        self.color_l = None

    def _meta(self, clip=True):
        aabb = [float(v) for v in self.aabb_train.tolist()]
        lo, hi = (float(self.args.sigma_clip_min), float(self.args.sigma_clip_max)) if clip else (-3.0e38, 3.0e38)
        return (self.plenoxel_degree, aabb, lo, hi)

    def forward(self, x, d):
        sigma, color = _FusedTensorsField.apply(x, d, self.tensor_volume[0], self._meta())
        self.sigma = sigma
        self.feature_sigma_color = None                                                      # network.py:407
        self.color_l = color
        return sigma, color

    def density(self, x):
        """network.py:461-476: the reference ends up with the UNCLAMPED trunc_exp(h[..., 0]) here (the clamped value is overwritten)."""
        x = x.reshape(-1, 3)
        with torch.no_grad():
            sigma, _ = _FusedTensorsField.apply(x, torch.zeros_like(x), self.tensor_volume[0], self._meta(clip=False))
        return {"sigma": sigma}

    def get_params(self, lr, lr2=1e-3):
        """network.py:677-681."""
        return [{"params": self.tensor_volume.parameters(), "lr": lr}, {"params": self.encoder_dir.parameters(), "lr": lr}]
