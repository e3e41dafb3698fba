"""Host side of the fused "vm" (TensoRF vector-matrix) field: autograd op + network module.

`VMNeRFField` mirrors `NeRFNetwork(model_type="vm")` of the reference (distill_mutual/network.py:72-90,193-214,344-382):
parameter names and shapes are the reference's (`sigma_mat.{0,1,2}` [1,16,R,R], `sigma_vec.{0,1,2}` [1,16,R,1],
`color_mat.{0,1,2}` [1,48,R,R], `color_vec.{0,1,2}` [1,48,R,1], `basis_mat.weight` [15,144], `color_net.{0,1,2}.weight`), so
checkpoints load; the tensors are merely kept in torch.channels_last memory format, which is what lets one bilinear tap be a
contiguous 64/192-byte read in the kernel (include/pvd_b200_fused.h).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _native as nv
from .fused import GW_WS_FLOATS, _Args, frozen_key
from .renderer import NeRFRenderer

VM_WBLOB_BYTES = 18944


class PvdVmField(C.Structure):
    _fields_ = [("sigma_mat", C.c_void_p * 3), ("sigma_vec", C.c_void_p * 3), ("color_mat", C.c_void_p * 3),
                ("color_vec", C.c_void_p * 3), ("wblob", C.c_void_p), ("res", C.c_uint32 * 3), ("aabb", C.c_float * 6),
                ("sigma_clip_min", C.c_float), ("sigma_clip_max", C.c_float), ("density_scale", C.c_float), ("plane_dtype", C.c_int32)]


class PvdVmGrads(C.Structure):
    _fields_ = [("sigma_mat", C.c_void_p * 3), ("sigma_vec", C.c_void_p * 3), ("color_mat", C.c_void_p * 3),
                ("color_vec", C.c_void_p * 3)]


def _cl(t):
    """channels-last view/copy of a [1,R,H,W] tensor (no copy when it already is)."""
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


def _ptrs3(ts):
    return (C.c_void_p * 3)(*[t.data_ptr() for t in ts])


def _vm_struct(planes, wblob, res, aabb, clip_min, clip_max, density_scale):
    smat, svec, cmat, cvec = planes
    dts = {t.dtype for grp in planes for t in grp}
    assert len(dts) == 1 and dts <= {torch.float32, torch.float16}, "planes / lines must all be fp32 or all be fp16 shadows"
    return PvdVmField(plane_dtype=nv.F16 if torch.float16 in dts else nv.F32, sigma_mat=_ptrs3(smat), sigma_vec=_ptrs3(svec), color_mat=_ptrs3(cmat), color_vec=_ptrs3(cvec),
                      wblob=wblob.data_ptr(), res=(C.c_uint32 * 3)(*res), aabb=(C.c_float * 6)(*aabb),
                      sigma_clip_min=clip_min, sigma_clip_max=clip_max, density_scale=density_scale)


class StagedVmWeights:
    def __init__(self):
        self._key = None
        self.wblob = None

    def invalidate(self):
        self._key = None

    def get(self, ws):
        key = frozen_key(ws)   # None for trainable weights: re-packed on every call (fused.frozen_key)
        if key is None or key != self._key or self.wblob is None:
            if self.wblob is None:
                self.wblob = torch.empty(VM_WBLOB_BYTES, dtype=torch.uint8, device=ws[0].device)
            w32 = [w.detach().float().contiguous() for w in ws]
            with nv.on_device(self.wblob):
                nv.check(nv.lib().pvd_vm_pack_weights(nv.ptr(w32[0]), nv.ptr(w32[1]), nv.ptr(w32[2]), nv.ptr(w32[3]),
                                                      nv.ptr(self.wblob), nv.stream_of(self.wblob)))
            self._key = key
        return self.wblob


class _FusedVmField(Function):
    @staticmethod
    def forward(ctx, xyzs, dirs, meta, *params):
        # params: 3 sigma_mat, 3 sigma_vec, 3 color_mat, 3 color_vec, basis_w, wc0, wc1, wc2
        res, aabb, clip_min, clip_max, staged = meta
        xyzs = xyzs.detach().float().contiguous()
        dirs = dirs.detach().float().contiguous()
        planes = [[_cl(p.detach()) for p in params[3 * k:3 * k + 3]] for k in range(4)]
        wblob = staged.get(params[12:16])
        M, dev = xyzs.shape[0], xyzs.device
        sigmas = torch.empty(M, dtype=torch.float32, device=dev)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        feat = torch.empty(M, 16, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        f = _vm_struct(planes, wblob, res, aabb, clip_min, clip_max, 1.0)
        with nv.on_device(xyzs):
            nv.check(nv.lib().pvd_vm_field_forward(C.byref(f), nv.ptr(xyzs), nv.ptr(dirs), C.c_uint32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                   nv.ptr(feat), nv.ptr(status), nv.stream_of(xyzs)))
        ctx.save_for_backward(xyzs, dirs, wblob, *params)
        ctx.meta = meta
        ctx.status = status
        return sigmas, rgbs, feat

    @staticmethod
    def backward(ctx, grad_sigmas, grad_rgbs, grad_feat):
        xyzs, dirs, wblob, *params = ctx.saved_tensors
        res, aabb, clip_min, clip_max, staged = ctx.meta
        M, dev = xyzs.shape[0], xyzs.device
        gs = (grad_sigmas if grad_sigmas is not None else torch.zeros(M, device=dev)).float().contiguous()
        gc = (grad_rgbs if grad_rgbs is not None else torch.zeros(M, 3, device=dev)).float().contiguous()
        gf = grad_feat.float().contiguous() if grad_feat is not None else None
        planes = [[_cl(p.detach()) for p in params[3 * k:3 * k + 3]] for k in range(4)]
        grads = [[torch.zeros_like(p, memory_format=torch.preserve_format) for p in grp] for grp in planes]
        gw_ws = torch.zeros(GW_WS_FLOATS, dtype=torch.float32, device=dev)
        f = _vm_struct(planes, wblob, res, aabb, clip_min, clip_max, 1.0)
        g = PvdVmGrads(sigma_mat=_ptrs3(grads[0]), sigma_vec=_ptrs3(grads[1]), color_mat=_ptrs3(grads[2]), color_vec=_ptrs3(grads[3]))
        with nv.on_device(xyzs):
            nv.check(nv.lib().pvd_vm_field_backward(C.byref(f), C.byref(g), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(gs), nv.ptr(gc),
                                                    nv.ptr(gf), C.c_uint32(M), None, nv.ptr(gw_ws), nv.ptr(ctx.status),
                                                    nv.stream_of(xyzs)))
            wg = [torch.zeros_like(w, dtype=torch.float32) for w in params[12:16]]
            nv.check(nv.lib().pvd_vm_unpack_wgrads(nv.ptr(gw_ws), nv.ptr(wg[0]), nv.ptr(wg[1]), nv.ptr(wg[2]), nv.ptr(wg[3]),
                                                   nv.stream_of(xyzs)))
        flat = [t for grp in grads for t in grp] + [w.to(p.dtype) for w, p in zip(wg, params[12:16])]
        return (None, None, None, *flat)


class VMNeRFField(NeRFRenderer):
    sigma_rank = 16   # network.py:73
    color_rank = 48   # network.py:74

    def __init__(self, resolution0=300, bound=1, args=None, density_scale=1.0, is_teacher=False, scale=0.1, **renderer_kwargs):
        super().__init__(bound=bound, density_scale=density_scale, **renderer_kwargs)
        from shencoder import SHEncoder
        self.is_teacher = is_teacher
        self.model_type = "vm"
        self.args = args or _Args()
        self.resolution = [resolution0] * 3 if isinstance(resolution0, int) else list(resolution0)
        self.mat_ids = [[0, 1], [0, 2], [1, 2]]
        self.vec_ids = [2, 1, 0]
        self.sigma_mat, self.sigma_vec = self._init_one(self.sigma_rank, scale)
        self.color_mat, self.color_vec = self._init_one(self.color_rank, scale)
        self.basis_mat = nn.Linear(3 * self.color_rank, 15, bias=False)
        self.encoder_dir = SHEncoder(degree=4)
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 3, bias=False)])
        self._staged = StagedVmWeights()
        self.feature_sigma_color = None
        self.sigma_l = None
        self.color_l = None

    def _init_one(self, rank, scale):
        mats, vecs = [], []
        for i in range(3):
            m0, m1 = self.mat_ids[i]
            mat = scale * torch.randn(1, rank, self.resolution[m1], self.resolution[m0])       # network.py:200-207
            vec = scale * torch.randn(1, rank, self.resolution[self.vec_ids[i]], 1)           # network.py:208-212
            mats.append(nn.Parameter(mat.contiguous(memory_format=torch.channels_last)))
            vecs.append(nn.Parameter(vec.contiguous(memory_format=torch.channels_last)))
        return nn.ParameterList(mats), nn.ParameterList(vecs)

    def _params(self):
        return (*self.sigma_mat, *self.sigma_vec, *self.color_mat, *self.color_vec, self.basis_mat.weight,
                self.color_net[0].weight, self.color_net[1].weight, self.color_net[2].weight)

    def forward(self, x, d):
        aabb = [float(v) for v in self.aabb_train.tolist()]
        meta = (self.resolution, aabb, float(self.args.sigma_clip_min), float(self.args.sigma_clip_max), self._staged)
        sigma, color, feat = _FusedVmField.apply(x, d, meta, *self._params())
        self.feature_sigma_color = feat
        if self.training and self.args.global_step < self.args.stage_iters["stage1"]:
            return None, None
        self.sigma_l = feat[..., 0]
        self.color_l = color
        return sigma, color

    def density(self, x):
        x = x.reshape(-1, 3)
        with torch.no_grad():
            aabb = [float(v) for v in self.aabb_train.tolist()]
            meta = (self.resolution, aabb, float(self.args.sigma_clip_min), float(self.args.sigma_clip_max), self._staged)
            sigma, _, _ = _FusedVmField.apply(x, torch.zeros_like(x), meta, *self._params())
        return {"sigma": sigma}

    def density_loss(self):
        """L1 penalty on the sigma planes and lines (network.py:549-557)."""
        loss = 0
        for i in range(3):
            loss = loss + torch.mean(torch.abs(self.sigma_mat[i])) + torch.mean(torch.abs(self.sigma_vec[i]))
        return loss

    def get_params(self, lr, lr2=1e-3):
        """Optimizer groups of the reference's vm model (network.py:661-669)."""
        return [{"params": self.color_net.parameters(), "lr": lr2}, {"params": self.sigma_mat, "lr": lr},
                {"params": self.sigma_vec, "lr": lr}, {"params": self.color_mat, "lr": lr}, {"params": self.color_vec, "lr": lr},
                {"params": self.basis_mat.parameters(), "lr": lr2}]
