"""Host side of the fused field query (include/pvd_b200_fused.h): autograd op, parameter staging, network module.

`fused_hash_field(xyzs, dirs, embeddings, w_sigma0, w_sigma1, w_color0, w_color1, w_color2, cfg)` is one autograd node
that stands in for the whole of `NeRFNetwork.forward` for model_type "hash" (distill_mutual/network.py:335-437): autograd,
AdamW and GradScaler see the same leaf parameters with the same names and shapes as in the reference.

`HashNeRFField` is an nn.Module with the reference's parameter names (`encoder.embeddings`, `sigma_net.{0,1}.weight`,
`color_net.{0,1,2}.weight`), so reference checkpoints load into it, and with the side-channel attributes the distillation
trainer reads (`feature_sigma_color`, `sigma_l`, `color_l`, distill_mutual/utils.py:1049-1083).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _native as nv

WBLOB_BYTES = 20480
GW_FLOATS = 10240
GW_COPIES = 16                   # PVD_FIELD_GW_COPIES: replicas of the weight-gradient workspace
GW_WS_FLOATS = GW_FLOATS * GW_COPIES
LOSS_SLOTS = 64                  # PVD_LOSS_SLOTS
ENC_STRIDE = 32
# Table-gradient scatter as its own kernel (1) or by scatter warps inside the MLP backward kernel (0, default: one launch, the
# reductions of tile i under the tensor-core chain of tile i+1; measured 54 us against 31 + 31 us at 73 k samples)
SPLIT_SCATTER = os.environ.get("PVD_SPLIT_SCATTER", "0") != "0"


class PvdHashField(C.Structure):
    _fields_ = [("table", C.c_void_p), ("offsets", C.c_void_p), ("wblob", C.c_void_p), ("table_dtype", C.c_int32),
                ("L", C.c_uint32), ("H", C.c_uint32), ("S", C.c_float), ("bound", C.c_float), ("sigma_clip_min", C.c_float),
                ("sigma_clip_max", C.c_float), ("density_scale", C.c_float)]


@dataclass
class HashFieldConfig:
    num_levels: int
    base_resolution: int
    per_level_scale: float
    bound: float = 1.0
    sigma_clip_min: float = -2.0   # main_distill_mutual.py:182
    sigma_clip_max: float = 7.0    # main_distill_mutual.py:183
    density_scale: float = 1.0
    table_fp16: bool = True        # gather from the fp16 shadow (what autocast does in the reference, grid.py:51-52)


def _cstruct(cfg: HashFieldConfig, table, offsets, wblob):
    return PvdHashField(table=table.data_ptr(), offsets=offsets.data_ptr(), wblob=wblob.data_ptr(),
                        table_dtype=nv.F16 if table.dtype == torch.float16 else nv.F32, L=cfg.num_levels,
                        H=cfg.base_resolution, S=float(np.log2(cfg.per_level_scale)), bound=cfg.bound,
                        sigma_clip_min=cfg.sigma_clip_min, sigma_clip_max=cfg.sigma_clip_max,
                        density_scale=cfg.density_scale)


def frozen_key(params):
    """Cache key for staged copies of FROZEN parameters (a distillation teacher: requires_grad False on every tensor,
    main_distill_mutual.py:320-321), None for anything trainable.

    Staged copies of trainable parameters are never cached: a key built from `param._version` is not safe for them -- in-place
    writes through `.data` (torch_ema copy_to() / restore(), distill_mutual/utils.py:1210-1212, reset_parameters) do not bump it,
    so an evaluation after an EMA swap would run on stale weights.  Nobody optimises or EMA-swaps a frozen teacher; loading a
    checkpoint into it goes through `param.copy_`, which does bump the version (`invalidate()` exists for anything more exotic)."""
    params = list(params)
    if any(p.requires_grad for p in params):
        return None
    return tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in params)


class StagedParams:
    """fp16 table shadow + packed tensor-core weight tiles of one model, in persistent buffers.

    Trainable parameters are re-staged on every call (one cast kernel / one pack kernel), see `frozen_key`.  The training engines
    call `stage()` once per optimizer step, or not at all when the fused optimizer (pvd_b200/optim.py) writes the shadow and the
    packed tiles itself."""

    def __init__(self):
        self.table = None
        self.wblob = None
        self._table_key = None
        self._w_key = None

    def invalidate(self):
        self._table_key = self._w_key = None

    def table_for(self, emb: torch.Tensor, fp16: bool):
        if not fp16:
            return emb.detach()
        key = frozen_key([emb])
        if key is not None and key == self._table_key and self.table is not None:
            return self.table
        if self.table is None or self.table.shape != emb.shape or self.table.device != emb.device:
            self.table = torch.empty_like(emb, dtype=torch.float16)
        src = emb.detach()
        if src.dtype == torch.float32 and src.is_contiguous() and src.numel() % 4 == 0:
            with nv.on_device(src):
                nv.check(nv.lib().pvd_cast_f32_to_f16(nv.ptr(src), nv.ptr(self.table), C.c_uint64(src.numel()), nv.stream_of(src)))
        else:
            self.table.copy_(src)
        self._table_key = key
        return self.table

    def wblob_for(self, ws, in_dim: int):
        key = frozen_key(ws)
        if key is not None and key == self._w_key and self.wblob is not None:
            return self.wblob
        if self.wblob is None or self.wblob.device != ws[0].device:
            self.wblob = torch.empty(WBLOB_BYTES, dtype=torch.uint8, device=ws[0].device)
        w32 = [w.detach().float().contiguous() for w in ws]
        with nv.on_device(self.wblob):
            nv.check(nv.lib().pvd_field_pack_weights(nv.ptr(w32[0]), nv.ptr(w32[1]), nv.ptr(w32[2]), nv.ptr(w32[3]),
                                                     nv.ptr(w32[4]), C.c_uint32(in_dim), nv.ptr(self.wblob),
                                                     nv.stream_of(self.wblob)))
        self._w_key = key
        return self.wblob


def hash_field_forward_raw(cfg, table, offsets, wblob, xyzs, dirs, want_enc=True, want_feat=False, status=None):
    M = xyzs.shape[0]
    dev = xyzs.device
    sigmas = torch.empty(M, dtype=torch.float32, device=dev)
    rgbs = torch.empty(M, 3, dtype=torch.float32, device=dev)
    enc = torch.empty(M, ENC_STRIDE, dtype=torch.float16, device=dev) if want_enc else None
    feat = torch.empty(M, 16, dtype=torch.float32, device=dev) if want_feat else None
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    f = _cstruct(cfg, table, offsets, wblob)
    with nv.on_device(xyzs):
        nv.check(nv.lib().pvd_hash_field_forward(C.byref(f), nv.ptr(xyzs), nv.ptr(dirs), C.c_uint32(M), nv.ptr(sigmas),
                                                 nv.ptr(rgbs), nv.ptr(enc), nv.ptr(feat), nv.ptr(status), nv.stream_of(xyzs)))
    return sigmas, rgbs, enc, feat, status


def hash_field_backward_raw(cfg, table, offsets, wblob, xyzs, dirs, enc, grad_sigmas, grad_rgbs, grad_table, gw_ws, n_valid=None,
                            status=None, grad_feat=None, dx_ws=None):
    M = xyzs.shape[0]
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=xyzs.device)
    f = _cstruct(cfg, table, offsets, wblob)
    with nv.on_device(xyzs):
        nv.check(nv.lib().pvd_hash_field_backward(C.byref(f), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(enc), nv.ptr(grad_sigmas),
                                                  nv.ptr(grad_rgbs), nv.ptr(grad_feat), C.c_uint32(M), nv.ptr(n_valid),
                                                  nv.ptr(grad_table),
                                                  nv.ptr(gw_ws), nv.ptr(dx_ws), nv.ptr(status), nv.stream_of(xyzs)))
    return status


def unpack_wgrads(gw_ws, in_dim, like):
    outs = [torch.zeros_like(w, dtype=torch.float32) for w in like]
    with nv.on_device(gw_ws):
        nv.check(nv.lib().pvd_field_unpack_wgrads(nv.ptr(gw_ws), C.c_uint32(in_dim), nv.ptr(outs[0]), nv.ptr(outs[1]),
                                                  nv.ptr(outs[2]), nv.ptr(outs[3]), nv.ptr(outs[4]), nv.stream_of(gw_ws)))
    return outs


class _FusedHashField(Function):
    @staticmethod
    def forward(ctx, xyzs, dirs, embeddings, w0, w1, w2, w3, w4, offsets, cfg, staged, want_feat):
        xyzs = xyzs.detach().float().contiguous()
        dirs = dirs.detach().float().contiguous()
        table = staged.table_for(embeddings, cfg.table_fp16)
        wblob = staged.wblob_for((w0, w1, w2, w3, w4), 2 * cfg.num_levels)
        need_bwd = any(ctx.needs_input_grad)  # a frozen teacher under no_grad does not save the encoding
        sigmas, rgbs, enc, feat, status = hash_field_forward_raw(cfg, table, offsets, wblob, xyzs, dirs, need_bwd, want_feat)
        if enc is None:
            enc = torch.empty(0, dtype=torch.float16, device=xyzs.device)
        # the staged table / weight tiles are shared buffers that the next forward re-writes: not autograd-saved tensors.  The
        # backward reads the tiles only (the encoding was saved), so it re-packs them from the saved parameters.
        ctx.save_for_backward(xyzs, dirs, enc, offsets, embeddings, w0, w1, w2, w3, w4)
        ctx.table = table
        ctx.staged = staged
        ctx.cfg = cfg
        ctx.status = status
        if want_feat:
            return sigmas, rgbs, feat
        return sigmas, rgbs

    @staticmethod
    def backward(ctx, grad_sigmas, grad_rgbs, grad_feat=None):
        xyzs, dirs, enc, offsets, embeddings, w0, w1, w2, w3, w4 = ctx.saved_tensors
        cfg = ctx.cfg
        dev = xyzs.device
        table = ctx.table   # pointer only: the backward kernel does not read table entries
        wblob = ctx.staged.wblob_for((w0, w1, w2, w3, w4), 2 * cfg.num_levels)
        gs = (grad_sigmas if grad_sigmas is not None else torch.zeros(xyzs.shape[0], device=dev)).float().contiguous()
        gc = (grad_rgbs if grad_rgbs is not None else torch.zeros(xyzs.shape[0], 3, device=dev)).float().contiguous()
        grad_table = torch.zeros(embeddings.shape, dtype=torch.float32, device=dev)
        gw_ws = torch.zeros(GW_WS_FLOATS, dtype=torch.float32, device=dev)
        gf = grad_feat.float().contiguous() if grad_feat is not None else None
        dx_ws = torch.empty(xyzs.shape[0], ENC_STRIDE, dtype=torch.float16, device=dev) if SPLIT_SCATTER else None
        hash_field_backward_raw(cfg, table, offsets, wblob, xyzs, dirs, enc, gs, gc, grad_table, gw_ws, None, ctx.status, gf, dx_ws)
        g = unpack_wgrads(gw_ws, 2 * cfg.num_levels, (w0, w1, w2, w3, w4))
        g = [gi.to(w.dtype) for gi, w in zip(g, (w0, w1, w2, w3, w4))]
        return (None, None, grad_table.to(embeddings.dtype), g[0], g[1], g[2], g[3], g[4], None, None, None, None)


class PvdFieldWeightsF32(C.Structure):
    _fields_ = [("sigma0", C.c_void_p), ("sigma1", C.c_void_p), ("color0", C.c_void_p), ("color1", C.c_void_p), ("color2", C.c_void_p)]


class _FusedHashFieldF32(Function):
    """The hash field in fp32 end to end (csrc/field_hash_f32.cu): fp32 table gather, fp32 MLPs on the CUDA cores, fp32 gradients.
    What the reference computes without --fp16; the path north_star's 1e-4 fp32 bound is checked on."""

    @staticmethod
    def _structs(cfg, table, offsets, ws):
        f = _cstruct(cfg, table, offsets, table)          # wblob is not used by the fp32 kernels
        w = PvdFieldWeightsF32(*[x.data_ptr() for x in ws])
        return f, w

    @staticmethod
    def forward(ctx, xyzs, dirs, embeddings, w0, w1, w2, w3, w4, offsets, cfg):
        xyzs = xyzs.detach().float().contiguous().view(-1, 3)
        dirs = dirs.detach().float().contiguous().view(-1, 3)
        table = embeddings.detach().float().contiguous()
        ws = [w.detach().float().contiguous() for w in (w0, w1, w2, w3, w4)]
        M, dev = xyzs.shape[0], xyzs.device
        sigmas = torch.empty(M, dtype=torch.float32, device=dev)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        feat = torch.empty(M, 16, dtype=torch.float32, device=dev)
        f, w = _FusedHashFieldF32._structs(cfg, table, offsets, ws)
        with nv.on_device(xyzs):
            nv.check(nv.lib().pvd_hash_field_forward_f32(C.byref(f), C.byref(w), nv.ptr(xyzs), nv.ptr(dirs), C.c_uint32(M), nv.ptr(sigmas),
                                                         nv.ptr(rgbs), nv.ptr(feat), nv.stream_of(xyzs)))
        ctx.save_for_backward(xyzs, dirs, offsets, embeddings, w0, w1, w2, w3, w4)
        ctx.cfg = cfg
        return sigmas, rgbs, feat

    @staticmethod
    def backward(ctx, grad_sigmas, grad_rgbs, grad_feat):
        xyzs, dirs, offsets, embeddings, w0, w1, w2, w3, w4 = ctx.saved_tensors
        cfg, dev, M = ctx.cfg, xyzs.device, xyzs.shape[0]
        table = embeddings.detach().float().contiguous()
        ws = [w.detach().float().contiguous() for w in (w0, w1, w2, w3, w4)]
        gs = (grad_sigmas if grad_sigmas is not None else torch.zeros(M, device=dev)).float().contiguous()
        gc = (grad_rgbs if grad_rgbs is not None else torch.zeros(M, 3, device=dev)).float().contiguous()
        gf = grad_feat.float().contiguous() if grad_feat is not None else None
        grad_table = torch.zeros(embeddings.shape, dtype=torch.float32, device=dev)
        gw_ws = torch.zeros(GW_WS_FLOATS, dtype=torch.float32, device=dev)
        f, w = _FusedHashFieldF32._structs(cfg, table, offsets, ws)
        with nv.on_device(xyzs):
            nv.check(nv.lib().pvd_hash_field_backward_f32(C.byref(f), C.byref(w), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(gs), nv.ptr(gc), nv.ptr(gf),
                                                          C.c_uint32(M), None, nv.ptr(grad_table), nv.ptr(gw_ws), nv.stream_of(xyzs)))
        g = unpack_wgrads(gw_ws, 2 * cfg.num_levels, (w0, w1, w2, w3, w4))
        g = [gi.to(x.dtype) for gi, x in zip(g, (w0, w1, w2, w3, w4))]
        return (None, None, grad_table.to(embeddings.dtype), g[0], g[1], g[2], g[3], g[4], None, None)


def fused_hash_field(xyzs, dirs, embeddings, w0, w1, w2, w3, w4, offsets, cfg, staged, want_feat=False):
    return _FusedHashField.apply(xyzs, dirs, embeddings, w0, w1, w2, w3, w4, offsets, cfg, staged, want_feat)


from .renderer import NeRFRenderer  # noqa: E402


class _Args:  # the few attributes of the reference's argparse namespace that forward()/run_cuda read (network.py:358,422)
    def __init__(self, sigma_clip_min=-2.0, sigma_clip_max=7.0, global_step=10 ** 9, stage_iters=None, render_stu_first=True):
        self.sigma_clip_min = sigma_clip_min
        self.sigma_clip_max = sigma_clip_max
        self.global_step = global_step
        self.stage_iters = stage_iters or {"stage1": -1, "stage2": -1}
        self.render_stu_first = render_stu_first


class HashNeRFField(NeRFRenderer):
    """`NeRFNetwork(model_type="hash")` of the reference (distill_mutual/network.py:12-182) with a fused forward/backward.

    Same parameter names/shapes: encoder.embeddings [5303704, 2] (L=14), sigma_net.{0,1}.weight, color_net.{0,1,2}.weight;
    same renderer buffers (it IS a NeRFRenderer, like the reference's NeRFNetwork).
    """

    def __init__(self, num_levels=14, desired_resolution=2048, bound=1, hidden_dim=64, geo_feat_dim=15, args=None,
                 density_scale=1.0, table_fp16=True, is_teacher=False, fp32=False, **renderer_kwargs):
        super().__init__(bound=bound, density_scale=density_scale, **renderer_kwargs)
        self.is_teacher = is_teacher
        self.model_type = "hash"
        from gridencoder import GridEncoder
        from shencoder import SHEncoder
        assert hidden_dim == 64 and geo_feat_dim == 15, "the fused kernel is built for PVD's 64-wide / 15-feature heads"
        self.bound = bound
        self.args = args or _Args()
        self.encoder = GridEncoder(num_levels=num_levels, desired_resolution=desired_resolution * bound)
        self.in_dim = self.encoder.output_dim
        self.encoder_dir = SHEncoder(degree=4)
        self.sigma_net = nn.ModuleList([nn.Linear(self.in_dim, 64, bias=False), nn.Linear(64, 16, bias=False)])
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False),
                                        nn.Linear(64, 3, bias=False)])
        self.table_fp16 = table_fp16
        self.fp32 = bool(fp32)   # fp32 end to end (csrc/field_hash_f32.cu): the reference without --fp16; fp16 tensor-core path otherwise
        self._staged = StagedParams()
        self.feature_sigma_color = None
        self.sigma_l = None
        self.color_l = None

    def config(self) -> HashFieldConfig:
        e = self.encoder
        return HashFieldConfig(num_levels=e.num_levels, base_resolution=e.base_resolution, per_level_scale=float(e.per_level_scale),
                               bound=float(self.bound), sigma_clip_min=float(self.args.sigma_clip_min),
                               sigma_clip_max=float(self.args.sigma_clip_max), density_scale=1.0, table_fp16=self.table_fp16)

    def forward(self, x, d):
        # x [M,3] in [-bound, bound], d [M,3] unit -> sigma [M], color [M,3]   (network.py:335-437, hash branch)
        if self.fp32:
            shape = x.shape[:-1]
            sigma, color, feat = _FusedHashFieldF32.apply(x, d, self.encoder.embeddings, self.sigma_net[0].weight, self.sigma_net[1].weight,
                                                          self.color_net[0].weight, self.color_net[1].weight, self.color_net[2].weight,
                                                          self.encoder.offsets, self.config())
            out = (sigma.view(shape), color.view(*shape, 3), feat.view(*shape, 16))
        else:
            out = self._fp16_forward(x, d)
        sigma, color, feat = out
        self.feature_sigma_color = feat
        if self.training and self.args.global_step < self.args.stage_iters["stage1"]:
            return None, None
        self.sigma_l = feat[..., 0]
        self.color_l = color
        return sigma, color

    def _fp16_forward(self, x, d):
        return fused_hash_field(x, d, self.encoder.embeddings, self.sigma_net[0].weight, self.sigma_net[1].weight,
                                self.color_net[0].weight, self.color_net[1].weight, self.color_net[2].weight,
                                self.encoder.offsets, self.config(), self._staged, True)

    @torch.no_grad()
    def render_persistent(self, rays_o, rays_d, nears, fars, dt_gamma=0.0, max_steps=1024):
        """The whole evaluation loop of run_cuda (renderer.py:450-543) in one persistent kernel: raw weights_sum [N], depth [N],
        image [N,3] accumulators for rays_o / rays_d [N,3] (include/pvd_b200_fused.h::pvd_hash_render_persistent)."""
        rays_o = rays_o.detach().float().contiguous()
        rays_d = rays_d.detach().float().contiguous()
        N, dev = rays_o.shape[0], rays_o.device
        cfg = self.config()
        cfg.density_scale = float(self.density_scale)      # renderer.py:517: sigmas = self.density_scale * sigmas
        table = self._staged.table_for(self.encoder.embeddings, cfg.table_fp16)
        wblob = self._staged.wblob_for((self.sigma_net[0].weight, self.sigma_net[1].weight, self.color_net[0].weight,
                                        self.color_net[1].weight, self.color_net[2].weight), 2 * cfg.num_levels)
        f = _cstruct(cfg, table, self.encoder.offsets, wblob)
        ws = torch.zeros(N, dtype=torch.float32, device=dev)
        depth = torch.zeros(N, dtype=torch.float32, device=dev)
        image = torch.zeros(N, 3, dtype=torch.float32, device=dev)
        queue = torch.zeros(1, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with nv.on_device(rays_o):
            nv.check(nv.lib().pvd_hash_render_persistent(C.byref(f), nv.ptr(rays_o), nv.ptr(rays_d), nv.ptr(self.density_bitfield), nv.ptr(nears),
                                                         nv.ptr(fars), C.c_float(float(self.bound)), C.c_float(float(dt_gamma)), C.c_uint32(int(max_steps)),
                                                         C.c_uint32(int(self.cascade)), C.c_uint32(int(self.grid_size)), C.c_uint32(N), nv.ptr(queue),
                                                         nv.ptr(ws), nv.ptr(depth), nv.ptr(image), nv.ptr(status), nv.stream_of(rays_o)))
        self._render_status = status
        return ws, depth, image

    def density(self, x):
        """sigma only, for the density-grid upkeep (network.py:439-494; the colour half of the kernel output is discarded)."""
        x = x.reshape(-1, 3)
        d = torch.zeros_like(x)
        with torch.no_grad():
            sigma, _, _ = fused_hash_field(x, d, self.encoder.embeddings, self.sigma_net[0].weight, self.sigma_net[1].weight,
                                           self.color_net[0].weight, self.color_net[1].weight, self.color_net[2].weight,
                                           self.encoder.offsets, self.config(), self._staged, True)
        return {"sigma": sigma}

    def get_params(self, lr, lr2=1e-3):
        """Optimizer groups of the reference's hash model (network.py:646-653)."""
        return [{"params": self.encoder.parameters(), "lr": lr}, {"params": self.sigma_net.parameters(), "lr": lr},
                {"params": self.encoder_dir.parameters(), "lr": lr}, {"params": self.color_net.parameters(), "lr": lr}]
