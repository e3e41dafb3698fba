"""Host side of the fused "mlp" (NeRF) field.

`MLPNeRFField` mirrors `NeRFNetwork(model_type="mlp")` of the reference (distill_mutual/network.py:56-70,103-152,324-333,
413-437): `nerf_mlp.{0..7}` (Linear with bias; 63->256, 256->256 x3, 319->256, 256->256 x2, 256->28), `sigma_net.{0,1}`,
`color_net.{0,1,2}`.  In PVD this model type is the frozen TEACHER of mlp->hash distillation and is evaluated under
torch.no_grad (distill_mutual/utils.py:1008-1018): that forward is one fused tcgen05 kernel (csrc/field_mlp.cu).
When gradients are required (training an mlp model from images, main_just_train_tea.py --model_type mlp) the forward is the same
kernel saving its operand tiles, and the backward is three tcgen05 kernels (csrc/field_mlp_bwd.cu; `_FusedMlpField` below):
tail backward -> trunk data gradients -> weight gradients.  `PVD_MLP_TORCH_TRAIN=1` selects the plain torch composition of the
same layers instead (autograd through cuBLAS, what the reference does) -- kept as the in-process comparison for the tests.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _native as nv
from . import fused
from .fused import StagedParams, _Args, frozen_key
from .renderer import NeRFRenderer

MLP_WBLOB_BYTES = 876544
MLP_WBLOB_T_BYTES = 49 * 16384
MLP_SAVE_TILE_BYTES = 475136
MLP_GRAD_TILE_BYTES = 466944
MLP_GW_FLOATS = 7 * 256 * 320 + 256 * 32 + 8 * 256
TORCH_TRAIN = os.environ.get("PVD_MLP_TORCH_TRAIN", "0") == "1"
# copies of the packed weight streams in global memory; CTA b reads copy b mod R (include/pvd_b200_fused.h::PvdMlpField.replicas)
REPLICAS = max(1, int(os.environ.get("PVD_MLP_REPLICAS", "1")))


def replicate(blob, bytes_each):
    """Fill replicas 1.. of a packed stream from replica 0 (stream-ordered device copies)."""
    if REPLICAS > 1:
        with nv.on_device(blob):
            nv.check(nv.lib().pvd_mlp_replicate_weights(nv.ptr(blob), C.c_uint64(bytes_each), C.c_uint32(REPLICAS), nv.stream_of(blob)))


class PvdMlpField(C.Structure):
    _fields_ = [("wblob", C.c_void_p), ("tail_wblob", C.c_void_p), ("sigma_clip_min", C.c_float), ("sigma_clip_max", C.c_float),
                ("density_scale", C.c_float), ("replicas", C.c_uint32)]


def _ptr_array(tensors, dev):
    """Device array of device pointers (the C ABI's `const float* const*`)."""
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=dev)


def mlp_tail_backward(field_args, tail_blob, xyzs, dirs, enc, gs, gc, gf, status):
    """d(loss)/d(x28) [M,32] fp16 and the tail's weight-gradient workspace, through the hash model's tcgen05 tail backward in its
    d(encoding)-export mode (include/pvd_b200_fused.h::pvd_hash_field_backward_rows, PVD_BWD_MLP): the 28-wide trunk output plays
    the role of the hash encoding; no table is involved."""
    M, dev = xyzs.shape[0], xyzs.device
    offsets = torch.arange(17, dtype=torch.int32, device=dev) * 8          # a syntactically valid level table; never used for addressing
    cfg = fused.HashFieldConfig(num_levels=14, base_resolution=16, per_level_scale=2.0, bound=1.0, sigma_clip_min=float(field_args.sigma_clip_min),
                                sigma_clip_max=float(field_args.sigma_clip_max), density_scale=1.0)
    gw_tail = torch.zeros(fused.GW_WS_FLOATS, dtype=torch.float32, device=dev)
    d_x28 = torch.empty(M, fused.ENC_STRIDE, dtype=torch.float16, device=dev)
    f = fused._cstruct(cfg, enc, offsets, tail_blob)
    with nv.on_device(xyzs):
        nv.check(nv.lib().pvd_hash_field_backward_rows(C.byref(f), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(enc), nv.ptr(gs), nv.ptr(gc), nv.ptr(gf),
                                                       C.c_uint32(0), C.c_uint32(M), None, nv.ptr(gw_tail), nv.ptr(gw_tail), nv.ptr(d_x28),
                                                       nv.ptr(status), C.c_uint32(1), nv.stream_of(xyzs)))
    return d_x28, gw_tail


class _FusedMlpField(Function):
    """NeRFNetwork.forward for model_type "mlp" with gradients (network.py:324-333,413-437): fused forward that saves its operand
    tiles, backward = tail kernel -> k_mlp_trunk_bwd -> k_mlp_wgrad -> unpack.  params = nerf_mlp.{0..7}.{weight,bias} (16), then
    sigma_net.{0,1}.weight, color_net.{0,1,2}.weight."""

    @staticmethod
    def forward(ctx, x, d, field, *params):
        x = x.detach().float().contiguous().view(-1, 3)
        d = d.detach().float().contiguous().view(-1, 3)
        M, dev = x.shape[0], x.device
        tiles = (M + 127) // 128
        save_ws = torch.empty(max(tiles, 1) * MLP_SAVE_TILE_BYTES, dtype=torch.uint8, device=dev)
        enc = torch.empty(M, fused.ENC_STRIDE, dtype=torch.float16, device=dev)
        sigmas, rgbs, feat = field._fused_forward(x, d, save_ws=save_ws, enc=enc)
        ctx.save_for_backward(x, d, enc, *params)
        ctx.save_ws, ctx.field = save_ws, field
        return sigmas, rgbs, feat

    @staticmethod
    def backward(ctx, g_sigma, g_rgb, g_feat):
        x, d, enc, *params = ctx.saved_tensors
        field, save_ws = ctx.field, ctx.save_ws
        M, dev = x.shape[0], x.device
        ws, bs, tail = list(params[0:16:2]), list(params[1:16:2]), params[16:]
        gs = (g_sigma if g_sigma is not None else torch.zeros(M, device=dev)).float().contiguous()
        gc = (g_rgb if g_rgb is not None else torch.zeros(M, 3, device=dev)).float().contiguous()
        gf = g_feat.float().contiguous() if g_feat is not None else None
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        tail_blob = field._staged.wblob_for(tail, field.in_dim)
        d_x28, gw_tail = mlp_tail_backward(field.args, tail_blob, x, d, enc, gs, gc, gf, status)
        w32 = [w.detach().float().contiguous() for w in ws]
        wblob_t = torch.empty(MLP_WBLOB_T_BYTES * REPLICAS, dtype=torch.uint8, device=dev)
        grad_ws = torch.empty(max((M + 127) // 128, 1) * MLP_GRAD_TILE_BYTES, dtype=torch.uint8, device=dev)
        gw = torch.zeros(MLP_GW_FLOATS, dtype=torch.float32, device=dev)
        gws = [torch.zeros(w.shape, dtype=torch.float32, device=dev) for w in ws]
        gbs = [torch.zeros(b.shape, dtype=torch.float32, device=dev) for b in bs]
        wp, gwp, gbp = _ptr_array(w32, dev), _ptr_array(gws, dev), _ptr_array(gbs, dev)
        with nv.on_device(x):
            st = nv.stream_of(x)
            nv.check(nv.lib().pvd_mlp_pack_weights_t(nv.ptr(wp), nv.ptr(wblob_t), st))
            replicate(wblob_t, MLP_WBLOB_T_BYTES)
            nv.check(nv.lib().pvd_mlp_trunk_backward(nv.ptr(wblob_t), C.c_uint32(REPLICAS), nv.ptr(save_ws), nv.ptr(d_x28), C.c_uint32(M), None, nv.ptr(grad_ws), nv.ptr(status), st))
            nv.check(nv.lib().pvd_mlp_weight_grads(nv.ptr(save_ws), nv.ptr(grad_ws), C.c_uint32(M), nv.ptr(gw), nv.ptr(status), st))
            nv.check(nv.lib().pvd_mlp_unpack_wgrads(nv.ptr(gw), nv.ptr(gwp), nv.ptr(gbp), st))
        gt = fused.unpack_wgrads(gw_tail, field.in_dim, tail)
        field._bwd_status = status
        ctx.save_ws = None
        grads = []
        for gw_l, gb_l, w, b in zip(gws, gbs, ws, bs):
            grads += [gw_l.to(w.dtype), gb_l.to(b.dtype)]
        grads += [g.to(w.dtype) for g, w in zip(gt, tail)]
        return (None, None, None, *grads)


class MLPNeRFField(NeRFRenderer):
    def __init__(self, bound=1, args=None, density_scale=1.0, is_teacher=True, PE=10, width=256, **renderer_kwargs):
        super().__init__(bound=bound, density_scale=density_scale, **renderer_kwargs)
        from shencoder import SHEncoder
        from tools.encoding import get_encoder
        assert PE == 10 and width == 256, "the fused kernel is built for PVD's defaults (PE=10, 8 x 256, skip at 3)"
        self.is_teacher = is_teacher
        self.model_type = "mlp"
        self.args = args or _Args()
        self.encoder_nerf_pe, self.in_dim_nerf = get_encoder("frequency", multires=PE)
        self.skips = 3
        self.in_dim = 28
        layers = [nn.Linear(self.in_dim_nerf, width)]
        for i in range(6):
            layers.append(nn.Linear(width + self.in_dim_nerf, width) if i == self.skips else nn.Linear(width, width))
        layers.append(nn.Linear(width, self.in_dim))
        self.nerf_mlp = nn.ModuleList(layers)
        self.encoder_dir = SHEncoder(degree=4)
        self.sigma_net = nn.ModuleList([nn.Linear(self.in_dim, 64, bias=False), nn.Linear(64, 16, bias=False)])
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 3, bias=False)])
        self._staged = StagedParams()
        self._mlp_key = None
        self._mlp_blob = None
        self._ptr_key = None
        self.feature_sigma_color = None
        self.sigma_l = None
        self.color_l = None

    # ---------------------------------------------------------------- fused (no-grad) forward
    def _blob(self):
        """The packed forward weight stream.  Frozen parameters: packed once (fused.frozen_key).  Trainable: re-packed on every call by
        ONE kernel reading the fp32 parameters in place through a cached device array of their pointers -- no temporaries, no host
        synchronisation, CUDA-graph capturable."""
        ps = [p for l in self.nerf_mlp for p in (l.weight, l.bias)]
        key = frozen_key(ps)   # None for a trainable model
        if key is not None and key == self._mlp_key and self._mlp_blob is not None:
            return self._mlp_blob
        dev = ps[0].device
        if self._mlp_blob is None or self._mlp_blob.device != dev:
            self._mlp_blob = torch.empty(MLP_WBLOB_BYTES * REPLICAS, dtype=torch.uint8, device=dev)
            self._ptr_key = None
        in_place = all(p.dtype == torch.float32 and p.is_contiguous() for p in ps)
        if in_place:
            pkey = tuple(p.data_ptr() for p in ps)
            if pkey != self._ptr_key:
                self._wp = _ptr_array([l.weight for l in self.nerf_mlp], dev)
                self._bp = _ptr_array([l.bias for l in self.nerf_mlp], dev)
                self._ptr_key = pkey
            with nv.on_device(self._mlp_blob):
                nv.check(nv.lib().pvd_mlp_pack_weights(nv.ptr(self._wp), nv.ptr(self._bp), nv.ptr(self._mlp_blob), nv.stream_of(self._mlp_blob)))
        else:
            ws = [l.weight.detach().float().contiguous() for l in self.nerf_mlp]
            bs = [l.bias.detach().float().contiguous() for l in self.nerf_mlp]
            wp, bp = _ptr_array(ws, dev), _ptr_array(bs, dev)
            with nv.on_device(self._mlp_blob):
                nv.check(nv.lib().pvd_mlp_pack_weights(nv.ptr(wp), nv.ptr(bp), nv.ptr(self._mlp_blob), nv.stream_of(self._mlp_blob)))
            torch.cuda.current_stream(dev).synchronize()  # the temporaries may be freed after this scope
        replicate(self._mlp_blob, MLP_WBLOB_BYTES)
        self._mlp_key = key
        return self._mlp_blob

    def _fused_forward(self, x, d, save_ws=None, enc=None):
        x = x.detach().float().contiguous().view(-1, 3)
        d = d.detach().float().contiguous().view(-1, 3)
        M, dev = x.shape[0], x.device
        tail = self._staged.wblob_for((self.sigma_net[0].weight, self.sigma_net[1].weight, self.color_net[0].weight,
                                       self.color_net[1].weight, self.color_net[2].weight), self.in_dim)
        f = PvdMlpField(wblob=self._blob().data_ptr(), tail_wblob=tail.data_ptr(), sigma_clip_min=float(self.args.sigma_clip_min),
                        sigma_clip_max=float(self.args.sigma_clip_max), density_scale=1.0, replicas=REPLICAS)
        sigmas = torch.empty(M, dtype=torch.float32, device=dev)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        feat = torch.empty(M, 16, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with nv.on_device(x):
            if save_ws is None:
                nv.check(nv.lib().pvd_mlp_field_forward(C.byref(f), nv.ptr(x), nv.ptr(d), C.c_uint32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                        nv.ptr(feat), nv.ptr(status), nv.stream_of(x)))
            else:
                nv.check(nv.lib().pvd_mlp_field_forward_train(C.byref(f), nv.ptr(x), nv.ptr(d), C.c_uint32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                              nv.ptr(feat), nv.ptr(save_ws), nv.ptr(enc), nv.ptr(status), nv.stream_of(x)))
        self._status = status
        return sigmas, rgbs, feat

    # ---------------------------------------------------------------- torch composition (gradients required)
    def _torch_forward(self, x, d):
        from tools.activation import trunc_exp
        h = self.encoder_nerf_pe(x)
        in_pts = h
        for i, layer in enumerate(self.nerf_mlp):          # network.py:324-333
            h = layer(h)
            if i != len(self.nerf_mlp) - 1:
                h = F.relu(h)
            if i == self.skips:
                h = torch.cat([in_pts, h], -1)
        h = F.relu(self.sigma_net[0](h))                   # network.py:413-420
        h = self.sigma_net[1](h)
        h0 = torch.clamp(h[..., 0], self.args.sigma_clip_min, self.args.sigma_clip_max)
        feat = torch.cat([h0.unsqueeze(-1), h[..., 1:]], dim=-1)
        sigma = trunc_exp(h0)
        c = torch.cat([self.encoder_dir(d), h[..., 1:]], dim=-1)
        c = F.relu(self.color_net[0](c))
        c = F.relu(self.color_net[1](c))
        return sigma, torch.sigmoid(self.color_net[2](c)), feat

    def forward(self, x, d):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not needs_grad:
            sigma, color, feat = self._fused_forward(x, d)
        elif TORCH_TRAIN:
            sigma, color, feat = self._torch_forward(x, d)
        else:
            ps = [p for l in self.nerf_mlp for p in (l.weight, l.bias)]
            ps += [self.sigma_net[0].weight, self.sigma_net[1].weight, self.color_net[0].weight, self.color_net[1].weight, self.color_net[2].weight]
            shape = x.shape[:-1]
            sigma, color, feat = _FusedMlpField.apply(x, d, self, *ps)
            sigma, color, feat = sigma.view(shape), color.view(*shape, 3), feat.view(*shape, 16)
        self.feature_sigma_color = feat
        if self.training and self.args.global_step < self.args.stage_iters["stage1"]:
            return None, None
        self.sigma_l = feat[..., 0]
        self.color_l = color
        return sigma, color

    def density(self, x):
        with torch.no_grad():
            x = x.reshape(-1, 3)
            sigma, _, _ = self._fused_forward(x, torch.zeros_like(x))
        return {"sigma": sigma}

    def get_params(self, lr, lr2=1e-3):
        """Optimizer groups of the reference's mlp model (network.py:654-660)."""
        return [{"params": self.sigma_net.parameters(), "lr": lr}, {"params": self.encoder_dir.parameters(), "lr": lr},
                {"params": self.color_net.parameters(), "lr": lr}, {"params": self.nerf_mlp.parameters(), "lr": lr}]
