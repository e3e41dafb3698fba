"""Host side of the fused "mlp" (NeRF) field.

`MLPNeRFField` mirrors `NeRFNetwork(model_type="mlp")` of the reference (distill_mutual/network.py:56-70,103-152,324-333,
413-437): `nerf_mlp.{0..7}` (Linear with bias; 63->256, 256->256 x3, 319->256, 256->256 x2, 256->28), `sigma_net.{0,1}`,
`color_net.{0,1,2}`.  In PVD this model type is the frozen TEACHER of mlp->hash distillation and is evaluated under
torch.no_grad (distill_mutual/utils.py:1008-1018): that forward is one fused tcgen05 kernel (csrc/field_mlp.cu).
When gradients are required (training an mlp model from images) the forward is the plain torch composition of the same
layers -- autograd through cuBLAS, exactly what the reference does; it is not accelerated here.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as nv
from .fused import StagedParams, _Args, frozen_key
from .renderer import NeRFRenderer

MLP_WBLOB_BYTES = 876544


class PvdMlpField(C.Structure):
    _fields_ = [("wblob", C.c_void_p), ("tail_wblob", C.c_void_p), ("sigma_clip_min", C.c_float), ("sigma_clip_max", C.c_float),
                ("density_scale", C.c_float)]


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
        self.feature_sigma_color = None
        self.sigma_l = None
        self.color_l = None

    # ---------------------------------------------------------------- fused (no-grad) forward
    def _blob(self):
        ps = [p for l in self.nerf_mlp for p in (l.weight, l.bias)]
        key = frozen_key(ps)   # None for a trainable model: re-packed on every call (fused.frozen_key)
        if key is None or key != self._mlp_key or self._mlp_blob is None:
            dev = ps[0].device
            if self._mlp_blob is None:
                self._mlp_blob = torch.empty(MLP_WBLOB_BYTES, dtype=torch.uint8, device=dev)
            ws = [l.weight.detach().float().contiguous() for l in self.nerf_mlp]
            bs = [l.bias.detach().float().contiguous() for l in self.nerf_mlp]
            wp = torch.tensor([w.data_ptr() for w in ws], dtype=torch.int64, device=dev)
            bp = torch.tensor([b.data_ptr() for b in bs], dtype=torch.int64, device=dev)
            with nv.on_device(self._mlp_blob):
                nv.check(nv.lib().pvd_mlp_pack_weights(nv.ptr(wp), nv.ptr(bp), nv.ptr(self._mlp_blob), nv.stream_of(self._mlp_blob)))
            torch.cuda.current_stream(dev).synchronize()  # ws / bs / pointer arrays may be freed after this scope
            self._mlp_key = key
        return self._mlp_blob

    def _fused_forward(self, x, d):
        x = x.detach().float().contiguous().view(-1, 3)
        d = d.detach().float().contiguous().view(-1, 3)
        M, dev = x.shape[0], x.device
        tail = self._staged.wblob_for((self.sigma_net[0].weight, self.sigma_net[1].weight, self.color_net[0].weight,
                                       self.color_net[1].weight, self.color_net[2].weight), self.in_dim)
        f = PvdMlpField(wblob=self._blob().data_ptr(), tail_wblob=tail.data_ptr(), sigma_clip_min=float(self.args.sigma_clip_min),
                        sigma_clip_max=float(self.args.sigma_clip_max), density_scale=1.0)
        sigmas = torch.empty(M, dtype=torch.float32, device=dev)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        feat = torch.empty(M, 16, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with nv.on_device(x):
            nv.check(nv.lib().pvd_mlp_field_forward(C.byref(f), nv.ptr(x), nv.ptr(d), C.c_uint32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                    nv.ptr(feat), nv.ptr(status), nv.stream_of(x)))
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
        sigma, color, feat = self._torch_forward(x, d) if needs_grad else self._fused_forward(x, d)
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
