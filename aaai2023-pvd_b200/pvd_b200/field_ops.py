"""Per-field-type launchers the training engines are built from: what `NeRFNetwork.forward` / its autograd backward amount to for
one model type (distill_mutual/network.py:335-437), on pre-allocated buffers, no autograd, no allocation, CUDA-graph safe.

    HashOps  model_type "hash"  (fused.HashNeRFField)     forward + backward   csrc/field_hash.cu
    VmOps    model_type "vm"    (fused_vm.VMNeRFField)    forward + backward   csrc/field_vm.cu
    MlpOps   model_type "mlp"   (fused_mlp.MLPNeRFField)  forward (frozen teacher of mlp -> hash distillation) + backward   csrc/field_mlp*.cu

Common protocol (all sample tensors are the engine's): `stage(density_scale)` refreshes staged parameters (fp16 table shadow,
packed weight tiles); `alloc(M)` sizes per-sample scratch; `forward(st, xyzs, dirs, M, sigmas, rgbs, feat, status)`;
`backward(st, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat, M, n_valid, gw_ws, status)`; `clear_grads()` zeroes the big
parameter-gradient buffers (the engines run it on a side stream); `grads()` returns {reference parameter name: gradient}.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _native as nv
from . import fused

_u32, _f32 = C.c_uint32, C.c_float


class HashOps:
    kind = "hash"
    kernels_fwd = 1

    def __init__(self, field: "fused.HashNeRFField", dev, trainable: bool = True):
        self.field, self.dev, self.trainable = field, torch.device(dev), trainable
        self.grad_table = torch.zeros(field.encoder.embeddings.shape, dtype=torch.float32, device=self.dev) if trainable else None
        self.enc = self.dx_ws = None
        self.fp32 = bool(getattr(field, "fp32", False))   # fp32 end to end: csrc/field_hash_f32.cu (no saved encoding, no packed tiles)
        self.kernels_bwd = 1 if self.fp32 else (2 if fused.SPLIT_SCATTER else 1)
        self.table = self.wblob = self.cfield = None
        ws = self._weights()
        # the small weight gradients in parameter shapes, ONE flat buffer (one memset), for callers with an external optimizer
        self._wflat = torch.zeros(sum(w.numel() for w in ws), dtype=torch.float32, device=self.dev) if trainable else None
        self.wgrads, off = [], 0
        for w in ws if trainable else []:
            self.wgrads.append(self._wflat[off:off + w.numel()].view_as(w))
            off += w.numel()

    def _weights(self):
        f = self.field
        return (f.sigma_net[0].weight, f.sigma_net[1].weight, f.color_net[0].weight, f.color_net[1].weight, f.color_net[2].weight)

    def zero_weight_grads(self):
        self._wflat.zero_()

    def unpack_weight_grads(self, gw_ws, st, zero=True):
        """gw_ws (kernel-native, 16 replicas) -> self.wgrads (parameter shapes): memset (unless the caller has zeroed them already,
        `zero_weight_grads`, off the critical path) + k_unpack_wgrads, which accumulates."""
        if zero:
            self._wflat.zero_()
        nv.check(nv.lib().pvd_field_unpack_wgrads(nv.ptr(gw_ws), _u32(2 * self.cfg.num_levels), *[nv.ptr(g) for g in self.wgrads], st))

    def stage(self, density_scale=1.0):
        f = self.field
        cfg = f.config()
        cfg.density_scale = density_scale
        self.cfg = cfg
        if self.fp32:
            ws = self._weights()
            assert f.encoder.embeddings.dtype == torch.float32 and all(w.dtype == torch.float32 and w.is_contiguous() for w in ws)
            self.table = f.encoder.embeddings.detach()
            self.cfield = fused._cstruct(cfg, self.table, f.encoder.offsets, self.table)
            self.cweights = fused.PvdFieldWeightsF32(*[w.data_ptr() for w in ws])     # read in place: nothing to re-stage after an optimizer step
            return
        self.table = f._staged.table_for(f.encoder.embeddings, cfg.table_fp16)
        self.wblob = f._staged.wblob_for(self._weights(), 2 * cfg.num_levels)
        self.cfield = fused._cstruct(cfg, self.table, f.encoder.offsets, self.wblob)

    def prefetch(self, st):
        """Stream the table the forward gathers from into L2 (a side-stream TMA prefetch; see csrc/optim.cu::k_l2_prefetch)."""
        t = self.table
        nv.check(nv.lib().pvd_l2_prefetch(nv.ptr(t), C.c_uint64(t.numel() * t.element_size()), st))

    def alloc(self, M):
        if self.trainable and not self.fp32:
            self.enc = torch.empty(M, fused.ENC_STRIDE, dtype=torch.float16, device=self.dev)
            self.dx_ws = torch.empty(M, fused.ENC_STRIDE, dtype=torch.float16, device=self.dev) if fused.SPLIT_SCATTER else None

    def forward(self, st, xyzs, dirs, M, sigmas, rgbs, feat, status):
        if self.fp32:
            nv.check(nv.lib().pvd_hash_field_forward_f32(C.byref(self.cfield), C.byref(self.cweights), nv.ptr(xyzs), nv.ptr(dirs), _u32(M),
                                                         nv.ptr(sigmas), nv.ptr(rgbs), nv.ptr(feat), st))
            return
        nv.check(nv.lib().pvd_hash_field_forward(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), _u32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                 nv.ptr(self.enc), nv.ptr(feat), nv.ptr(status), st))

    def backward(self, st, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat, M, n_valid, gw_ws, status, phases=None):
        """phases: None = the whole backward; PVD_BWD_MLP (1) / PVD_BWD_SCATTER (2) = one of its two kernels (needs dx_ws)."""
        if self.fp32:
            nv.check(nv.lib().pvd_hash_field_backward_f32(C.byref(self.cfield), C.byref(self.cweights), nv.ptr(xyzs), nv.ptr(dirs),
                                                          nv.ptr(grad_sigmas), nv.ptr(grad_rgbs), nv.ptr(grad_feat), _u32(M), nv.ptr(n_valid),
                                                          nv.ptr(self.grad_table), nv.ptr(gw_ws), st))
            return
        if phases is None:
            nv.check(nv.lib().pvd_hash_field_backward(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(self.enc),
                                                      nv.ptr(grad_sigmas), nv.ptr(grad_rgbs), nv.ptr(grad_feat), _u32(M), nv.ptr(n_valid),
                                                      nv.ptr(self.grad_table), nv.ptr(gw_ws), nv.ptr(self.dx_ws), nv.ptr(status), st))
        else:
            nv.check(nv.lib().pvd_hash_field_backward_rows(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(self.enc),
                                                           nv.ptr(grad_sigmas), nv.ptr(grad_rgbs), nv.ptr(grad_feat), _u32(0), _u32(M),
                                                           nv.ptr(n_valid), nv.ptr(self.grad_table), nv.ptr(gw_ws), nv.ptr(self.dx_ws),
                                                           nv.ptr(status), _u32(phases), st))

    def clear_grads(self):
        self.grad_table.zero_()

    def big_grad(self):
        """The flat fp32 buffer that dominates the multi-GPU gradient exchange."""
        return self.grad_table.view(-1)

    def regularise(self, st, loss_scale, loss_slots, weight):
        pass

    def weight_grads(self, gw_ws):
        f = self.field
        like = (f.sigma_net[0].weight, f.sigma_net[1].weight, f.color_net[0].weight, f.color_net[1].weight, f.color_net[2].weight)
        g = fused.unpack_wgrads(gw_ws, 2 * self.cfg.num_levels, like)
        return dict(zip(("sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight", "color_net.2.weight"), g))

    def grads(self, gw_ws):
        out = {"encoder.embeddings": self.grad_table}
        out.update(self.weight_grads(gw_ws))
        return out

    def algorithmic_bytes(self):
        """(forward, backward) bytes per sample: L x 8 corners x 2 features x 2 B gathered; fp32 reductions + the saved encoding."""
        L = self.cfg.num_levels
        entry = 2 if self.cfg.table_fp16 else 4   # bytes per feature gathered: fp16 shadow or the fp32 master
        return L * 8 * 2 * entry, (L * 8 * 2 * 4 + 64) if self.trainable else 0


# vm backward as two kernels (MLP kernel -> workspace -> high-occupancy scatter kernel; default) or as one (0)
VM_SPLIT_SCATTER = os.environ.get("PVD_VM_SPLIT_SCATTER", "1") != "0"
# PVD_VM_PLANE_F16=1: the engines gather the vm planes / lines from an fp16 channels-last SHADOW (2304 B/sample instead of 4608),
# refreshed by stage() or written by the fused optimizer; masters and gradients stay fp32.  MEASURED on B200 (4096 rays, 300^3,
# profiles/README.md): forward 45.6 -> 39.9 us, backward 123.7 -> 120.6 us -- the gather is bound by requests, not bytes -- while
# re-casting 69 MB of planes per step costs 16 us when an external optimizer owns the parameters (step 209.9 -> 225.7 us) and the
# fused optimizer only gains 1 % (324.0 -> 321.2 us).  So the default gathers the fp32 parameters in place, like the module path.
VM_PLANE_F16 = os.environ.get("PVD_VM_PLANE_F16", "0") != "0"


class PvdCastDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_uint64)]


class VmOps:
    kind = "vm"
    kernels_fwd = 1

    def __init__(self, field, dev, trainable: bool = True):
        from . import fused_vm
        self._vm = fused_vm
        self.field, self.dev, self.trainable = field, torch.device(dev), trainable
        self.groups = [list(field.sigma_mat), list(field.sigma_vec), list(field.color_mat), list(field.color_vec)]
        for grp in self.groups:
            for p in grp:
                assert p.is_contiguous(memory_format=torch.channels_last), "vm planes/lines must be torch.channels_last"
        self._flat = None
        self.scatter_ws = None
        self.kernels_bwd = 2 if VM_SPLIT_SCATTER else 1
        self.wblob = self.cfield = None
        self.plane_f16 = VM_PLANE_F16
        self._shadow = self._cast_descs = None
        self.shadow_groups = None
        self._shadow_key = None
        self._aabb = [float(v) for v in field.aabb_train.tolist()]   # read once: stage() must stay free of host syncs (graph capture)
        ws = self._weights()
        self._wflat = torch.zeros(sum(w.numel() for w in ws), dtype=torch.float32, device=self.dev) if trainable else None
        self.wgrads, off = [], 0
        for w in ws if trainable else []:
            self.wgrads.append(self._wflat[off:off + w.numel()].view_as(w))
            off += w.numel()
        if trainable:
            # every plane/line gradient lives in ONE flat buffer (one memset per step); each view has its parameter's shape and
            # channels-last strides ([H][W][R] in memory), sigma planes and lines first (the L1 penalty covers exactly that prefix)
            total = sum(p.numel() for grp in self.groups for p in grp)
            self._flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
            off, self.grad_groups = 0, []
            for grp in (self.groups[0], self.groups[1], self.groups[2], self.groups[3]):
                views = []
                for p in grp:
                    _, R, H, W = p.shape
                    views.append(self._flat[off:off + p.numel()].view(1, H, W, R).permute(0, 3, 1, 2))
                    off += p.numel()
                self.grad_groups.append(views)
            self._cgrads = fused_vm.PvdVmGrads(sigma_mat=fused_vm._ptrs3(self.grad_groups[0]), sigma_vec=fused_vm._ptrs3(self.grad_groups[1]),
                                               color_mat=fused_vm._ptrs3(self.grad_groups[2]), color_vec=fused_vm._ptrs3(self.grad_groups[3]))

    def _weights(self):
        f = self.field
        return (f.basis_mat.weight, f.color_net[0].weight, f.color_net[1].weight, f.color_net[2].weight)

    def zero_weight_grads(self):
        self._wflat.zero_()

    def unpack_weight_grads(self, gw_ws, st, zero=True):
        if zero:
            self._wflat.zero_()
        nv.check(nv.lib().pvd_vm_unpack_wgrads(nv.ptr(gw_ws), *[nv.ptr(g) for g in self.wgrads], st))

    def stage(self, density_scale=1.0):
        f = self.field
        self.wblob = f._staged.get(self._weights())
        aabb = self._aabb
        if self.plane_f16:
            planes = self._stage_shadow()
        else:
            planes = [[p.detach() for p in grp] for grp in self.groups]
        self.cfield = self._vm._vm_struct(planes, self.wblob, f.resolution, aabb, float(f.args.sigma_clip_min), float(f.args.sigma_clip_max),
                                          float(density_scale))

    def _stage_shadow(self):
        """fp16 channels-last shadows of the 12 plane / line tensors in ONE flat buffer (same order as the gradient buffer), cast by
        one multi-tensor launch.  Frozen parameters (a teacher) are cast once (fused.frozen_key)."""
        params = [p for grp in self.groups for p in grp]
        if self._shadow is None:
            total = sum(p.numel() for p in params)
            self._shadow = torch.empty(total, dtype=torch.float16, device=self.dev)
            descs = (PvdCastDesc * len(params))()
            off, self.shadow_groups, flat = 0, [], []
            for grp in self.groups:
                views = []
                for p in grp:
                    assert p.dtype == torch.float32
                    v = self._shadow[off:off + p.numel()]
                    descs[len(flat)] = PvdCastDesc(src=p.data_ptr(), dst=v.data_ptr(), n=p.numel())
                    views.append(v); flat.append(v)
                    off += p.numel()
                self.shadow_groups.append(views)
            self.shadow_flat = flat
            raw = bytes(memoryview(descs).cast("B"))
            self._cast_descs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.dev)
            self._cast_n = (len(params), max(p.numel() for p in params))
        key = fused.frozen_key(params)
        if key is None or key != self._shadow_key:
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            nv.check(nv.lib().pvd_cast_f32_to_f16_multi(nv.ptr(self._cast_descs), _u32(self._cast_n[0]), C.c_uint64(self._cast_n[1]), st))
            self._shadow_key = key
        return self.shadow_groups

    def prefetch(self, st):
        groups = self.shadow_groups if (self.plane_f16 and self.shadow_groups) else self.groups
        for grp in groups:
            for p in grp:
                nv.check(nv.lib().pvd_l2_prefetch(nv.ptr(p), C.c_uint64(p.numel() * p.element_size()), st))

    def alloc(self, M):
        if self.trainable and VM_SPLIT_SCATTER:
            l = nv.lib()
            l.pvd_vm_backward_workspace_bytes.restype = C.c_uint64
            self.scatter_ws = torch.empty(int(l.pvd_vm_backward_workspace_bytes(_u32(M))), dtype=torch.uint8, device=self.dev)

    def forward(self, st, xyzs, dirs, M, sigmas, rgbs, feat, status):
        nv.check(nv.lib().pvd_vm_field_forward(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), _u32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                               nv.ptr(feat), nv.ptr(status), st))

    def backward(self, st, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat, M, n_valid, gw_ws, status):
        if self.scatter_ws is not None:
            nv.check(nv.lib().pvd_vm_field_backward_ws(C.byref(self.cfield), C.byref(self._cgrads), nv.ptr(xyzs), nv.ptr(dirs),
                                                       nv.ptr(grad_sigmas), nv.ptr(grad_rgbs), nv.ptr(grad_feat), _u32(M), nv.ptr(n_valid),
                                                       nv.ptr(gw_ws), nv.ptr(self.scatter_ws), nv.ptr(status), st))
            return
        nv.check(nv.lib().pvd_vm_field_backward(C.byref(self.cfield), C.byref(self._cgrads), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(grad_sigmas),
                                                nv.ptr(grad_rgbs), nv.ptr(grad_feat), _u32(M), nv.ptr(n_valid), nv.ptr(gw_ws),
                                                nv.ptr(status), st))

    def clear_grads(self):
        self._flat.zero_()

    def big_grad(self):
        return self._flat

    def regularise(self, st, loss_scale, loss_slots, weight):
        """l1_reg_weight * density_loss() (network.py:549-557; added to the loss for vm models, utils.py:1135-1136): its gradient
        goes onto the sigma plane / line gradients, its value into the step's loss slots."""
        if not weight:
            return
        for p, g in zip(self.groups[0] + self.groups[1], self.grad_groups[0] + self.grad_groups[1]):
            nv.check(nv.lib().pvd_l1_mean_reg(nv.ptr(p), C.c_uint64(p.numel()), _f32(weight), _f32(loss_scale), nv.ptr(g), nv.ptr(loss_slots), st))

    def weight_grads(self, gw_ws):
        f = self.field
        ws = (f.basis_mat.weight, f.color_net[0].weight, f.color_net[1].weight, f.color_net[2].weight)
        wg = [torch.zeros_like(w, dtype=torch.float32) for w in ws]
        with nv.on_device(gw_ws):
            nv.check(nv.lib().pvd_vm_unpack_wgrads(nv.ptr(gw_ws), nv.ptr(wg[0]), nv.ptr(wg[1]), nv.ptr(wg[2]), nv.ptr(wg[3]), nv.stream_of(gw_ws)))
        return dict(zip(("basis_mat.weight", "color_net.0.weight", "color_net.1.weight", "color_net.2.weight"), wg))

    def grads(self, gw_ws):
        out = {}
        for name, views in zip(("sigma_mat", "sigma_vec", "color_mat", "color_vec"), self.grad_groups):
            for i, v in enumerate(views):
                out[f"{name}.{i}"] = v
        out.update(self.weight_grads(gw_ws))
        return out

    def algorithmic_bytes(self):
        """(forward gather, backward reductions) per sample: 3 (plane, line) pairs x (4 + 2) taps x (16 + 48) components x 2 B (fp16
        shadow) or 4 B gathered; always 4 B per component reduced (fp32 gradients)."""
        taps = 3 * (4 + 2) * (16 + 48)
        return taps * (2 if self.plane_f16 else 4), (taps * 4 if self.trainable else 0)


class MlpOps:
    """model_type "mlp" (fused_mlp.MLPNeRFField).  Frozen teacher: forward only (csrc/field_mlp.cu).  Trainable (main_just_train_tea.py
    --model_type mlp): the forward saves its operand tiles and the backward is tail kernel -> k_mlp_trunk_bwd -> k_mlp_wgrad
    (csrc/field_mlp_bwd.cu); all scratch is allocated in alloc(M), nothing in the step."""
    kind = "mlp"
    kernels_fwd = 1
    NAMES = tuple(f"nerf_mlp.{i}.{k}" for i in range(8) for k in ("weight", "bias")) + \
        ("sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight", "color_net.2.weight")

    def __init__(self, field, dev, trainable: bool = False):
        from . import fused_mlp
        self._fm = fused_mlp
        self.field, self.dev, self.trainable = field, torch.device(dev), trainable
        self.kernels_bwd = 3 if trainable else 0
        self.save_ws = self.grad_ws = self.enc = self.d_x28 = None
        self.grad_table = None
        self.wgrads = []
        if trainable:
            d = self.dev
            ps = self._params()
            assert all(p.dtype == torch.float32 and p.is_contiguous() for p in ps), "trainable mlp parameters must be contiguous fp32"
            self._wflat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=d)
            off = 0
            for p_ in ps:
                self.wgrads.append(self._wflat[off:off + p_.numel()].view_as(p_))
                off += p_.numel()
            self.gw_mlp = torch.zeros(fused_mlp.MLP_GW_FLOATS, dtype=torch.float32, device=d)
            self.wblob_t = torch.empty(fused_mlp.MLP_WBLOB_T_BYTES * fused_mlp.REPLICAS, dtype=torch.uint8, device=d)
            # device arrays of device pointers (the ABI's float* const*): parameters and their gradient views never move
            self._wp = fused_mlp._ptr_array([l.weight for l in field.nerf_mlp], d)
            self._gwp = fused_mlp._ptr_array(self.wgrads[0:16:2], d)
            self._gbp = fused_mlp._ptr_array(self.wgrads[1:16:2], d)
            self._offsets = torch.arange(17, dtype=torch.int32, device=d) * 8
            torch.cuda.current_stream(d).synchronize()

    def _params(self):
        f = self.field
        return [p for l in f.nerf_mlp for p in (l.weight, l.bias)] + [f.sigma_net[0].weight, f.sigma_net[1].weight, f.color_net[0].weight,
                                                                      f.color_net[1].weight, f.color_net[2].weight]

    def _tail(self):
        f = self.field
        return (f.sigma_net[0].weight, f.sigma_net[1].weight, f.color_net[0].weight, f.color_net[1].weight, f.color_net[2].weight)

    def stage(self, density_scale=1.0):
        f = self.field
        self.tail = f._staged.wblob_for(self._tail(), f.in_dim)
        self.blob = f._blob()
        self.cfield = self._fm.PvdMlpField(wblob=self.blob.data_ptr(), tail_wblob=self.tail.data_ptr(), sigma_clip_min=float(f.args.sigma_clip_min),
                                           sigma_clip_max=float(f.args.sigma_clip_max), density_scale=float(density_scale), replicas=self._fm.REPLICAS)
        if self.trainable:
            nv.check(nv.lib().pvd_mlp_pack_weights_t(nv.ptr(self._wp), nv.ptr(self.wblob_t), nv.stream_of(self.wblob_t)))
            self._fm.replicate(self.wblob_t, self._fm.MLP_WBLOB_T_BYTES)
            cfg = fused.HashFieldConfig(num_levels=14, base_resolution=16, per_level_scale=2.0, bound=1.0, sigma_clip_min=float(f.args.sigma_clip_min),
                                        sigma_clip_max=float(f.args.sigma_clip_max), density_scale=float(density_scale))
            self._tail_cfg = cfg

    def alloc(self, M):
        if self.trainable:
            tiles = max((M + 127) // 128, 1)
            self.save_ws = torch.empty(tiles * self._fm.MLP_SAVE_TILE_BYTES, dtype=torch.uint8, device=self.dev)
            self.grad_ws = torch.empty(tiles * self._fm.MLP_GRAD_TILE_BYTES, dtype=torch.uint8, device=self.dev)
            self.enc = torch.empty(M, fused.ENC_STRIDE, dtype=torch.float16, device=self.dev)
            self.d_x28 = torch.zeros(M, fused.ENC_STRIDE, dtype=torch.float16, device=self.dev)

    def prefetch(self, st):
        pass   # 876 KB of weights: L2-resident after the first tile

    def forward(self, st, xyzs, dirs, M, sigmas, rgbs, feat, status):
        if self.trainable:
            nv.check(nv.lib().pvd_mlp_field_forward_train(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), _u32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                          nv.ptr(feat), nv.ptr(self.save_ws), nv.ptr(self.enc), nv.ptr(status), st))
        else:
            nv.check(nv.lib().pvd_mlp_field_forward(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), _u32(M), nv.ptr(sigmas), nv.ptr(rgbs),
                                                    nv.ptr(feat), nv.ptr(status), st))

    def backward(self, st, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat, M, n_valid, gw_ws, status, phases=None):
        l = nv.lib()
        f = fused._cstruct(self._tail_cfg, self.enc, self._offsets, self.tail)
        # 1. sigma / colour tail: the hash model's tail backward in its d(encoding)-export mode; x28 plays the encoding
        nv.check(l.pvd_hash_field_backward_rows(C.byref(f), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(self.enc), nv.ptr(grad_sigmas), nv.ptr(grad_rgbs),
                                                nv.ptr(grad_feat), _u32(0), _u32(M), nv.ptr(n_valid), nv.ptr(gw_ws), nv.ptr(gw_ws), nv.ptr(self.d_x28),
                                                nv.ptr(status), _u32(1), st))
        # 2. data gradients through layers 7..1; 3. weight / bias gradients (reduction over all samples in TMEM)
        nv.check(l.pvd_mlp_trunk_backward(nv.ptr(self.wblob_t), _u32(self._fm.REPLICAS), nv.ptr(self.save_ws), nv.ptr(self.d_x28), _u32(M), nv.ptr(n_valid),
                                          nv.ptr(self.grad_ws), nv.ptr(status), st))
        nv.check(l.pvd_mlp_weight_grads(nv.ptr(self.save_ws), nv.ptr(self.grad_ws), _u32(M), nv.ptr(self.gw_mlp), nv.ptr(status), st))

    def clear_grads(self):
        self.gw_mlp.zero_()

    def big_grad(self):
        return self.gw_mlp

    def regularise(self, st, loss_scale, loss_slots, weight):
        pass

    def zero_weight_grads(self):
        self._wflat.zero_()

    def unpack_weight_grads(self, gw_ws, st, zero=True):
        """kernel-native workspaces -> self.wgrads (parameter shapes, NAMES order)."""
        if zero:
            self._wflat.zero_()
        nv.check(nv.lib().pvd_mlp_unpack_wgrads(nv.ptr(self.gw_mlp), nv.ptr(self._gwp), nv.ptr(self._gbp), st))
        nv.check(nv.lib().pvd_field_unpack_wgrads(nv.ptr(gw_ws), _u32(self.field.in_dim), *[nv.ptr(g) for g in self.wgrads[16:]], st))

    def weight_grads(self, gw_ws):
        self.unpack_weight_grads(gw_ws, nv.stream_of(gw_ws))
        return {n: g.clone() for n, g in zip(self.NAMES, self.wgrads)}

    def grads(self, gw_ws):
        return self.weight_grads(gw_ws)

    def algorithmic_bytes(self):
        return 0, 0   # FLOP-bound: 865 280 FLOP/sample forward


class TensorsOps:
    """model_type "tensors" (fused_tensors.TensorsNeRFField): one trilinear gather of the channels-last-3d volume, no MLP.
    csrc/field_tensors.cu.  Trains from images (FieldTrainEngine); the reference leaves feature_sigma_color = None for this type
    (network.py:407), so it takes no part in the feature losses of a distillation pair."""
    kind = "tensors"
    kernels_fwd = 1
    kernels_bwd = 1

    def __init__(self, field, dev, trainable: bool = True):
        from . import fused_tensors
        self._ft = fused_tensors
        self.field, self.dev, self.trainable = field, torch.device(dev), trainable
        vol = field.tensor_volume[0]
        assert vol.is_contiguous(memory_format=torch.channels_last_3d), "the plenoxel volume must be torch.channels_last_3d"
        self._flat = torch.zeros(vol.numel(), dtype=torch.float32, device=self.dev) if trainable else None
        self.grad_volume = None
        if trainable:
            _, Cc, D, H, W = vol.shape
            self.grad_volume = self._flat.view(1, D, H, W, Cc).permute(0, 4, 1, 2, 3)   # the parameter's shape, channels-last-3d memory
        self._aabb = [float(v) for v in field.aabb_train.tolist()]
        self.cfield = None
        self.wgrads = []

    def stage(self, density_scale=1.0):
        f = self.field
        self.cfield = self._ft.tensors_struct(f.tensor_volume[0].detach(), f.plenoxel_degree, self._aabb, float(f.args.sigma_clip_min),
                                              float(f.args.sigma_clip_max), float(density_scale))

    def alloc(self, M):
        pass

    def prefetch(self, st):
        pass   # 235 MB at 128^3: larger than L2, gathered from HBM either way

    def forward(self, st, xyzs, dirs, M, sigmas, rgbs, feat, status):
        assert feat is None, "the tensors field has no feature_sigma_color (network.py:407)"
        nv.check(nv.lib().pvd_tensors_field_forward(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), _u32(M), nv.ptr(sigmas), nv.ptr(rgbs), st))

    def backward(self, st, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat, M, n_valid, gw_ws, status):
        nv.check(nv.lib().pvd_tensors_field_backward(C.byref(self.cfield), nv.ptr(xyzs), nv.ptr(dirs), nv.ptr(grad_sigmas), nv.ptr(grad_rgbs),
                                                     _u32(M), nv.ptr(n_valid), nv.ptr(self._flat), st))

    def clear_grads(self):
        self._flat.zero_()

    def big_grad(self):
        return self._flat

    def regularise(self, st, loss_scale, loss_slots, weight):
        pass

    def zero_weight_grads(self):
        pass

    def unpack_weight_grads(self, gw_ws, st, zero=True):
        pass

    def weight_grads(self, gw_ws):
        return {}

    def grads(self, gw_ws):
        return {"tensor_volume.0": self.grad_volume}

    def algorithmic_bytes(self):
        """8 corners x C channels x 4 B gathered forward; the same again re-gathered + reduced backward."""
        b = 8 * (3 * self.field.plenoxel_degree ** 2 + 1) * 4
        return b, (2 * b if self.trainable else 0)


def make_ops(field, dev, trainable):
    mt = getattr(field, "model_type", None)
    if mt == "hash":
        return HashOps(field, dev, trainable)
    if mt == "vm":
        return VmOps(field, dev, trainable)
    if mt == "mlp":
        return MlpOps(field, dev, trainable)
    if mt == "tensors":
        return TensorsOps(field, dev, trainable)
    raise ValueError(f"no fused field for model_type {mt!r} (hash | vm | mlp | tensors)")
