"""Fused optimizer step for the training engines (include/pvd_b200_optim.h, csrc/optim.cu).

What the reference does around every backward (distill_mutual/utils.py:802-819 `optimizer.zero_grad(); scaler.scale(loss).backward();
scaler.step(optimizer); scaler.update()` with `torch.optim.AdamW(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15)`,
main_distill_mutual.py:327-339; plus the encoder wrapper's per-forward `embeddings.to(half)` and per-backward `zeros_like`,
gridencoder/grid.py:52,106) is here ONE multi-tensor launch over the engine's parameters:

    unscale (x 1/loss_scale) | found_inf skip | AdamW on the fp32 masters | fp16 table shadow | gradient zeroing

followed by the re-pack of the five (four for vm) small weight matrices into tensor-core tiles.  With a `FusedAdamW` attached, an
engine needs neither `stage()` nor the per-step memset of its big gradient buffer: the optimizer leaves both in place.

Parameter groups and learning rates are the reference's (`field.get_params(lr, lr2)`, network.py:646-683).  State lives in
device memory, so `step()` can be captured into a CUDA graph together with the training step (`engine.capture_iteration`).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nv


class PvdAdamSlot(C.Structure):
    _fields_ = [("param", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("grad", C.c_void_p),
                ("grad_f16", C.c_void_p), ("shadow_f16", C.c_void_p), ("n", C.c_uint64), ("lr", C.c_double),
                ("weight_decay", C.c_double), ("neg_step_size", C.c_float), ("decay", C.c_float), ("zero_grad", C.c_uint32),
                ("grad_mul", C.c_float)]


class PvdAdamState(C.Structure):
    _fields_ = [("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_float), ("grad_scale", C.c_float), ("step", C.c_int32),
                ("found_inf", C.c_int32), ("skipped", C.c_int32), ("flags", C.c_uint32), ("w1", C.c_float), ("w2", C.c_float),
                ("beta2_f", C.c_float), ("inv_bc2_sqrt", C.c_float)]


ADDCMUL_LEFT = 1


def _dense(p: torch.Tensor) -> bool:
    """True when the tensor's elements occupy one gap-free block of memory (any permutation of a contiguous layout: row-major,
    channels_last, channels_last_3d, ...)."""
    dims = sorted(((st, sz) for st, sz in zip(p.stride(), p.shape) if sz > 1), key=lambda t: t[0])
    expect = 1
    for st, sz in dims:
        if st != expect:
            return False
        expect *= sz
    return True


class FusedAdamW:
    """AdamW + GradScaler.unscale_/step over explicit (param, grad) pairs; see module docstring.

    `entries`: list of dicts {param, grad, lr, shadow (optional fp16 tensor), grad_f16 (optional), zero_grad (bool)} whose tensors
    are dense fp32 blocks of equal numel (layouts of param and grad must match element for element); `grad_mul` (default 1) is a
    per-tensor factor on the gradient (rank-count factors of the multi-GPU exchange)."""

    def __init__(self, entries, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.01, loss_scale=1.0, check_finite=True, flags=0, device=None):
        assert entries, "nothing to optimise"
        self.dev = torch.device(device) if device is not None else entries[0]["param"].device
        self.entries = entries
        self.check_finite = bool(check_finite)
        self.exp_avg, self.exp_avg_sq = [], []
        slots = (PvdAdamSlot * len(entries))()
        for i, e in enumerate(entries):
            p, g = e["param"], e.get("grad")
            assert p.dtype == torch.float32 and _dense(p), "fused AdamW needs dense fp32 parameters"
            n = p.numel()
            assert g is None or (g.dtype == torch.float32 and g.numel() == n)
            m = torch.zeros(n, dtype=torch.float32, device=self.dev)
            v = torch.zeros(n, dtype=torch.float32, device=self.dev)
            self.exp_avg.append(m); self.exp_avg_sq.append(v)
            sh, gh = e.get("shadow"), e.get("grad_f16")
            assert sh is None or (sh.dtype == torch.float16 and sh.numel() == n)
            assert gh is None or (gh.dtype == torch.float16 and gh.numel() == n)
            slots[i] = PvdAdamSlot(param=p.data_ptr(), exp_avg=m.data_ptr(), exp_avg_sq=v.data_ptr(), grad=g.data_ptr() if g is not None else None,
                                   grad_f16=gh.data_ptr() if gh is not None else None, shadow_f16=sh.data_ptr() if sh is not None else None,
                                   n=n, lr=float(e["lr"]), weight_decay=float(e.get("weight_decay", weight_decay)), neg_step_size=0.0,
                                   decay=1.0, zero_grad=1 if e.get("zero_grad", True) else 0, grad_mul=float(e.get("grad_mul", 1.0)))
        self.n_slots = len(entries)
        self.max_n = max(e["param"].numel() for e in entries)
        self._slots_host = slots
        self.slots = torch.empty(C.sizeof(slots), dtype=torch.uint8, device=self.dev)
        self._upload(self.slots, slots)
        st = PvdAdamState(beta1=float(betas[0]), beta2=float(betas[1]), eps=float(eps), grad_scale=1.0 / float(loss_scale),
                          step=0, found_inf=0, skipped=0, flags=int(flags), w1=0.0, w2=0.0, beta2_f=0.0, inv_bc2_sqrt=1.0)
        self._state_host = st
        self.state = torch.empty(C.sizeof(st), dtype=torch.uint8, device=self.dev)
        self._upload(self.state, st)
        self.post_step = []   # callables(stream) run after the update (re-pack of the weight tiles)

    def _upload(self, dst: torch.Tensor, cstruct):
        raw = bytes(memoryview(cstruct).cast("B"))
        dst.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))

    # ------------------------------------------------------------------ host-visible knobs (a sync each: not for the hot loop)
    def read_state(self) -> PvdAdamState:
        raw = bytes(self.state.cpu().numpy().tobytes())
        return PvdAdamState.from_buffer_copy(raw)

    def set_lr(self, lrs):
        """New learning rate per slot (the reference's LambdaLR / CosineAnnealingLR schedulers, utils.py:604-610, update lr per step)."""
        for i, lr in enumerate(lrs):
            self._slots_host[i].lr = float(lr)
        self._upload(self.slots, self._slots_host)   # the derived fields are recomputed by the next step's advance kernel

    def set_loss_scale(self, loss_scale: float):
        st = self.read_state()
        st.grad_scale = 1.0 / float(loss_scale)
        self._upload(self.state, st)

    # ------------------------------------------------------------------ the step
    def step(self, stream=None):
        """Enqueue one optimizer step on `stream` (default: the current stream).  CUDA-graph capturable."""
        l = nv.lib()
        st = C.c_void_p((stream or torch.cuda.current_stream(self.dev)).cuda_stream)
        sp, ss = C.c_void_p(self.state.data_ptr()), C.c_void_p(self.slots.data_ptr())
        if self.check_finite:   # GradScaler's found_inf: one read pass over every gradient before anything is updated
            nv.check(l.pvd_grad_nonfinite_slots(sp, ss, C.c_uint32(self.n_slots), C.c_uint64(self.max_n), st))
        nv.check(l.pvd_adamw_advance(sp, ss, C.c_uint32(self.n_slots), st))
        nv.check(l.pvd_adamw_step(sp, ss, C.c_uint32(self.n_slots), C.c_uint64(self.max_n), st))
        for fn in self.post_step:
            fn(st)

    @property
    def kernels_per_step(self):
        return (1 if self.check_finite else 0) + 3 + getattr(self, "kernels_extra", 0)   # check | advance, step, finish | unpack, pack


def _lr_of(field, lr, lr2):
    """{id(param): lr} from the reference's optimizer groups (network.py:646-683)."""
    out = {}
    for g in field.get_params(lr, lr2):
        for p in g["params"]:
            out[id(p)] = g["lr"]
    return out


def for_engine(engine, lr=1e-2, lr2=1e-3, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.01, check_finite=True, exchange=None, flags=0):
    """FusedAdamW over every trainable parameter of `engine`'s model (hash | vm), wired to the engine's gradient buffers:
    the big buffer (hash table / vm planes) is read in place (or from the exchanged fp16 payload), zeroed in the same pass, and
    for a hash model the fp16 shadow the field kernels gather from is written too; the small weight gradients are unpacked from the
    kernels' replicated workspace first and the weight tiles re-packed afterwards.

    Multi-GPU (`exchange` = dist.TableGradExchange of W ranks): the payload holds sum_r g_r / W, the small workspace sum_r g_r.
    engine.grad_reduction "mean" (per-ray mean losses, every rank normalises by its local ray count: the global-batch gradient is
    the MEAN of the ranks') or "sum" (global-norm losses whose coefficients already are global, engine.PairDistillEngine)."""
    ops, field = engine.ops, engine.field
    if getattr(ops, "cfield", None) is None:
        engine.stage()   # creates the staged buffers (fp16 shadow, weight tiles) the optimizer keeps current from here on
    lrs = _lr_of(field, lr, lr2)
    l = nv.lib()
    entries = []
    W = int(exchange.world) if exchange is not None else 1
    mean = getattr(engine, "grad_reduction", "mean") == "mean"
    big_mul = 1.0 if (exchange is None or mean) else float(W)
    small_mul = (1.0 / W) if (exchange is not None and mean) else 1.0
    gh_all = exchange.payload if exchange is not None else None
    if ops.kind == "hash":
        emb = field.encoder.embeddings
        shadow = ops.table if ops.table.dtype == torch.float16 else None
        entries.append(dict(param=emb.data, grad=ops.grad_table, lr=lrs[id(emb)], shadow=shadow, grad_f16=gh_all, zero_grad=True, grad_mul=big_mul))
        ws = [field.sigma_net[0].weight, field.sigma_net[1].weight, field.color_net[0].weight, field.color_net[1].weight, field.color_net[2].weight]
        wg = [torch.zeros_like(w, dtype=torch.float32) for w in ws]
        in_dim = 2 * ops.cfg.num_levels

        def pre(st):
            nv.check(l.pvd_field_unpack_wgrads(nv.ptr(engine.gw_ws), C.c_uint32(in_dim), *[nv.ptr(g) for g in wg], st))

        def post(st):
            if not getattr(ops, "fp32", False):     # the fp32 kernels read the parameters in place
                nv.check(l.pvd_field_pack_weights(*[nv.ptr(w.data) for w in ws], C.c_uint32(in_dim), nv.ptr(ops.wblob), st))
    elif ops.kind == "vm":
        off, k = 0, 0
        for grp, views in zip(ops.groups, ops.grad_groups):
            for p, gv in zip(grp, views):
                n = p.numel()
                gh = gh_all[off:off + n] if gh_all is not None else None
                assert gv.data_ptr() == ops._flat[off:off + n].data_ptr()
                sh = ops.shadow_flat[k] if ops.plane_f16 else None   # the fp16 shadow the field kernels gather from
                entries.append(dict(param=p.data, grad=ops._flat[off:off + n], lr=lrs[id(p)], grad_f16=gh, shadow=sh, zero_grad=True, grad_mul=big_mul))
                off += n
                k += 1
        ws = [field.basis_mat.weight, field.color_net[0].weight, field.color_net[1].weight, field.color_net[2].weight]
        wg = [torch.zeros_like(w, dtype=torch.float32) for w in ws]

        def pre(st):
            nv.check(l.pvd_vm_unpack_wgrads(nv.ptr(engine.gw_ws), *[nv.ptr(g) for g in wg], st))

        def post(st):
            nv.check(l.pvd_vm_pack_weights(*[nv.ptr(w.data) for w in ws], nv.ptr(ops.wblob), st))
    elif ops.kind == "tensors":
        vol = field.tensor_volume[0]
        entries.append(dict(param=vol.data, grad=ops._flat, lr=lrs[id(vol)], grad_f16=gh_all, zero_grad=True, grad_mul=big_mul))
        ws, wg = [], []
        pre = post = lambda st: None
    elif ops.kind == "mlp":
        assert exchange is None, "the mlp model's fused optimizer is single-GPU"
        from .fused_mlp import _ptr_array
        ws = ops._params()                     # nerf_mlp.{0..7}.{weight,bias}, sigma_net.{0,1}.weight, color_net.{0,1,2}.weight
        wg = [torch.zeros_like(w, dtype=torch.float32) for w in ws]
        gwp, gbp = _ptr_array(wg[0:16:2], engine.dev), _ptr_array(wg[1:16:2], engine.dev)
        torch.cuda.current_stream(engine.dev).synchronize()
        opt_ptrs = (gwp, gbp)                  # kept alive by the closure below

        def pre(st):
            nv.check(l.pvd_mlp_unpack_wgrads(nv.ptr(ops.gw_mlp), nv.ptr(opt_ptrs[0]), nv.ptr(opt_ptrs[1]), st))
            nv.check(l.pvd_field_unpack_wgrads(nv.ptr(engine.gw_ws), C.c_uint32(field.in_dim), *[nv.ptr(g) for g in wg[16:]], st))
            ops.gw_mlp.zero_()                 # the weight-gradient kernel accumulates into it

        def post(st):
            ops.stage(engine.density_scale)    # forward stream, transposed stream and tail tiles re-packed from the updated masters
    else:
        raise ValueError(f"no fused optimizer wiring for model_type {ops.kind!r}")
    for w, g in zip(ws, wg):
        assert w.is_contiguous() and w.dtype == torch.float32
        entries.append(dict(param=w.data, grad=g, lr=lrs[id(w)], zero_grad=True, grad_mul=small_mul))   # unpack ACCUMULATES: cleared here
    opt = FusedAdamW(entries, betas=betas, eps=eps, weight_decay=weight_decay, loss_scale=engine.loss_scale, check_finite=check_finite,
                     flags=flags, device=engine.dev)
    opt.weight_grads = wg
    inner = opt.step

    def step(stream=None):
        pre(C.c_void_p((stream or torch.cuda.current_stream(engine.dev)).cuda_stream))
        inner(stream)

    opt.step = step
    opt.post_step.append(post)
    opt.kernels_extra = 2   # unpack + pack
    return opt
