"""The fused optimizer step (csrc/optim.cu, pvd_b200/optim.py) on the GPU.

* bit-exactness of the fp32 masters / moments against torch.optim.AdamW(betas=(0.9, 0.99), eps=1e-15) -- the reference's optimizer
  (main_distill_mutual.py:327-339) -- in its single-tensor CUDA path, after k steps on the same gradients, including the
  GradScaler semantics (unscale by 1/loss_scale, skip the step on a non-finite gradient);
* the fp16 shadow, gradient zeroing and weight-tile re-pack an engine relies on once the optimizer is attached;
* a captured training ITERATION (step + optimizer in one graph) against the same iteration driven by torch's optimizer.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BETAS, EPS, WD = (0.9, 0.99), 1e-15, 0.01


def _mem(t):
    """elements in MEMORY order (a channels-last 4-D tensor is [N][H][W][C] in memory)"""
    return (t.permute(0, 2, 3, 1) if t.dim() == 4 else t).reshape(-1)


def _run_pair(flags, steps=12, loss_scale=65536.0, inf_at=None, seed=0, trace=None):
    """(ours, torch) parameter / moment tensors after `steps` AdamW steps on identical gradient sequences."""
    from pvd_b200.optim import FusedAdamW
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(100003,), (64, 28), (4096, 4), (1, 16, 12, 12)]
    lrs = [1e-2, 1e-2, 1e-3, 2e-2]
    ps = [torch.randn(s, device="cuda", generator=g) * 0.3 for s in shapes]
    ps[3] = ps[3].contiguous(memory_format=torch.channels_last)
    ours = [p.clone(memory_format=torch.preserve_format) for p in ps]
    ref = [torch.nn.Parameter(p.clone(memory_format=torch.preserve_format)) for p in ps]
    grads = [torch.zeros_like(p, memory_format=torch.preserve_format) for p in ours]
    shadow = torch.zeros(100003, dtype=torch.float16, device="cuda")
    entries = [dict(param=p, grad=gr, lr=lr, zero_grad=True) for p, gr, lr in zip(ours, grads, lrs)]
    entries[0]["shadow"] = shadow
    opt = FusedAdamW(entries, betas=BETAS, eps=EPS, weight_decay=WD, loss_scale=loss_scale, check_finite=True, flags=flags)
    topt = torch.optim.AdamW([{"params": [r], "lr": lr} for r, lr in zip(ref, lrs)], betas=BETAS, eps=EPS, weight_decay=WD, foreach=False, fused=False)
    for s in range(steps):
        raw = [torch.randn(p.shape, device="cuda", generator=g).contiguous(memory_format=torch.preserve_format) * (10.0 ** float(-(s % 4))) for p in ps]
        raw[3] = raw[3].contiguous(memory_format=torch.channels_last)
        raw[1][0, 0] = 0.0   # an exactly-zero gradient entry
        bad = inf_at is not None and s == inf_at
        for gr, r in zip(grads, raw):
            gr.copy_(r * loss_scale)      # what a loss-scaled backward leaves in the accumulator (exact: power of two)
        if bad:
            grads[2][7, 1] = float("inf")
        opt.step()
        torch.cuda.synchronize()
        for gr in grads:
            assert float(gr.abs().max()) == 0.0, "gradient accumulators must be zero after the step"
        if not bad:                        # GradScaler.step skips the optimizer when found_inf
            for r, rg in zip(ref, raw):
                r.grad = rg.clone(memory_format=torch.preserve_format)
            topt.step()
        if trace is not None:              # fraction of bit-identical elements after this step: (param, exp_avg, exp_avg_sq) per tensor
            row = []
            for i, (p, r) in enumerate(zip(ours, ref)):
                stt = topt.state[r]
                row.append((round(float((p == r.detach()).float().mean()), 6), round(float((opt.exp_avg[i] == _mem(stt["exp_avg"])).float().mean()), 6),
                            round(float((opt.exp_avg_sq[i] == _mem(stt["exp_avg_sq"])).float().mean()), 6)))
            trace.append(row)
    st = opt.read_state()
    return opt, ours, ref, topt, shadow, st


def test_fused_adamw_is_bitwise_torch_adamw():
    from pvd_b200.optim import ADDCMUL_LEFT
    verdict = {}
    traces = {}
    for flags in (0, ADDCMUL_LEFT):
        traces[flags] = []
        opt, ours, ref, topt, shadow, st = _run_pair(flags, trace=traces[flags])
        assert st.step == 12 and st.skipped == 0 and st.found_inf == 0
        same = True
        mem = _mem
        for i, (p, r) in enumerate(zip(ours, ref)):
            state = topt.state[r]
            same &= torch.equal(p, r.detach())
            same &= torch.equal(opt.exp_avg[i], mem(state["exp_avg"])) and torch.equal(opt.exp_avg_sq[i], mem(state["exp_avg_sq"]))
        verdict[flags] = bool(same)
        if flags == 0:
            assert torch.equal(shadow, ours[0].to(torch.float16)), "fp16 shadow != half(master)"
            worst = max(float((p - r.detach()).abs().max()) for p, r in zip(ours, ref))
    assert verdict[0], (f"default arithmetic is not bit-identical to torch.optim.AdamW (single-tensor): {verdict}, max |diff| {worst:.3g}; "
                        f"per step [(param, m, v) equal fractions per tensor], default flags: {traces[0][:4]} ... {traces[0][-1]}")


def test_fused_adamw_skips_nonfinite_steps_like_gradscaler():
    opt, ours, ref, topt, shadow, st = _run_pair(0, steps=6, inf_at=3)
    assert st.step == 5 and st.skipped == 1 and st.found_inf == 0
    for p, r in zip(ours, ref):
        assert torch.equal(p, r.detach())


def _engine(scene, seed, n_rays=1024):
    from pvd_b200.engine import HashTrainEngine
    from pvd_b200.fused import HashNeRFField
    torch.manual_seed(seed)
    net = HashNeRFField(num_levels=14, desired_resolution=2048).cuda()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    eng = HashTrainEngine(net, torch.from_numpy(scene["bitfield"]), n_rays, loss_scale=1024.0)
    eng.stage()
    return net, eng


def test_engine_iteration_with_fused_optimizer_matches_torch_loop(scene):
    """3 iterations of step + fused AdamW (captured in ONE graph) vs the same engine step followed by torch.optim.AdamW + stage()."""
    from pvd_b200 import optim as pvd_optim
    n = 1024
    batches = []
    for b in range(3):
        ro, rd = scene["batches"][b]
        gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(30 + b))
        batches.append((ro[:n].contiguous().cuda(), rd[:n].contiguous().cuda(), gt.cuda()))
    net_a, eng_a = _engine(scene, 4)
    net_b, eng_b = _engine(scene, 4)
    for eng in (eng_a, eng_b):
        eng.rays_o.copy_(batches[0][0]); eng.rays_d.copy_(batches[0][1]); eng.gt.copy_(batches[0][2])
        eng.step(warmup=True)
        eng.finish_warmup()
    p0 = {k: v.detach().clone() for k, v in net_a.named_parameters()}
    # A: fused optimizer inside the captured step
    opt = pvd_optim.for_engine(eng_a, lr=1e-2, check_finite=True)
    eng_a.attach_optimizer(opt)
    eng_a.rays_o.copy_(batches[0][0]); eng_a.rays_d.copy_(batches[0][1]); eng_a.gt.copy_(batches[0][2])
    # (capture() runs the step once eagerly and once under capture without replaying: undo those two optimizer steps afterwards)
    snap = {k: v.detach().clone() for k, v in net_a.named_parameters()}
    eng_a.capture()
    with torch.no_grad():
        for k, v in net_a.named_parameters():
            v.copy_(snap[k])
    for m, v in zip(opt.exp_avg, opt.exp_avg_sq):
        m.zero_(); v.zero_()
    st = opt.read_state(); st.step = 0; opt._upload(opt.state, st)
    eng_a.stage()
    eng_a.ops.clear_grads()
    losses_a = []
    for ro, rd, gt in batches:
        eng_a.rays_o.copy_(ro); eng_a.rays_d.copy_(rd); eng_a.gt.copy_(gt)
        eng_a.replay()
        torch.cuda.synchronize()
        losses_a.append(float(eng_a.loss[0]))
    # B: the same step, torch's optimizer on the unscaled gradients, explicit re-stage
    params_b = dict(net_b.named_parameters())
    topt = torch.optim.AdamW([p for p in net_b.parameters() if p.requires_grad], lr=1e-2, betas=BETAS, eps=EPS, weight_decay=WD, foreach=False)
    losses_b = []
    for ro, rd, gt in batches:
        eng_b.rays_o.copy_(ro); eng_b.rays_d.copy_(rd); eng_b.gt.copy_(gt)
        eng_b.stage()
        eng_b.step()
        torch.cuda.synchronize()
        losses_b.append(float(eng_b.loss[0]))
        for k, g in eng_b.grads().items():
            params_b[k].grad = (g / eng_b.loss_scale).to(params_b[k].dtype).view_as(params_b[k]).clone()
        topt.step()
    assert int(eng_a.status.item()) == 0 and opt.read_state().step == 3
    np.testing.assert_allclose(losses_a, losses_b, rtol=2e-3)
    for k, pa in net_a.named_parameters():
        pb = params_b[k].detach()
        da, db = (pa.detach() - p0[k]).double(), (pb - p0[k]).double()
        if float(db.norm()) == 0:
            assert float(da.norm()) == 0, k
            continue
        # float atomics order the gradient sums differently in the two engines; Adam's first steps move every touched entry by
        # ~lr * sign(g), so entries whose gradient cancels to ~0 may flip: bound the relative L2 distance of the UPDATES
        assert float((da - db).norm() / db.norm()) < 2e-2, k
    # what the field kernels read is current: shadow == half(master), weight tiles == a fresh pack
    assert torch.equal(eng_a.ops.table, net_a.encoder.embeddings.detach().to(torch.float16))
    blob = eng_a.ops.wblob.clone()
    eng_a.stage()
    torch.cuda.synchronize()
    assert torch.equal(blob, eng_a.ops.wblob)
    assert float(eng_a.grad_table.abs().max()) == 0.0


def test_vm_engine_iteration_with_fused_optimizer(scene):
    """The vm wiring: 12 channels-last plane / line tensors + 4 weight matrices with the reference's two learning rates."""
    from pvd_b200 import optim as pvd_optim
    from pvd_b200.engine import VMTrainEngine
    from pvd_b200.fused_vm import VMNeRFField
    torch.manual_seed(3)
    net = VMNeRFField(resolution0=48, scale=0.4).cuda()
    n = 512
    eng = VMTrainEngine(net, torch.from_numpy(scene["bitfield"]), n, loss_scale=256.0, l1_reg_weight=1e-4)
    eng.stage()
    ro, rd = scene["batches"][0]
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5)).cuda()
    eng.rays_o.copy_(ro[:n]); eng.rays_d.copy_(rd[:n]); eng.gt.copy_(gt)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    g_ref = {k: (v / eng.loss_scale).clone() for k, v in eng.grads().items()}
    p0 = {k: v.detach().clone() for k, v in net.named_parameters()}
    lr, lr2 = 2e-2, 1e-3
    opt = pvd_optim.for_engine(eng, lr=lr, lr2=lr2)
    eng.attach_optimizer(opt)
    eng.step()
    torch.cuda.synchronize()
    assert opt.read_state().step == 1 and int(eng.status.item()) == 0
    lrs = pvd_optim._lr_of(net, lr, lr2)
    for k, p in net.named_parameters():
        g = g_ref[k].view_as(p) if k in g_ref else None
        assert g is not None, k
        # first AdamW step from zero moments: p1 = p0 (1 - lr wd) - lr * g / (|g| + eps)
        want = p0[k] * (1 - lrs[id(p)] * WD) - lrs[id(p)] * torch.sign(g)
        touched = g.abs() > 1e-9
        if not bool(touched.any()):
            continue
        # (a gradient that cancels to ~0 may change sign between the two backward passes -- float atomics -- and then moves the other
        # way: allow a vanishing fraction of such entries)
        wrong = ((p.detach() - want)[touched].abs() > 1e-3 * lrs[id(p)] + 1e-7).float().mean()
        assert float(wrong) < 1e-3, (k, float(wrong))
    assert float(eng.ops._flat.abs().max()) == 0.0
