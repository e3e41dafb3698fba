"""The distillation step engine (pvd_b200.engine.PairDistillEngine) and its loss kernels (csrc/pair_loss.cu) on the GPU, against
the CPU oracle's pair step (oracle/field.py::pair_distill_step = Trainer.train_step of distill_mutual/utils.py:954-1189), and the
vm teacher-training engine (BASELINE configs 3, 4, 5).  Tolerances: fp16 tables / tensor-core MLPs -> 1e-2-class (north_star)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RATES = (1.0, 0.002, 0.002, 0.002)   # main_distill_mutual.py:174-177


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ------------------------------------------------------------------------------------------------ kernels alone
@pytest.mark.parametrize("stage", [1, 2, 3])
def test_pair_kernels_match_autograd(stage):
    from pvd_b200 import _native as nv
    from pvd_b200.engine import PvdPairRates
    from test_pair_loss_cpu import _autograd, _fake_pair
    rays, deltas, stu, tea, total, M = _fake_pair(10 + stage, n_rays=301, max_cnt=150, pad=77)
    rates = (1.0, 0.5, 0.25, 0.125)
    want = _autograd(rays, deltas, stu, tea, rates, stage)
    dev = "cuda"
    g = lambda t: t.detach().to(dev).contiguous()
    ft, fs, ct, cs, st_, ss = g(tea["feat"]), g(stu["feat"]), g(tea["rgb"]), g(stu["rgb"]), g(tea["sigma"]), g(stu["sigma"])
    drays, ddl = g(rays), g(deltas)
    N = rays.shape[0]
    sums = torch.zeros(64 * 4, device=dev)
    gs, gc, gf = torch.full((M,), 7.0, device=dev), torch.full((M, 3), 7.0, device=dev), torch.empty(M, 16, device=dev)
    pred_t, ws, depth, img = torch.empty(N, 3, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
    loss_out = torch.zeros(5, device=dev)
    counter = torch.tensor([total, N], dtype=torch.int32, device=dev)
    bg = torch.ones(3, device=dev)
    r = list(rates)
    if stage == 1:
        r = [0, r[1], 0, 0]
    if stage == 2:
        r[0] = 0
    l, s = nv.lib(), nv.stream_of(ft)
    nv.check(l.pvd_pair_sample_sq(nv.ptr(ft), nv.ptr(fs), nv.ptr(ct), nv.ptr(cs), C.c_uint32(M), nv.ptr(sums), s))
    if stage == 3:
        nv.check(l.pvd_pair_composite(nv.ptr(bg), nv.ptr(st_), nv.ptr(ct), nv.ptr(ss), nv.ptr(cs), nv.ptr(ddl), nv.ptr(drays), C.c_uint32(M),
                                      C.c_uint32(N), nv.ptr(pred_t), nv.ptr(ws), nv.ptr(depth), nv.ptr(img), nv.ptr(gs), nv.ptr(gc), nv.ptr(sums), s))
    nv.check(l.pvd_pair_combine(nv.ptr(ft), nv.ptr(fs), nv.ptr(ct), nv.ptr(cs), nv.ptr(sums), C.byref(PvdPairRates(*r)), C.c_float(4.0),
                                C.c_uint32(M), nv.ptr(counter) if stage == 3 else None, nv.ptr(gs), nv.ptr(gc), nv.ptr(gf), nv.ptr(loss_out), s))
    torch.cuda.synchronize()
    assert abs(float(loss_out[0]) - want[0]) < 1e-4 * max(1.0, want[0])
    np.testing.assert_allclose(gs.cpu().numpy() / 4.0, want[1], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(gc.cpu().numpy() / 4.0, want[2], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(gf.cpu().numpy() / 4.0, want[3], rtol=2e-3, atol=2e-6)
    if stage == 3:
        from oracle import cpu
        ows, od, oi = cpu.composite_rays_train_forward(stu["sigma"].detach().numpy(), stu["rgb"].detach().numpy(), deltas.numpy(), rays.numpy())
        np.testing.assert_allclose(img.cpu().numpy(), oi, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(depth.cpu().numpy(), od, rtol=2e-5, atol=2e-6)
        twt, _, tit = cpu.composite_rays_train_forward(tea["sigma"].numpy(), tea["rgb"].numpy(), deltas.numpy(), rays.numpy())
        np.testing.assert_allclose(pred_t.cpu().numpy(), tit + (1 - twt)[:, None], rtol=2e-5, atol=2e-6)


def test_zero_sample_tail_and_l1_reg():
    from pvd_b200 import _native as nv
    dev = "cuda"
    N, M = 5, 40
    # ray 3 does not fit (offset 30 + 12 >= 40) and is dropped; ray 4 starts past M
    rays = torch.tensor([[0, 0, 10], [1, 10, 0], [2, 10, 20], [3, 30, 12], [4, 42, 5]], dtype=torch.int32, device=dev)
    counter = torch.tensor([47, 5], dtype=torch.int32, device=dev)
    x, d, dl = torch.ones(M, 3, device=dev), torch.ones(M, 3, device=dev), torch.ones(M, 2, device=dev)
    nv.check(nv.lib().pvd_zero_sample_tail(nv.ptr(rays), nv.ptr(counter), C.c_uint32(N), C.c_uint32(M), nv.ptr(x), nv.ptr(d), nv.ptr(dl), nv.stream_of(x)))
    assert float(x[:30].min()) == 1.0 and float(x[30:].abs().max()) == 0.0 and float(d[30:].abs().max()) == 0.0 and float(dl[30:].abs().max()) == 0.0
    counter[0] = 25   # nothing dropped, samples end at row 25 (rays above are inconsistent with it on purpose: tail rule alone)
    rays[3, 2] = 0
    x.fill_(1.0)
    nv.check(nv.lib().pvd_zero_sample_tail(nv.ptr(rays), nv.ptr(counter), C.c_uint32(N), C.c_uint32(M), nv.ptr(x), nv.ptr(d), nv.ptr(dl), nv.stream_of(x)))
    assert float(x[:25].min()) == 1.0 and float(x[25:].abs().max()) == 0.0
    # L1 penalty: value and gradient of w * mean|p|
    p = torch.randn(3001, device=dev)
    p[7] = 0.0
    grad = torch.full_like(p, 0.5)
    slots = torch.zeros(128, device=dev)
    nv.check(nv.lib().pvd_l1_mean_reg(nv.ptr(p), C.c_uint64(p.numel()), C.c_float(1e-2), C.c_float(8.0), nv.ptr(grad), nv.ptr(slots), nv.stream_of(p)))
    assert abs(float(slots.view(64, 2).sum(0)[0]) - 1e-2 * float(p.abs().mean())) < 1e-7
    torch.testing.assert_close(grad, 0.5 + 8.0 * 1e-2 / p.numel() * torch.sign(p), rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------------------------------------ engines
def _hash_net(seed, is_teacher=False, levels=14):
    from pvd_b200.fused import HashNeRFField, _Args
    torch.manual_seed(seed)
    net = HashNeRFField(num_levels=levels, desired_resolution=2048, is_teacher=is_teacher, args=_Args()).cuda()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    return net


def _hash_oracle(net, trainable):
    from oracle import field
    ws = [m.weight.detach().cpu().clone().requires_grad_(trainable) for m in list(net.sigma_net) + list(net.color_net)]
    emb = net.encoder.embeddings.detach().cpu().clone().requires_grad_(trainable)
    e = net.encoder
    fn = lambda x, d: field.hash_field_forward(x, d, emb, e.offsets.cpu().numpy(), float(e.per_level_scale), e.base_resolution, ws,
                                               quantize_fp16=True)
    names = ["encoder.embeddings", "sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight", "color_net.2.weight"]
    return fn, dict(zip(names, [emb] + ws))


def _vm_net(seed, res=48):
    from pvd_b200.fused import _Args
    from pvd_b200.fused_vm import VMNeRFField
    torch.manual_seed(seed)
    net = VMNeRFField(resolution0=res, scale=0.4, args=_Args()).cuda()
    for m in list(net.color_net) + [net.basis_mat]:
        m.weight.data.mul_(1.5)
    return net


def _vm_oracle(net):
    from oracle import field
    c = lambda p_: p_.detach().cpu().contiguous().clone().requires_grad_(True)
    P = dict(sm=[c(p_) for p_ in net.sigma_mat], sv=[c(p_) for p_ in net.sigma_vec], cm=[c(p_) for p_ in net.color_mat],
             cv=[c(p_) for p_ in net.color_vec], bw=c(net.basis_mat.weight), cw=[c(m.weight) for m in net.color_net])
    aabb = net.aabb_train.cpu()
    fn = lambda x, d: field.vm_field_forward(x, d, P["sm"], P["sv"], P["cm"], P["cv"], P["bw"], P["cw"], aabb, quantize_fp16=True)
    named = {}
    for k, name in (("sm", "sigma_mat"), ("sv", "sigma_vec"), ("cm", "color_mat"), ("cv", "color_vec")):
        for i in range(3):
            named[f"{name}.{i}"] = P[k][i]
    named["basis_mat.weight"] = P["bw"]
    for i in range(3):
        named[f"color_net.{i}.weight"] = P["cw"][i]
    return fn, named, P


def _run_engine(eng, ro, rd, gt=None):
    eng.stage()
    for rs in eng.sets:
        rs.rays_o.copy_(ro); rs.rays_d.copy_(rd)
        if gt is not None:
            rs.gt.copy_(gt)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0, "tensor-core pipeline reported a timeout"


def _check_grads(eng, named, tol=3e-2):
    """engine gradients vs the oracle's.  The oracle's loss is multiplied by the SAME loss scale before backward(): its
    quantize_fp16 casts round the gradient to fp16 on the way back, exactly like autocast, so an unscaled oracle underflows."""
    got = eng.grads()
    assert set(got) == set(named), (sorted(got), sorted(named))
    for k, ref in named.items():
        if ref.grad is None:   # a parameter the stage's loss does not reach (stage 1: the colour head)
            assert float(got[k].abs().max()) == 0.0, k
            continue
        e = _rel_l2(got[k], ref.grad)
        assert e < tol, f"{k}: rel-L2 {e}"


def test_pair_engine_hash_to_vm_stage3(scene):
    """BASELINE config 4 (hash teacher -> vm student): loss terms, both images and every student gradient vs the oracle pair step."""
    from oracle import field
    from pvd_b200.engine import PairDistillEngine
    tea, stu = _hash_net(5, True), _vm_net(6)
    ro, rd = scene["batches"][2]
    ro, rd = ro[:640].contiguous(), rd[:640].contiguous()
    eng = PairDistillEngine(tea, stu, torch.from_numpy(scene["bitfield"]), 640, rates=RATES, stage=3, l1_reg_weight=1e-2, loss_scale=128.0)
    _run_engine(eng, ro.cuda(), rd.cuda())
    f_t, _ = _hash_oracle(tea, False)
    f_s, named, P = _vm_oracle(stu)
    o = field.pair_distill_step(ro, rd, scene["bitfield"], f_s, f_t, RATES, stage=3, M=eng.M,
                                l1_reg=1e-2 * field.vm_density_loss(P["sm"], P["sv"]))
    (o["loss"] * 128.0).backward()
    assert torch.equal(eng.rays.cpu(), o["rays"])
    assert torch.equal(eng.xyzs.cpu(), o["xyzs"]), "sample rows (padding included) differ from the oracle"
    terms = eng.loss_terms()
    for k in ("rgb", "fea", "color", "sigma"):
        assert abs(terms[k] - float(o["terms"][k])) < 2e-2 * float(o["terms"][k]), (k, terms[k], float(o["terms"][k]))
    assert abs(float(eng.loss[0]) - float(o["loss"])) < 2e-2 * float(o["loss"])
    pred_s, pred_t = eng.final_images()
    torch.testing.assert_close(pred_s.cpu(), o["image"].detach(), rtol=1e-2, atol=5e-3)
    torch.testing.assert_close(pred_t.cpu(), o["image_tea"], rtol=1e-2, atol=5e-3)
    _check_grads(eng, named)


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_pair_engine_hash_to_hash_stages(scene, stage):
    from oracle import field
    from pvd_b200.engine import PairDistillEngine
    tea, stu = _hash_net(2, True), _hash_net(1)
    ro, rd = scene["batches"][1]
    ro, rd = ro[:512].contiguous(), rd[:512].contiguous()
    eng = PairDistillEngine(tea, stu, torch.from_numpy(scene["bitfield"]), 512, rates=RATES, stage=stage, loss_scale=64.0)
    _run_engine(eng, ro.cuda(), rd.cuda())
    f_t, _ = _hash_oracle(tea, False)
    f_s, named = _hash_oracle(stu, True)
    o = field.pair_distill_step(ro, rd, scene["bitfield"], f_s, f_t, RATES, stage=stage, M=eng.M)
    (o["loss"] * 64.0).backward()
    assert abs(float(eng.loss[0]) - float(o["loss"])) < 2e-2 * float(o["loss"])
    _check_grads(eng, named)


def test_pair_engine_mlp_to_hash(scene):
    """BASELINE config 5 (NeRF-MLP teacher -> hash student)."""
    from oracle import field
    from pvd_b200.engine import PairDistillEngine
    from pvd_b200.fused import _Args
    from pvd_b200.fused_mlp import MLPNeRFField
    torch.manual_seed(11)
    tea = MLPNeRFField(args=_Args()).cuda()
    stu = _hash_net(12)
    ro, rd = scene["batches"][0]
    ro, rd = ro[:384].contiguous(), rd[:384].contiguous()
    eng = PairDistillEngine(tea, stu, torch.from_numpy(scene["bitfield"]), 384, rates=RATES, stage=3, loss_scale=64.0)
    _run_engine(eng, ro.cuda(), rd.cuda())
    nw = [l.weight.detach().cpu() for l in tea.nerf_mlp]
    nb = [l.bias.detach().cpu() for l in tea.nerf_mlp]
    tw = [m.weight.detach().cpu() for m in list(tea.sigma_net) + list(tea.color_net)]
    f_t = lambda x, d: field.mlp_field_forward(x, d, nw, nb, tw, quantize_fp16=True)
    f_s, named = _hash_oracle(stu, True)
    o = field.pair_distill_step(ro, rd, scene["bitfield"], f_s, f_t, RATES, stage=3, M=eng.M)
    (o["loss"] * 64.0).backward()
    assert abs(float(eng.loss[0]) - float(o["loss"])) < 2e-2 * float(o["loss"])
    _check_grads(eng, named)


def test_vm_train_engine(scene):
    """BASELINE config 3 (vm teacher training): MSE + l1_reg_weight * density_loss, every gradient vs the oracle step."""
    from oracle import field
    from pvd_b200.engine import VMTrainEngine
    net = _vm_net(21)
    ro, rd = scene["batches"][0]
    ro, rd = ro[:640].contiguous(), rd[:640].contiguous()
    gt = torch.rand(640, 3, generator=torch.Generator().manual_seed(3))
    LS = 65536.0   # GradScaler's initial scale (the reference trains under torch.cuda.amp.GradScaler)
    eng = VMTrainEngine(net, torch.from_numpy(scene["bitfield"]), 640, loss_scale=LS, l1_reg_weight=1e-2)
    _run_engine(eng, ro.cuda(), rd.cuda(), gt.cuda())
    f_s, named, P = _vm_oracle(net)
    o = field.render_train_step(ro, rd, scene["bitfield"], gt, lambda x, d: f_s(x, d)[:2], M=eng.M, aabb=None)
    loss = o["loss"] + 1e-2 * field.vm_density_loss(P["sm"], P["sv"])
    (loss * LS).backward()
    assert torch.equal(eng.rays.cpu(), o["rays"])
    assert abs(float(eng.loss[0]) - float(loss)) < 2e-2 * float(loss)
    pred, _ = eng.final_image()
    torch.testing.assert_close(pred.cpu(), o["image"].detach(), rtol=1e-2, atol=5e-3)
    _check_grads(eng, named)


def test_pair_engine_graph_and_pipeline_equal_eager(scene):
    """One captured graph and the two-graph pipelined schedule give the gradients of the eager step (same batch)."""
    from pvd_b200.engine import PairDistillEngine
    tea, stu = _hash_net(5, True), _vm_net(6)
    ro, rd = scene["batches"][1]
    ro, rd = ro[:1024].cuda().contiguous(), rd[:1024].cuda().contiguous()
    eng = PairDistillEngine(tea, stu, torch.from_numpy(scene["bitfield"]), 1024, rates=RATES, stage=3, loss_scale=128.0)
    _run_engine(eng, ro, rd)
    want = {k: v.clone() for k, v in eng.grads().items()}
    loss = float(eng.loss[0])
    eng.capture()
    eng.replay()
    torch.cuda.synchronize()
    assert abs(float(eng.loss[0]) - loss) < 1e-5 * loss
    for k, v in eng.grads().items():
        assert _rel_l2(v, want[k]) < 1e-3, k
    eng.capture_pipelined()
    eng.march(0)
    for i in range(3):
        eng.replay_pipelined(i)
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0
    assert abs(float(eng.loss[0]) - loss) < 1e-5 * loss
    for k, v in eng.grads().items():
        assert _rel_l2(v, want[k]) < 1e-3, k


def test_pair_engine_hash_to_mlp(scene):
    """hash teacher -> NeRF-MLP STUDENT (INGP -> NeRF, one of the paper's conversions): the student's backward is the fused mlp
    backward (csrc/field_mlp_bwd.cu) driven by the pair losses, including d(loss)/d(feature_sigma_color).  Checked against the
    oracle's pair step with the same loss scale; the bound per tensor is 3e-2 or 1.25 x the deviation of the oracle's OWN fp16
    model from its fp32 evaluation (the deep trunk amplifies fp16 rounding more than the two-layer heads do)."""
    from oracle import field
    from pvd_b200.engine import PairDistillEngine
    from pvd_b200.fused import _Args
    from pvd_b200.fused_mlp import MLPNeRFField
    tea = _hash_net(21, True)
    torch.manual_seed(22)
    stu = MLPNeRFField(args=_Args(), is_teacher=False).cuda()
    ro, rd = scene["batches"][1]
    ro, rd = ro[:384].contiguous(), rd[:384].contiguous()
    scale = 4096.0
    eng = PairDistillEngine(tea, stu, torch.from_numpy(scene["bitfield"]), 384, rates=RATES, stage=3, loss_scale=scale)
    _run_engine(eng, ro.cuda(), rd.cuda())
    f_t, _ = _hash_oracle(tea, False)

    def oracle_grads(q):
        nw = [l.weight.detach().cpu().clone().requires_grad_(True) for l in stu.nerf_mlp]
        nb = [l.bias.detach().cpu().clone().requires_grad_(True) for l in stu.nerf_mlp]
        tw = [m.weight.detach().cpu().clone().requires_grad_(True) for m in list(stu.sigma_net) + list(stu.color_net)]
        f_s = lambda x, d: field.mlp_field_forward(x, d, nw, nb, tw, quantize_fp16=q)
        o = field.pair_distill_step(ro, rd, scene["bitfield"], f_s, f_t, RATES, stage=3, M=eng.M)
        (o["loss"] * scale).backward()
        named = {}
        for i in range(8):
            named[f"nerf_mlp.{i}.weight"], named[f"nerf_mlp.{i}.bias"] = nw[i].grad, nb[i].grad
        for n, w in zip(("sigma_net.0", "sigma_net.1", "color_net.0", "color_net.1", "color_net.2"), tw):
            named[n + ".weight"] = w.grad
        return o, named

    o16, g16 = oracle_grads(True)
    _, g32 = oracle_grads(False)
    assert abs(float(eng.loss[0]) - float(o16["loss"])) < 2e-2 * float(o16["loss"])
    got = eng.grads()
    assert set(got) == set(g32)
    for k in g32:
        ours, noise = _rel_l2(got[k], g32[k]), _rel_l2(g16[k], g32[k])
        assert ours < max(3e-2, 1.25 * noise), f"{k}: ours vs fp32 oracle {ours:.3e}, fp16 oracle vs fp32 oracle {noise:.3e}"
