"""GPU parity of the fused NeRF-MLP teacher forward (tcgen05, streamed weights) against the torch-CPU oracle, and the
mlp -> hash distillation step of BASELINE config 5 through the renderer contract."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(seed=0):
    from pvd_b200.fused_mlp import MLPNeRFField
    torch.manual_seed(seed)
    return MLPNeRFField().cuda()


def test_mlp_teacher_forward_matches_oracle():
    from oracle import field
    net = _make(1)
    g = torch.Generator(device="cuda").manual_seed(2)
    M = 128 * 11 + 40
    x = torch.rand(M, 3, device="cuda", generator=g) * 2 - 1
    d = torch.randn(M, 3, device="cuda", generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    net.eval()
    with torch.no_grad():
        sigma, color = net(x, d)
        feat = net.feature_sigma_color
        s_t, c_t, f_t = net._torch_forward(x, d)          # the same layers through cuBLAS in fp32
    assert int(net._status.item()) == 0
    nw = [l.weight.detach().cpu() for l in net.nerf_mlp]
    nb = [l.bias.detach().cpu() for l in net.nerf_mlp]
    tw = [m.weight.detach().cpu() for m in list(net.sigma_net) + list(net.color_net)]
    so, co, fo = field.mlp_field_forward(x.cpu(), d.cpu(), nw, nb, tw, quantize_fp16=True)
    torch.testing.assert_close(feat.cpu(), fo, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(color.cpu(), co, rtol=1e-2, atol=5e-3)
    torch.testing.assert_close(sigma.cpu(), so, rtol=2e-2, atol=2e-3)
    # and within the fp16 tolerance of the pure-fp32 torch composition on the GPU
    torch.testing.assert_close(color, c_t, rtol=2e-2, atol=1e-2)
    torch.testing.assert_close(feat, f_t, rtol=2e-2, atol=2e-2)


def test_mlp_to_hash_distillation_step(scene):
    """Config 5: frozen mlp teacher, hash student, shared samples, norm losses; only the student receives gradients."""
    from pvd_b200.fused import HashNeRFField, _Args
    tea = _make(3)
    torch.manual_seed(4)
    stu = HashNeRFField(num_levels=14, desired_resolution=2048, args=_Args()).cuda()
    for net in (tea, stu):
        net.density_bitfield.copy_(torch.from_numpy(scene["bitfield"]))
        net.train()
    for p in tea.parameters():
        p.requires_grad_(False)   # main_distill_mutual.py:320-321
    ro, rd = scene["batches"][0]
    ro, rd = ro[:512].cuda().unsqueeze(0), rd[:512].cuda().unsqueeze(0)
    o_s = stu.render(ro, rd, bg_color=1, perturb=True)
    with torch.no_grad():
        o_t = tea.render(ro, rd, bg_color=1, perturb=True, inherited_params=o_s["inherited_params"])
    loss = torch.norm(o_t["image"] - o_s["image"]) + 0.002 * torch.norm(stu.feature_sigma_color - tea.feature_sigma_color) \
        + 0.002 * torch.norm(stu.color_l - tea.color_l) + 0.002 * torch.norm(stu.sigma_l - tea.sigma_l)
    loss.backward()
    assert torch.isfinite(loss) and stu.encoder.embeddings.grad.abs().sum() > 0
    assert all(p.grad is None for p in tea.parameters())
    assert tea.feature_sigma_color.shape == stu.feature_sigma_color.shape


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("M", [128 * 5 + 17, 128 * 300])
def test_mlp_training_backward_matches_torch_autograd(M):
    """model_type mlp TRAINED (network.py:324-333 under autograd in the reference): the fused forward + the three backward kernels
    (tail, trunk data gradients, weight gradients on the tensor core) against autograd through the plain fp32 torch composition of
    the same layers, every parameter gradient.  The bound per tensor is north_star's fp16 tolerance (1e-2, relative L2) or 1.25 x
    what torch.autocast(fp16) -- the reference's arithmetic -- deviates from fp32 on the same tensor, whichever is larger."""
    net = _make(5)
    net.train()
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.rand(M, 3, device="cuda", generator=g) * 2 - 1
    d = torch.randn(M, 3, device="cuda", generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    cs = torch.randn(M, device="cuda", generator=g) * 0.05
    cc = torch.randn(M, 3, device="cuda", generator=g)
    cf = torch.randn(M, 16, device="cuda", generator=g) * 0.1

    def run(kind):
        net.zero_grad(set_to_none=True)
        if kind == "fused":
            sigma, color = net(x, d)
            feat = net.feature_sigma_color
        elif kind == "fp32":
            sigma, color, feat = net._torch_forward(x, d)
        else:
            with torch.autocast("cuda", dtype=torch.float16):
                sigma, color, feat = net._torch_forward(x, d)
        loss = (sigma.float() * cs).sum() + (color.float() * cc).sum() + (feat.float() * cf).sum()
        loss.backward()
        return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}, sigma.detach().float(), color.detach().float()

    g_fused, s_f, c_f = run("fused")
    assert int(net._status.item()) == 0 and int(net._bwd_status.item()) == 0, "a tensor-core wait timed out"
    g_ref, s_r, c_r = run("fp32")
    g_amp, _, _ = run("amp")
    assert set(g_fused) == set(g_ref), set(g_ref) ^ set(g_fused)
    torch.testing.assert_close(c_f, c_r, rtol=2e-2, atol=1e-2)
    worst = {}
    for n in g_ref:
        assert g_fused[n].shape == g_ref[n].shape and torch.isfinite(g_fused[n]).all(), n
        oa, af = _rel(g_fused[n], g_ref[n]), _rel(g_amp[n], g_ref[n])
        worst[n] = (oa, af)
        assert oa <= max(1e-2, 1.25 * af), f"{n}: fused vs fp32 {oa:.3e}, autocast vs fp32 {af:.3e}"
    print("max fused-vs-fp32", max(v[0] for v in worst.values()), "max autocast-vs-fp32", max(v[1] for v in worst.values()))


def test_mlp_training_through_the_renderer(scene):
    """One optimisation step of an mlp model through NeRFRenderer.run_cuda (march -> fused field -> composite) reduces nothing to
    torch: the loss back-propagates into every nerf_mlp / sigma_net / color_net parameter through the fused kernels."""
    net = _make(7)
    net.density_bitfield.copy_(torch.from_numpy(scene["bitfield"]))
    net.train()
    ro, rd = scene["batches"][0]
    ro, rd = ro[:1024].cuda().unsqueeze(0), rd[:1024].cuda().unsqueeze(0)
    out = net.render(ro, rd, bg_color=1, perturb=True)
    loss = ((out["image"] - 0.5) ** 2).mean()
    (loss * 4096.0).backward()     # GradScaler's role: the data gradients travel as fp16 operand tiles, as under the reference's autocast
    assert int(net._bwd_status.item()) == 0
    for n, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert net.nerf_mlp[0].weight.grad.abs().sum() > 0 and net.nerf_mlp[7].bias.grad.abs().sum() > 0


def test_mlp_engine_step_matches_module_path_and_fp32_autograd(scene):
    """MLPTrainEngine (no autograd, buffers allocated once, CUDA-graph capturable) == the module path through the drop-in operators
    (same kernels), and both within the fp16 bound of fp32 torch autograd on the engine's own samples; then the captured graph
    reproduces the eager step."""
    import raymarching
    from pvd_b200.engine import MLPTrainEngine
    n_rays, scale = 1024, 16384.0      # GradScaler's role: the data gradients travel as fp16 operand tiles
    net = _make(8)
    net.train()
    eng = MLPTrainEngine(net, torch.from_numpy(scene["bitfield"]), n_rays, loss_scale=scale)
    eng.stage()
    ro, rd = scene["batches"][0]
    ro, rd = ro[:n_rays].contiguous(), rd[:n_rays].contiguous()
    gt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(3))
    eng.rays_o.copy_(ro); eng.rays_d.copy_(rd); eng.gt.copy_(gt)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0
    loss_e = float(eng.loss[0].item())
    g_e = {k: v.clone() for k, v in eng.grads().items()}
    assert set(g_e) == {n for n, _ in net.named_parameters()}

    def reference(kind):
        net.zero_grad(set_to_none=True)
        xyzs, dirs, deltas, rays = eng.xyzs[:eng.M], eng.dirs[:eng.M], eng.deltas[:eng.M], eng.rays
        if kind == "module":
            sigma, color = net(xyzs, dirs)
        elif kind == "amp":
            with torch.autocast("cuda", dtype=torch.float16):
                sigma, color, _ = net._torch_forward(xyzs, dirs)
        else:
            sigma, color, _ = net._torch_forward(xyzs, dirs)
        ws, depth, image = raymarching.composite_rays_train(sigma.float(), color.float(), deltas, rays)
        pred = image + (1 - ws).unsqueeze(-1)
        loss = torch.mean((pred - gt.cuda()) ** 2)
        (loss * scale).backward()
        return float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters()}

    loss_m, g_m = reference("module")
    loss_t, g_t = reference("fp32")
    _, g_a = reference("amp")
    assert abs(loss_e - loss_m) < 1e-5 * max(1.0, loss_m)
    assert abs(loss_e - loss_t) < 1e-2 * max(1e-3, loss_t)
    for n in g_t:
        assert _rel(g_e[n], g_m[n]) < 1e-3, (n, _rel(g_e[n], g_m[n]))          # same kernels; float reductions reorder
        oa, af = _rel(g_e[n], g_t[n]), _rel(g_a[n], g_t[n])                   # ours vs fp32; the reference's autocast vs fp32
        assert oa <= max(1e-2, 1.25 * af), (n, oa, af)
    # CUDA graph of the step
    eng.unpack_each_step = True
    eng.capture()
    eng.replay()
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0
    assert abs(float(eng.loss[0].item()) - loss_e) < 1e-5 * max(1.0, loss_e)
    for n, g in zip(eng.ops.NAMES, eng.ops.wgrads):
        assert _rel(g, g_e[n]) < 1e-3, n
