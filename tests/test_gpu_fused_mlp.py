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
