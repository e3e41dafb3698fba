"""GPU parity of the fused "vm" (TensoRF) field against the torch-CPU oracle (F.grid_sample on the same parameters)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _make(res, seed):
    from pvd_b200.fused_vm import VMNeRFField
    torch.manual_seed(seed)
    net = VMNeRFField(resolution0=res, scale=0.4).cuda()
    for m in list(net.color_net) + [net.basis_mat]:
        m.weight.data.mul_(1.5)
    return net


def _oracle_params(net):
    c = lambda p: p.detach().cpu().contiguous().clone().requires_grad_(True)
    return dict(sm=[c(p) for p in net.sigma_mat], sv=[c(p) for p in net.sigma_vec], cm=[c(p) for p in net.color_mat],
                cv=[c(p) for p in net.color_vec], bw=c(net.basis_mat.weight), cw=[c(m.weight) for m in net.color_net])


def _oracle_forward(P, x, d, aabb):
    from oracle import field
    return field.vm_field_forward(x, d, P["sm"], P["sv"], P["cm"], P["cv"], P["bw"], P["cw"], aabb, quantize_fp16=True)


@pytest.mark.parametrize("res", [64, [40, 56, 72]])
def test_vm_forward_backward_matches_oracle(res):
    net = _make(res, 1)
    assert tuple(net.sigma_mat[1].shape) == (1, 16, net.resolution[2], net.resolution[0])  # [1,R,res[mat_id_1],res[mat_id_0]]
    assert net.color_mat[0].is_contiguous(memory_format=torch.channels_last)
    g = torch.Generator(device="cuda").manual_seed(2)
    M = 128 * 9 + 17
    x = torch.rand(M, 3, device="cuda", generator=g) * 2 - 1
    x[0] = torch.tensor([1.0, -1.0, 0.3], device="cuda")   # on the boundary: taps beyond the grid are zero-padded
    d = torch.randn(M, 3, device="cuda", generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    net.train()
    sigma, color = net(x, d)
    feat = net.feature_sigma_color
    P = _oracle_params(net)
    so, co, fo = _oracle_forward(P, x.cpu(), d.cpu(), net.aabb_train.cpu())
    torch.testing.assert_close(feat.detach().cpu(), fo.detach(), rtol=5e-3, atol=5e-3)
    torch.testing.assert_close(color.detach().cpu(), co.detach(), rtol=5e-3, atol=3e-3)
    torch.testing.assert_close(sigma.detach().cpu(), so.detach(), rtol=1e-2, atol=1e-3)
    gs = torch.randn(M, device="cuda", generator=g) * 0.1
    gc = torch.randn(M, 3, device="cuda", generator=g)
    gf = torch.randn(M, 16, device="cuda", generator=g) * 0.05
    (sigma * gs).sum().add((color * gc).sum()).add((feat * gf).sum()).backward()
    (so * gs.cpu()).sum().add((co * gc.cpu()).sum()).add((fo * gf.cpu()).sum()).backward()
    pairs = [("basis", net.basis_mat.weight, P["bw"])]
    pairs += [(f"color_net.{i}", m.weight, w) for i, (m, w) in enumerate(zip(net.color_net, P["cw"]))]
    for k, (plist, olist) in {"sigma_mat": (net.sigma_mat, P["sm"]), "sigma_vec": (net.sigma_vec, P["sv"]),
                              "color_mat": (net.color_mat, P["cm"]), "color_vec": (net.color_vec, P["cv"])}.items():
        pairs += [(f"{k}.{i}", p, o) for i, (p, o) in enumerate(zip(plist, olist))]
    for name, p, o in pairs:
        assert p.grad is not None and p.grad.shape == o.grad.shape, name
        assert _rel_l2(p.grad.cpu(), o.grad) < 2e-2, f"{name}: rel L2 {_rel_l2(p.grad.cpu(), o.grad):.4f}"
        assert _rel(p.grad.cpu(), o.grad) < 1e-1, f"{name}: max {_rel(p.grad.cpu(), o.grad):.4f}"


def test_vm_in_a_training_step(scene):
    """march -> vm field -> composite -> MSE + L1 regulariser (just_train_tea/utils.py:841-846), renderer contract."""
    from oracle import field
    net = _make(48, 3)
    net.density_bitfield.copy_(torch.from_numpy(scene["bitfield"]))
    net.train()
    ro, rd = scene["batches"][0]
    ro, rd = ro[:512].contiguous(), rd[:512].contiguous()
    gt = torch.rand(512, 3, generator=torch.Generator().manual_seed(4))
    out = net.render(ro.cuda().unsqueeze(0), rd.cuda().unsqueeze(0), bg_color=1, perturb=True)
    loss = torch.mean((out["image"][0] - gt.cuda()) ** 2) + net.density_loss() * 1e-4
    (loss * 256.0).backward()
    P = _oracle_params(net)
    fn = lambda x, d: _oracle_forward(P, x, d, net.aabb_train.cpu())[:2]
    o = field.render_train_step(ro, rd, scene["bitfield"], gt, fn)
    reg = sum(torch.mean(torch.abs(P["sm"][i])) + torch.mean(torch.abs(P["sv"][i])) for i in range(3))
    ((o["loss"] + reg * 1e-4) * 256.0).backward()
    assert torch.equal(o["rays"], out["rays"].cpu())
    torch.testing.assert_close(out["image"][0].detach().cpu(), o["image"].detach(), rtol=1e-2, atol=5e-3)
    for i in range(3):
        assert _rel_l2(net.sigma_mat[i].grad.cpu(), P["sm"][i].grad) < 3e-2
        assert _rel_l2(net.color_vec[i].grad.cpu(), P["cv"][i].grad) < 3e-2
    assert _rel_l2(net.basis_mat.weight.grad.cpu(), P["bw"].grad) < 3e-2
