"""BASELINE configs[0]: mlp model, 256 rays x 64 uniform samples, the pure-PyTorch renderer path (`NeRFRenderer.run`,
distill_mutual/renderer.py:139-317) on synthetic 'chair'-shaped rays.  CPU: the oracle's restatement of `run` against the C oracle's
compositing kernel (two independent formulations of the same quadrature) and its gradients; GPU: the same configuration through the
CUDA operators (fused NeRF-MLP field + composite_rays_train) against the restatement."""
import numpy as np
import pytest
import torch

from oracle import cpu, field

N, T = 256, 64
AABB = [-1.0, -1, -1, 1, 1, 1]


def _rays(seed=0):
    """256 rays from a random pose on a sphere of radius 2 looking at a chair-sized object around the origin."""
    g = torch.Generator().manual_seed(seed)
    theta, phi = torch.rand(1, generator=g) * 6.283, 0.3 + torch.rand(1, generator=g) * 0.8
    eye = 2.0 * torch.tensor([float(torch.sin(phi) * torch.cos(theta)), float(torch.cos(phi)), float(torch.sin(phi) * torch.sin(theta))])
    target = (torch.rand(N, 3, generator=g) - 0.5) * 0.8
    d = target - eye
    return eye.expand(N, 3).contiguous(), (d / d.norm(dim=-1, keepdim=True)).contiguous()


def _mlp(seed=1):
    torch.manual_seed(seed)
    dims = [(63, 256)] + [(256, 256)] * 3 + [(319, 256)] + [(256, 256)] * 2 + [(256, 28)]
    ls = [torch.nn.Linear(i, o) for i, o in dims]
    nw, nb = [l.weight.detach().requires_grad_(True) for l in ls], [l.bias.detach().requires_grad_(True) for l in ls]
    tail = [torch.nn.Linear(i, o, bias=False).weight.detach().requires_grad_(True) for i, o in ((28, 64), (64, 16), (31, 64), (64, 64), (64, 3))]
    return nw, nb, tail


def test_config0_uniform_renderer_cpu():
    ro, rd = _rays()
    nw, nb, tail = _mlp()
    fn = lambda x, d: field.mlp_field_forward(x, d, nw, nb, tail)[:2]
    out = field.render_uniform(ro, rd, fn, AABB, num_steps=T)
    assert out["image"].shape == (N, 3) and out["xyzs"].shape == (N * T, 3)
    assert float(out["weights_sum"].min()) >= 0 and float(out["weights_sum"].max()) <= 1 + 1e-5
    assert float(out["xyzs"].abs().max()) <= 1.0 and bool(torch.isfinite(out["image"]).all())
    # the same quadrature through the C oracle's composite_rays_train kernel (raymarching.cu:505-582): per-ray segments of T samples
    rays = np.stack([np.arange(N), np.arange(N) * T, np.full(N, T)], 1).astype(np.int32)
    deltas = torch.stack([out["deltas"], out["deltas"]], -1).reshape(-1, 2).detach().numpy()
    ws, _, img = cpu.composite_rays_train_forward(out["sigma"].reshape(-1).detach().numpy(), out["rgb"].reshape(-1, 3).detach().numpy(),
                                                  np.ascontiguousarray(deltas), np.concatenate([rays, np.zeros((1, 3), np.int32)])[:N])
    # M must exceed the last ray's end (offset + count >= M drops it, raymarching.cu:525): the oracle wrapper takes M = rows
    np.testing.assert_allclose(ws[:-1], out["weights_sum"].detach().numpy()[:-1], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(img[:-1] + (1 - ws[:-1])[:, None], out["image"].detach().numpy()[:-1], rtol=2e-5, atol=2e-6)
    # trainable end to end: MSE against random colours reaches every parameter of the NeRF MLP
    loss = ((out["image"] - torch.rand(N, 3, generator=torch.Generator().manual_seed(3))) ** 2).mean()
    loss.backward()
    assert all(w.grad is not None and bool(torch.isfinite(w.grad).all()) for w in nw + nb + tail)
    assert float(nw[0].grad.abs().sum()) > 0 and float(tail[4].grad.abs().sum()) > 0
    # jitter moves the samples by at most half a step
    noise = torch.rand(N, T, generator=torch.Generator().manual_seed(4))
    out2 = field.render_uniform(ro, rd, fn, AABB, num_steps=T, noise=noise)
    step = (out["z_vals"][:, 1] - out["z_vals"][:, 0]).unsqueeze(-1)
    assert float(((out2["z_vals"] - out["z_vals"]).abs() - 0.5 * step * (T - 1) / T).max()) <= 1e-6


@pytest.mark.gpu
def test_config0_uniform_renderer_through_cuda_ops():
    import raymarching
    from pvd_b200.fused import _Args
    from pvd_b200.fused_mlp import MLPNeRFField
    ro, rd = _rays()
    torch.manual_seed(1)
    net = MLPNeRFField(args=_Args()).cuda()
    nw = [l.weight.detach().cpu() for l in net.nerf_mlp]
    nb = [l.bias.detach().cpu() for l in net.nerf_mlp]
    tail = [m.weight.detach().cpu() for m in list(net.sigma_net) + list(net.color_net)]
    fn = lambda x, d: field.mlp_field_forward(x, d, nw, nb, tail, quantize_fp16=True)[:2]
    o = field.render_uniform(ro, rd, fn, AABB, num_steps=T)
    with torch.no_grad():
        sigma, rgb = net(o["xyzs"].cuda(), o["dirs"].cuda().contiguous())     # fused tcgen05 forward
    M = N * T + 128
    pad = lambda t, w: torch.cat([t, torch.zeros((M - N * T,) + t.shape[1:], device="cuda")])
    deltas = torch.stack([o["deltas"], o["deltas"]], -1).reshape(-1, 2).cuda()
    rays = torch.stack([torch.arange(N), torch.arange(N) * T, torch.full((N,), T)], 1).to(torch.int32).cuda()
    ws, depth, image = raymarching.composite_rays_train(pad(sigma.float(), 1), pad(rgb.float(), 3), pad(deltas, 2), rays)
    image = image + (1 - ws).unsqueeze(-1)
    torch.testing.assert_close(ws.cpu(), o["weights_sum"].detach(), rtol=2e-2, atol=5e-3)
    torch.testing.assert_close(image.cpu(), o["image"].detach(), rtol=2e-2, atol=5e-3)
