"""SURVEY 8f-4 on the GPU: the "tensors" (Plenoxels-style) field -- module path and training engine -- against the torch-CPU oracle
(F.grid_sample 3-D, oracle/field.py::tensors_field_forward, itself pinned to the reference's network.py by the golden vectors), and the
device ray generator against the reference's get_rays formula.  fp32 end to end: tolerance 1e-4 relative (north_star, fp32)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _net(res=(20, 24, 28), seed=0, scale=0.8):
    from pvd_b200.fused_tensors import TensorsNeRFField
    torch.manual_seed(seed)
    return TensorsNeRFField(plenoxel_degree=3, plenoxel_res=res, scale=scale).cuda()


def test_tensors_field_forward_backward_vs_oracle():
    from oracle import field
    net = _net()
    g = torch.Generator().manual_seed(1)
    n = 3001                                                   # not a multiple of the 4 samples a warp holds
    x = torch.rand(n, 3, generator=g) * 2.2 - 1.1              # some samples outside the volume: zero padding, partial corners
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    cs, cc = torch.randn(n, generator=g) * 0.1, torch.randn(n, 3, generator=g)
    sigma, color = net(x.cuda(), d.cuda())
    ((sigma * cs.cuda()).sum() + (color * cc.cuda()).sum()).backward()
    vol = net.tensor_volume[0].detach().cpu().contiguous().clone().requires_grad_(True)
    so, co = field.tensors_field_forward(x, d, vol, 3, torch.tensor([-1.0, -1, -1, 1, 1, 1]))
    ((so * cs).sum() + (co * cc).sum()).backward()
    assert _rel(sigma, so) < 1e-5 and _rel(color, co) < 1e-5
    assert _rel(net.tensor_volume[0].grad, vol.grad) < 1e-4
    # density(): the reference's unclamped trunc_exp(h0) (network.py:461-476)
    vol2 = net.tensor_volume[0].detach().cpu().contiguous()
    so2, _ = field.tensors_field_forward(x, d, vol2, 3, torch.tensor([-1.0, -1, -1, 1, 1, 1]), clip_min=-1e30, clip_max=1e30)
    assert _rel(net.density(x.cuda())["sigma"], so2) < 1e-5


def test_tensors_engine_step_vs_oracle(scene):
    """FieldTrainEngine over the tensors field: march -> field -> composite -> MSE -> backward, against the oracle's CPU step."""
    from oracle import field
    from pvd_b200.engine import TensorsTrainEngine
    net = _net(res=(32, 32, 32), seed=2, scale=1.5)
    n = 512
    ro, rd = scene["batches"][1]
    ro, rd = ro[:n].contiguous(), rd[:n].contiguous()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(4))
    eng = TensorsTrainEngine(net, torch.from_numpy(scene["bitfield"]), n, loss_scale=64.0)
    eng.stage()
    for rs in eng.sets:
        rs.rays_o.copy_(ro); rs.rays_d.copy_(rd); rs.gt.copy_(gt)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    vol = net.tensor_volume[0].detach().cpu().contiguous().clone().requires_grad_(True)
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1])
    o = field.render_train_step(ro, rd, scene["bitfield"], gt, lambda x, d: field.tensors_field_forward(x, d, vol, 3, aabb), M=eng.M)
    (o["loss"] * 64.0).backward()
    assert torch.equal(eng.rays.cpu(), o["rays"])
    assert abs(float(eng.loss[0]) - float(o["loss"])) < 1e-5 * float(o["loss"])
    pred, _ = eng.final_image()
    torch.testing.assert_close(pred.cpu(), o["image"].detach(), rtol=1e-4, atol=1e-5)
    assert _rel(eng.grads()["tensor_volume.0"], vol.grad) < 1e-4
    # graph replay gives the same step
    g0 = eng.grads()["tensor_volume.0"].clone()
    eng.capture()
    eng.replay()
    torch.cuda.synchronize()
    assert _rel(eng.grads()["tensor_volume.0"], g0) < 1e-5


def test_get_rays_vs_oracle():
    from oracle import field
    from pvd_b200.rays import get_rays
    g = torch.Generator().manual_seed(3)
    B, H, W = 3, 800, 800
    poses = torch.eye(4).repeat(B, 1, 1)
    poses[:, :3, :3] = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    poses[:, :3, 3] = torch.randn(B, 3, generator=g)
    intr = np.array([1111.11, 1111.11, 400.0, 400.0])
    torch.manual_seed(7)
    out = get_rays(poses.cuda(), intr, H, W, N=4096)
    inds = out["inds"].cpu()
    assert inds.shape == (B, 4096) and out["rays_o"].shape == (B, 4096, 3)
    ro, rd = field.get_rays(poses, tuple(float(v) for v in intr), H, W, inds)
    torch.testing.assert_close(out["rays_d"].cpu(), rd, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(out["rays_o"].cpu(), ro.contiguous(), rtol=0, atol=0)
    # the same seed draws the same pixels as the reference's torch.randint call
    torch.manual_seed(7)
    assert torch.equal(inds[0], torch.randint(0, H * W, size=[4096], device="cuda").cpu())
    # full frame
    small = get_rays(poses[:1].cuda(), intr, 40, 50, N=-1)
    ro2, rd2 = field.get_rays(poses[:1], tuple(float(v) for v in intr), 40, 50, None)
    torch.testing.assert_close(small["rays_d"].cpu(), rd2, rtol=1e-6, atol=1e-6)
    # error-map sampling keeps the reference's result keys and ranges
    em = torch.rand(B, 128 * 128, generator=g)
    o3 = get_rays(poses.cuda(), intr, H, W, N=1024, error_map=em)
    assert set(o3) == {"inds_coarse", "inds", "rays_o", "rays_d"} and int(o3["inds"].max()) < H * W
    ro3, rd3 = field.get_rays(poses, tuple(float(v) for v in intr), H, W, o3["inds"].cpu())
    torch.testing.assert_close(o3["rays_d"].cpu(), rd3, rtol=1e-6, atol=1e-6)
