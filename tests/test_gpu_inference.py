"""SURVEY 8f-2 on the GPU: the persistent inference kernel (march + hash field + composite in ONE launch, rays pulled from a device
queue; csrc/field_hash.cu::k_hash_render_persistent) against the host loop of the reference's evaluation branch
(distill_mutual/renderer.py:450-543 = march_rays -> forward -> composite_rays -> compact_rays, rebuilt from the drop-in kernels).
With perturb off a ray's result does not depend on how its samples are chunked, and both paths run the same arithmetic:
the per-ray outputs must be IDENTICAL."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(scene, seed=0, fp16=True):
    from pvd_b200.fused import HashNeRFField
    torch.manual_seed(seed)
    net = HashNeRFField(num_levels=14, desired_resolution=2048, table_fp16=fp16, density_scale=1).cuda().eval()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    net.density_bitfield.copy_(torch.from_numpy(scene["bitfield"]))
    return net


def _render(net, ro, rd, persistent, **kw):
    old = os.environ.get("PVD_PERSISTENT_INFER")
    os.environ["PVD_PERSISTENT_INFER"] = "1" if persistent else "0"
    try:
        with torch.no_grad():
            out = net.render(ro, rd, bg_color=1, perturb=False, **kw)
        torch.cuda.synchronize()
        return out
    finally:
        if old is None:
            os.environ.pop("PVD_PERSISTENT_INFER", None)
        else:
            os.environ["PVD_PERSISTENT_INFER"] = old


@pytest.mark.parametrize("fp16", [True, False])
def test_persistent_inference_equals_the_host_loop(scene, fp16):
    net = _net(scene, fp16=fp16)
    ros, rds = zip(*scene["batches"])
    ro, rd = torch.cat(ros)[:6000].cuda().unsqueeze(0), torch.cat(rds)[:6000].cuda().unsqueeze(0)   # not a multiple of 16 rays
    loop = _render(net, ro, rd, persistent=False)
    pers = _render(net, ro, rd, persistent=True)
    assert int(net._render_status.item()) == 0, "tensor-core pipeline reported a timeout"
    assert set(pers) == {"depth", "image", "inherited_params"}
    hit = (loop["image"] != 1).any(-1)
    assert 0.05 < float(hit.float().mean()) < 0.95, "the batch should mix rays that hit the scene and rays that miss it"
    torch.testing.assert_close(pers["image"], loop["image"], rtol=0, atol=0)
    # (rays that miss the scene box have near == far: the reference's depth normalisation (d - near) / (far - near) is 0 / 0 for them)
    torch.testing.assert_close(pers["depth"], loop["depth"], rtol=0, atol=0, equal_nan=True)


def test_persistent_inference_dt_gamma_and_short_budget(scene):
    """cone-like step growth (dt_gamma > 0) and a max_steps budget that cuts rays short"""
    net = _net(scene, seed=1)
    ro, rd = scene["batches"][1]
    ro, rd = ro[:2048].cuda(), rd[:2048].cuda()
    a = _render(net, ro, rd, persistent=False, dt_gamma=1.0 / 128, max_steps=512)
    b = _render(net, ro, rd, persistent=True, dt_gamma=1.0 / 128, max_steps=512)
    torch.testing.assert_close(b["image"], a["image"], rtol=0, atol=0)
    torch.testing.assert_close(b["depth"], a["depth"], rtol=0, atol=0, equal_nan=True)
