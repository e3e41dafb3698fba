"""The fused CUDA fields against golden vectors produced by THE REFERENCE's `distill_mutual/network.py::NeRFNetwork.forward`
(tests/golden/ref_network_golden.npz, generated in the build container by tests/golden/make_network_golden.py: fp32, CPU).

The fused fields compute in the reference's autocast precision (fp16 table / fp16 tensor-core MLP, fp32 accumulation), the golden
is the fp32 result, so the bound is north_star's fp16 one: relative L2 <= 1e-2 per tensor (recorded in
gpurun_out/network_golden.json).
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import netgold  # noqa: E402

TOL = 1e-2
_report = {}


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(HERE, "golden", "ref_network_golden.npz")))


def _rel(a, b):
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _net(mt):
    from pvd_b200.fused import HashNeRFField
    from pvd_b200.fused_mlp import MLPNeRFField
    from pvd_b200.fused_vm import VMNeRFField
    if mt == "hash":
        net = HashNeRFField(num_levels=14, desired_resolution=2048)
    elif mt == "vm":
        net = VMNeRFField(resolution0=netgold.VM_RES)
    else:
        net = MLPNeRFField(is_teacher=False)
    net = net.cuda()
    netgold.load_into(net, netgold.seeded_params(mt))
    return net


def _dump():
    try:
        os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "network_golden.json"), "w") as f:
            json.dump(_report, f, indent=1)
    except OSError:
        pass


@pytest.mark.parametrize("mt", ["hash", "vm", "mlp"])
def test_fused_field_matches_reference_network_golden(gold, mt):
    net = _net(mt)
    net.train()
    x, d, cs, cc, cf = (t.cuda() for t in netgold.query_points(mt))
    sigma, color = net(x, d)
    feat = net.feature_sigma_color
    netgold.scalar(sigma, color, feat, cs, cc, cf).backward()
    torch.cuda.synchronize()
    rec = {"sigma": _rel(sigma.detach().cpu().numpy(), gold[f"{mt}/sigma"]), "color": _rel(color.detach().cpu().numpy(), gold[f"{mt}/color"]),
           "feat": _rel(feat.detach().cpu().numpy(), gold[f"{mt}/feat"])}
    offsets = net.encoder.offsets.cpu().numpy() if mt == "hash" else None
    seen = 0
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        for k, v in netgold.summarise_grad(name, p.grad, offsets).items():
            g = gold[f"{mt}/{k}"]
            if k.startswith("gradsum/"):    # per level: the two feature sums are signed and cancel; the norm column is the robust one
                rec[k + ":norm"] = _rel(v[:, 2], g[:, 2])
                rec[k + ":sums"] = float(np.abs(v[:, :2] - g[:, :2]).max() / (np.abs(g[:, 2]).max() + 1e-30))
            else:
                rec[k] = _rel(v, g)
            seen += 1
    assert seen == sum(1 for k in gold if k.startswith(mt + "/grad")), "a parameter received no gradient"
    _report[mt] = rec
    _dump()
    bad = {k: v for k, v in rec.items() if not v <= TOL}
    assert not bad, f"{mt}: rel-L2 vs the reference network's golden above {TOL}: {bad}"


def test_fused_mlp_teacher_forward_matches_golden(gold):
    """The NeRF-MLP teacher's fused tcgen05 forward (no_grad, as distill_mutual/utils.py:1008-1018 evaluates it)."""
    net = _net("mlp")
    net.eval()
    x, d, *_ = (t.cuda() for t in netgold.query_points("mlp"))
    with torch.no_grad():
        sigma, color = net(x, d)
    feat = net.feature_sigma_color
    torch.cuda.synchronize()
    rec = {"sigma": _rel(sigma.cpu().numpy(), gold["mlp/sigma"]), "color": _rel(color.cpu().numpy(), gold["mlp/color"]),
           "feat": _rel(feat.cpu().numpy(), gold["mlp/feat"])}
    _report["mlp_fused_forward"] = rec
    _dump()
    assert all(v <= TOL for v in rec.values()), rec
