"""The fused CUDA fields against golden vectors produced by THE REFERENCE's `distill_mutual/network.py::NeRFNetwork.forward`
(tests/golden/ref_network_golden.npz, generated in the build container by tests/golden/make_network_golden.py: fp32, CPU).

The fused fields compute in the reference's autocast precision (fp16 table / fp16 tensor-core MLP, fp32 accumulation), the golden
is the fp32 result, so the bound is north_star's fp16 one: relative L2 <= 1e-2 per tensor -- or, where the REFERENCE's own
autocast path (the same network through oracle/ref_pipeline.py's modules = the reference's kernels + cuBLAS under autocast, run here
on the same seeded parameters and points) is further than that from the fp32 golden, no further than 1.25 x its distance.  All three
distances are recorded in gpurun_out/network_golden.json.
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
    from pvd_b200.fused_tensors import TensorsNeRFField
    from pvd_b200.fused_vm import VMNeRFField
    if mt == "tensors":
        net = TensorsNeRFField(plenoxel_degree=3, plenoxel_res=netgold.TENSORS_RES)
    elif mt == "hash":
        net = HashNeRFField(num_levels=14, desired_resolution=2048)
    elif mt == "vm":
        net = VMNeRFField(resolution0=netgold.VM_RES)
    else:
        net = MLPNeRFField(is_teacher=False)
    net = net.cuda()
    netgold.load_into(net, netgold.seeded_params(mt))
    return net


def _reference_autocast(mt, ref_ext):
    """{record: value} of the reference's OWN autocast path on the golden's parameters / points (None without oracle/_ref)."""
    if ref_ext is None or mt == "tensors":
        return None
    from oracle import ref_pipeline as rp
    P = netgold.seeded_params(mt)
    if mt == "hash":
        from oracle import cpu
        offsets, pls = cpu.grid_offsets(3, 14, 16, 19, desired_resolution=2048)
        net = rp.RefHashNetwork(ref_ext, offsets, pls).cuda()
        P = {("embeddings" if k == "encoder.embeddings" else k): v for k, v in P.items()}
    elif mt == "mlp":
        net = rp.RefMlpNetwork(ref_ext).cuda()      # a TRAINED mlp model: autograd through cuBLAS under autocast in the reference
        offsets = None
    else:
        net = rp.RefVmNetwork(ref_ext, resolution=netgold.VM_RES).cuda()
        offsets = None
    netgold.load_into(net, P)
    x, d, cs, cc, cf = (t.cuda() for t in netgold.query_points(mt))
    with torch.autocast("cuda", dtype=torch.float16):
        sigma, color = net(x, d)
        feat = net.feature_sigma_color
        loss = netgold.scalar(sigma.float(), color.float(), feat.float(), cs, cc, cf)
    scale = 128.0                       # GradScaler-style: keeps the fp16 gradients of the autocast backward out of the subnormals
    (loss * scale).backward()
    out = {"sigma": sigma.detach().float().cpu().numpy(), "color": color.detach().float().cpu().numpy(), "feat": feat.detach().float().cpu().numpy()}
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        name = "encoder.embeddings" if name == "embeddings" else name
        out.update(netgold.summarise_grad(name, p.grad.float() / scale, offsets))
    return out


def _dump():
    try:
        os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "network_golden.json"), "w") as f:
            json.dump(_report, f, indent=1)
    except OSError:
        pass


@pytest.mark.parametrize("mt", ["hash", "vm", "mlp", "tensors"])
def test_fused_field_matches_reference_network_golden(gold, ref_ext, mt):
    amp = _reference_autocast(mt, ref_ext)
    net = _net(mt)
    net.train()
    x, d, cs, cc, cf = (t.cuda() for t in netgold.query_points(mt))
    sigma, color = net(x, d)
    feat = net.feature_sigma_color
    scale = 128.0 if mt == "mlp" else 1.0   # mlp: the data gradients travel between layers as fp16 tiles, as under the reference's autocast + GradScaler
    (netgold.scalar(sigma, color, feat, cs, cc, cf) * scale).backward()
    torch.cuda.synchronize()
    rec, noise = {}, {}

    def judge(key, ours, golden, ref_amp):
        if key.startswith("gradsum/"):    # per level (sum f0, sum f1, norm): the sums are signed and cancel; measure them against the norms
            dist = lambda a: max(_rel(a[:, 2], golden[:, 2]), float(np.abs(a[:, :2] - golden[:, :2]).max() / (np.abs(golden[:, 2]).max() + 1e-30)))
        else:
            dist = lambda a: _rel(a, golden)
        rec[key] = dist(ours)
        if ref_amp is not None:
            noise[key] = dist(ref_amp)

    judge("sigma", sigma.detach().cpu().numpy(), gold[f"{mt}/sigma"], amp and amp["sigma"])
    judge("color", color.detach().cpu().numpy(), gold[f"{mt}/color"], amp and amp["color"])
    if feat is not None:
        judge("feat", feat.detach().cpu().numpy(), gold[f"{mt}/feat"], amp and amp["feat"])
    offsets = net.encoder.offsets.cpu().numpy() if mt == "hash" else None
    seen = 0
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        for k, v in netgold.summarise_grad(name, p.grad / scale, offsets).items():
            judge(k, v, gold[f"{mt}/{k}"], amp and amp.get(k))
            seen += 1
    assert seen == sum(1 for k in gold if k.startswith(mt + "/grad")), "a parameter received no gradient"
    _report[mt] = {"ours_vs_golden_fp32": rec, "reference_autocast_vs_golden_fp32": noise}
    _dump()
    tol = 2e-5 if mt == "tensors" else TOL     # the tensors field is fp32 end to end: north_star's 1e-4 fp32 bound with room to spare
    # whole tensors: 1.25 x the reference's own autocast deviation.  Row / column SUMS of a large gradient (gradrow / gradcol records) are
    # signed sums of 256 entries that largely cancel: their relative deviation is itself noisy from run to run, so they get 2 x
    factor = lambda k: 2.0 if k.startswith(("gradrow/", "gradcol/")) else 1.25
    bad = {k: (v, noise.get(k)) for k, v in rec.items() if not (v <= tol or v <= factor(k) * noise.get(k, 0.0))}
    assert not bad, f"{mt}: (ours, reference-autocast) rel-L2 vs the reference network's fp32 golden: {bad}"


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
