#!/usr/bin/env python
"""Generate tests/golden/ref_network_golden.npz by running THE REFERENCE's `distill_mutual/network.py::NeRFNetwork.forward`
(imported from /root/reference over this repo's drop-in packages, tests/refnet.py) in fp32 on the CPU, for model types hash, vm
and mlp, with the seeded parameters of tests/golden/netgold.py.

Only runs in the build container (the reference cannot travel); the .npz it writes is committed and is what pins
  * oracle/field.py's restatements (tests/test_reference_dropin_cpu.py, CPU, tight tolerance), and
  * the fused CUDA fields (tests/test_gpu_network_golden.py, B200, fp16 tolerance)
to the reference's own network code.

    python tests/golden/make_network_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import netgold  # noqa: E402
import refnet  # noqa: E402


def build(net_mod, model_type):
    args = refnet.make_args(resolution0=netgold.VM_RES, plenoxel_degree=3, plenoxel_res=str(list(netgold.TENSORS_RES)))
    net = refnet.construct(net_mod.NeRFNetwork, encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10, bg_radius=-1,
                              model_type=model_type, args=args, is_teacher=False)
    netgold.load_into(net, netgold.seeded_params(model_type))
    return refnet.cpu_standins(net)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    net_mod, _ = refnet.load()
    out = {}
    for mt in ("hash", "vm", "mlp", "tensors"):
        net = build(net_mod, mt)
        net.train()
        x, d, cs, cc, cf = netgold.query_points(mt)
        sigma, color = net(x, d)
        feat = net.feature_sigma_color
        assert net.color_l is color and (feat is None if mt == "tensors" else torch.equal(net.sigma_l, feat[..., 0]))
        netgold.scalar(sigma, color, feat, cs, cc, cf).backward()
        out[f"{mt}/sigma"] = sigma.detach().numpy()
        out[f"{mt}/color"] = color.detach().numpy()
        if feat is not None:
            out[f"{mt}/feat"] = feat.detach().numpy()
        offsets = net.encoder.enc.offsets.numpy() if mt == "hash" else None
        for name, p in net.named_parameters():
            name = name.replace("encoder.enc.", "encoder.")
            if p.grad is None:
                continue
            for k, v in netgold.summarise_grad(name, p.grad, offsets).items():
                out[f"{mt}/{k}"] = v
        print(mt, "sigma", float(sigma.mean()), "color", float(color.mean()), "keys", sum(k.startswith(mt) for k in out))
    path = os.path.join(HERE, "ref_network_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; stubbed third-party modules:", refnet.stubbed())


if __name__ == "__main__":
    main()
