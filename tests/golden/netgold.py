"""Seeded parameters / query points shared by make_network_golden.py (which runs the REFERENCE's network.py in the build container)
and by the tests that check the oracle and the fused CUDA fields against the committed golden outputs.

numpy's legacy RandomState is bit-stable across numpy versions, so the (large) parameters never have to be stored: only the query
points, the outputs and the gradients of the small tensors are committed (tests/golden/ref_network_golden.npz).
"""
from __future__ import annotations

import numpy as np
import torch

N_POINTS = 384
VM_RES = 24
SHAPES = {
    "hash": [("encoder.embeddings", (5303704, 2)), ("sigma_net.0.weight", (64, 28)), ("sigma_net.1.weight", (16, 64)),
             ("color_net.0.weight", (64, 31)), ("color_net.1.weight", (64, 64)), ("color_net.2.weight", (3, 64))],
    "vm": [(f"{g}.{i}", (1, r, VM_RES, w)) for g, r, w in (("sigma_mat", 16, VM_RES), ("sigma_vec", 16, 1), ("color_mat", 48, VM_RES),
                                                           ("color_vec", 48, 1)) for i in range(3)]
          + [("basis_mat.weight", (15, 144)), ("color_net.0.weight", (64, 31)), ("color_net.1.weight", (64, 64)), ("color_net.2.weight", (3, 64))],
    "mlp": [(f"nerf_mlp.{i}.{k}", s) for i, (o, n) in enumerate([(256, 63), (256, 256), (256, 256), (256, 256), (256, 319), (256, 256),
                                                                 (256, 256), (28, 256)]) for k, s in (("weight", (o, n)), ("bias", (o,)))]
           + [("sigma_net.0.weight", (64, 28)), ("sigma_net.1.weight", (16, 64)), ("color_net.0.weight", (64, 31)),
              ("color_net.1.weight", (64, 64)), ("color_net.2.weight", (3, 64))],
}
TENSORS_RES = (12, 10, 14)   # D, H, W of the golden's plenoxel volume (degree 3 -> 28 channels)
SHAPES["tensors"] = [("tensor_volume.0", (1, 28) + TENSORS_RES)]
SEEDS = {"hash": 101, "vm": 202, "mlp": 303, "tensors": 404}


def seeded_params(model_type: str) -> dict:
    """{reference parameter name: float32 tensor}, drawn in the fixed order of SHAPES from RandomState(SEEDS[model_type])."""
    rs = np.random.RandomState(SEEDS[model_type])
    out = {}
    for name, shape in SHAPES[model_type]:
        if name == "encoder.embeddings":
            a = rs.uniform(-0.5, 0.5, shape)
        elif name.startswith(("sigma_mat", "sigma_vec", "color_mat", "color_vec")):
            a = 0.4 * rs.standard_normal(shape)
        elif name.startswith("tensor_volume"):
            a = 0.8 * rs.standard_normal(shape)
        elif name.endswith("bias"):
            a = rs.uniform(-0.1, 0.1, shape)
        else:  # Linear weight [out, in]: keeps activations of order one through the ReLU stacks
            a = rs.uniform(-1.0, 1.0, shape) * np.sqrt(3.0 / shape[1]) * 1.4
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def query_points(model_type: str):
    """(x [N,3] inside the unit cube, d [N,3] unit, cotangents for sigma [N], color [N,3], feat [N,16])."""
    rs = np.random.RandomState(SEEDS[model_type] + 1)
    x = rs.uniform(-0.95, 0.95, (N_POINTS, 3)).astype(np.float32)
    d = rs.standard_normal((N_POINTS, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    cs = rs.standard_normal(N_POINTS).astype(np.float32) * 0.1
    cc = rs.standard_normal((N_POINTS, 3)).astype(np.float32)
    cf = rs.standard_normal((N_POINTS, 16)).astype(np.float32) * 0.3
    return tuple(torch.from_numpy(a) for a in (x, d, cs, cc, cf))


def load_into(module: torch.nn.Module, params: dict):
    """Copy seeded parameters into a module with the reference's parameter names (ours or the reference's own NeRFNetwork)."""
    own = dict(module.named_parameters())
    with torch.no_grad():
        for k, v in params.items():
            own[k].copy_(v.to(own[k].device))


def scalar(sigma, color, feat, cs, cc, cf):
    """The fixed scalar whose gradients are committed: <sigma, cs> + <color, cc> + <feat, cf> (feat None for the tensors model)."""
    out = (sigma * cs).sum() + (color * cc).sum()
    return out if feat is None else out + (feat * cf).sum()


SMALL = 8192   # gradients of tensors up to this many elements are stored whole; larger ones as summaries


def summarise_grad(name: str, g: torch.Tensor, offsets=None) -> dict:
    """What is committed for one parameter gradient: the tensor itself when small; per-level (sum f0, sum f1, L2 norm) for the hash
    table; (row sums, column sums, L2 norm) of the [out | rank, rest] matrix view for the other large ones."""
    g = g.detach().double().cpu()
    if g.numel() <= SMALL:
        return {f"grad/{name}": g.float().numpy()}
    if offsets is not None and name == "encoder.embeddings":
        rows = []
        for l in range(len(offsets) - 1):
            s = g[int(offsets[l]):int(offsets[l + 1])]
            rows.append([float(s[:, 0].sum()), float(s[:, 1].sum()), float(s.norm())])
        return {f"gradsum/{name}": np.array(rows, np.float64)}
    g2 = g.reshape(g.shape[0] if g.dim() == 2 else g.shape[1], -1)
    return {f"gradrow/{name}": g2.sum(1).numpy(), f"gradcol/{name}": g2.sum(0).numpy(), f"gradnorm/{name}": np.array([float(g2.norm())])}
