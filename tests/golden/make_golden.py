"""Generate tests/golden/ref_golden.npz from the REFERENCE's own CUDA kernels (oracle/_ref) on a GPU box.

Run (on the B200 box, from the repo root):   python tests/golden/make_golden.py gpurun_out/ref_golden.npz
then copy the file to tests/golden/ref_golden.npz and commit it.  The fixture pins the CPU oracle
(tests/test_oracle_golden.py) to outputs of the reference itself, since the reference ships no golden vectors.

Every input is stored next to the reference's output, so the CPU test needs neither a GPU nor /root/reference.
Sizes are small (a few hundred KB compressed).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aaai2023-pvd_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))


def main(out_path):
    import _raymarching, _gridencoder, _shencoder  # the unmodified reference, compiled by oracle/build_ref.py
    from oracle import ref_glue, cpu
    from pvd_b200 import synthetic as syn

    ext = {"raymarching": _raymarching, "gridencoder": _gridencoder, "shencoder": _shencoder}
    out = {}
    dev = "cuda"
    grid, bitfield, sha = syn.lego_bitfield()
    out["bitfield_sha256"] = np.frombuffer(bytes.fromhex(sha), dtype=np.uint8)
    ro, rd = syn.make_ray_batches(1, 4096, seed=0)[0]
    ro, rd = ro[:768].contiguous(), rd[:768].contiguous()
    out["rays_o"], out["rays_d"] = ro.numpy(), rd.numpy()
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32)
    gro, grd, gbf = ro.to(dev), rd.to(dev), torch.from_numpy(bitfield).to(dev)
    nears, fars = ref_glue.near_far_from_aabb(ext, gro, grd, aabb.to(dev), 0.2)
    out["nears"], out["fars"] = nears.cpu().numpy(), fars.cpu().numpy()

    for tag, perturb, dt_gamma, max_steps in (("p1", True, 0.0, 1024), ("p0", False, 0.0, 1024), ("g", True, 1.0 / 128, 1024)):
        M = 768 * max_steps
        r = ref_glue.march_rays_train_raw(ext, gro, grd, 1.0, gbf, 1, 128, nears, fars, M, perturb, dt_gamma, max_steps)
        total = int(r[4][0].item())
        Mc = total + (128 - total % 128)
        cx, cd, cl, cr = ref_glue.canonicalize(r[0], r[1], r[2], r[3], M_out=Mc)
        out[f"march_{tag}_counter"] = r[4].cpu().numpy()
        out[f"march_{tag}_rays"] = cr.cpu().numpy()
        out[f"march_{tag}_xyzs"] = cx.cpu().numpy()
        out[f"march_{tag}_deltas"] = cl.cpu().numpy()
        if tag == "p1":
            g = torch.Generator().manual_seed(7)
            sig = (torch.rand(Mc, generator=g) * 30)
            rgb = torch.rand(Mc, 3, generator=g)
            gws = torch.rand(768, generator=g)
            gim = torch.rand(768, 3, generator=g)
            ws, depth, image = ref_glue.composite_rays_train_forward(ext, sig.to(dev), rgb.to(dev), cl, cr)
            gs, gc = ref_glue.composite_rays_train_backward(ext, gws.to(dev), gim.to(dev), sig.to(dev), rgb.to(dev), cl, cr, ws, image)
            out["comp_sigmas"], out["comp_rgbs"], out["comp_gws"], out["comp_gim"] = sig.numpy(), rgb.numpy(), gws.numpy(), gim.numpy()
            out["comp_ws"], out["comp_depth"], out["comp_image"] = ws.cpu().numpy(), depth.cpu().numpy(), image.cpu().numpy()
            out["comp_grad_sigmas"], out["comp_grad_rgbs"] = gs.cpu().numpy(), gc.cpu().numpy()

    # grid encoder: small table with 1 dense + 7 hashed levels; the table itself is regenerated from the seed by the test
    offsets, pls = cpu.grid_offsets(3, 8, 16, 13, desired_resolution=512)
    g = torch.Generator().manual_seed(11)
    emb = torch.rand(int(offsets[-1]), 2, generator=g) * 2 - 1
    out["grid_emb_checksum"] = np.array([float(emb.double().sum()), float(emb.double().abs().sum())])
    x = torch.rand(2048, 3, generator=g)
    x[0] = torch.tensor([1.5, 0.2, 0.2])
    gg = torch.randn(2048, 16, generator=g)
    goff = torch.from_numpy(offsets).to(dev)
    o, j = ref_glue.grid_encode_forward(ext, x.to(dev), emb.to(dev), goff, pls, 16, True, 0, False)
    ge, gi = ref_glue.grid_encode_backward(ext, gg.to(dev), x.to(dev), emb.to(dev), goff, pls, 16, j, 0, False)
    from gridencoder.grid import level_table
    sc, rs = level_table(goff, pls, 16)  # this library's device table; the reference computes the same expression inline
    out["grid_scales"], out["grid_res"] = sc.cpu().numpy(), rs.cpu().numpy()
    out["grid_offsets"], out["grid_pls"] = offsets, np.array([pls], np.float64)
    out["grid_x"], out["grid_g"] = x.numpy(), gg.numpy()
    out["grid_out"], out["grid_dydx"] = o.cpu().numpy(), j.cpu().numpy()
    out["grid_gemb"], out["grid_gin"] = ge.cpu().numpy(), gi.cpu().numpy()
    # per-level scale as the device computes it (exp2f on the GPU) via a 1-point probe is not observable; store resolution table
    # SH
    d = torch.randn(1024, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    out["sh_dirs"] = d.numpy()
    for deg in (4, 8):
        so, _ = ref_glue.sh_encode_forward(ext, d.to(dev), deg, False)
        out[f"sh_out{deg}"] = so.cpu().numpy()
    # morton / packbits
    rng = np.random.default_rng(5)
    coords = rng.integers(0, 128, size=(512, 3)).astype(np.int32)
    ind = torch.empty(512, dtype=torch.int32, device=dev)
    _raymarching.morton3D(torch.from_numpy(coords).to(dev), 512, ind)
    out["morton_coords"], out["morton_ind"] = coords, ind.cpu().numpy()
    dens = rng.random(8 * 4096).astype(np.float32)
    bits = torch.empty(4096, dtype=torch.uint8, device=dev)
    _raymarching.packbits(torch.from_numpy(dens).to(dev), 4096, 0.5, bits)
    out["pack_grid"], out["pack_bits"] = dens, bits.cpu().numpy()
    # pcg32 jitter as the device produces it: t0 - near for rays with perturb, recovered from the first sample is fragile;
    # instead march a fully occupied grid from near = 0 with dt_gamma = 0 and read xyz of the first sample.
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_golden.npz"))
