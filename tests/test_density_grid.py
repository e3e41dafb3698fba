"""Density-grid upkeep (NeRFRenderer.update_extra_state, distill_mutual/renderer.py:647-773): the numpy restatement
against torch on the CPU, and -- on the GPU -- the fused kernels of csrc/density_grid.cu against the restatement and against the
reference's own torch flow."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import cpu


def _torch_points(coords, noise, H, bound):
    """renderer.py:679-694 verbatim in torch (CPU: true division, see below)."""
    xyzs = 2 * coords.float() / (H - 1) - 1
    half = bound / H
    p = xyzs * (bound - half)
    p = p + (noise * 2 - 1) * half
    return p


def test_oracle_points_and_update_rule_cpu():
    H, bound = 128, 1.0
    g = torch.Generator().manual_seed(0)
    ind = torch.randint(0, H ** 3, (5000,), generator=g)
    coords = torch.from_numpy(cpu.morton3D_invert(ind.numpy().astype(np.int32)).astype(np.int64))
    noise = torch.rand(5000, 3, generator=g)
    got = cpu.density_grid_points(ind.numpy(), noise.numpy(), H, bound)
    want = _torch_points(coords, noise, H, bound).numpy()
    # torch on the CPU divides (a / b), on the GPU it multiplies by the fp32 reciprocal (the oracle follows the GPU): <= 1 ulp apart
    np.testing.assert_allclose(got, want, rtol=0, atol=2.4e-7)
    # every point lies inside its own cell: centre +- half a cell
    cell = np.floor((got / bound + 1) / 2 * H).clip(0, H - 1).astype(np.int64)
    assert np.array_equal(cell, coords.numpy())
    # update rule against torch's masked form
    grid = torch.rand(4096, generator=g)
    grid[::7] = -1.0            # cells marked untrained stay untouched
    sig = torch.rand(1500, generator=g) * 3
    idx = torch.randperm(4096, generator=g)[:1500]
    tmp = -torch.ones(4096)
    tmp[idx] = sig * 2.0
    ref = grid.clone()
    valid = (ref >= 0) & (tmp >= 0)
    ref[valid] = torch.maximum(ref[valid] * 0.95, tmp[valid])
    new, total = cpu.density_grid_update(grid.numpy(), sig.numpy(), idx.numpy(), 2.0, 0.95)
    assert np.array_equal(new, ref.numpy())
    assert abs(total - float(ref.clamp(min=0).double().sum())) < 1e-6
    bits, mean = cpu.packbits_mean(new, total, 0.01)
    assert abs(mean - float(ref.clamp(min=0).mean())) < 1e-6
    assert np.array_equal(bits, cpu.packbits(new, min(mean, 0.01)))


@pytest.mark.gpu
def test_upkeep_kernels_match_oracle():
    from pvd_b200 import _native as nv
    l = nv.lib()
    dev = "cuda"
    H, bound = 128, 2.0
    g = torch.Generator().manual_seed(1)
    for full in (True, False):
        n = H ** 3 if full else 70001
        ind = None if full else torch.randint(0, H ** 3, (n,), generator=g, dtype=torch.int32)
        noise = torch.rand(n, 3, generator=g)
        xyz = torch.empty(n, 3, device=dev)
        dn, di = noise.to(dev), None if ind is None else ind.to(dev)
        nv.check(l.pvd_density_grid_points(nv.ptr(di), nv.ptr(dn), C.c_uint32(n), C.c_uint32(H), C.c_float(bound), nv.ptr(xyz), nv.stream_of(xyz)))
        want = cpu.density_grid_points(None if ind is None else ind.numpy(), noise.numpy(), H, bound)
        assert np.array_equal(xyz.cpu().numpy(), want), "query points differ from the oracle"
        if not full:   # and from torch's own arithmetic on the GPU, bit for bit
            coords = torch.from_numpy(cpu.morton3D_invert(ind.numpy()).astype(np.int64)).to(dev)
            assert torch.equal(xyz, _torch_points(coords, dn, H, bound))
        grid = torch.rand(H ** 3, generator=g)
        grid[::5] = -1.0
        sig = torch.rand(n, generator=g) * 4
        dgrid, dsig = grid.to(dev), sig.to(dev)
        tmp = torch.empty(H ** 3, device=dev)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        nv.check(l.pvd_density_grid_update(nv.ptr(dgrid), nv.ptr(tmp), nv.ptr(di), nv.ptr(dsig), C.c_uint32(n), C.c_uint32(H ** 3),
                                           C.c_float(1.5), C.c_float(0.95), nv.ptr(acc), nv.stream_of(dgrid)))
        new, total = cpu.density_grid_update(grid.numpy(), sig.numpy(), None if ind is None else ind.numpy(), 1.5, 0.95)
        assert np.array_equal(dgrid.cpu().numpy(), new), "EMA update differs from the oracle"
        assert abs(float(acc) - total) < 1e-6 * total
        bf = torch.empty(H ** 3 // 8, dtype=torch.uint8, device=dev)
        mean = torch.zeros(1, device=dev)
        nv.check(l.pvd_packbits_mean(nv.ptr(dgrid), C.c_uint32(H ** 3 // 8), nv.ptr(acc), C.c_uint32(H ** 3), C.c_float(0.3), nv.ptr(bf),
                                     nv.ptr(mean), nv.stream_of(dgrid)))
        bits, m = cpu.packbits_mean(new, float(acc), 0.3)
        assert abs(float(mean) - m) < 1e-7 and np.array_equal(bf.cpu().numpy(), bits)


@pytest.mark.gpu
def test_fused_update_extra_state_equals_reference_flow(monkeypatch):
    """19 updates (16 full sweeps + 3 partial ones) of a field whose density is constant inside each grid cell: the fused path and
    the reference's torch flow draw the same random numbers and must leave the SAME grid, bitfield and mean density.  (The jitter
    is confined to the middle half of a cell so that fp32 rounding cannot push a point across a cell face: the two paths pair
    the noise with the cells in a different order.)"""
    from pvd_b200.renderer import NeRFRenderer
    rand, rand_like = torch.rand, torch.rand_like
    monkeypatch.setattr(torch, "rand", lambda *a, **k: 0.25 + 0.5 * rand(*a, **k))
    monkeypatch.setattr(torch, "rand_like", lambda *a, **k: 0.25 + 0.5 * rand_like(*a, **k))

    class CellField(NeRFRenderer):
        def density(self, x):
            H = self.grid_size
            c = ((x / self._b + 1) * 0.5 * H).floor().clamp(0, H - 1).long()
            h = (c[:, 0] * 73856093) ^ (c[:, 1] * 19349663) ^ (c[:, 2] * 83492791)
            return {"sigma": ((h % 1000).float() / 1000.0) ** 6 * 5.0 * self._gain}

    res = []
    for fused in (False, True):
        torch.manual_seed(0)
        net = CellField(bound=1).cuda()
        net._b = 1.0
        net.density_grid[0, ::11] = -1.0      # untrained cells
        for it in range(19):
            net._gain = 0.5 + 0.05 * ((it * 7) % 11)   # the field changes between updates: decay and max both matter
            net.local_step = 3
            net.step_counter[:3, 0] = torch.tensor([100, 200, 300], dtype=torch.int32)
            net.update_extra_state(fused=fused)
        res.append((net.density_grid.clone(), net.density_bitfield.clone(), net.mean_density, net.mean_count, net.iter_density))
    (g0, b0, m0, c0, i0), (g1, b1, m1, c1, i1) = res
    assert i0 == i1 == 19 and c0 == c1 == 200
    assert torch.equal(g0, g1), f"grids differ in {(g0 != g1).sum().item()} cells"
    assert torch.equal(b0, b1)
    assert abs(m0 - m1) < 1e-7 * max(m0, 1e-9)
    assert int(b1.count_nonzero()) > 0
