"""Pin the CPU oracle to the reference itself: tests/golden/ref_golden.npz holds outputs of the reference's own CUDA
kernels (oracle/_ref = /root/reference compiled unmodified for sm_100a) run on a B200 by tests/golden/make_golden.py.
Integer / exactly-rounded results must match bit for bit; results that pass through __expf, exp2f or float atomics carry
the tolerance written next to the check."""
import hashlib
import os

import numpy as np
import pytest

from oracle import cpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(G), reason="golden fixture not generated yet")

AABB = np.array([-1, -1, -1, 1, 1, 1], np.float32)


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(G))


def test_scene_fixture_is_the_same(gold, scene):
    assert hashlib.sha256(scene["bitfield"].tobytes()).digest() == gold["bitfield_sha256"].tobytes()


def test_near_far_bit_exact(gold):
    n, f = cpu.near_far_from_aabb(gold["rays_o"], gold["rays_d"], AABB, 0.2)
    assert np.array_equal(n, gold["nears"]) and np.array_equal(f, gold["fars"])


@pytest.mark.parametrize("tag,perturb,dt_gamma", [("p1", True, 0.0), ("p0", False, 0.0), ("g", True, 1.0 / 128)])
def test_march_bit_exact(gold, scene, tag, perturb, dt_gamma):
    M = gold[f"march_{tag}_xyzs"].shape[0]
    xyzs, dirs, deltas, rays, counter = cpu.march_rays_train(gold["rays_o"], gold["rays_d"], 1.0, scene["bitfield"], 1, 128,
                                                            gold["nears"], gold["fars"], M=M, perturb=perturb,
                                                            dt_gamma=dt_gamma, max_steps=1024)
    assert np.array_equal(counter, gold[f"march_{tag}_counter"])
    assert np.array_equal(rays, gold[f"march_{tag}_rays"])  # per-ray sample counts and canonical offsets
    assert np.array_equal(xyzs, gold[f"march_{tag}_xyzs"])
    assert np.array_equal(deltas, gold[f"march_{tag}_deltas"])


def test_composite_forward_backward(gold):
    rays, deltas = gold["march_p1_rays"], gold["march_p1_deltas"]
    ws, depth, image = cpu.composite_rays_train_forward(gold["comp_sigmas"], gold["comp_rgbs"], deltas, rays)
    # the reference uses __expf (ex2.approx); the oracle expf: ~1e-6 relative per sample
    np.testing.assert_allclose(ws, gold["comp_ws"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(depth, gold["comp_depth"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(image, gold["comp_image"], rtol=1e-5, atol=1e-6)
    gs, gc = cpu.composite_rays_train_backward(gold["comp_gws"], gold["comp_gim"], gold["comp_sigmas"], gold["comp_rgbs"], deltas,
                                               rays, gold["comp_ws"], gold["comp_image"])
    np.testing.assert_allclose(gc, gold["comp_grad_rgbs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gs, gold["comp_grad_sigmas"], rtol=1e-3, atol=2e-7)


def test_grid_encoder(gold):
    import torch
    offsets, pls = gold["grid_offsets"], float(gold["grid_pls"][0])
    if "grid_emb" in gold:
        emb = gold["grid_emb"]
    else:
        g = torch.Generator().manual_seed(11)
        emb = (torch.rand(int(offsets[-1]), 2, generator=g) * 2 - 1).numpy()
        assert np.isclose(float(emb.astype(np.float64).sum()), gold["grid_emb_checksum"][0])
    x, gg = gold["grid_x"], gold["grid_g"]
    pinned = "grid_scales" in gold
    if pinned:  # device-computed exp2f scales -> bit-exact features
        cpu.set_level_scales(gold["grid_scales"])
    try:
        out, dy_dx = cpu.grid_encode_forward(x, emb, offsets, pls, 16, True, 0, False)
        if pinned:
            assert np.array_equal(out, gold["grid_out"])
        else:
            np.testing.assert_allclose(out, gold["grid_out"], rtol=1e-4, atol=5e-4)
        np.testing.assert_allclose(dy_dx, gold["grid_dydx"], rtol=1e-3, atol=2e-2)
        ge, gi = cpu.grid_encode_backward(gg, x, emb.shape, offsets, pls, 16, gold["grid_dydx"], 0, False)
        np.testing.assert_allclose(ge, gold["grid_gemb"], rtol=1e-4, atol=1e-3)  # float atomics: summation order differs
        np.testing.assert_allclose(gi, gold["grid_gin"], rtol=1e-4, atol=1e-2)
    finally:
        cpu.set_level_scales(None)


def test_sh_morton_packbits(gold):
    from oracle import sh_reference
    got = cpu.sh_encode_forward(gold["sh_dirs"], 4)
    np.testing.assert_allclose(got, gold["sh_out4"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sh_reference.real_sh(gold["sh_dirs"], 8), gold["sh_out8"], rtol=1e-4, atol=2e-5)
    assert np.array_equal(cpu.morton3D(gold["morton_coords"]), gold["morton_ind"])
    assert np.array_equal(cpu.packbits(gold["pack_grid"], 0.5), gold["pack_bits"])
