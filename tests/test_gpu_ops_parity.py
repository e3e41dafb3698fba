"""GPU parity of the drop-in operators (through the C ABI) against the CPU oracle and, when oracle/_ref was built,
against the reference's own CUDA extensions.  Integer results (sample counts, offsets, indices, bitfields) and
everything derived from exact arithmetic must be bit-exact; float results carry the tolerance written at the check.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

AABB = np.array([-1, -1, -1, 1, 1, 1], np.float32)


def dev(a):
    return torch.as_tensor(a).cuda()


def _setup(scene, b=0, bound=1.0):
    ro, rd = scene["batches"][b]
    return ro.cuda(), rd.cuda(), dev(scene["bitfield"])


def test_native_library_is_loaded():
    from pvd_b200 import _native as nv
    assert nv.lib().pvd_abi_version() == nv.ABI_VERSION
    with open("/proc/self/maps") as f:
        assert "libpvd_b200.so" in f.read()


def test_near_far_bit_exact(scene, ref_ext):
    import raymarching
    from oracle import cpu, ref_glue
    ro, rd, _ = _setup(scene)
    aabb = dev(AABB)
    n, f = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    on, of = cpu.near_far_from_aabb(ro.cpu().numpy(), rd.cpu().numpy(), AABB, 0.2)
    assert np.array_equal(n.cpu().numpy(), on) and np.array_equal(f.cpu().numpy(), of)
    if ref_ext:
        rn, rf = ref_glue.near_far_from_aabb(ref_ext, ro, rd, aabb, 0.2)
        assert torch.equal(rn, n) and torch.equal(rf, f)


@pytest.mark.parametrize("perturb,dt_gamma,max_steps", [(True, 0.0, 1024), (False, 0.0, 1024), (True, 1.0 / 128, 1024),
                                                         (True, 0.0, 512)])
def test_march_rays_train_bit_exact(scene, ref_ext, perturb, dt_gamma, max_steps):
    import raymarching
    from oracle import cpu, ref_glue
    ro, rd, bf = _setup(scene)
    aabb = dev(AABB)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, counter, -1, perturb, 128,
                                                            False, dt_gamma, max_steps)
    oxyzs, odirs, odeltas, orays, ocnt = cpu.march_rays_train(ro.cpu().numpy(), rd.cpu().numpy(), 1.0, scene["bitfield"], 1, 128,
                                                             nears.cpu().numpy(), fars.cpu().numpy(), M=xyzs.shape[0],
                                                             perturb=perturb, dt_gamma=dt_gamma, max_steps=max_steps)
    total = int(counter[0].item())
    assert total == int(ocnt[0]) and int(counter[1].item()) == ro.shape[0]
    assert total > 5000
    assert xyzs.shape[0] == total + (128 - total % 128)  # strict round-up, raymarching.py:278-279
    assert np.array_equal(rays.cpu().numpy(), orays)
    assert np.array_equal(xyzs.cpu().numpy(), oxyzs)
    assert np.array_equal(dirs.cpu().numpy(), odirs)
    assert np.array_equal(deltas.cpu().numpy(), odeltas)
    if ref_ext:
        M = ro.shape[0] * max_steps
        r = ref_glue.march_rays_train_raw(ref_ext, ro, rd, 1.0, bf, 1, 128, nears, fars, M, perturb, dt_gamma, max_steps)
        assert r[4].tolist() == [total, ro.shape[0]]
        cx, cd, cl, cr = ref_glue.canonicalize(r[0], r[1], r[2], r[3], M_out=xyzs.shape[0])
        assert torch.equal(cr, rays)
        assert torch.equal(cx, xyzs) and torch.equal(cd, dirs) and torch.equal(cl, deltas)


def test_march_rays_train_cascade2_bit_exact(ref_ext):
    """bound = 2 -> two cascades: exercises mip_from_pos / mip_from_dt and non-unit mip_bound."""
    import raymarching
    from oracle import cpu, ref_glue
    from pvd_b200 import synthetic as syn
    g = syn.lego_density_grid(128, 2.0, 2)
    bfn = syn.pack_bitfield(g)
    ro, rd = syn.make_ray_batches(1, 2048, seed=3)[0]
    ro, rd, bf = ro.cuda(), rd.cuda(), dev(bfn)
    aabb = dev(np.array([-2, -2, -2, 2, 2, 2], np.float32))
    nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 2.0, bf, 2, 128, nears, fars, counter, -1, True, 128, False,
                                                            0.0, 1024)
    o = cpu.march_rays_train(ro.cpu().numpy(), rd.cpu().numpy(), 2.0, bfn, 2, 128, nears.cpu().numpy(), fars.cpu().numpy(),
                             M=xyzs.shape[0], perturb=True)
    assert int(counter[0].item()) == int(o[4][0]) > 1000
    assert np.array_equal(rays.cpu().numpy(), o[3])
    assert np.array_equal(xyzs.cpu().numpy(), o[0]) and np.array_equal(deltas.cpu().numpy(), o[2])
    if ref_ext:
        r = ref_glue.march_rays_train_raw(ref_ext, ro, rd, 2.0, bf, 2, 128, nears, fars, 2048 * 1024, True, 0.0, 1024)
        cx, cd, cl, cr = ref_glue.canonicalize(r[0], r[1], r[2], r[3], M_out=xyzs.shape[0])
        assert torch.equal(cr, rays) and torch.equal(cx, xyzs) and torch.equal(cl, deltas)


def test_march_overflow_drops_highest_offsets(scene):
    """mean_count path: rays with offset + n >= M write nothing and composite returns zeros for them."""
    import raymarching
    ro, rd, bf = _setup(scene)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, dev(AABB), 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, counter, 20000, True, 128,
                                                            False, 0.0, 1024)
    M = xyzs.shape[0]
    assert M == 20000 + (128 - 20000 % 128)
    r = rays.long()
    kept = (r[:, 2] > 0) & (r[:, 1] + r[:, 2] < M)
    last = int((r[kept, 1] + r[kept, 2]).max().item())
    assert torch.all(xyzs[last:] == 0) and torch.all(deltas[last:] == 0)
    assert torch.all(deltas[:last, 0] > 0)
    sig = torch.rand(M, device="cuda")
    rgb = torch.rand(M, 3, device="cuda")
    ws, depth, img = raymarching.composite_rays_train(sig, rgb, deltas, rays)
    assert torch.all(ws[~kept] == 0) and torch.all(img[~kept] == 0) and torch.all(ws[kept] > 0)


def _march(scene, b=0):
    import raymarching
    ro, rd, bf = _setup(scene, b)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, dev(AABB), 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    return raymarching.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, counter, -1, True, 128, False, 0.0, 1024)


def test_composite_train_forward_backward(scene, ref_ext):
    import raymarching
    from oracle import cpu, ref_glue
    xyzs, dirs, deltas, rays = _march(scene, 1)
    M = xyzs.shape[0]
    g = torch.Generator(device="cuda").manual_seed(1)
    sigmas = (torch.rand(M, device="cuda", generator=g) * 30).requires_grad_(True)
    rgbs = torch.rand(M, 3, device="cuda", generator=g).requires_grad_(True)
    ws, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays)
    ows, odepth, oimage = cpu.composite_rays_train_forward(sigmas.detach().cpu().numpy(), rgbs.detach().cpu().numpy(),
                                                           deltas.cpu().numpy(), rays.cpu().numpy())
    # warp-scan vs serial order and __expf vs expf: 2e-5 relative on O(1) quantities
    np.testing.assert_allclose(ws.detach().cpu().numpy(), ows, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), odepth, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(image.detach().cpu().numpy(), oimage, rtol=2e-5, atol=2e-6)
    gws = torch.rand(ws.shape, device="cuda", generator=g)
    gim = torch.rand(image.shape, device="cuda", generator=g)
    (ws * gws).sum().add((image * gim).sum()).backward()
    ogs, ogc = cpu.composite_rays_train_backward(gws.cpu().numpy(), gim.cpu().numpy(), sigmas.detach().cpu().numpy(),
                                                 rgbs.detach().cpu().numpy(), deltas.cpu().numpy(), rays.cpu().numpy(), ows, oimage)
    np.testing.assert_allclose(rgbs.grad.cpu().numpy(), ogc, rtol=2e-5, atol=2e-6)
    # grad_sigma is a difference of nearly equal running sums: absolute tolerance scaled by delta (~3e-3)
    np.testing.assert_allclose(sigmas.grad.cpu().numpy(), ogs, rtol=1e-3, atol=2e-7)
    if ref_ext:
        rws, rdepth, rimage = ref_glue.composite_rays_train_forward(ref_ext, sigmas.detach(), rgbs.detach(), deltas, rays)
        torch.testing.assert_close(ws.detach(), rws, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(depth.detach(), rdepth, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(image.detach(), rimage, rtol=1e-5, atol=1e-6)
        rgs, rgc = ref_glue.composite_rays_train_backward(ref_ext, gws, gim, sigmas.detach(), rgbs.detach(), deltas, rays, rws, rimage)
        torch.testing.assert_close(rgbs.grad, rgc, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(sigmas.grad, rgs, rtol=1e-3, atol=2e-7)


def test_morton_packbits_exact(scene):
    import raymarching
    from oracle import cpu
    rng = np.random.default_rng(0)
    coords = rng.integers(0, 128, size=(10000, 3)).astype(np.int32)
    ind = raymarching.morton3D(dev(coords))
    assert np.array_equal(ind.cpu().numpy(), cpu.morton3D(coords))
    back = raymarching.morton3D_invert(ind)
    assert np.array_equal(back.cpu().numpy(), coords)
    grid = rng.random((1, 128 ** 3)).astype(np.float32)
    bits = raymarching.packbits(dev(grid), 0.37)
    assert np.array_equal(bits.cpu().numpy(), cpu.packbits(grid, 0.37))
    bits2 = raymarching.packbits(dev(scene["grid"]), 0.01)
    assert np.array_equal(bits2.cpu().numpy(), scene["bitfield"])


@pytest.mark.parametrize("L,desired", [(14, 2048), (16, 2048)])
def test_grid_encode_fp32(ref_ext, L, desired):
    from gridencoder import GridEncoder, grid_encode
    from oracle import cpu, ref_glue
    torch.manual_seed(0)
    enc = GridEncoder(num_levels=L, desired_resolution=desired).cuda()
    enc.embeddings.data.uniform_(-1, 1)
    B = 20000
    x = torch.rand(B, 3, device="cuda")
    x[:7] = torch.tensor([1.5, 0.5, 0.5], device="cuda")  # out-of-range rows -> zero features (gridencoder.cu:99-123)
    x[7] = 0.0
    x[8] = 1.0
    out = grid_encode(x, enc.embeddings, enc.offsets, enc.per_level_scale, enc.base_resolution, False, 0, False)
    assert torch.all(out[:7] == 0)
    from gridencoder.grid import level_table
    scales, res = level_table(enc.offsets, enc.per_level_scale, enc.base_resolution)
    lscales, lres = cpu.grid_level_info(enc.offsets.cpu().numpy(), np.float32(np.log2(enc.per_level_scale)), enc.base_resolution)
    assert np.array_equal(res.cpu().numpy(), lres)                       # integer resolutions agree with libm exp2f
    np.testing.assert_allclose(scales.cpu().numpy(), lscales, rtol=3e-7)  # scales within an ulp or two
    cpu.set_level_scales(scales.cpu().numpy())  # pin the oracle to the device's exp2f: features become bit-exact
    oout, _ = cpu.grid_encode_forward(x.cpu().numpy(), enc.embeddings.detach().cpu().numpy(), enc.offsets.cpu().numpy(),
                                      enc.per_level_scale, enc.base_resolution)
    assert np.array_equal(out.detach().cpu().numpy(), oout)  # same FMA order, same scales: bit-exact
    g = torch.randn(B, L * 2, device="cuda")
    out.backward(g)
    oge, _ = cpu.grid_encode_backward(g.cpu().numpy(), x.cpu().numpy(), tuple(enc.embeddings.shape), enc.offsets.cpu().numpy(),
                                      enc.per_level_scale, enc.base_resolution)
    cpu.set_level_scales(None)
    np.testing.assert_allclose(enc.embeddings.grad.cpu().numpy(), oge, rtol=1e-4, atol=2e-4)  # float atomics: order differs
    if ref_ext:
        rout, _ = ref_glue.grid_encode_forward(ref_ext, x, enc.embeddings.detach(), enc.offsets, enc.per_level_scale,
                                               enc.base_resolution)
        assert torch.equal(rout, out.detach())  # same device exp2f, same FMA order: bit-exact
        rge, _ = ref_glue.grid_encode_backward(ref_ext, g, x, enc.embeddings.detach(), enc.offsets, enc.per_level_scale,
                                               enc.base_resolution)
        torch.testing.assert_close(enc.embeddings.grad, rge, rtol=1e-4, atol=1e-5)  # atomic order differs


def test_grid_encode_fp16_autocast(ref_ext):
    from gridencoder import GridEncoder
    from oracle import ref_glue
    torch.manual_seed(0)
    enc = GridEncoder(num_levels=14, desired_resolution=2048).cuda()
    enc.embeddings.data.uniform_(-1, 1)
    x = torch.rand(8192, 3, device="cuda") * 2 - 1
    with torch.autocast("cuda", dtype=torch.float16):
        out = enc(x, bound=1)
    assert out.dtype == torch.float16
    ref32 = enc(x, bound=1)
    torch.testing.assert_close(out.float(), ref32, rtol=1e-2, atol=2e-3)
    g = torch.randn_like(out)
    out.backward(g)
    assert enc.embeddings.grad.dtype == torch.float32
    if ref_ext:
        xin = (x + 1) / 2
        emb16 = enc.embeddings.detach().half()
        rout, _ = ref_glue.grid_encode_forward(ref_ext, xin, emb16, enc.offsets, enc.per_level_scale, enc.base_resolution)
        torch.testing.assert_close(out.float(), rout.float(), rtol=1e-2, atol=2e-3)
        rge, _ = ref_glue.grid_encode_backward(ref_ext, g, xin, emb16, enc.offsets, enc.per_level_scale, enc.base_resolution)
        # fp16 atomics in both; compare where gradients are not tiny
        torch.testing.assert_close(enc.embeddings.grad, rge.float(), rtol=1e-2, atol=2e-2)


def test_grid_encode_input_grad_and_tiled(ref_ext):
    from gridencoder import GridEncoder
    from oracle import ref_glue
    torch.manual_seed(1)
    for gridtype, D, C in (("tiled", 2, 4), ("hash", 3, 1), ("hash", 3, 8)):
        enc = GridEncoder(input_dim=D, num_levels=6, level_dim=C, base_resolution=8, log2_hashmap_size=12,
                          desired_resolution=256, gridtype=gridtype).cuda()
        enc.embeddings.data.uniform_(-1, 1)
        x = (torch.rand(3000, D, device="cuda") * 2 - 1).requires_grad_(True)
        out = enc(x, bound=1)
        g = torch.randn_like(out)
        out.backward(g)
        assert x.grad is not None and torch.isfinite(x.grad).all()
        if ref_ext:
            xin = ((x.detach() + 1) / 2)
            rout, rj = ref_glue.grid_encode_forward(ref_ext, xin, enc.embeddings.detach(), enc.offsets, enc.per_level_scale,
                                                    enc.base_resolution, True, enc.gridtype_id, False)
            assert torch.equal(rout, out.detach())
            rge, rgi = ref_glue.grid_encode_backward(ref_ext, g, xin, enc.embeddings.detach(), enc.offsets, enc.per_level_scale,
                                                     enc.base_resolution, rj, enc.gridtype_id, False)
            torch.testing.assert_close(enc.embeddings.grad, rge, rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(x.grad, rgi / 2, rtol=1e-4, atol=1e-4)  # d/dx of (x+1)/2


def test_grid_encode_rejects_bad_level_dim():
    from gridencoder import grid_encode
    emb = torch.zeros(64, 3, device="cuda")
    offs = torch.tensor([0, 64], dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError):  # "GridEncoding: C must be 1, 2, 4, or 8." gridencoder.cu:355
        grid_encode(torch.rand(4, 3, device="cuda"), emb, offs, 2.0, 16)


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_encode(ref_ext, degree):
    from shencoder import SHEncoder, sh_encode
    from oracle import ref_glue, sh_reference
    g = torch.Generator(device="cuda").manual_seed(degree)
    d = torch.randn(5000, 3, device="cuda", generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    out = SHEncoder(degree=degree)(d)
    ref = sh_reference.real_sh(d.cpu().numpy(), degree)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=2e-5)
    x = (torch.rand(2000, 3, device="cuda", generator=g) * 2 - 1).requires_grad_(True)  # non-unit: polynomial forms
    o2 = sh_encode(x, degree, True)
    gr = torch.randn(o2.shape, device="cuda", generator=g)
    o2.backward(gr)
    if ref_ext:
        rout, rj = ref_glue.sh_encode_forward(ref_ext, x.detach(), degree, True)
        torch.testing.assert_close(o2.detach(), rout, rtol=1e-5, atol=1e-5)
        gi = torch.einsum("bc,bdc->bd", gr, rj.view(-1, 3, degree ** 2))
        torch.testing.assert_close(x.grad, gi, rtol=1e-4, atol=1e-4)


def test_inference_march_composite_compact(scene):
    """Eval-path kernels vs the oracle on one compaction round."""
    import raymarching
    from oracle import cpu
    ro, rd, bf = _setup(scene, 2)
    N = ro.shape[0]
    nears, fars = raymarching.near_far_from_aabb(ro, rd, dev(AABB), 0.2)
    rays_alive = torch.arange(N, dtype=torch.int32, device="cuda")
    rays_t = nears.clone()
    n_step = 4
    xyzs, dirs, deltas = raymarching.march_rays(N, n_step, rays_alive, rays_t, ro, rd, 1.0, bf, 1, 128, nears, fars, 128, False,
                                                0.0, 1024)
    ox, od, ol = cpu.march_rays(N, n_step, rays_alive.cpu().numpy(), rays_t.cpu().numpy(), ro.cpu().numpy(), rd.cpu().numpy(), 1.0,
                                scene["bitfield"], 1, 128, nears.cpu().numpy(), fars.cpu().numpy())
    assert np.array_equal(xyzs[: N * n_step].cpu().numpy(), ox) and np.array_equal(deltas[: N * n_step].cpu().numpy(), ol)
    g = torch.Generator(device="cuda").manual_seed(5)
    sig = torch.rand(xyzs.shape[0], device="cuda", generator=g) * 200
    rgb = torch.rand(xyzs.shape[0], 3, device="cuda", generator=g)
    ws = torch.zeros(N, device="cuda"); dp = torch.zeros(N, device="cuda"); im = torch.zeros(N, 3, device="cuda")
    ows, odp, oim, ort = (np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32),
                          rays_t.cpu().numpy().copy())
    raymarching.composite_rays(N, n_step, rays_alive, rays_t, sig, rgb, deltas, ws, dp, im)
    cpu.composite_rays(N, n_step, rays_alive.cpu().numpy(), ort, sig.cpu().numpy(), rgb.cpu().numpy(), deltas.cpu().numpy(),
                       ows, odp, oim)
    np.testing.assert_allclose(ws.cpu().numpy(), ows, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(im.cpu().numpy(), oim, rtol=1e-5, atol=1e-6)
    dead = ort < 0
    assert np.array_equal(rays_t.cpu().numpy() < 0, dead) and dead.sum() > 0
    new_alive = torch.zeros_like(rays_alive); new_t = torch.zeros_like(rays_t)
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    raymarching.compact_rays(N, new_alive, rays_alive, new_t, rays_t, cnt)
    oa, ot, oc = cpu.compact_rays(N, rays_alive.cpu().numpy(), rays_t.cpu().numpy())
    assert int(cnt.item()) == oc
    assert np.array_equal(new_alive[:oc].cpu().numpy(), oa[:oc]) and np.array_equal(new_t[:oc].cpu().numpy(), ot[:oc])


def test_march_terminates_on_garbage_rays(scene):
    """Non-finite or degenerate rays must not hang the marcher (the reference's loop would spin on them): a window cap bounds it."""
    import raymarching
    N = 256
    g = torch.Generator().manual_seed(0)
    ro = torch.randn(N, 3, generator=g)
    rd = torch.randn(N, 3, generator=g)
    rd[0] = 0.0
    rd[1] = float("nan")
    ro[2] = float("inf")
    rd[3] = torch.tensor([1e-30, 0.0, 0.0])
    rd[4] = float("inf")
    ro[5] = float("nan")
    ro, rd = ro.cuda(), rd.cuda()
    bf = torch.from_numpy(scene["bitfield"]).cuda()
    nears = torch.full((N,), 0.2, device="cuda")
    fars = torch.full((N,), float("inf"), device="cuda")   # worst case: no far bound at all
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    out = raymarching.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, counter, -1, True, 128, False, 0.0, 1024)
    torch.cuda.synchronize()
    assert out[3].shape == (N, 3) and int(counter[1]) == N
