"""Properties of the CPU oracle itself (no GPU): these pin the restatement to invariants of the reference algorithm
that hold regardless of hardware -- Morton round trip, PCG32 against an independent big-integer implementation,
sample-count bookkeeping, gradients against finite differences, SH against an independent construction."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import cpu, sh_reference

AABB = np.array([-1, -1, -1, 1, 1, 1], np.float32)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 1023), st.integers(0, 1023), st.integers(0, 1023)), min_size=1, max_size=64))
def test_morton_round_trip(coords):
    c = np.array(coords, np.int32)
    ind = cpu.morton3D(c)
    assert np.array_equal(cpu.morton3D_invert(ind), c)
    # bit interleave definition
    x, y, z = (int(v) for v in c[0])
    want = 0
    for b in range(10):
        want |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
    assert int(ind[0]) == want


def _pcg32_python(seed, n):
    """PCG32 XSH-RR with big integers (O'Neill's definition), independent of the C code."""
    mask = (1 << 64) - 1
    mult, inc = 0x5851F42D4C957F2D, (1 << 1) | 1
    out = []
    for i in range(n):
        state = 0
        state = (state * mult + inc) & mask
        state = (state + seed) & mask
        state = (state * mult + inc) & mask
        for _ in range(i):  # advance(i) the slow way
            state = (state * mult + inc) & mask
        old = state
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        r = ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF
        bits = (r >> 9) | 0x3F800000
        out.append(np.frombuffer(np.uint32(bits).tobytes(), np.float32)[0] - np.float32(1.0))
    return np.array(out, np.float32)


def test_pcg32_matches_independent_implementation():
    got = cpu.pcg32_jitter(42, 200)
    assert np.array_equal(got, _pcg32_python(42, 200))
    assert (got >= 0).all() and (got < 1).all()


def test_packbits_rule():
    rng = np.random.default_rng(0)
    g = rng.random(8 * 1000).astype(np.float32)
    bits = cpu.packbits(g, 0.5)
    assert np.array_equal(np.unpackbits(bits[:, None], axis=1, bitorder="little").ravel(), (g > 0.5).astype(np.uint8))


def test_near_far_against_closed_form():
    o = np.array([[0, 0, -3.0], [0, 0, -3.0], [5, 5, 5], [0, 0, 0]], np.float32)
    d = np.array([[0, 0, 1.0], [0, 1, 0], [1, 0, 0], [0.6, 0, 0.8]], np.float32)
    n, f = cpu.near_far_from_aabb(o, d, AABB, 0.2)
    assert n[0] == 2.0 and f[0] == 4.0
    assert n[1] == np.finfo(np.float32).max and f[2] == np.finfo(np.float32).max  # misses
    assert n[3] == np.float32(0.2) and abs(f[3] - 1.25) < 1e-6  # starts inside: clamped to min_near


def test_march_bookkeeping(scene):
    ro, rd = scene["batches"][0]
    ro, rd = ro.numpy()[:1024], rd.numpy()[:1024]
    n, f = cpu.near_far_from_aabb(ro, rd, AABB, 0.2)
    xyzs, dirs, deltas, rays, counter = cpu.march_rays_train(ro, rd, 1.0, scene["bitfield"], 1, 128, n, f, perturb=True)
    assert counter[1] == 1024 and counter[0] == rays[:, 2].sum() > 1000
    assert np.array_equal(rays[:, 0], np.arange(1024))
    assert np.array_equal(rays[:, 1], np.cumsum(rays[:, 2]) - rays[:, 2])
    S = int(counter[0])
    assert (deltas[:S, 0] > 0).all() and (deltas[S:] == 0).all()
    assert (np.abs(xyzs[:S]) <= 1).all()
    dt_min = np.float32(2 * 1.7320508075688772) / np.float32(1024)
    assert np.allclose(deltas[:S, 0], dt_min)
    # every sample lies in an occupied cell of the bitfield (the marcher only emits occupied cells)
    from pvd_b200 import synthetic as syn
    cell = np.clip((0.5 * (xyzs[:S] + 1) * 128).astype(np.int64), 0, 127)
    m = syn.morton3d_np(cell[:, 0], cell[:, 1], cell[:, 2]).astype(np.int64)
    occ = (scene["bitfield"][m // 8] >> (m % 8)) & 1
    assert occ.all()
    # samples of a ray are on the ray and ordered
    k = int(np.argmax(rays[:, 2]))
    off, cnt = rays[k, 1], rays[k, 2]
    t = ((xyzs[off:off + cnt] - ro[k]) @ rd[k]) / (rd[k] @ rd[k])
    assert (np.diff(t) > 0).all()
    # no perturbation: first sample of any ray starts at >= near
    x2, _, dl2, r2, c2 = cpu.march_rays_train(ro, rd, 1.0, scene["bitfield"], 1, 128, n, f, perturb=False)
    assert c2[0] > 1000 and abs(int(c2[0]) - S) < 0.05 * S


def test_march_overflow_rule(scene):
    ro, rd = scene["batches"][0]
    ro, rd = ro.numpy()[:1024], rd.numpy()[:1024]
    n, f = cpu.near_far_from_aabb(ro, rd, AABB, 0.2)
    M = 4096
    xyzs, dirs, deltas, rays, counter = cpu.march_rays_train(ro, rd, 1.0, scene["bitfield"], 1, 128, n, f, M=M, perturb=True)
    kept = (rays[:, 2] > 0) & (rays[:, 1] + rays[:, 2] < M)  # drop rule is >= M (raymarching.cu:419)
    assert kept.sum() > 0 and (~kept & (rays[:, 2] > 0)).sum() > 0
    last = (rays[kept, 1] + rays[kept, 2]).max()
    assert (deltas[last:] == 0).all()


def test_composite_matches_closed_form_and_finite_differences():
    rng = np.random.default_rng(1)
    rays = np.array([[0, 0, 5], [1, 5, 0], [2, 5, 7]], np.int32)
    M = 16
    sig = (rng.random(M) * 20).astype(np.float32)
    rgb = rng.random((M, 3)).astype(np.float32)
    dl = np.stack([np.full(M, 0.01), np.full(M, 0.01)], 1).astype(np.float32)
    ws, depth, img = cpu.composite_rays_train_forward(sig, rgb, dl, rays)
    a = 1 - np.exp(-sig[:5].astype(np.float64) * 0.01)
    T = np.concatenate([[1], np.cumprod(1 - a)[:-1]])
    assert np.allclose(ws[0], (a * T).sum(), rtol=1e-5) and ws[1] == 0
    assert np.allclose(img[0], ((a * T)[:, None] * rgb[:5]).sum(0), rtol=1e-5)
    assert np.allclose(depth[0], (a * T * 0.01 * np.arange(1, 6)).sum(), rtol=1e-5)
    gws = rng.random(3).astype(np.float32)
    gim = rng.random((3, 3)).astype(np.float32)
    gs, gc = cpu.composite_rays_train_backward(gws, gim, sig, rgb, dl, rays, ws, img)

    def loss(s, c):
        w, _, i = cpu.composite_rays_train_forward(s, c, dl, rays)
        return float((w.astype(np.float64) * gws).sum() + (i.astype(np.float64) * gim).sum())

    for k in (0, 3, 6, 11):
        e = np.zeros(M, np.float32); e[k] = 1e-2
        fd = (loss(sig + e, rgb) - loss(sig - e, rgb)) / 2e-2
        assert abs(fd - gs[k]) < 2e-3 * max(1.0, abs(fd)), (k, fd, gs[k])
    assert gs[15] == 0 and (gc[12:] == 0).all()  # padding rows untouched


@pytest.mark.parametrize("gridtype,D,C", [(0, 3, 2), (1, 2, 4), (0, 3, 1)])
def test_grid_encoder_gradients_by_finite_differences(gridtype, D, C):
    rng = np.random.default_rng(2)
    offsets, pls = cpu.grid_offsets(D, 5, 4, 8, desired_resolution=64)
    emb = rng.uniform(-1, 1, size=(int(offsets[-1]), C)).astype(np.float32)
    x = rng.random((64, D)).astype(np.float32)
    out, dy_dx = cpu.grid_encode_forward(x, emb, offsets, pls, 4, True, gridtype, False)
    g = rng.standard_normal(out.shape).astype(np.float32)
    ge, gi = cpu.grid_encode_backward(g, x, emb.shape, offsets, pls, 4, dy_dx, gridtype, False)
    # linear in the table: <out, g> == <emb, ge>
    assert np.isclose((out.astype(np.float64) * g).sum(), (emb.astype(np.float64) * ge).sum(), rtol=1e-4)
    # input gradient by central differences (piecewise multilinear -> exact away from cell borders)
    eps = 2e-4
    for d in range(D):
        xp, xm = x.copy(), x.copy()
        xp[:, d] += eps; xm[:, d] -= eps
        fp, _ = cpu.grid_encode_forward(xp, emb, offsets, pls, 4, False, gridtype, False)
        fm, _ = cpu.grid_encode_forward(xm, emb, offsets, pls, 4, False, gridtype, False)
        fd = (((fp - fm) / (2 * eps)).astype(np.float64) * g).sum(1)
        ok = np.abs(fd - gi[:, d]) < 0.05 * (np.abs(fd) + 1)
        assert ok.mean() > 0.8  # a few points straddle a cell border within +-eps
    # out-of-range inputs produce zeros (gridencoder.cu:99-123)
    xo = x.copy(); xo[0, 0] = 1.2
    oo, _ = cpu.grid_encode_forward(xo, emb, offsets, pls, 4, False, gridtype, False)
    assert (oo[0] == 0).all() and np.array_equal(oo[1:], out[1:])


def test_grid_level_table_of_the_reference_configs():
    """Hash level table of SURVEY 8a: L=14 (PVD) and L=16 (BASELINE text)."""
    off14, pls14 = cpu.grid_offsets(3, 14, 16, 19, desired_resolution=2048)
    assert off14[-1] == 5303704 and list(np.diff(off14)[:5]) == [4920, 15632, 42880, 132656, 389024]
    assert (np.diff(off14)[5:] == 524288).all()
    sc, res = cpu.grid_level_info(off14, np.float32(np.log2(pls14)), 16)
    assert list(res[:6]) == [16, 24, 34, 50, 72, 104] and res[13] == 2049
    off16, pls16 = cpu.grid_offsets(3, 16, 16, 19, desired_resolution=2048)
    assert off16[-1] == 6119864 and list(np.diff(off16)[:5]) == [4920, 13824, 32768, 85184, 216000]


@pytest.mark.parametrize("degree", [1, 2, 3, 4])
def test_sh_against_independent_construction(degree):
    rng = np.random.default_rng(degree)
    d = rng.standard_normal((500, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    got = cpu.sh_encode_forward(d.astype(np.float32), degree)
    want = sh_reference.real_sh(d.astype(np.float32), degree)
    assert np.abs(got - want).max() < 2e-6
    # prefix property (shencoder.cu:50-122): lower-degree outputs are a prefix of higher-degree ones
    if degree > 1:
        assert np.array_equal(got[:, :(degree - 1) ** 2], cpu.sh_encode_forward(d.astype(np.float32), degree - 1))
