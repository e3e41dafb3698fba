"""The (teacher, student) distillation loss (distill_mutual/utils.py:954-1189, normL2) on the CPU oracle, and the decomposition the
CUDA pair kernels implement (csrc/pair_loss.cu): sums of squares -> un-normalised composite backward -> coefficients rate / ||.||.
No GPU: torch autograd through the oracle's composite is the reference, `decomposed_pair_grads` restates the kernels' algorithm."""
import numpy as np
import torch

from oracle import cpu, field


def _fake_pair(seed, n_rays=37, max_cnt=70, pad=19):
    g = torch.Generator().manual_seed(seed)
    cnts = torch.randint(0, max_cnt, (n_rays,), generator=g)
    cnts[3] = 0
    total = int(cnts.sum())
    M = total + pad
    offs = torch.cumsum(cnts, 0) - cnts
    rays = torch.stack([torch.arange(n_rays), offs, cnts], 1).to(torch.int32)
    deltas = torch.zeros(M, 2)
    deltas[:total, 0] = 0.003 + 0.002 * torch.rand(total, generator=g)
    deltas[:total, 1] = deltas[:total, 0]
    mk = lambda *s: torch.rand(*s, generator=g)
    stu = dict(sigma=(30 * mk(M)).requires_grad_(True), rgb=mk(M, 3).requires_grad_(True), feat=(mk(M, 16) - 0.5).requires_grad_(True))
    tea = dict(sigma=30 * mk(M), rgb=mk(M, 3), feat=mk(M, 16) - 0.5)
    return rays, deltas, stu, tea, total, M


def decomposed_pair_grads(rays, deltas, stu, tea, rates, total, bg=1.0, loss_scale=1.0, stage=3, reduce_sums=None):
    """What k_pair_sample_sq -> k_pair_composite -> k_pair_combine compute, restated with numpy + the C composite oracle."""
    r_rgb, r_fea, r_col, r_sig = rates
    if stage == 1:
        r_rgb = r_col = r_sig = 0.0
    if stage == 2:
        r_rgb = 0.0
    fs, ft = stu["feat"].detach().numpy(), tea["feat"].numpy()
    cs, ct = stu["rgb"].detach().numpy(), tea["rgb"].numpy()
    ss, st_ = stu["sigma"].detach().numpy(), tea["sigma"].numpy()
    sums = dict(fea=((fs - ft) ** 2).sum(), color=((cs - ct) ** 2).sum(), sigma=((fs[:, 0] - ft[:, 0]) ** 2).sum(), rgb=0.0)
    M = fs.shape[0]
    gs_u, gc_u = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    if stage == 3:
        r, d = rays.numpy(), deltas.numpy()
        ws, _, img = cpu.composite_rays_train_forward(ss, cs, d, r)
        wt, _, imt = cpu.composite_rays_train_forward(st_, ct, d, r)
        diff = (img + (1 - ws)[:, None] * bg) - (imt + (1 - wt)[:, None] * bg)
        sums["rgb"] = (diff ** 2).sum()
        gws = -(diff * bg).sum(1).astype(np.float32)
        gs_u, gc_u = cpu.composite_rays_train_backward(gws, diff.astype(np.float32), ss, cs, d, r, ws, img)
        gs_u[total:] = 0
        gc_u[total:] = 0
    if reduce_sums is not None:   # ray-sharded run: the sums of squares are all-reduced before the coefficients are formed
        sums = reduce_sums(sums)
    coef = {k: (loss_scale * rate / np.sqrt(sums[k]) if sums[k] > 0 and rate else 0.0)
            for k, rate in (("rgb", r_rgb), ("fea", r_fea), ("color", r_col), ("sigma", r_sig))}
    g_sigma = coef["rgb"] * gs_u
    g_rgb = coef["rgb"] * gc_u + coef["color"] * (cs - ct)
    g_feat = coef["fea"] * (fs - ft)
    g_feat[:, 0] += coef["sigma"] * (fs[:, 0] - ft[:, 0])
    loss = sum(rate * np.sqrt(sums[k]) for k, rate in (("rgb", r_rgb), ("fea", r_fea), ("color", r_col), ("sigma", r_sig)))
    return loss, g_sigma, g_rgb, g_feat


def _autograd(rays, deltas, stu, tea, rates, stage, bg=1.0):
    r_rgb, r_fea, r_col, r_sig = rates
    loss = r_fea * torch.norm(stu["feat"] - tea["feat"])
    if stage >= 2:
        loss = loss + r_col * torch.norm(stu["rgb"] - tea["rgb"]) + r_sig * torch.norm(stu["feat"][:, 0] - tea["feat"][:, 0])
    if stage == 3:
        ws, _, im = field.composite(stu["sigma"], stu["rgb"], deltas, rays)
        wt, _, it = field.composite(tea["sigma"], tea["rgb"], deltas, rays)
        loss = loss + r_rgb * torch.norm((it + (1 - wt).unsqueeze(-1) * bg) - (im + (1 - ws).unsqueeze(-1) * bg))
    g = torch.autograd.grad(loss, [stu["sigma"], stu["rgb"], stu["feat"]], allow_unused=True)
    z = lambda t, like: torch.zeros_like(like) if t is None else t
    return float(loss), z(g[0], stu["sigma"]).numpy(), z(g[1], stu["rgb"]).numpy(), z(g[2], stu["feat"]).numpy()


def test_decomposed_pair_gradients_equal_autograd():
    rates = (1.0, 0.002, 0.002, 0.002)
    for stage in (1, 2, 3):
        rays, deltas, stu, tea, total, M = _fake_pair(stage)
        want = _autograd(rays, deltas, stu, tea, rates, stage)
        got = decomposed_pair_grads(rays, deltas, stu, tea, rates, total, stage=stage)
        assert abs(got[0] - want[0]) < 1e-5 * max(1.0, abs(want[0]))
        for a, b, name in zip(got[1:], want[1:], ("sigma", "rgb", "feat")):
            np.testing.assert_allclose(a, b, rtol=2e-4, atol=1e-7, err_msg=f"stage {stage} grad {name}")
        # rows past the last sample get per-sample gradients only (no ray owns them)
        assert np.all(got[1][total:] == 0) and np.any(got[3][total:] != 0)


def test_pair_step_oracle_terms_and_stages():
    """pair_distill_step: stage gating and the four norms, against an independent evaluation on the same march."""
    from pvd_b200 import synthetic as syn
    _, bitfield, _ = syn.lego_bitfield()
    ro, rd = syn.make_ray_batches(1, 256, seed=3)[0]
    offsets, pls = cpu.grid_offsets(3, 4, 16, 12, desired_resolution=64)
    torch.manual_seed(0)
    mk = lambda: (torch.empty(int(offsets[-1]), 2).uniform_(-0.5, 0.5),
                  [torch.randn(o, i) * 0.3 for o, i in ((64, 8), (16, 64), (64, 31), (64, 64), (3, 64))])
    (es, ws_s), (et, ws_t) = mk(), mk()
    es.requires_grad_(True)
    f_s = lambda x, d: field.hash_field_forward(x, d, es, offsets, pls, 16, ws_s)
    f_t = lambda x, d: field.hash_field_forward(x, d, et, offsets, pls, 16, ws_t)
    rates = (1.0, 0.5, 0.25, 0.125)
    o3 = field.pair_distill_step(ro, rd, bitfield, f_s, f_t, rates, stage=3)
    o2 = field.pair_distill_step(ro, rd, bitfield, f_s, f_t, rates, stage=2)
    o1 = field.pair_distill_step(ro, rd, bitfield, f_s, f_t, rates, stage=1)
    assert set(o1["terms"]) == {"fea"} and set(o2["terms"]) == {"fea", "color", "sigma"} and set(o3["terms"]) == {"fea", "color", "sigma", "rgb"}
    assert o3["xyzs"].shape[0] % 128 == 0 and int(o3["counter"][0]) < o3["xyzs"].shape[0]
    t = o3["terms"]
    want = rates[0] * t["rgb"] + rates[1] * t["fea"] + rates[2] * t["color"] + rates[3] * t["sigma"]
    assert abs(float(o3["loss"]) - float(want)) < 1e-6 * float(want)
    assert abs(float(o2["loss"]) - float(want - rates[0] * t["rgb"])) < 1e-5 * float(want)
    assert abs(float(o1["loss"]) - float(rates[1] * t["fea"])) < 1e-6
    # padding rows are evaluated at the origin by both networks and enter the feature norm
    pad = o3["xyzs"][int(o3["counter"][0]):]
    assert pad.numel() > 0 and float(pad.abs().max()) == 0.0
    assert float((o3["feat"][-1] - o3["feat_tea"][-1]).abs().sum()) > 0
    o3["loss"].backward()
    assert es.grad is not None and float(es.grad.abs().sum()) > 0
    # a student equal to its teacher: every term vanishes
    same = field.pair_distill_step(ro, rd, bitfield, f_t, f_t, rates, stage=3)
    assert float(same["loss"]) == 0.0


def test_sharded_pair_gradients_with_global_sums():
    """Two shards of the rays (PairDistillEngine under ray sharding): each shard runs the decomposition on its own rays
    and samples, the four sums of squares are added across the shards before the coefficients are formed (the engine's 1 KB
    all-reduce); the concatenated per-sample gradients then equal those of the single-process step on the whole batch."""
    rates = (1.0, 0.5, 0.25, 0.125)
    rays, deltas, stu, tea, total, M = _fake_pair(7, n_rays=40, pad=1)
    for k in stu:   # the one padding row carries no difference (both networks agree there), so it can be replicated per shard
        stu[k].data[-1] = tea[k][-1]
    want = _autograd(rays, deltas, stu, tea, rates, 3)
    cut_ray = 17
    cut = int(rays[cut_ray, 1])
    shards = []
    for lo_r, hi_r, lo_s, hi_s in ((0, cut_ray, 0, cut), (cut_ray, rays.shape[0], cut, total)):
        r = rays[lo_r:hi_r].clone()
        r[:, 0] -= lo_r   # a shard numbers its own rays and samples from zero
        r[:, 1] -= lo_s
        # the shard's samples + one padding row (a ray whose samples end exactly at M is dropped, raymarching.cu:419)
        rows = torch.cat([torch.arange(lo_s, hi_s), torch.tensor([M - 1])])
        sl = lambda d: {k: (v[rows].detach().clone().requires_grad_(v.requires_grad)) for k, v in d.items()}
        shards.append((r, deltas[rows], sl(stu), sl(tea), hi_s - lo_s))
    local = []
    for r, d, s_, t_, tot in shards:   # first pass: every shard's own sums
        box = {}
        decomposed_pair_grads(r, d, s_, t_, rates, tot, reduce_sums=lambda sm: box.update(sm) or sm)
        local.append(dict(box))
    glob = {k: sum(l[k] for l in local) for k in local[0]}
    parts = [decomposed_pair_grads(r, d, s_, t_, rates, tot, reduce_sums=lambda sm: glob) for r, d, s_, t_, tot in shards]
    assert abs(parts[0][0] - want[0]) < 1e-5 * want[0] and abs(parts[1][0] - want[0]) < 1e-5 * want[0]
    for i, name in ((1, "sigma"), (2, "rgb"), (3, "feat")):
        got = np.concatenate([p_[i][:-1] for p_ in parts], axis=0)   # without the shards' padding rows
        np.testing.assert_allclose(got, want[i][:-1], rtol=2e-4, atol=1e-7, err_msg=name)
