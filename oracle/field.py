"""torch-CPU restatement of the field networks and of one training step.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's field is built from ATen ops (F.linear, relu, sigmoid, exp, clamp, cat: distill_mutual/network.py:413-437)
around its three extensions; ATen on the CPU is its own oracle, the extensions are restated in pvd_oracle.c.  This module
glues them with autograd so that `loss.backward()` yields reference gradients for every parameter:

    hash_field_forward   network.py:335-343,413-437  (GridEncoder -> sigma_net -> clamp -> trunc_exp -> SH -> color_net)
    composite            raymarching.py:292-360
    render_train_step    renderer.py:359-448 + the MSE criterion of just_train_tea/utils.py:841-846
All math in float32 (the reference's autocast path differs from it by fp16 rounding, tolerance 1e-2 in the tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import cpu


class _GridEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x01, emb, offsets, per_level_scale, base_resolution):
        out, _ = cpu.grid_encode_forward(x01.detach().numpy(), emb.detach().numpy(), offsets, per_level_scale, base_resolution)
        ctx.save_for_backward(x01)
        ctx.meta = (tuple(emb.shape), offsets, per_level_scale, base_resolution)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, g):
        (x01,) = ctx.saved_tensors
        shape, offsets, pls, H = ctx.meta
        ge, _ = cpu.grid_encode_backward(g.contiguous().numpy(), x01.numpy(), shape, offsets, pls, H)
        return None, torch.from_numpy(ge), None, None, None


class _TruncExp(torch.autograd.Function):  # tools/activation.py:6-21
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-12, 12))


class _Composite(torch.autograd.Function):  # raymarching.py:292-360
    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays):
        ws, depth, image = cpu.composite_rays_train_forward(sigmas.detach().numpy(), rgbs.detach().numpy(), deltas.numpy(),
                                                            rays.numpy())
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, torch.from_numpy(ws), torch.from_numpy(image))
        return torch.from_numpy(ws), torch.from_numpy(depth), torch.from_numpy(image)

    @staticmethod
    def backward(ctx, gws, gdepth, gimage):
        sigmas, rgbs, deltas, rays, ws, image = ctx.saved_tensors
        gs, gc = cpu.composite_rays_train_backward(gws.contiguous().numpy(), gimage.contiguous().numpy(), sigmas.detach().numpy(),
                                                   rgbs.detach().numpy(), deltas.numpy(), rays.numpy(), ws.numpy(), image.numpy())
        return torch.from_numpy(gs), torch.from_numpy(gc), None, None


def hash_field_forward(x, d, emb, offsets, per_level_scale, base_resolution, weights, bound=1.0, clip_min=-2.0, clip_max=7.0,
                       quantize_fp16=False):
    """x [M,3] in [-bound,bound], d [M,3] -> sigma [M], color [M,3], feat [M,16].  weights = (ws0, ws1, wc0, wc1, wc2).

    quantize_fp16 rounds table, weights and activations to fp16 where the reference's autocast does (grid.py:51-52 and
    torch autocast's F.linear rule), keeping the arithmetic itself in fp32 -- a tight model of the fused kernel."""
    q = (lambda t: t.half().float()) if quantize_fp16 else (lambda t: t)
    ws0, ws1, wc0, wc1, wc2 = weights
    x01 = (x + bound) / (2 * bound)  # grid.py:211
    h = _GridEncode.apply(x01, q(emb), offsets, per_level_scale, base_resolution)
    h = q(F.relu(F.linear(q(h), q(ws0))))
    h = q(F.linear(h, q(ws1)))
    h0 = torch.clamp(h[..., 0], clip_min, clip_max)  # network.py:418-420
    feat = torch.cat([h0.unsqueeze(-1), h[..., 1:]], dim=-1)
    sigma = _TruncExp.apply(h0)
    sh = torch.from_numpy(cpu.sh_encode_forward(d.detach().numpy(), 4))
    c = torch.cat([sh, h[..., 1:]], dim=-1)  # network.py:428-429
    c = q(F.relu(F.linear(q(c), q(wc0))))
    c = q(F.relu(F.linear(c, q(wc1))))
    color = torch.sigmoid(F.linear(c, q(wc2)))
    return sigma, color, feat


def composite(sigmas, rgbs, deltas, rays):
    return _Composite.apply(sigmas, rgbs, deltas, rays)


def march_samples(rays_o, rays_d, bitfield, bound=1.0, cascade=1, grid_size=128, min_near=0.2, M=None, perturb=True, dt_gamma=0.0,
                  max_steps=1024, aabb=None):
    """near/far + march_rays_train on the CPU (renderer.py:342-391): torch tensors xyzs, dirs, deltas, rays + counter, nears, fars.
    M=None is the warm-up sizing: the total rounded up strictly to 128 (raymarching.py:276-282); padding rows are zeros."""
    ro, rd = rays_o.numpy(), rays_d.numpy()
    if aabb is None:
        aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = cpu.near_far_from_aabb(ro, rd, aabb, min_near)
    xyzs, dirs, deltas, rays, counter = cpu.march_rays_train(ro, rd, bound, bitfield, cascade, grid_size, nears, fars, M=M,
                                                            perturb=perturb, dt_gamma=dt_gamma, max_steps=max_steps)
    if M is None:
        m = int(counter[0])
        m += 128 - m % 128
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    txyz, tdir, tdl, trays = (torch.from_numpy(np.ascontiguousarray(a)) for a in (xyzs, dirs, deltas, rays))
    return txyz, tdir, tdl, trays, counter, nears, fars


def render_train_step(rays_o, rays_d, bitfield, gt_rgb, field_fn, bound=1.0, cascade=1, grid_size=128, min_near=0.2, bg_color=1.0,
                      density_scale=1.0, M=None, perturb=True, dt_gamma=0.0, max_steps=1024, aabb=None):
    """One reference training step on the CPU: near/far -> march -> field -> composite -> bg mix -> MSE.

    Returns dict(loss, image, depth, weights_sum, xyzs, dirs, deltas, rays, counter).  `field_fn(xyzs, dirs) -> (sigma, rgb)`.
    """
    txyz, tdir, tdl, trays, counter, nears, fars = march_samples(rays_o, rays_d, bitfield, bound, cascade, grid_size, min_near, M,
                                                                 perturb, dt_gamma, max_steps, aabb)
    sigma, rgb = field_fn(txyz, tdir)
    ws, depth, image = composite(density_scale * sigma, rgb, tdl, trays)
    pred = image + (1 - ws).unsqueeze(-1) * bg_color  # renderer.py:445
    depth_n = torch.clamp(depth - torch.from_numpy(nears), min=0) / (torch.from_numpy(fars - nears) + 1e-6)  # renderer.py:446
    loss = torch.mean((pred - gt_rgb) ** 2)
    return dict(loss=loss, image=pred, depth=depth_n, weights_sum=ws, xyzs=txyz, dirs=tdir, deltas=tdl, rays=trays,
                counter=counter, sigma=sigma, rgb=rgb)


def render_uniform(rays_o, rays_d, field_fn, aabb, num_steps=64, min_near=0.2, bg_color=1.0, density_scale=1.0, noise=None):
    """`NeRFRenderer.run` without upsampling (distill_mutual/renderer.py:139-317, upsample_steps=0) -- the pure-PyTorch renderer of
    BASELINE configs[0] (mlp model, 256 rays x 64 uniform samples, CPU).  The reference's own `run` ends in `self.color(...)`, which
    asserts False (network.py:516); as SURVEY.md 8c prescribes, `field_fn(xyzs, dirs) -> (sigma, rgb)` (= `forward`) stands in for
    the dead density / color split, everything else is restated line by line:
        z = near + (far - near) * linspace(0, 1, T)  [+ (noise - 0.5) * sample_dist]          :168-182
        xyz = clip(o + d z, aabb)                                                               :185-188
        delta_i = z_{i+1} - z_i, last = sample_dist ; alpha = 1 - exp(-delta * density_scale * sigma)   :262-268
        w = alpha * cumprod([1, 1 - alpha + 1e-15])[:-1] ; image = sum w rgb + (1 - sum w) * bg ; depth = sum w * clamp((z - near) / (far - near))  :269-311
    Returns dict(image [N,3], depth [N], weights_sum [N], weights [N,T], z_vals, xyzs [N*T,3], dirs, sigma, rgb)."""
    N = rays_o.shape[0]
    nears, fars = cpu.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), np.asarray(aabb, np.float32), min_near)
    nears, fars = torch.from_numpy(nears).unsqueeze(-1), torch.from_numpy(fars).unsqueeze(-1)
    z = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand(N, num_steps)
    z = nears + (fars - nears) * z
    sample_dist = (fars - nears) / num_steps
    if noise is not None:
        z = z + (noise - 0.5) * sample_dist
    aabb_t = torch.as_tensor(aabb, dtype=torch.float32)
    xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z.unsqueeze(-1)
    xyzs = torch.min(torch.max(xyzs, aabb_t[:3]), aabb_t[3:])
    dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
    sigma, rgb = field_fn(xyzs.reshape(-1, 3), dirs.reshape(-1, 3))
    sigma, rgb = sigma.view(N, num_steps), rgb.view(N, num_steps, 3)
    deltas = torch.cat([z[..., 1:] - z[..., :-1], sample_dist * torch.ones_like(z[..., :1])], dim=-1)
    alphas = 1 - torch.exp(-deltas * density_scale * sigma)
    shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
    weights = alphas * torch.cumprod(shifted, dim=-1)[..., :-1]
    ws = weights.sum(dim=-1)
    depth = torch.sum(weights * ((z - nears) / (fars - nears)).clamp(0, 1), dim=-1)
    image = torch.sum(weights.unsqueeze(-1) * rgb, dim=-2) + (1 - ws).unsqueeze(-1) * bg_color
    return dict(image=image, depth=depth, weights_sum=ws, weights=weights, z_vals=z, deltas=deltas, xyzs=xyzs.reshape(-1, 3),
                dirs=dirs.reshape(-1, 3), sigma=sigma, rgb=rgb)


def pair_distill_step(rays_o, rays_d, bitfield, student_fn, teacher_fn, rates=(1.0, 0.002, 0.002, 0.002), stage=3, l1_reg=None,
                      bg_color=1.0, density_scale=1.0, M=None, **march_kw):
    """One distillation step of `Trainer.train_step` (distill_mutual/utils.py:954-1189, loss_type normL2) on the CPU.

    The student marches, the teacher is evaluated under no_grad at the SAME samples (renderer.py:374-394), both see all M rows
    (padding rows are zeros).  student_fn / teacher_fn: (xyzs, dirs) -> (sigma [M], rgb [M,3], feat [M,16]).
    rates = (rgb, fea_sc, color, sigma) (main_distill_mutual.py:174-177); stage 1 = feature term only (:1046-1060), 2 = + colour and
    sigma without compositing (:1061-1108), 3 = all four (:1110-1176); l1_reg = weight * density_loss() of a vm student, a scalar
    tensor (:1135-1136, stage 3 only).  Returns dict(loss, terms, image, image_tea, rays, xyzs, ...); loss.backward() gives the
    student's reference gradients."""
    txyz, tdir, tdl, trays, counter, nears, fars = march_samples(rays_o, rays_d, bitfield, M=M, **march_kw)
    s_s, c_s, f_s = student_fn(txyz, tdir)
    with torch.no_grad():
        s_t, c_t, f_t = teacher_fn(txyz, tdir)
    r_rgb, r_fea, r_col, r_sig = rates
    terms = {"fea": torch.norm(f_s - f_t)}
    out = dict(xyzs=txyz, dirs=tdir, deltas=tdl, rays=trays, counter=counter, feat=f_s, feat_tea=f_t)
    if stage == 1:
        out.update(loss=r_fea * terms["fea"], terms=terms)
        return out
    terms["color"] = torch.norm(c_s - c_t)
    terms["sigma"] = torch.norm(f_s[:, 0] - f_t[:, 0])
    loss = r_fea * terms["fea"] + r_col * terms["color"] + r_sig * terms["sigma"]
    if stage == 2:
        out.update(loss=loss, terms=terms)
        return out
    ws, depth, image = composite(density_scale * s_s, c_s, tdl, trays)
    pred = image + (1 - ws).unsqueeze(-1) * bg_color
    with torch.no_grad():
        wt, _, it = composite(density_scale * s_t, c_t, tdl, trays)
        pred_t = it + (1 - wt).unsqueeze(-1) * bg_color
    terms["rgb"] = torch.norm(pred_t - pred)
    loss = loss + r_rgb * terms["rgb"]
    if l1_reg is not None:
        loss = loss + l1_reg
    out.update(loss=loss, terms=terms, image=pred, image_tea=pred_t, weights_sum=ws)
    return out


def vm_density_loss(sigma_mat, sigma_vec):
    """NeRFNetwork.density_loss (network.py:549-557): sum of mean |.| over the sigma planes and lines."""
    loss = 0
    for m, v in zip(sigma_mat, sigma_vec):
        loss = loss + torch.mean(torch.abs(m)) + torch.mean(torch.abs(v))
    return loss


# ------------------------------------------------------------------------------------------------ VM (TensoRF) field
MAT_IDS = [[0, 1], [0, 2], [1, 2]]  # network.py:76
VEC_IDS = [2, 1, 0]                 # network.py:77


def vm_plane_line(x, mats, vecs):
    """Per-component plane and line samples: [3, R, N] each (F.grid_sample, align_corners=True, zero padding;
    network.py:222-260 for sigma, :269-300 for colour)."""
    N = x.shape[0]
    mat_coord = torch.stack([x[..., MAT_IDS[0]], x[..., MAT_IDS[1]], x[..., MAT_IDS[2]]]).detach().view(3, -1, 1, 2)
    vec_coord = torch.stack([x[..., VEC_IDS[0]], x[..., VEC_IDS[1]], x[..., VEC_IDS[2]]])
    vec_coord = torch.stack((torch.zeros_like(vec_coord), vec_coord), dim=-1).detach().view(3, -1, 1, 2)
    pm, pv = [], []
    for i in range(3):
        pm.append(F.grid_sample(mats[i], mat_coord[[i]], align_corners=True).view(-1, N))
        pv.append(F.grid_sample(vecs[i], vec_coord[[i]], align_corners=True).view(-1, N))
    return pm, pv


def vm_field_forward(x, d, sigma_mat, sigma_vec, color_mat, color_vec, basis_w, color_ws, aabb, clip_min=-2.0, clip_max=7.0,
                     quantize_fp16=False):
    """`NeRFNetwork.forward` for model_type "vm" (network.py:344-382).  Returns sigma [N], color [N,3], feat [N,16].
    sigma_mat/color_mat: 3 x [1,R,H,W]; sigma_vec/color_vec: 3 x [1,R,D,1]; basis_w [15,144]; color_ws = (wc0, wc1, wc2)."""
    q = (lambda t: t.half().float()) if quantize_fp16 else (lambda t: t)
    xn = 2 * (x - aabb[:3]) / (aabb[3:] - aabb[:3]) - 1
    sm, sv = vm_plane_line(xn, sigma_mat, sigma_vec)
    sigma_feat = torch.zeros(x.shape[0])
    for i in range(3):
        sigma_feat = sigma_feat + torch.sum(sm[i] * sv[i], dim=0)
    cm, cv = vm_plane_line(xn, color_mat, color_vec)
    app = (torch.cat(cm, dim=0) * torch.cat(cv, dim=0)).T  # [N, 144]
    color_feat = q(F.linear(q(app), q(basis_w)))
    sigma_feat = torch.clamp(sigma_feat, clip_min, clip_max)
    color_feat = torch.clamp(color_feat, clip_min, clip_max)
    feat = torch.cat([sigma_feat.unsqueeze(-1), color_feat], dim=-1)
    sigma = _TruncExp.apply(sigma_feat)
    sh = torch.from_numpy(cpu.sh_encode_forward(d.detach().numpy(), 4))
    wc0, wc1, wc2 = color_ws
    h = torch.cat([sh, color_feat], dim=-1)
    h = q(F.relu(F.linear(q(h), q(wc0))))
    h = q(F.relu(F.linear(h, q(wc1))))
    color = torch.sigmoid(F.linear(h, q(wc2)))
    return sigma, color, feat


# ------------------------------------------------------------------------------------------------ NeRF MLP field
def mlp_field_forward(x, d, nerf_w, nerf_b, tail_ws, PE=10, skips=3, clip_min=-2.0, clip_max=7.0, quantize_fp16=False):
    """`NeRFNetwork.forward` for model_type "mlp": FreqEncoder (tools/encoding.py:36-49) -> nerf_mlp with a skip concat
    (network.py:324-333) -> the sigma_net / color_net tail (network.py:413-437).  nerf_w / nerf_b: 8 weights / biases."""
    q = (lambda t: t.half().float()) if quantize_fp16 else (lambda t: t)
    parts = [x]
    for k in range(PE):
        f = 2.0 ** k
        parts += [torch.sin(x * f), torch.cos(x * f)]
    h = torch.cat(parts, dim=-1)
    in_pts = h
    n = len(nerf_w)
    for i in range(n):
        h = F.linear(q(h), q(nerf_w[i])) + nerf_b[i]   # fp16 operands, fp32 accumulate, fp32 bias (as the fused kernel)
        if i != n - 1:
            h = F.relu(h)
        if i == skips:
            h = torch.cat([in_pts, h], dim=-1)
    ws0, ws1, wc0, wc1, wc2 = tail_ws
    h = q(F.relu(F.linear(q(h), q(ws0))))
    h = q(F.linear(h, q(ws1)))
    h0 = torch.clamp(h[..., 0], clip_min, clip_max)
    feat = torch.cat([h0.unsqueeze(-1), h[..., 1:]], dim=-1)
    sigma = torch.exp(h0)
    sh = torch.from_numpy(cpu.sh_encode_forward(d.detach().numpy(), 4))
    c = torch.cat([sh, h[..., 1:]], dim=-1)
    c = q(F.relu(F.linear(q(c), q(wc0))))
    c = q(F.relu(F.linear(c, q(wc1))))
    return sigma, torch.sigmoid(F.linear(c, q(wc2))), feat


# ------------------------------------------------------------------------------------------------ tensors (Plenoxels-style) field
def tensors_field_forward(x, d, volume, degree, aabb, clip_min=-2.0, clip_max=7.0):
    """`NeRFNetwork.forward` for model_type "tensors" (network.py:311-322 compute_plenoxel_fea, :383-409): trilinear
    F.grid_sample (align_corners=True) of volume [1, C, D, H, W], C = 3 degree^2 + 1; sigma = trunc_exp(clamp(h0));
    rgb_k = sigmoid(sum_j h[1 + k degree^2 + j] SH_j(d)).  Returns sigma [N], color [N, 3]."""
    xn = 2 * (x - aabb[:3]) / (aabb[3:] - aabb[:3]) - 1
    h = F.grid_sample(volume, xn.view(1, 1, -1, 1, 3), align_corners=True).view(-1, x.shape[0]).permute(1, 0)   # [N, C]
    h0 = torch.clamp(h[..., 0], clip_min, clip_max)
    sigma = _TruncExp.apply(h0)
    sh = h[..., 1:].view(-1, 3, degree ** 2)
    enc = torch.from_numpy(cpu.sh_encode_forward(d.detach().numpy(), degree)).unsqueeze(1)
    color = torch.sigmoid((sh * enc).sum(-1))
    return sigma, color


def get_rays(poses, intrinsics, H, W, inds=None):
    """get_rays (distill_mutual/utils.py:324-404) for given pixel indices inds [B, N] (None: all pixels): rays_o, rays_d [B, N, 3]."""
    fx, fy, cx, cy = intrinsics
    B = poses.shape[0]
    if inds is None:
        inds = torch.arange(H * W).expand([B, H * W])
    i = (inds % W).float() + 0.5
    j = torch.div(inds, W, rounding_mode="floor").float() + 0.5
    zs = torch.ones_like(i)
    directions = torch.stack(((i - cx) / fx * zs, (j - cy) / fy * zs, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions @ poses[:, :3, :3].transpose(-1, -2)
    rays_o = poses[..., :3, 3][..., None, :].expand_as(rays_d)
    return rays_o, rays_d
