"""The reference's training step, driven through ITS OWN compiled CUDA extensions (oracle/_ref).  TEST / BENCH INFRASTRUCTURE ONLY.

`/root/reference` cannot travel to the GPU box and its sources must not be copied, so the thin Python layer the reference
wraps around its native kernels is restated here, each piece citing the lines it follows:

    _RefGridEncode     gridencoder/grid.py:20-139        (half cast of the whole table per call, [L,B,C] + permute, zeros_like)
    _RefSH             shencoder/sphere_harmonics.py:15-64
    _RefTruncExp       tools/activation.py:6-21
    _RefComposite      raymarching/raymarching.py:292-360
    ref_march          raymarching/raymarching.py:176-289  (zero-filled M-row buffers, mean_count sizing)
    RefHashNetwork     distill_mutual/network.py:103-152,335-343,413-437  (hash branch: nn.Linear stacks under autocast)
    RefVmNetwork       distill_mutual/network.py:72-90,193-309,344-382    (vm branch: 12 F.grid_sample on NCHW planes / lines)
    RefTrainer.step    distill_mutual/renderer.py:359-448 + just_train_tea/utils.py:588-606,841-846 (autocast, MSE, GradScaler)
    RefPairTrainer.step distill_mutual/utils.py:954-1189 (student renders and marches, the frozen teacher renders on the student's
                       samples under no_grad, normL2 losses on rgb / feature_sigma_color / color_l / sigma_l, vm L1 penalty)

Everything numerical (ray marching, hash gather/scatter, SH, compositing) runs in the reference's kernels, the GEMMs in
cuBLAS through F.linear, exactly as in the reference.  This is the "reference extensions rebuilt for sm_100a" baseline of
BASELINE.md section 2a.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.amp import custom_bwd, custom_fwd

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_ext():
    d = os.path.join(_HERE, "_ref")
    if d not in sys.path:
        sys.path.insert(0, d)
    import _gridencoder
    import _raymarching
    import _shencoder
    return {"raymarching": _raymarching, "gridencoder": _gridencoder, "shencoder": _shencoder}


def make_ops(ext):
    rm, ge, sh = ext["raymarching"], ext["gridencoder"], ext["shencoder"]

    class _RefGridEncode(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda")
        def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution):
            inputs = inputs.contiguous()
            B, D = inputs.shape
            L, C = offsets.shape[0] - 1, embeddings.shape[1]
            S = np.log2(per_level_scale)
            if torch.is_autocast_enabled("cuda") and C % 2 == 0:
                embeddings = embeddings.to(torch.half)  # grid.py:51-52: whole table, every forward
            outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
            dy_dx = torch.empty(1, device=inputs.device, dtype=embeddings.dtype)
            ge.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, base_resolution, False, dy_dx, 0, False)
            outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
            ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
            ctx.dims = [B, D, C, L, S, base_resolution]
            return outputs

        @staticmethod
        @custom_bwd(device_type="cuda")
        def backward(ctx, grad):
            inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
            B, D, C, L, S, H = ctx.dims
            grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
            grad_embeddings = torch.zeros_like(embeddings)  # grid.py:106
            grad_inputs = torch.zeros(1, device=inputs.device, dtype=embeddings.dtype)
            ge.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, False, dy_dx, grad_inputs,
                                    0, False)
            return None, grad_embeddings, None, None, None

    class _RefSH(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
        def forward(ctx, inputs, degree):
            inputs = inputs.contiguous()
            B, D = inputs.shape
            outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
            dy_dx = torch.empty(1, dtype=inputs.dtype, device=inputs.device)
            sh.sh_encode_forward(inputs, outputs, B, D, degree, False, dy_dx)
            return outputs

        @staticmethod
        def backward(ctx, grad):
            return None, None

    class _RefTruncExp(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
        def forward(ctx, x):
            ctx.save_for_backward(x)
            return torch.exp(x)

        @staticmethod
        @custom_bwd(device_type="cuda")
        def backward(ctx, g):
            return g * torch.exp(ctx.saved_tensors[0].clamp(-12, 12))

    class _RefComposite(torch.autograd.Function):
        @staticmethod
        @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
        def forward(ctx, sigmas, rgbs, deltas, rays):
            sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
            M, N = sigmas.shape[0], rays.shape[0]
            ws = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
            depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
            image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
            rm.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, ws, depth, image)
            ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, depth, image)
            ctx.dims = [M, N]
            return ws, depth, image

        @staticmethod
        @custom_bwd(device_type="cuda")
        def backward(ctx, gws, gdepth, gimage):
            gws, gimage = gws.contiguous(), gimage.contiguous()
            sigmas, rgbs, deltas, rays, ws, depth, image = ctx.saved_tensors
            M, N = ctx.dims
            gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
            rm.composite_rays_train_backward(gws, gimage, sigmas, rgbs, deltas, rays, ws, image, M, N, gs, gc)
            return gs, gc, None, None

    return _RefGridEncode, _RefSH, _RefTruncExp, _RefComposite


class RefHashNetwork(nn.Module):
    """The hash branch of the reference's NeRFNetwork (network.py:47-51,103-152,335-343,413-437)."""

    def __init__(self, ext, offsets, per_level_scale, base_resolution=16, bound=1.0, clip_min=-2.0, clip_max=7.0):
        super().__init__()
        self.ops = make_ops(ext)
        self.register_buffer("offsets", torch.as_tensor(offsets, dtype=torch.int32))
        self.per_level_scale, self.base_resolution, self.bound = per_level_scale, base_resolution, bound
        self.clip_min, self.clip_max = clip_min, clip_max
        L = len(offsets) - 1
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), 2).uniform_(-1e-4, 1e-4))
        self.sigma_net = nn.ModuleList([nn.Linear(2 * L, 64, bias=False), nn.Linear(64, 16, bias=False)])
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 3, bias=False)])

    def forward(self, x, d):
        GridEncode, SH, TruncExp, _ = self.ops
        x = (x + self.bound) / (2 * self.bound)
        h = GridEncode.apply(x.view(-1, 3), self.embeddings, self.offsets, self.per_level_scale, self.base_resolution)
        for l in range(2):
            h = self.sigma_net[l](h)
            if l != 1:
                h = F.relu(h, inplace=True)
        h[..., 0] = torch.clamp(h[..., 0].clone(), self.clip_min, self.clip_max)
        self.feature_sigma_color = h           # network.py:421
        self.sigma_l = h[..., 0]               # network.py:424
        sigma = TruncExp.apply(h[..., 0])
        geo_feat = h[..., 1:]
        d = SH.apply(d.reshape(-1, 3), 4)
        h = torch.cat([d, geo_feat], dim=-1)
        for l in range(3):
            h = self.color_net[l](h)
            if l != 2:
                h = F.relu(h, inplace=True)
        self.color_l = torch.sigmoid(h)        # network.py:436
        return sigma, self.color_l


class RefVmNetwork(nn.Module):
    """The vm branch of the reference's NeRFNetwork (network.py:72-90 sizes, :193-214 init, :216-309 features, :344-382 forward)."""
    MAT_IDS = [[0, 1], [0, 2], [1, 2]]
    VEC_IDS = [2, 1, 0]

    def __init__(self, ext, resolution=300, bound=1.0, clip_min=-2.0, clip_max=7.0, scale=0.1):
        super().__init__()
        self.ops = make_ops(ext)
        self.clip_min, self.clip_max = clip_min, clip_max
        self.register_buffer("aabb_train", torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32))
        res = [resolution] * 3

        def init(rank):
            mats = [nn.Parameter(scale * torch.randn(1, rank, res[m1], res[m0])) for m0, m1 in self.MAT_IDS]
            vecs = [nn.Parameter(scale * torch.randn(1, rank, res[v], 1)) for v in self.VEC_IDS]
            return nn.ParameterList(mats), nn.ParameterList(vecs)

        self.sigma_mat, self.sigma_vec = init(16)
        self.color_mat, self.color_vec = init(48)
        self.basis_mat = nn.Linear(144, 15, bias=False)
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 3, bias=False)])

    def _coords(self, x):
        mat = torch.stack([x[..., ids] for ids in self.MAT_IDS]).detach().view(3, -1, 1, 2)
        vec = torch.stack([x[..., i] for i in self.VEC_IDS])
        vec = torch.stack((torch.zeros_like(vec), vec), dim=-1).detach().view(3, -1, 1, 2)
        return mat, vec

    def forward(self, x, d):
        _, SH, TruncExp, _ = self.ops
        N = x.shape[0]
        x = 2 * (x - self.aabb_train[:3]) / (self.aabb_train[3:] - self.aabb_train[:3]) - 1
        mat, vec = self._coords(x)
        sigma_feat = torch.zeros(N, device=x.device)
        for i in range(3):                                                     # get_sigma_feat, network.py:216-262
            mf = F.grid_sample(self.sigma_mat[i], mat[[i]], align_corners=True).view(-1, N)
            vf = F.grid_sample(self.sigma_vec[i], vec[[i]], align_corners=True).view(-1, N)
            sigma_feat = sigma_feat + torch.sum(mf * vf, dim=0)
        mfs = [F.grid_sample(self.color_mat[i], mat[[i]], align_corners=True).view(-1, N) for i in range(3)]   # get_color_feat
        vfs = [F.grid_sample(self.color_vec[i], vec[[i]], align_corners=True).view(-1, N) for i in range(3)]
        color_feat = self.basis_mat((torch.cat(mfs, dim=0) * torch.cat(vfs, dim=0)).T)
        sigma_feat = torch.clamp(sigma_feat, self.clip_min, self.clip_max)
        color_feat = torch.clamp(color_feat, self.clip_min, self.clip_max)
        self.feature_sigma_color = torch.cat([sigma_feat.unsqueeze(-1), color_feat], dim=-1)
        self.sigma_l = sigma_feat
        sigma = TruncExp.apply(sigma_feat)
        h = torch.cat([SH.apply(d.reshape(-1, 3), 4), color_feat], dim=-1)
        for l in range(3):
            h = self.color_net[l](h)
            if l != 2:
                h = F.relu(h, inplace=True)
        self.color_l = torch.sigmoid(h)
        return sigma, self.color_l

    def density_loss(self):                                                    # network.py:549-557
        loss = 0
        for i in range(3):
            loss = loss + torch.mean(torch.abs(self.sigma_mat[i])) + torch.mean(torch.abs(self.sigma_vec[i]))
        return loss


class RefMlpNetwork(nn.Module):
    """The mlp branch (NeRF): FreqEncoder PE=10 (tools/encoding.py:6-49) -> nerf_mlp, 8 Linear layers 256 wide with the skip concat
    after layer 3 (network.py:56-70,324-333) -> the same sigma_net / color_net tail as the hash model (:413-437)."""

    def __init__(self, ext, PE=10, W=256, layers=8, skips=3, clip_min=-2.0, clip_max=7.0):
        super().__init__()
        self.ops = make_ops(ext)
        self.clip_min, self.clip_max, self.skips = clip_min, clip_max, skips
        self.freqs = (2.0 ** torch.linspace(0.0, PE - 1, PE)).tolist()
        d_in = 3 + 3 * 2 * PE
        mlp = [nn.Linear(d_in, W)]
        for i in range(layers - 2):
            mlp.append(nn.Linear(W + d_in, W) if i == skips else nn.Linear(W, W))
        mlp.append(nn.Linear(W, 28))
        self.nerf_mlp = nn.ModuleList(mlp)
        self.sigma_net = nn.ModuleList([nn.Linear(28, 64, bias=False), nn.Linear(64, 16, bias=False)])
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 3, bias=False)])

    def forward(self, x, d):
        _, SH, TruncExp, _ = self.ops
        out = [x]
        for f in self.freqs:
            out += [torch.sin(x * f), torch.cos(x * f)]
        h = torch.cat(out, dim=-1)
        in_pts = h
        for i, layer in enumerate(self.nerf_mlp):
            h = layer(h)
            if i != len(self.nerf_mlp) - 1:
                h = F.relu(h, inplace=True)
            if i == self.skips:
                h = torch.cat([in_pts, h], -1)
        for l in range(2):
            h = self.sigma_net[l](h)
            if l != 1:
                h = F.relu(h, inplace=True)
        h[..., 0] = torch.clamp(h[..., 0].clone(), self.clip_min, self.clip_max)
        self.feature_sigma_color = h
        self.sigma_l = h[..., 0]
        sigma = TruncExp.apply(h[..., 0])
        h = torch.cat([SH.apply(d.reshape(-1, 3), 4), h[..., 1:]], dim=-1)
        for l in range(3):
            h = self.color_net[l](h)
            if l != 2:
                h = F.relu(h, inplace=True)
        self.color_l = torch.sigmoid(h)
        return sigma, self.color_l


class RefTrainer:
    """run_cuda (training branch) + MSE + scaled backward, as the reference's train_one_epoch does per iteration."""

    def __init__(self, ext, net, bitfield, bound=1.0, cascade=1, grid_size=128, min_near=0.2, max_steps=1024,
                 loss_scale=65536.0, l1_reg_weight=0.0, autocast=True):
        self.ext, self.net = ext, net
        self.autocast = autocast   # False = the same flow in fp32 (parity tests: the noise floor of the reference's own fp16 path)
        self.last = {}             # per-ray outputs / samples of the last step (parity tests)
        self.bitfield = bitfield
        self.bound, self.cascade, self.grid_size, self.min_near, self.max_steps = bound, cascade, grid_size, min_near, max_steps
        dev = bitfield.device
        self.aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32, device=dev)  # renderer.py:110-113
        self.mean_count, self.local_step = 0, 0
        self.loss_scale = loss_scale
        self.l1_reg_weight = l1_reg_weight

    def march(self, rays_o, rays_d, nears, fars, counter, perturb=True):
        rm = self.ext["raymarching"]
        N = rays_o.shape[0]
        M = N * self.max_steps
        if self.mean_count > 0:  # raymarching.py:235-238
            M = self.mean_count + (128 - self.mean_count % 128)
        dev = rays_o.device
        xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        rm.march_rays_train(rays_o, rays_d, self.bitfield, self.bound, 0.0, self.max_steps, N, self.cascade, self.grid_size, M, nears,
                            fars, xyzs, dirs, deltas, rays, counter, int(perturb))
        if self.mean_count <= 0:  # raymarching.py:276-284
            m = counter[0].item()
            m += 128 - m % 128
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
            torch.cuda.empty_cache()
        return xyzs, dirs, deltas, rays

    def step(self, rays_o, rays_d, gt, bg_color=1.0):
        rm = self.ext["raymarching"]
        Composite = self.net.ops[3]
        for p in self.net.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.autocast):  # fp16=True is forced by both CLIs (main_distill_mutual.py:251-254)
            N = rays_o.shape[0]
            nears = torch.empty(N, device=rays_o.device)
            fars = torch.empty(N, device=rays_o.device)
            rm.near_far_from_aabb(rays_o, rays_d, self.aabb, N, self.min_near, nears, fars)
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = self.march(rays_o, rays_d, nears, fars, counter)
            sigmas, rgbs = self.net(xyzs, dirs)
            ws, depth, image = Composite.apply(sigmas, rgbs, deltas, rays)
            image = image + (1 - ws).unsqueeze(-1) * bg_color
            depth = torch.clamp(depth - nears, min=0) / (fars - nears + 1e-6)
            loss = torch.mean((image - gt) ** 2)
            if self.l1_reg_weight and hasattr(self.net, "density_loss"):      # just_train_tea/utils.py:843-844
                loss = loss + self.net.density_loss() * self.l1_reg_weight
            self.last = {"image": image.detach(), "depth": depth.detach(), "weights_sum": ws.detach(), "rays": rays, "xyzs": xyzs,
                         "dirs": dirs, "deltas": deltas, "nears": nears, "fars": fars, "total": counter.clone()}
        (loss * self.loss_scale).backward()
        return loss

    def update_mean_count(self):
        total_step = min(16, self.local_step)  # renderer.py:768-773
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0


class RefPairTrainer(RefTrainer):
    """Trainer.train_step of distill_mutual/utils.py:954-1189 with render_stu_first (forced True, main_distill_mutual.py:240): the
    student renders (and marches), the teacher renders under no_grad on the student's `inherited_params` (renderer.py:374-394), then
    the stage-3 normL2 losses (:1110-1176) with the reference's default rates (main_distill_mutual.py:174-178)."""

    def __init__(self, ext, student, teacher, bitfield, rates=(1.0, 0.002, 0.002, 0.002), l1_reg_weight=1e-4, **kw):
        super().__init__(ext, student, bitfield, l1_reg_weight=l1_reg_weight, **kw)
        self.teacher = teacher
        self.rates = rates
        for p in teacher.parameters():
            p.requires_grad_(False)                                            # main_distill_mutual.py:320-321

    def _render(self, net, xyzs, dirs, deltas, rays, nears, fars, bg_color):
        Composite = net.ops[3]
        sigmas, rgbs = net(xyzs, dirs)
        ws, depth, image = Composite.apply(sigmas, rgbs, deltas, rays)
        image = image + (1 - ws).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears + 1e-6)
        return image

    def step(self, rays_o, rays_d, gt=None, bg_color=1.0):
        rm = self.ext["raymarching"]
        stu, tea = self.net, self.teacher
        for p in stu.parameters():
            p.grad = None
        r_rgb, r_fea, r_col, r_sig = self.rates
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.autocast):
            N = rays_o.shape[0]
            nears = torch.empty(N, device=rays_o.device)
            fars = torch.empty(N, device=rays_o.device)
            rm.near_far_from_aabb(rays_o, rays_d, self.aabb, N, self.min_near, nears, fars)
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = self.march(rays_o, rays_d, nears, fars, counter)
            pred_stu = self._render(stu, xyzs, dirs, deltas, rays, nears, fars, bg_color)
            with torch.no_grad():
                # the teacher's run_cuda recomputes near/far and then re-uses the samples (renderer.py:342,393-394)
                rm.near_far_from_aabb(rays_o, rays_d, self.aabb, N, self.min_near, nears, fars)
                pred_tea = self._render(tea, xyzs, dirs, deltas, rays, nears, fars, bg_color)
            t_rgb = torch.norm(pred_tea - pred_stu)
            t_fea = torch.norm(stu.feature_sigma_color - tea.feature_sigma_color)
            t_col = torch.norm(stu.color_l - tea.color_l)
            t_sig = torch.norm(stu.sigma_l - tea.sigma_l)
            loss = r_rgb * t_rgb                                                                      # :1110-1111
            if self.l1_reg_weight and hasattr(stu, "density_loss"):
                loss = loss + stu.density_loss() * self.l1_reg_weight                                 # :1135-1136
            loss = loss + r_fea * t_fea                                                               # :1137-1149
            loss = loss + r_col * t_col                                                               # :1158-1165
            loss = loss + r_sig * t_sig                                                               # :1166-1173
            self.last = {"image": pred_stu.detach(), "image_tea": pred_tea.detach(), "rays": rays, "xyzs": xyzs, "dirs": dirs,
                         "deltas": deltas, "total": counter.clone(),
                         "terms": {"rgb": t_rgb.detach(), "fea": t_fea.detach(), "color": t_col.detach(), "sigma": t_sig.detach()}}
        (loss * self.loss_scale).backward()
        return loss

