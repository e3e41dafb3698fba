"""Training-step engine for the hot path: rays -> near/far -> march -> fused field -> composite -> loss -> backward,
on caller-provided parameters, with pre-allocated device buffers, no host synchronisation and no autograd graph.

This is what `NeRFRenderer.run_cuda` (training branch, distill_mutual/renderer.py:359-448) + `Trainer.train_step`'s
criterion (just_train_tea/utils.py:841-846) + `loss.backward()` amount to for a "hash" model, expressed as 10 kernel
launches on one stream:

    near_far | march count | offset scan | sample expand | field fwd | composite fwd | composite bwd (+MSE) | field bwd
    (+ two memsets of the gradient accumulators)

Sample-buffer sizing follows the reference: the first `warmup` steps read the sample counter back (raymarching.py:277) and
their mean becomes `mean_count`; afterwards M = mean_count rounded up strictly to 128 and rays that do not fit are dropped
(raymarching.cu:419).  Gradients are left in `grad_table` / `grad_weights()`; an optimizer step is the caller's business
(it is outside the timed region of the benchmark).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nv
from . import fused

_u32, _f32 = C.c_uint32, C.c_float


class HashTrainEngine:
    def __init__(self, field: "fused.HashNeRFField", bitfield: torch.Tensor, n_rays: int, bound: float = 1.0, cascade: int = 1,
                 grid_size: int = 128, min_near: float = 0.2, max_steps: int = 1024, dt_gamma: float = 0.0, bg_color=(1.0, 1.0, 1.0),
                 loss_scale: float = 1.0, density_scale: float = 1.0, device="cuda"):
        self.field = field
        self.dev = torch.device(device)
        self.N = int(n_rays)
        self.bound, self.cascade, self.grid_size = float(bound), int(cascade), int(grid_size)
        self.min_near, self.max_steps, self.dt_gamma = float(min_near), int(max_steps), float(dt_gamma)
        self.loss_scale = float(loss_scale)
        self.density_scale = float(density_scale)
        self.bitfield = bitfield.to(self.dev).contiguous()
        d = self.dev
        N = self.N
        self.aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=d)
        self.bg = torch.tensor(list(bg_color), dtype=torch.float32, device=d)
        self.rays_o = torch.empty(N, 3, device=d)
        self.rays_d = torch.empty(N, 3, device=d)
        self.gt = torch.empty(N, 3, device=d)
        self.nears = torch.empty(N, device=d)
        self.fars = torch.empty(N, device=d)
        self.rays = torch.empty(N, 3, dtype=torch.int32, device=d)
        # everything that must be zero at the start of a step lives in ONE buffer (one memset node instead of four):
        # counter[2] int32 | loss[2] f32 | gw_ws
        self._zeros = torch.zeros(4 + fused.GW_WS_FLOATS, dtype=torch.float32, device=d)
        self.counter = self._zeros[0:2].view(torch.int32)
        self.loss = self._zeros[2:4]
        self.gw_ws = self._zeros[4:]
        self._side = torch.cuda.Stream(device=d)
        self.ws_march = torch.empty(int(nv.lib().pvd_march_rays_train_workspace_words(N, self.max_steps)), dtype=torch.int32, device=d)
        self.weights_sum = torch.empty(N, device=d)
        self.depth = torch.empty(N, device=d)
        self.image = torch.empty(N, 3, device=d)
        self.status = torch.zeros(1, dtype=torch.int32, device=d)
        emb = field.encoder.embeddings
        self.grad_table = torch.zeros(emb.shape, dtype=torch.float32, device=d)
        self.M = 0
        self.mean_count = 0
        self._counts = []
        self._alloc_samples(N * 32)
        self._coarse_valid = False
        self.launches_per_step = 8 if fused.SPLIT_SCATTER else 7  # kernels of libpvd_b200.so only (torch memsets not counted)

    # ------------------------------------------------------------------ buffers sized by M
    def _alloc_samples(self, M: int):
        d = self.dev
        self.M = int(M)
        self.xyzs = torch.zeros(M, 3, device=d)
        self.dirs = torch.zeros(M, 3, device=d)
        self.deltas = torch.zeros(M, 2, device=d)
        self.sigmas = torch.empty(M, device=d)
        self.rgbs = torch.empty(M, 3, device=d)
        self.enc = torch.empty(M, fused.ENC_STRIDE, dtype=torch.float16, device=d)
        self.grad_sigmas = torch.zeros(M, device=d)
        self.grad_rgbs = torch.zeros(M, 3, device=d)
        self.dx_ws = torch.empty(M, fused.ENC_STRIDE, dtype=torch.float16, device=d) if fused.SPLIT_SCATTER else None

    def set_bitfield(self, bitfield: torch.Tensor):
        """New occupancy bitfield (after a density-grid update): the cached coarse rejection mask is invalid."""
        self.bitfield = bitfield.to(self.dev).contiguous()
        self._coarse_valid = False

    def set_mean_count(self, mean_count: int):
        """M = mean_count rounded up strictly to a multiple of 128 (raymarching.py:235-238)."""
        self.mean_count = int(mean_count)
        M = self.mean_count + (128 - self.mean_count % 128)
        self._alloc_samples(M)

    def stage(self):
        """Refresh the fp16 table shadow and the packed weight tiles if the parameters changed."""
        f = self.field
        cfg = f.config()
        cfg.density_scale = self.density_scale
        self.cfg = cfg
        self.table = f._staged.table_for(f.encoder.embeddings, cfg.table_fp16)
        self.wblob = f._staged.wblob_for((f.sigma_net[0].weight, f.sigma_net[1].weight, f.color_net[0].weight,
                                          f.color_net[1].weight, f.color_net[2].weight), 2 * cfg.num_levels)
        self.cfield = fused._cstruct(cfg, self.table, f.encoder.offsets, self.wblob)

    # ------------------------------------------------------------------ one step
    def _march_count(self, st):
        l = nv.lib()
        # near/far fused into the count kernel; the coarse rejection mask is rebuilt only when the bitfield changed
        nv.check(l.pvd_march_rays_train_count_aabb(nv.ptr(self.rays_o), nv.ptr(self.rays_d), nv.ptr(self.bitfield), nv.ptr(self.aabb),
                                                   _f32(self.min_near), _f32(self.bound), _f32(self.dt_gamma), _u32(self.max_steps),
                                                   _u32(self.N), _u32(self.cascade), _u32(self.grid_size), nv.ptr(self.nears),
                                                   nv.ptr(self.fars), nv.ptr(self.rays), nv.ptr(self.counter), _u32(1),
                                                   _u32(1 if self._coarse_valid else 0), nv.ptr(self.ws_march), st))
        self._coarse_valid = True

    def step(self, warmup: bool = False):
        """Forward + backward for the rays currently in self.rays_o / rays_d / gt.  Leaves loss in self.loss[0]."""
        l = nv.lib()
        st = nv.stream_of(self.rays_o)
        cur = torch.cuda.current_stream(self.dev)
        self._zeros.zero_()                    # counter, loss, weight-gradient workspace
        self._side.wait_stream(cur)            # fork: the 42 MB table-gradient memset runs beside the march (HBM vs. latency bound)
        with torch.cuda.stream(self._side):
            self.grad_table.zero_()
        self._march_count(st)
        if warmup:  # size the sample buffers from this step's count (one D2H read, raymarching.py:277)
            total = int(self.counter[0].item())
            self._counts.append(total)
            need = total + (128 - total % 128)
            if need > self.M:
                self._alloc_samples(need)
            M_drop = self.N * self.max_steps
            M = need
        else:
            M = M_drop = self.M
        N = self.N
        nv.check(l.pvd_march_rays_train_write(nv.ptr(self.rays_o), nv.ptr(self.rays_d), _f32(self.bound), _u32(self.max_steps),
                                              _u32(N), _u32(M_drop), nv.ptr(self.rays), nv.ptr(self.ws_march), nv.ptr(self.xyzs),
                                              nv.ptr(self.dirs), nv.ptr(self.deltas), st))
        nv.check(l.pvd_hash_field_forward(C.byref(self.cfield), nv.ptr(self.xyzs), nv.ptr(self.dirs), _u32(M), nv.ptr(self.sigmas),
                                          nv.ptr(self.rgbs), nv.ptr(self.enc), None, nv.ptr(self.status), st))
        nv.check(l.pvd_composite_rays_train_forward(nv.ptr(self.sigmas), nv.ptr(self.rgbs), nv.ptr(self.deltas), nv.ptr(self.rays),
                                                    _u32(M_drop), _u32(N), nv.ptr(self.weights_sum), nv.ptr(self.depth),
                                                    nv.ptr(self.image), st))
        # backward (grad_sigmas / grad_rgbs need no clearing: every row below n_valid is written by the composite backward)
        nv.check(l.pvd_composite_rays_train_backward_mse(nv.ptr(self.gt), nv.ptr(self.bg), _f32(self.loss_scale), nv.ptr(self.sigmas),
                                                         nv.ptr(self.rgbs), nv.ptr(self.deltas), nv.ptr(self.rays),
                                                         nv.ptr(self.weights_sum), nv.ptr(self.image), _u32(M_drop), _u32(N),
                                                         nv.ptr(self.grad_sigmas), nv.ptr(self.grad_rgbs), nv.ptr(self.loss), st))
        cur.wait_stream(self._side)            # join: the table gradient is clear before the first reduction into it
        nv.check(l.pvd_hash_field_backward(C.byref(self.cfield), nv.ptr(self.xyzs), nv.ptr(self.dirs), nv.ptr(self.enc),
                                           nv.ptr(self.grad_sigmas), nv.ptr(self.grad_rgbs), None, _u32(M), nv.ptr(self.counter),
                                           nv.ptr(self.grad_table), nv.ptr(self.gw_ws), nv.ptr(self.dx_ws), nv.ptr(self.status), st))

    # ------------------------------------------------------------------ CUDA graph of one steady-state step
    def capture(self):
        """Capture `step()` (fixed M, static input buffers rays_o / rays_d / gt) into a CUDA graph; `replay()` then costs one
        launch.  Inputs must be written INTO self.rays_o / self.rays_d / self.gt (copy_), not rebound."""
        self.step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step()
        self.graph = g
        return g

    def replay(self):
        self.graph.replay()

    def finish_warmup(self):
        """mean_count = mean of the warm-up sample counts (renderer.py:768-772)."""
        if self._counts:
            self.set_mean_count(int(sum(self._counts) / len(self._counts)))
        self._counts = []

    def grad_weights(self):
        f = self.field
        like = (f.sigma_net[0].weight, f.sigma_net[1].weight, f.color_net[0].weight, f.color_net[1].weight, f.color_net[2].weight)
        return fused.unpack_wgrads(self.gw_ws, 2 * self.cfg.num_levels, like)

    def final_image(self):
        """pred rgb [N,3] and normalised depth [N] as run_cuda returns them (renderer.py:445-446)."""
        pred = self.image + (1 - self.weights_sum).unsqueeze(-1) * self.bg
        depth = torch.clamp(self.depth - self.nears, min=0) / (self.fars - self.nears + 1e-6)
        return pred, depth
