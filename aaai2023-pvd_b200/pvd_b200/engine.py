"""Training-step engine for the hot path: rays -> near/far -> march -> fused field -> composite -> loss -> backward,
on caller-provided parameters, with pre-allocated device buffers, no host synchronisation and no autograd graph.

This is what `NeRFRenderer.run_cuda` (training branch, distill_mutual/renderer.py:359-448) + `Trainer.train_step`'s
criterion (just_train_tea/utils.py:841-846) + `loss.backward()` amount to for a "hash" or "vm" model, expressed as 6-7 kernel
launches on one stream (+ two memsets of the gradient accumulators on a side branch):

    march count (+ near/far) | offset scan | sample expand | field fwd | composite fwd + MSE + composite bwd | field bwd (| scatter)

`FieldTrainEngine` is that step for one model (HashTrainEngine / VMTrainEngine); `PairDistillEngine` is the distillation step of
`main_distill_mutual.py` for a (teacher, student) pair evaluated at the SAME samples (distill_mutual/utils.py:954-1189).

Two ray sets (inputs + march outputs) are kept so that the march of batch i+1 -- a latency-bound kernel chain that depends on
nothing the model computes -- can run BESIDE the field backward of batch i (`capture_pipelined` / `replay_pipelined`: the
prefetch a training loop's data loader would do, expressed as a second branch of the step's CUDA graph).

Sample-buffer sizing follows the reference: the first `warmup` steps read the sample counter back (raymarching.py:277) and
their mean becomes `mean_count`; afterwards M = mean_count rounded up strictly to 128 and rays that do not fit are dropped
(raymarching.cu:419).  Gradients are left in `grad_table` / `grad_weights()`; an optimizer step is the caller's business
(it is outside the timed region of the benchmark).
"""
from __future__ import annotations

import ctypes as C
import contextlib
import gc
import os

import numpy as np
import torch

from . import _native as nv
from . import field_ops, fused

_u32, _f32 = C.c_uint32, C.c_float
# opt-in experiment (measured SLOWER on B200, 0.129 vs 0.120 ms/step: the full-occupancy scatter CTAs crowd out the MLP CTAs of the
# other half instead of overlapping with them); the default issues MLP backward and scatter back to back on one stream
SPLIT_HALVES = os.environ.get("PVD_SPLIT_HALVES", "0") == "1"
PACK_LAST = os.environ.get("PVD_PACK_LAST", "1") != "0"             # weight-pack kernel immediately before the forward (PDL edge)
LOSS_D2H_EARLY = os.environ.get("PVD_LOSS_D2H_EARLY", "1") != "0"   # host-fed graphs: read the loss back beside the backward


@contextlib.contextmanager
def _no_gc():
    """No cyclic garbage collection while a stream is capturing: collecting some unrelated object that owns pinned host memory (or
    anything else whose release records a CUDA event) in the middle of a capture invalidates it."""
    was = gc.isenabled()
    gc.collect()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class _RaySet:
    """What belongs to one batch of rays: the inputs and everything the march produces from them."""

    def __init__(self, N, dev):
        # rays_o | rays_d | gt live in ONE buffer: a batch arriving from the host is a single H2D copy (147 KB at 4096 rays)
        self.inputs = torch.empty(3, N, 3, device=dev)
        self.host = None   # pinned staging twin of `inputs` (FieldTrainEngine.enable_host_io)
        self.reset_views()
        self.nears = torch.empty(N, device=dev)
        self.fars = torch.empty(N, device=dev)
        self.rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        self.counter = torch.zeros(2, dtype=torch.int32, device=dev)
        self.xyzs = self.dirs = self.deltas = None

    def reset_views(self):
        """rays_o / rays_d / gt are views of `inputs` (they may have been re-bound to caller tensors in eager use)."""
        self.rays_o, self.rays_d, self.gt = self.inputs[0], self.inputs[1], self.inputs[2]

    def alloc_samples(self, M, dev):
        self.xyzs = torch.zeros(M, 3, device=dev)
        self.dirs = torch.zeros(M, 3, device=dev)
        self.deltas = torch.zeros(M, 2, device=dev)


def _set_attr(name):
    def get(self):
        return getattr(self.sets[self.cur], name)

    def put(self, v):
        setattr(self.sets[self.cur], name, v)
    return property(get, put)


class FieldTrainEngine:
    """Training from images of ONE model of type hash or vm (`main_just_train_tea.py`; BASELINE configs 2 and 3): MSE against
    gt_rgb (just_train_tea/utils.py:841-846), plus l1_reg_weight * density_loss() for vm models (:843-844)."""
    # how rank gradients combine into the global-batch gradient under ray sharding: every rank's MSE is a mean over ITS rays
    grad_reduction = "mean"
    # the "current" ray set is what step() and every single-set accessor works on
    rays_o, rays_d, gt = _set_attr("rays_o"), _set_attr("rays_d"), _set_attr("gt")
    nears, fars, rays, counter = _set_attr("nears"), _set_attr("fars"), _set_attr("rays"), _set_attr("counter")
    xyzs, dirs, deltas = _set_attr("xyzs"), _set_attr("dirs"), _set_attr("deltas")

    def __init__(self, field, bitfield: torch.Tensor, n_rays: int, bound: float = 1.0, cascade: int = 1,
                 grid_size: int = 128, min_near: float = 0.2, max_steps: int = 1024, dt_gamma: float = 0.0, bg_color=(1.0, 1.0, 1.0),
                 loss_scale: float = 1.0, density_scale: float = 1.0, l1_reg_weight: float = 0.0, perturb: bool = True,
                 device="cuda"):
        self.field = field
        self.perturb = bool(perturb)   # per-ray jitter of the first sample (renderer.py:387: perturb=True in training)
        self.dev = torch.device(device)
        self.N = int(n_rays)
        self.bound, self.cascade, self.grid_size = float(bound), int(cascade), int(grid_size)
        self.min_near, self.max_steps, self.dt_gamma = float(min_near), int(max_steps), float(dt_gamma)
        self.loss_scale = float(loss_scale)
        self.density_scale = float(density_scale)
        self.l1_reg_weight = float(l1_reg_weight)
        self.bitfield = bitfield.to(self.dev).contiguous().clone()   # persistent: set_bitfield() copies into it
        d = self.dev
        N = self.N
        self.aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=d)
        self.bg = torch.tensor(list(bg_color), dtype=torch.float32, device=d)
        self.sets = [_RaySet(N, d), _RaySet(N, d)]
        self.cur = 0
        self._init_small_buffers()
        self._side = torch.cuda.Stream(device=d)    # parameter-gradient memset
        self._side2 = torch.cuda.Stream(device=d)   # march of the next batch (pipelined mode)
        self._side3 = torch.cuda.Stream(device=d)   # table-gradient scatter beside the MLP backward of the other half
        self.ws_march = torch.empty(int(nv.lib().pvd_march_rays_train_workspace_words(N, self.max_steps)), dtype=torch.int32, device=d)
        self.weights_sum = torch.empty(N, device=d)
        self.depth = torch.empty(N, device=d)
        self.image = torch.empty(N, 3, device=d)
        self.status = torch.zeros(1, dtype=torch.int32, device=d)
        self._init_fields()
        self.M = 0
        self.mean_count = 0
        self._counts = []
        self._alloc_samples(N * 32)
        self._coarse_valid = False
        # what surrounds forward + backward in a training iteration (all of it CUDA-graph capturable, see `_prologue` / `_epilogue`)
        # PVD_PREFETCH_L2=1: stream the tables / planes the forward gathers from into L2 at the top of the step (TMA bulk prefetch on the
        # memset branch).  MEASURED with bench.py's cold L2 (256 MiB flush between steps): hash step 112.8 vs 112.9 us (no effect: the
        # forward's cold misses are not what bounds it), vm 242.5 vs 212.8 us and hash -> vm 281 vs 251 us (WORSE: 69 MB of planes + 69 MB
        # of gradients do not fit L2 together, the prefetch evicts what the backward needs).  Off by default.
        self.prefetch_l2 = os.environ.get("PVD_PREFETCH_L2", "0") != "0"
        self.restage_each_step = False   # parameters are changed by an EXTERNAL optimizer between steps: re-cast / re-pack at the top
        self.unpack_each_step = False    # leave the small weight gradients in parameter shapes (ops.wgrads) at the end of the step
        self.exchange = None             # dist.TableGradExchange: the one all-reduce of the multi-GPU path, after the backward
        self.optimizer = None            # optim.FusedAdamW: updates parameters, fp16 shadow and weight tiles, zeroes the gradients

    def _init_small_buffers(self):
        # what must be zero before a backward lives in ONE buffer (one memset node): loss slots [64][2] f32 | gw_ws
        self._zeros = torch.zeros(2 * fused.LOSS_SLOTS + fused.GW_WS_FLOATS, dtype=torch.float32, device=self.dev)
        self.loss_slots = self._zeros[0:2 * fused.LOSS_SLOTS]   # PVD_LOSS_SLOTS pairs (loss, rays); see `loss`
        self.gw_ws = self._zeros[2 * fused.LOSS_SLOTS:]

    def _init_fields(self):
        self.ops = field_ops.make_ops(self.field, self.dev, trainable=True)
        # kernels of libpvd_b200.so only (torch memsets not counted): count, scan, expand, fwd, composite (fwd+bwd), field backward
        self.launches_per_step = 3 + self.ops.kernels_fwd + 1 + self.ops.kernels_bwd
        if self.ops.kind == "vm" and self.l1_reg_weight:
            self.launches_per_step += 6

    @property
    def grad_table(self):   # hash models: the table gradient [entries, 2] fp32
        return self.ops.grad_table

    @property
    def enc(self):
        return self.ops.enc

    @property
    def dx_ws(self):
        return self.ops.dx_ws

    @property
    def cfield(self):
        return self.ops.cfield

    @property
    def cfg(self):
        return self.ops.cfg

    # ------------------------------------------------------------------ buffers sized by M
    def _alloc_samples(self, M: int):
        d = self.dev
        self.M = int(M)
        for rs in self.sets:
            rs.alloc_samples(M, d)
        self.sigmas = torch.empty(M, device=d)
        self.rgbs = torch.empty(M, 3, device=d)
        self.grad_sigmas = torch.zeros(M, device=d)
        self.grad_rgbs = torch.zeros(M, 3, device=d)
        self._alloc_field_samples(M)

    def _alloc_field_samples(self, M: int):
        self.ops.alloc(M)

    def set_bitfield(self, bitfield: torch.Tensor):
        """New occupancy bitfield (after a density-grid update, renderer.py:647-773).  The bytes are copied INTO the engine's
        persistent buffer -- captured graphs hold its address -- and the coarse rejection mask in the march workspace is rebuilt
        right away on the current stream, so that a replay of an existing graph marches against the new grid."""
        bitfield = bitfield.to(self.dev)
        if bitfield.shape != self.bitfield.shape or bitfield.dtype != self.bitfield.dtype:
            raise ValueError(f"bitfield must stay {tuple(self.bitfield.shape)} {self.bitfield.dtype} (captured graphs hold its address)")
        self.bitfield.copy_(bitfield)
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        nv.check(nv.lib().pvd_march_coarse_mask(nv.ptr(self.bitfield), _u32(self.cascade), _u32(self.grid_size), _f32(self.bound),
                                                nv.ptr(self.ws_march), st))
        self._coarse_valid = True

    def set_mean_count(self, mean_count: int):
        """M = mean_count rounded up strictly to a multiple of 128 (raymarching.py:235-238)."""
        self.mean_count = int(mean_count)
        M = self.mean_count + (128 - self.mean_count % 128)
        self._alloc_samples(M)

    def stage(self):
        """Refresh the staged parameters (fp16 table shadow, packed weight tiles) if the parameters changed."""
        self.ops.stage(self.density_scale)

    # ------------------------------------------------------------------ one step
    def _march_count(self, st, rs=None):
        rs = rs or self.sets[self.cur]
        l = nv.lib()
        rs.counter.zero_()
        # near/far fused into the count kernel; the coarse rejection mask is rebuilt only when the bitfield changed
        nv.check(l.pvd_march_rays_train_count_aabb(nv.ptr(rs.rays_o), nv.ptr(rs.rays_d), nv.ptr(self.bitfield), nv.ptr(self.aabb),
                                                   _f32(self.min_near), _f32(self.bound), _f32(self.dt_gamma), _u32(self.max_steps),
                                                   _u32(self.N), _u32(self.cascade), _u32(self.grid_size), nv.ptr(rs.nears),
                                                   nv.ptr(rs.fars), nv.ptr(rs.rays), nv.ptr(rs.counter), _u32(1 if self.perturb else 0),
                                                   _u32(1 if self._coarse_valid else 0), nv.ptr(self.ws_march), st))
        self._coarse_valid = True

    def _march_write(self, st, rs, M_drop):
        nv.check(nv.lib().pvd_march_rays_train_write(nv.ptr(rs.rays_o), nv.ptr(rs.rays_d), _f32(self.bound), _u32(self.max_steps),
                                                     _u32(self.N), _u32(M_drop), nv.ptr(rs.rays), nv.ptr(self.ws_march),
                                                     nv.ptr(rs.xyzs), nv.ptr(rs.dirs), nv.ptr(rs.deltas), st))

    def _forward(self, st, rs, M, M_drop):
        self.ops.forward(st, rs.xyzs, rs.dirs, M, self.sigmas, self.rgbs, None, self.status)

    def _loss_backward(self, st, rs, M, M_drop):
        """composite forward + MSE + composite backward in one launch.  grad_sigmas / grad_rgbs need no clearing: every row below
        n_valid is written here."""
        nv.check(nv.lib().pvd_composite_rays_train_mse(
            nv.ptr(rs.gt), nv.ptr(self.bg), _f32(self.loss_scale), nv.ptr(self.sigmas), nv.ptr(self.rgbs), nv.ptr(rs.deltas),
            nv.ptr(rs.rays), _u32(M_drop), _u32(self.N), nv.ptr(self.weights_sum), nv.ptr(self.depth), nv.ptr(self.image),
            nv.ptr(self.grad_sigmas), nv.ptr(self.grad_rgbs), nv.ptr(self.loss_slots), st))

    def _field_backward(self, st, rs, M, cur=None):
        """Field backward (hash: MLP backward + table-gradient scatter).  With a stream handle for the main branch (`cur`) and
        PVD_SPLIT_HALVES=1, the rows of a hash model are cut in two halves: MLP(A), MLP(B) run on the main branch and scatter(A),
        scatter(B) on a side branch, so that the atomic-bound scatter of one half overlaps the latency-bound tcgen05 chain of the other."""
        ops = self.ops
        Mh = (M // 256) * 128
        if ops.kind != "hash" or cur is None or ops.dx_ws is None or Mh == 0 or not SPLIT_HALVES:
            ops.backward(st, rs.xyzs, rs.dirs, self.grad_sigmas, self.grad_rgbs, None, M, rs.counter, self.gw_ws, self.status)
            return
        l = nv.lib()
        args = (C.byref(ops.cfield), nv.ptr(rs.xyzs), nv.ptr(rs.dirs), nv.ptr(ops.enc), nv.ptr(self.grad_sigmas),
                nv.ptr(self.grad_rgbs), None)
        tail = (nv.ptr(rs.counter), nv.ptr(ops.grad_table), nv.ptr(self.gw_ws), nv.ptr(ops.dx_ws), nv.ptr(self.status))
        st3 = C.c_void_p(self._side3.cuda_stream)
        nv.check(l.pvd_hash_field_backward_rows(*args, _u32(0), _u32(Mh), *tail, _u32(1), st))
        self._side3.wait_stream(cur)
        with torch.cuda.stream(self._side3):
            nv.check(l.pvd_hash_field_backward_rows(*args, _u32(0), _u32(Mh), *tail, _u32(2), st3))
        nv.check(l.pvd_hash_field_backward_rows(*args, _u32(Mh), _u32(M - Mh), *tail, _u32(1), st))
        self._side3.wait_stream(cur)
        with torch.cuda.stream(self._side3):
            nv.check(l.pvd_hash_field_backward_rows(*args, _u32(Mh), _u32(M - Mh), *tail, _u32(2), st3))
        cur.wait_stream(self._side3)

    def _clear_big(self, cur):
        """fork: the parameter-gradient memset (42 MB hash table / 69 MB vm planes) runs beside the march and the forward (HBM- vs.
        latency-bound); parameter-only loss terms (the vm L1 penalty) follow it on the same branch."""
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            if self.prefetch_l2:         # first on this branch: the forward starts gathering within microseconds
                self._prefetch(C.c_void_p(self._side.cuda_stream))
            if self.optimizer is None:   # a fused optimizer leaves the big gradient buffer zeroed (same pass as the update)
                self.ops.clear_grads()
                if self.unpack_each_step:    # the parameter-shaped weight gradients the epilogue's unpack kernel accumulates into
                    self.ops.zero_weight_grads()
            if self.l1_reg_weight:
                self.ops.regularise(C.c_void_p(self._side.cuda_stream), self.loss_scale, self.loss_slots, self.l1_reg_weight)

    def step(self, warmup: bool = False):
        """Forward + backward for the rays of the current set (self.rays_o / rays_d / gt).  Leaves the loss in self.loss[0]."""
        rs = self.sets[self.cur]
        cur = torch.cuda.current_stream(self.dev)
        st = C.c_void_p(cur.cuda_stream)
        self._prologue(cur)
        self._march_count(st, rs)
        if warmup:  # size the sample buffers from this step's count (one D2H read, raymarching.py:277)
            total = int(rs.counter[0].item())
            self._counts.append(total)
            need = total + (128 - total % 128)
            if need > self.M:
                self._alloc_samples(need)
            M_drop = self.N * self.max_steps
            M = need
        else:
            M = M_drop = self.M
        self._march_write(st, rs, M_drop)
        self._forward(st, rs, M, M_drop)
        # join BEFORE the loss kernel: the table gradient is clear before the first reduction into it, and the field backward's only
        # predecessor is then the loss kernel -- which lets it start early as a programmatic dependent launch (csrc/common.cuh)
        cur.wait_stream(self._side)
        self._loss_backward(st, rs, M, M_drop)
        self._field_backward(st, rs, M, cur)
        self._epilogue(st)

    def _prefetch(self, st):
        self.ops.prefetch(st)

    def _prologue(self, cur):
        restage = self.restage_each_step and self.optimizer is None
        if restage and not PACK_LAST:
            self.ops.stage(self.density_scale)     # trainable parameters only: a frozen teacher stays staged
        self._zeros.zero_()                    # loss, weight-gradient workspace
        self._clear_big(cur)
        if restage and PACK_LAST:
            # the pack kernel LAST: the field forward then follows a KERNEL in stream order and starts as its programmatic dependent
            # (it gathers its first tile while the tiles are packed); behind a memset node the launch is an ordinary full dependency
            self.ops.stage(self.density_scale)

    def _epilogue(self, st):
        """After the backward: gradient exchange (multi-GPU), then either the fused optimizer or -- for an external optimizer -- the
        small weight gradients in parameter shapes."""
        if self.exchange is not None:
            self.exchange()
        if self.optimizer is not None:
            self.optimizer.step()
        elif self.unpack_each_step:
            self.ops.unpack_weight_grads(self.gw_ws, st, zero=False)   # zeroed on the memset branch (_clear_big), off the critical path

    def attach_optimizer(self, optimizer):
        """From here on every step ends with `optimizer.step()` (graphs must be captured afterwards)."""
        self.optimizer = optimizer
        self.ops.clear_grads()                 # once: the optimizer keeps them zero from now on
        return optimizer

    # ------------------------------------------------------------------ CUDA graph of one steady-state step
    def capture(self):
        """Capture `step()` (fixed M, static input buffers rays_o / rays_d / gt) into a CUDA graph; `replay()` then costs one
        launch.  Inputs must be written INTO self.rays_o / self.rays_d / self.gt (copy_), not rebound."""
        self.step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with _no_gc(), torch.cuda.graph(g):
            self.step()
        self.graph = g
        return g

    def replay(self):
        self.graph.replay()

    # ------------------------------------------------------------------ pipelined steady state: march(i+1) beside backward(i)
    def march(self, k: int):
        """Eagerly march ray set k (priming the pipeline: the first batch has no previous step to hide behind)."""
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        self._march_count(st, self.sets[k])
        self._march_write(st, self.sets[k], self.M)

    # ------------------------------------------------------------------ host-resident batches (the end-to-end path)
    n_host_inputs = 3   # rays_o, rays_d, gt  (a distillation step has no gt: PairDistillEngine copies two)

    def _loss_dev(self):
        return self.loss_slots

    def enable_host_io(self):
        """Pinned staging buffers so that a step fed from the HOST costs no extra launches: `capture_pipelined(host_io=True)` puts
        the H2D copy of batch i+1 (sets[k].host -> sets[k].inputs, one copy) at the head of the march branch of step i's graph -- off
        the critical path -- and the D2H copy of the loss words (-> self.host_loss) at its end."""
        for rs in self.sets:
            if rs.host is None:
                rs.host = torch.empty(3, self.N, 3, dtype=torch.float32).pin_memory()
                rs.host.copy_(rs.inputs)   # never feed uninitialised rays to the marcher (the capture pass runs the step once)
        self.host_loss = torch.empty(self._loss_dev().numel(), dtype=torch.float32).pin_memory()

    def _pipelined_step(self, k: int, host_io: bool = False):
        """Field forward/backward of set k (already marched) with the march of set 1-k on a parallel branch, forked at the top of
        the step (the march is a latency-bound chain of one-warp CTAs that fits beside the field kernels)."""
        rs, nxt = self.sets[k], self.sets[1 - k]
        cur = torch.cuda.current_stream(self.dev)
        st = C.c_void_p(cur.cuda_stream)
        M = self.M

        def march_branch():
            self._side2.wait_stream(cur)
            with torch.cuda.stream(self._side2):
                st2 = C.c_void_p(self._side2.cuda_stream)
                if host_io:
                    n = self.n_host_inputs
                    nxt.inputs[:n].copy_(nxt.host[:n], non_blocking=True)
                self._march_count(st2, nxt)
                self._march_write(st2, nxt, M)

        # Where the next batch's march runs.  Single GPU: forked at the top, beside forward / backward (it fills the SMs' idle
        # issue slots).  With a gradient exchange attached: forked AFTER the backward, beside the exchange -- the collective leaves
        # the SMs almost idle, and forward / backward no longer share them with 4096 one-warp march CTAs.
        where = os.environ.get("PVD_PIPE_FORK", "auto")
        if where == "auto":
            where = "exchange" if self.exchange is not None else "early"
        if where == "early":
            march_branch()             # forked BEFORE the prologue: the march must not wait for the re-stage / memsets
        self._prologue(cur)
        self._forward(st, rs, M, M)
        if where == "late":
            march_branch()
        cur.wait_stream(self._side)    # before the loss kernel (see step()): the backward follows it as a programmatic dependent launch
        self._loss_backward(st, rs, M, M)
        if host_io and LOSS_D2H_EARLY:
            # the loss words are final once the loss kernel has run: their D2H copy goes on a side branch, beside the field backward,
            # instead of between the backward and the epilogue on the critical path
            self._side3.wait_stream(cur)
            with torch.cuda.stream(self._side3):
                self.host_loss.copy_(self._loss_dev(), non_blocking=True)
        self._field_backward(st, rs, M, cur)
        if host_io and not LOSS_D2H_EARLY:
            self.host_loss.copy_(self._loss_dev(), non_blocking=True)
        if where == "exchange":
            march_branch()
        self._epilogue(st)
        cur.wait_stream(self._side2)
        if host_io and LOSS_D2H_EARLY:
            cur.wait_stream(self._side3)

    def capture_pipelined(self, host_io: bool = False):
        """Two graphs (even / odd steps).  Protocol: write batch 0 into sets[0], call march(0); then for step i write batch i+1
        into sets[(i+1) % 2] (rays_o, rays_d, gt -- or, with host_io, into its pinned `host` twin) and replay_pipelined(i)."""
        if host_io:
            self.enable_host_io()
        graphs = []
        for k in (0, 1):
            self.march(k)
            self._pipelined_step(k, host_io)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with _no_gc(), torch.cuda.graph(g):
                self._pipelined_step(k, host_io)
            graphs.append(g)
        torch.cuda.synchronize()
        if host_io:
            self.graphs_host = graphs
        else:
            self.graphs = graphs
        return graphs

    def replay_pipelined(self, i: int, host_io: bool = False):
        self.cur = i & 1
        (self.graphs_host if host_io else self.graphs)[i & 1].replay()

    @property
    def loss(self):
        """[unscaled MSE loss, rays that contributed] of the last step (the kernel spreads them over PVD_LOSS_SLOTS slots)."""
        return self.loss_slots.view(fused.LOSS_SLOTS, 2).sum(0)

    def finish_warmup(self):
        """mean_count = mean of the warm-up sample counts (renderer.py:768-772)."""
        if self._counts:
            self.set_mean_count(int(sum(self._counts) / len(self._counts)))
        self._counts = []

    def grad_weights(self):
        """The MLP weight gradients in parameter shapes, reference order (hash: sigma_net.{0,1}, color_net.{0,1,2};
        vm: basis_mat, color_net.{0,1,2})."""
        return list(self.ops.weight_grads(self.gw_ws).values())

    def grads(self):
        """{reference parameter name: gradient} for every trainable parameter of the model (loss-scaled).  With a gradient exchange
        attached these are the REDUCED gradients: the fp16 payload is written back over the rank-local fp32 buffer first."""
        if self.exchange is not None:
            self.exchange.write_back(self.grad_reduction)
        return self.ops.grads(self.gw_ws)

    def final_image(self):
        """pred rgb [N,3] and normalised depth [N] as run_cuda returns them (renderer.py:445-446)."""
        pred = self.image + (1 - self.weights_sum).unsqueeze(-1) * self.bg
        depth = torch.clamp(self.depth - self.nears, min=0) / (self.fars - self.nears + 1e-6)
        return pred, depth


class HashTrainEngine(FieldTrainEngine):
    """hash (INGP) teacher training -- BASELINE configs[1], the configuration bench.py times by default."""

    def __init__(self, field: "fused.HashNeRFField", *a, **k):
        assert getattr(field, "model_type", None) == "hash"
        super().__init__(field, *a, **k)


class VMTrainEngine(FieldTrainEngine):
    """vm (TensoRF VM-48) teacher training -- BASELINE configs[2]; `l1_reg_weight` defaults to the reference's 1e-4
    (main_just_train_tea.py:170)."""

    def __init__(self, field, *a, l1_reg_weight: float = 1e-4, **k):
        assert getattr(field, "model_type", None) == "vm"
        super().__init__(field, *a, l1_reg_weight=l1_reg_weight, **k)



class TensorsTrainEngine(FieldTrainEngine):
    """tensors (Plenoxels-style dense volume) teacher training (`main_just_train_tea.py --model_type tensors`)."""

    def __init__(self, field, *a, **k):
        assert getattr(field, "model_type", None) == "tensors"
        super().__init__(field, *a, **k)


class MLPTrainEngine(FieldTrainEngine):
    """mlp (NeRF) teacher training (`main_just_train_tea.py --model_type mlp`): fused forward that saves its operand tiles, backward =
    tail kernel -> trunk data gradients -> weight gradients on the tensor core (csrc/field_mlp_bwd.cu)."""

    def __init__(self, field, *a, **k):
        assert getattr(field, "model_type", None) == "mlp"
        super().__init__(field, *a, **k)


PAIR_SUM_STRIDE = 4   # PVD_PAIR_SUM_STRIDE
PAIR_PARALLEL_FWD = os.environ.get("PVD_PAIR_PARALLEL_FWD", "1") != "0"


class PvdPairRates(C.Structure):
    _fields_ = [("rgb", C.c_float), ("fea", C.c_float), ("color", C.c_float), ("sigma", C.c_float)]


class PairDistillEngine(FieldTrainEngine):
    """The distillation step of `main_distill_mutual.py` (Trainer.train_step, distill_mutual/utils.py:954-1189) for one
    (teacher, student) pair: the student's rays are marched ONCE, the frozen teacher (hash | mlp | vm) and the student (hash | vm)
    are queried at the SAME samples back to back (renderer.py:374-394: `inherited_params`), and the normL2 losses

        rate_rgb ||pred_tea - pred_stu|| + rate_fea ||feat_stu - feat_tea|| + rate_color ||color_l|| + rate_sigma ||sigma_l||

    (+ l1_reg_weight * density_loss() for a vm student, :1135-1136) drive the student's backward -- 10 launches, no Python between
    them, no [M,16] activation ever leaves the device-side buffers.  `stage` follows the reference's schedule: 1 = feature loss
    only, 2 = + colour and sigma (neither composites, :1046-1108), 3 = everything.  Defaults are the reference's
    (main_distill_mutual.py:174-178).  BASELINE configs 4 (hash -> vm) and 5 (mlp -> hash).
    """

    def __init__(self, teacher, student, bitfield: torch.Tensor, n_rays: int, rates=(1.0, 0.002, 0.002, 0.002), stage: int = 3,
                 l1_reg_weight: float = 1e-4, group=None, dist_sync: bool = True, **kw):
        # Ray-sharded data parallelism: the normL2 losses are norms over the GLOBAL batch, so the four sums of squares
        # are all-reduced (1 KB, between k_pair_composite and k_pair_combine) before the coefficients rate / ||.|| are formed; summing
        # the ranks' parameter gradients afterwards reproduces the single-process gradient (pvd_b200/dist.py::ShardedNormL2).
        import torch.distributed as dist
        self._dist = dist if (dist_sync and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1) else None
        self._group = group
        self.teacher_field = teacher
        for p_ in teacher.parameters():
            p_.requires_grad_(False)               # main_distill_mutual.py:320-321 (and it lets the teacher's staged copies be cached)
        self.distill_stage = int(stage)
        r = [float(v) for v in rates]
        if self.distill_stage == 1:
            r = [0.0, r[1], 0.0, 0.0]
        elif self.distill_stage == 2:
            r[0] = 0.0
        self.rates = PvdPairRates(*r)
        if getattr(student, "model_type", None) != "vm" or self.distill_stage != 3:
            l1_reg_weight = 0.0
        super().__init__(student, bitfield, n_rays, l1_reg_weight=l1_reg_weight, **kw)
        if self._dist is not None and self.l1_reg_weight:
            # a parameter-only term must enter the SUM over ranks once: every rank contributes 1/W of it
            self.l1_reg_weight /= self._dist.get_world_size(self._group)

    n_host_inputs = 2   # rays_o, rays_d: the teacher's rendering is the target
    grad_reduction = "sum"   # the normL2 coefficients rate / ||.|| are formed from GLOBAL sums: rank gradients simply add up

    def _loss_dev(self):
        return self.loss_out

    def _init_small_buffers(self):
        n_sum = PAIR_SUM_STRIDE * fused.LOSS_SLOTS
        self._zeros = torch.zeros(n_sum + 2 * fused.LOSS_SLOTS + fused.GW_WS_FLOATS, dtype=torch.float32, device=self.dev)
        self.pair_sums = self._zeros[0:n_sum]                                  # [64][4]: rgb, feature, colour, sigma sums of squares
        self.loss_slots = self._zeros[n_sum:n_sum + 2 * fused.LOSS_SLOTS]      # parameter-only terms (vm L1 penalty)
        self.gw_ws = self._zeros[n_sum + 2 * fused.LOSS_SLOTS:]
        self.loss_out = torch.zeros(5, dtype=torch.float32, device=self.dev)   # total, ||rgb||, ||fea||, ||color||, ||sigma||
        self.pred_tea = torch.empty(self.N, 3, device=self.dev)
        if self._dist is not None:   # create the communicator now: it cannot be created inside a CUDA-graph capture
            self._dist.all_reduce(self.pair_sums, group=self._group)
            self.pair_sums.zero_()

    def _init_fields(self):
        self.ops = field_ops.make_ops(self.field, self.dev, trainable=True)
        self.tea = field_ops.make_ops(self.teacher_field, self.dev, trainable=False)
        # count, scan, expand, tail | teacher fwd, student fwd | sample_sq, (composite), combine | student bwd
        self.launches_per_step = 4 + self.tea.kernels_fwd + self.ops.kernels_fwd + (3 if self.distill_stage == 3 else 2) + self.ops.kernels_bwd
        if self.l1_reg_weight:
            self.launches_per_step += 6

    def _alloc_field_samples(self, M: int):
        d = self.dev
        self.ops.alloc(M)
        self.tea.alloc(M)
        self.sigmas_tea = torch.empty(M, device=d)
        self.rgbs_tea = torch.empty(M, 3, device=d)
        self.feat_tea = torch.empty(M, 16, device=d)
        self.feat = torch.empty(M, 16, device=d)
        self.grad_feat = torch.zeros(M, 16, device=d)

    def _prefetch(self, st):
        self.ops.prefetch(st)
        self.tea.prefetch(st)

    def stage(self):
        self.ops.stage(self.density_scale)
        self.tea.stage(self.density_scale)

    def _march_write(self, st, rs, M_drop):
        super()._march_write(st, rs, M_drop)
        # rows no surviving ray owns are zeros in the reference (fresh torch.zeros every call) and both networks see them
        M = rs.xyzs.shape[0]
        nv.check(nv.lib().pvd_zero_sample_tail(nv.ptr(rs.rays), nv.ptr(rs.counter), _u32(self.N), _u32(min(M, M_drop)), nv.ptr(rs.xyzs),
                                               nv.ptr(rs.dirs), nv.ptr(rs.deltas), st))

    def _forward(self, st, rs, M, M_drop):
        """Teacher and student are independent given the samples: the teacher's query runs on a parallel branch (its gathers /
        GEMMs and the student's are bound by different units).  PVD_PAIR_PARALLEL_FWD=0 issues them back to back."""
        if not PAIR_PARALLEL_FWD:
            self.tea.forward(st, rs.xyzs, rs.dirs, M, self.sigmas_tea, self.rgbs_tea, self.feat_tea, self.status)
            self.ops.forward(st, rs.xyzs, rs.dirs, M, self.sigmas, self.rgbs, self.feat, self.status)
            return
        cur = torch.cuda.current_stream(self.dev)
        if not hasattr(self, "_side4"):
            self._side4 = torch.cuda.Stream(device=self.dev)
        self._side4.wait_stream(cur)
        with torch.cuda.stream(self._side4):
            self.tea.forward(C.c_void_p(self._side4.cuda_stream), rs.xyzs, rs.dirs, M, self.sigmas_tea, self.rgbs_tea, self.feat_tea,
                             self.status)
        self.ops.forward(st, rs.xyzs, rs.dirs, M, self.sigmas, self.rgbs, self.feat, self.status)
        cur.wait_stream(self._side4)

    def _loss_backward(self, st, rs, M, M_drop):
        l = nv.lib()
        nv.check(l.pvd_pair_sample_sq(nv.ptr(self.feat_tea), nv.ptr(self.feat), nv.ptr(self.rgbs_tea), nv.ptr(self.rgbs), _u32(M),
                                      nv.ptr(self.pair_sums), st))
        if self.distill_stage == 3:
            nv.check(l.pvd_pair_composite(nv.ptr(self.bg), nv.ptr(self.sigmas_tea), nv.ptr(self.rgbs_tea), nv.ptr(self.sigmas),
                                          nv.ptr(self.rgbs), nv.ptr(rs.deltas), nv.ptr(rs.rays), _u32(M_drop), _u32(self.N),
                                          nv.ptr(self.pred_tea), nv.ptr(self.weights_sum), nv.ptr(self.depth), nv.ptr(self.image),
                                          nv.ptr(self.grad_sigmas), nv.ptr(self.grad_rgbs), nv.ptr(self.pair_sums), st))
        if self._dist is not None:
            self._dist.all_reduce(self.pair_sums, group=self._group)   # global sums of squares (the slots add up the same way)
        nv.check(l.pvd_pair_combine(nv.ptr(self.feat_tea), nv.ptr(self.feat), nv.ptr(self.rgbs_tea), nv.ptr(self.rgbs),
                                    nv.ptr(self.pair_sums), C.byref(self.rates), _f32(self.loss_scale), _u32(M),
                                    nv.ptr(rs.counter) if self.distill_stage == 3 else None, nv.ptr(self.grad_sigmas),
                                    nv.ptr(self.grad_rgbs), nv.ptr(self.grad_feat), nv.ptr(self.loss_out), st))

    def _field_backward(self, st, rs, M, cur=None):
        # all M rows: the per-sample losses cover the padding rows too
        self.ops.backward(st, rs.xyzs, rs.dirs, self.grad_sigmas, self.grad_rgbs, self.grad_feat, M, None, self.gw_ws, self.status)

    @property
    def loss(self):
        """[total loss (un-scaled, L1 penalty included), rays] -- same shape of answer as FieldTrainEngine.loss."""
        reg = self.loss_slots.view(fused.LOSS_SLOTS, 2).sum(0)[0]
        return torch.stack([self.loss_out[0] + reg, torch.tensor(float(self.N), device=self.dev)])

    def loss_terms(self):
        """The un-weighted norms the reference's trainer logs: rgb, feature, colour, sigma (utils.py:1177-1187)."""
        return {k: float(v) for k, v in zip(("rgb", "fea", "color", "sigma"), self.loss_out[1:5].tolist())}

    def final_images(self):
        """(student pred [N,3], teacher pred [N,3]) with the background mixed in (renderer.py:445)."""
        pred = self.image + (1 - self.weights_sum).unsqueeze(-1) * self.bg
        return pred, self.pred_tea
