"""Ray-sharded data parallelism for the hot path: one process per GPU, parameters replicated, rays split,
one all-reduce of the gradient buffers per step.  The reference has no multi-GPU path (tools/details.md:25).

Device-agnostic on purpose: the same helpers run over NCCL on the GPUs and over gloo in the CPU tests.

Loss normalisation.  Per-ray mean losses (MSE, just_train_tea/utils.py:841-846) decompose over shards: every rank uses
1/(3*N_global) as its normaliser and the gradient all-reduce is a plain SUM.  The distillation default
`torch.norm(a - b)` (distill_mutual/utils.py:945,1111) does not: ||x|| needs the global sum of squares first, so ranks
all-reduce one scalar (`global_l2_norm`) and back-propagate x / ||x||_global locally.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int):
    """Contiguous shard [lo, hi) of n rays for `rank`; the first n % world ranks take one extra ray."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, rank: int, world_size: int, *extra):
    """Slice [N,3] ray tensors (and any per-ray extras such as ground-truth colours) for this rank."""
    lo, hi = shard_bounds(rays_o.shape[0], rank, world_size)
    out = [rays_o[lo:hi].contiguous(), rays_d[lo:hi].contiguous()]
    out += [e[lo:hi].contiguous() for e in extra]
    return out


def allreduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of a list of gradient buffers (large ones individually, small ones coalesced)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    small = [t for t in tensors if t.numel() < (1 << 16)]
    for t in tensors:
        if t.numel() >= (1 << 16):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if small:
        flat = torch.cat([t.reshape(-1) for t in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t in small:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
    return tensors


def global_l2_norm(local_sq_sum: torch.Tensor, group=None) -> torch.Tensor:
    """sqrt of the all-reduced sum of squares: the value of torch.norm over the concatenation of every rank's shard."""
    s = local_sq_sum.detach().clone().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return torch.sqrt(s)[0]


class ShardedNormL2(torch.autograd.Function):
    """loss = ||concat_r(x_r)||_2 with x_r living on rank r.  forward returns the GLOBAL norm on every rank; backward gives
    d loss / d x_r = x_r / ||x||_global, so summing parameter gradients over ranks reproduces the single-process gradient."""

    @staticmethod
    def forward(ctx, x, group=None):
        n = global_l2_norm((x.detach().float() ** 2).sum(), group)
        ctx.save_for_backward(x, n)
        return n.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        x, n = ctx.saved_tensors
        return g * x / torch.clamp(n, min=1e-20).to(x.dtype), None


def sharded_norm_l2(x, group=None):
    return ShardedNormL2.apply(x, group)


class TableGradExchange:
    """The per-step gradient exchange of the ray-sharded training path: the big parameter gradient (hash table / vm planes) as an
    fp16 payload, the small MLP-gradient workspace in fp32.

    Payload convention: every rank casts `grad * (1 / W)` (saturating at +-65504: a finite gradient never becomes inf on the way) and
    the payloads are SUMMED, so `self.payload` ends up holding the MEAN over ranks of the (loss-scaled) rank gradients -- for per-ray
    mean losses (every rank normalises by its own ray count) that IS the global-batch gradient; for losses whose coefficients are
    already global (PairDistillEngine: rank gradients add up) the consumer multiplies by W.  `small` is summed as it is.
    Consumers: `write_back(reduction)` puts the reduced gradient back into the fp32 buffers (`engine.grads()` does it), and
    `optim.for_engine(..., exchange=...)` reads the payload directly with the right factors.

    Kinds: "nccl" (cast kernel + ncclAllReduce of the payload); "p2p" (the payload lives in a torch symmetric-memory buffer; each
    rank loads ITS 1/W shard from every rank over NVLink, sums in fp32 and stores the result into every rank's buffer -- one kernel,
    device-side barriers, csrc/collective.cu::k_p2p_allreduce_f16; what mode "auto" picks on 2 / 4 ranks); "multimem" (what "auto" picks on 8: the same
    shard reduced BY THE SWITCH through the buffer's multicast address, multimem.ld_reduce / multimem.st, barriers inside the kernel
    (`fused_barrier`) or as two signal-pad launches).  Anything that does not set up falls back to NCCL with the same result layout.
    """

    def __init__(self, grad_table: torch.Tensor, small: torch.Tensor, mode: str = "auto", group=None, fused_barrier: bool = True,
                 blocks: int = 0, unroll: int = 4):
        import ctypes as C
        self._C = C
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = grad_table.numel()
        self.small = small
        self.grad_table = grad_table
        self.kind = "nccl"
        self.why = ""
        self.pre_scale = 1.0 / self.world
        self.fused_barrier, self.blocks, self.unroll = bool(fused_barrier), int(blocks), int(unroll)
        self.weak = os.environ.get("PVD_P2P_WEAK", "1") != "0"
        self._side = torch.cuda.Stream(device=grad_table.device) if grad_table.is_cuda else None
        self.payload = None
        if mode in ("auto", "multimem", "p2p") and self.world > 1 and self.n % 8 == 0 and grad_table.is_cuda:
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(self.n, dtype=torch.float16, device=grad_table.device)
                hdl = symm.rendezvous(buf, self.group.group_name)
                vecs = self.n // 8
                lo = (vecs * self.rank) // self.world
                hi = (vecs * (self.rank + 1)) // self.world
                self._off, self._cnt = 8 * lo, 8 * (hi - lo)
                self._pads = int(getattr(hdl, "signal_pad_ptrs_dev", 0) or 0)
                self._bufs = int(getattr(hdl, "buffer_ptrs_dev", 0) or 0)
                self._local = torch.zeros(4, dtype=torch.int32, device=grad_table.device)
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                # "auto": measured totals (cast + kernel + small, scripts/micro/exchange_probe.py, profiles/r02_exchange_probe_N*):
                # 2 ranks p2p 73 vs multimem 98 us; 4 ranks 85 vs 88; 8 ranks 98 vs 90 -- the switch reduction wins once 7/8 of the
                # payload would have to cross the links in both directions
                p2p_ok = bool(self._pads and self._bufs and self.world in (2, 4, 8))
                mm_ok = bool(mc and self._pads)
                if mode == "auto":
                    want_p2p = p2p_ok and not (self.world >= 8 and mm_ok)
                    if not want_p2p and not mm_ok:
                        raise RuntimeError("neither peer pointers nor a multicast mapping are available")
                else:
                    want_p2p = mode == "p2p"
                if self._pads:
                    # the kernels' barriers are monotonic epoch flags in slots [12W, 13W) of the signal pads, counted from this
                    # rank's local[3] = 0: start from zeroed slots on every rank (a pad may be recycled from an earlier buffer)
                    pad = hdl.get_signal_pad(self.rank)
                    pad.view(-1).view(torch.int32)[12 * self.world:13 * self.world].zero_()
                    torch.cuda.synchronize(grad_table.device)
                    dist.barrier(group=self.group)
                if want_p2p:
                    if not p2p_ok:
                        raise RuntimeError("peer pointers / signal pads not exposed by this torch, or world not in (2, 4, 8)")
                    self.kind = "p2p"
                else:
                    if mc == 0:
                        raise RuntimeError("no multicast mapping (NVLS unavailable)")
                    if self._pads == 0:
                        self.fused_barrier = False
                    self.kind = "multimem"
                    if self.blocks == 0:
                        self.blocks, self.unroll = 64, 2
                self.payload, self._hdl, self._mc = buf, hdl, mc
            except Exception as ex:  # noqa: BLE001  (any failure: NCCL carries the exchange)
                self.why = repr(ex)[:160]
                if mode in ("multimem", "p2p"):
                    raise
        if self.payload is None:
            self.payload = torch.empty(self.n, dtype=torch.float16, device=grad_table.device)

    # the factor that turns the payload (mean over ranks) into what the optimizer needs
    def result_scale(self, reduction: str) -> float:
        return 1.0 if reduction == "mean" else float(self.world)

    def __call__(self):
        from . import _native as nv
        C = self._C
        cur = torch.cuda.current_stream(self.grad_table.device)
        st = C.c_void_p(cur.cuda_stream)
        # the small fp32 workspace: NCCL on a side stream, concurrent with the table exchange
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.small, group=self.group)
        nv.check(nv.lib().pvd_cast_f32_to_f16_scaled(nv.ptr(self.grad_table), nv.ptr(self.payload), C.c_uint64(self.n),
                                                     C.c_float(self.pre_scale), st))
        if self.kind == "p2p":
            nv.check(nv.lib().pvd_p2p_allreduce_f16(C.c_void_p(self._bufs), C.c_uint64(self._off), C.c_uint64(self._cnt), C.c_void_p(self._pads),
                                                    C.c_uint32(self.rank), C.c_uint32(self.world), nv.ptr(self._local), C.c_uint32(self.blocks),
                                                    C.c_uint32(self.unroll), C.c_uint32(1 if self.weak else 0), st))
        elif self.kind == "multimem" and self.fused_barrier:
            nv.check(nv.lib().pvd_multimem_allreduce_f16_fused(C.c_void_p(self._mc), C.c_uint64(self._off), C.c_uint64(self._cnt),
                                                               C.c_void_p(self._pads), C.c_uint32(self.rank), C.c_uint32(self.world),
                                                               nv.ptr(self._local), C.c_uint32(self.blocks), C.c_uint32(self.unroll), st))
        elif self.kind == "multimem":
            self._hdl.barrier(channel=0)     # every rank's payload is written
            nv.check(nv.lib().pvd_multimem_allreduce_f16(C.c_void_p(self._mc), C.c_uint64(self._off), C.c_uint64(self._cnt), st))
            self._hdl.barrier(channel=1)     # every rank's shard of sums is visible everywhere
        else:
            dist.all_reduce(self.payload, group=self.group)
        cur.wait_stream(self._side)

    def barrier_error(self) -> int:
        """Non-zero if a device-side barrier of the fused kernel timed out (host sync)."""
        return int(self._local[2].item()) if self.kind in ("multimem", "p2p") else 0

    def write_back(self, reduction: str = "mean"):
        """Reduced gradient -> the fp32 buffers a caller's optimizer reads: grad_table = payload * result_scale; the small workspace
        (already summed in fp32) is divided by W for mean-type losses.  Call once per exchange."""
        from . import _native as nv
        C = self._C
        st = C.c_void_p(torch.cuda.current_stream(self.grad_table.device).cuda_stream)
        nv.check(nv.lib().pvd_cast_f16_to_f32(nv.ptr(self.payload), nv.ptr(self.grad_table), C.c_uint64(self.n),
                                              C.c_float(self.result_scale(reduction)), st))
        if reduction == "mean":
            self.small.mul_(1.0 / self.world)


def reduce_gradients_reference(rank_grads, reduction: str):
    """What the exchange + write_back compute, in plain fp32 torch, for tests: list of per-rank tensors -> global-batch gradient."""
    total = torch.stack([g.float() for g in rank_grads]).sum(0)
    return total / len(rank_grads) if reduction == "mean" else total
