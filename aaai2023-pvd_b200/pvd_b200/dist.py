"""Ray-sharded data parallelism for the hot path: one process per GPU, parameters replicated, rays split,
one all-reduce of the gradient buffers per step.  The reference has no multi-GPU path (tools/details.md:25).

Device-agnostic on purpose: the same helpers run over NCCL on the GPUs and over gloo in the CPU tests.

Loss normalisation.  Per-ray mean losses (MSE, just_train_tea/utils.py:841-846) decompose over shards: every rank uses
1/(3*N_global) as its normaliser and the gradient all-reduce is a plain SUM.  The distillation default
`torch.norm(a - b)` (distill_mutual/utils.py:945,1111) does not: ||x|| needs the global sum of squares first, so ranks
all-reduce one scalar (`global_l2_norm`) and back-propagate x / ||x||_global locally.
"""
from __future__ import annotations

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
    """The per-step gradient exchange of the ray-sharded training path: sum over ranks of the hash table's gradient (fp16
    payload) and of the small MLP-gradient workspace (fp32).

    Fast path (NVSwitch, CUDA ranks): the payload lives in a torch symmetric-memory buffer that is also mapped through a
    multicast address; after the cast each rank reduces ITS 1/W of the buffer in the switch with `pvd_multimem_allreduce_f16`
    (multimem.ld_reduce / multimem.st, csrc/collective.cu) between two signal-pad barriers.  The small fp32 workspace goes
    through NCCL on a side stream at the same time.  Anything that does not set up (no NVLS, an older torch, one GPU) falls back
    to NCCL for both, with the same result layout: `self.payload` holds the reduced fp16 table gradient.
    """

    def __init__(self, grad_table: torch.Tensor, small: torch.Tensor, mode: str = "auto", group=None):
        import ctypes as C
        self._C = C
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = grad_table.numel()
        self.small = small
        self.grad_table = grad_table
        self.kind = "nccl"
        self.why = ""
        self._side = torch.cuda.Stream(device=grad_table.device)
        self.payload = None
        if mode in ("auto", "multimem") and self.world > 1 and self.n % 8 == 0:
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(self.n, dtype=torch.float16, device=grad_table.device)
                hdl = symm.rendezvous(buf, self.group.group_name)
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                if mc == 0:
                    raise RuntimeError("no multicast mapping (NVLS unavailable)")
                self.payload, self._hdl, self._mc = buf, hdl, mc
                vecs = self.n // 8
                lo = (vecs * self.rank) // self.world
                hi = (vecs * (self.rank + 1)) // self.world
                self._off, self._cnt = 8 * lo, 8 * (hi - lo)
                self.kind = "multimem"
            except Exception as ex:  # noqa: BLE001  (any failure: NCCL carries the exchange)
                self.why = repr(ex)[:160]
                if mode == "multimem":
                    raise
        if self.payload is None:
            self.payload = torch.empty(self.n, dtype=torch.float16, device=grad_table.device)

    def __call__(self):
        from . import _native as nv
        C = self._C
        cur = torch.cuda.current_stream(self.grad_table.device)
        st = C.c_void_p(cur.cuda_stream)
        # the small fp32 workspace: NCCL on a side stream, concurrent with the table exchange
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.small, group=self.group)
        nv.check(nv.lib().pvd_cast_f32_to_f16(nv.ptr(self.grad_table), nv.ptr(self.payload), C.c_uint64(self.n), st))
        if self.kind == "multimem":
            self._hdl.barrier(channel=0)     # every rank's payload is written
            nv.check(nv.lib().pvd_multimem_allreduce_f16(C.c_void_p(self._mc), C.c_uint64(self._off), C.c_uint64(self._cnt), st))
            self._hdl.barrier(channel=1)     # every rank's shard of sums is visible everywhere
        else:
            dist.all_reduce(self.payload, group=self.group)
        cur.wait_stream(self._side)
