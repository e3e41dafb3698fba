"""N-GPU timing of the table-gradient exchange variants (pvd_b200.dist.TableGradExchange) in isolation: 21 MB fp16 payload
(the L=14 hash table's gradient) + the 640 KB fp32 weight-gradient workspace.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 scripts/micro/exchange_probe.py

Prints a table on rank 0 and writes gpurun_out/exchange_probe_N<world>.json (median microseconds over 40 calls, CUDA events on the
launching stream, max over ranks)."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "aaai2023-pvd_b200"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from pvd_b200 import _native as nv  # noqa: E402
from pvd_b200.dist import TableGradExchange  # noqa: E402

g = torch.randn(5303704 * 2, device=dev)
small = torch.randn(163840, device=dev)


def timeit(fn, n=40, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ts[len(ts) // 2] * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"world": world, "payload_mb": g.numel() * 2 / 1e6}
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
# reference points
ex_nccl = TableGradExchange(g, small, mode="nccl")
out["nccl: cast + all-reduce(fp16) + small"] = timeit(ex_nccl)
out["  cast only"] = timeit(lambda: nv.check(nv.lib().pvd_cast_f32_to_f16_scaled(nv.ptr(g), nv.ptr(ex_nccl.payload), C.c_uint64(ex_nccl.n), C.c_float(1.0 / world), st())))
out["  nccl all-reduce(fp16 payload) only"] = timeit(lambda: dist.all_reduce(ex_nccl.payload))
out["  nccl all-reduce(small fp32) only"] = timeit(lambda: dist.all_reduce(small))
g32 = g.clone()
out["nccl all-reduce(fp32 table, 42 MB)"] = timeit(lambda: dist.all_reduce(g32))
del g32
try:
    ex = TableGradExchange(g, small, mode="multimem", fused_barrier=False)
    out["multimem, barriers as launches: total"] = timeit(ex)
    out["  signal-pad barrier launch only"] = timeit(lambda: ex._hdl.barrier(channel=0))
    out["  multimem kernel only (no barriers)"] = timeit(lambda: nv.check(nv.lib().pvd_multimem_allreduce_f16(C.c_void_p(ex._mc), C.c_uint64(ex._off), C.c_uint64(ex._cnt), st())))
    if ex._pads:
        ex.fused_barrier = True
        best = None
        for blocks in (16, 32, 64, 148, 296):
            for unroll in (2, 4, 8):
                ex.blocks, ex.unroll = blocks, unroll
                k = lambda: nv.check(nv.lib().pvd_multimem_allreduce_f16_fused(C.c_void_p(ex._mc), C.c_uint64(ex._off), C.c_uint64(ex._cnt), C.c_void_p(ex._pads),
                                                                               C.c_uint32(rank), C.c_uint32(world), nv.ptr(ex._local), C.c_uint32(blocks), C.c_uint32(unroll), st()))
                t = timeit(k, n=20, warm=3)
                out[f"  fused-barrier kernel only, {blocks} CTAs x unroll {unroll}"] = t
                if ex.barrier_error():
                    out[f"  !! barrier error at {blocks}x{unroll}"] = ex.barrier_error()
                if best is None or t < best[0]:
                    best = (t, blocks, unroll)
        ex.blocks, ex.unroll = best[1], best[2]
        out[f"multimem, barriers inside the kernel ({best[1]} CTAs x {best[2]}): total"] = timeit(ex)
        # correctness of the tuned configuration against NCCL on the same data
        ref = (g * ex.pre_scale).clamp(-65504, 65504).to(torch.float16)
        dist.all_reduce(ref)
        ex()
        torch.cuda.synchronize()
        out["fused-barrier result vs NCCL: max rel err"] = float((ex.payload.float() - ref.float()).abs().max() / ref.float().abs().max())
        out["best"] = {"blocks": best[1], "unroll": best[2], "kernel_us": best[0]}
    else:
        out["fused barrier"] = "signal_pad_ptrs_dev not exposed by this torch"
except Exception as e:  # noqa: BLE001
    out["multimem"] = f"unavailable: {e!r}"[:200]
try:
    exp = TableGradExchange(g, small, mode="p2p")
    bestp = None
    small_cnt = 8 * world      # barrier cost alone: a shard of one vector per rank
    exp.blocks, exp.unroll, exp.weak = 148, 1, True
    out["  p2p kernel, barriers only (one vector)"] = timeit(lambda: nv.check(nv.lib().pvd_p2p_allreduce_f16(
        C.c_void_p(exp._bufs), C.c_uint64(8 * rank), C.c_uint64(8), C.c_void_p(exp._pads), C.c_uint32(rank), C.c_uint32(world), nv.ptr(exp._local),
        C.c_uint32(148), C.c_uint32(1), C.c_uint32(1), st())), n=20, warm=3)
    for mode, name in ((0, "pull (remote loads)"), (1, "push (remote stores)"), (2, "pull + push at once")):
        t = timeit(lambda: nv.check(nv.lib().pvd_p2p_copy_probe(C.c_void_p(exp._bufs), C.c_uint64(exp.n), C.c_uint32(rank), C.c_uint32(world), C.c_uint32(mode),
                                                                C.c_uint32(148), st())), n=20, warm=3)
        out[f"  link probe, {name}: 21.2 MB to/from the next rank"] = t
        out[f"  link probe, {name}: GB/s per direction"] = exp.n * 2 / (t * 1e-6) / 1e9 / (2 if mode == 2 else 1)
    for weak in (1,):
        for unroll in (2,):
            for blocks in (74, 148, 296):
                k = lambda: nv.check(nv.lib().pvd_p2p_allreduce_f16(C.c_void_p(exp._bufs), C.c_uint64(exp._off), C.c_uint64(exp._cnt), C.c_void_p(exp._pads),
                                                                    C.c_uint32(rank), C.c_uint32(world), nv.ptr(exp._local), C.c_uint32(blocks),
                                                                    C.c_uint32(unroll), C.c_uint32(weak), st()))
                t = timeit(k, n=20, warm=3)
                out[f"  p2p kernel only, {'plain' if weak else 'relaxed.sys'} accesses, unroll {unroll}, {blocks} CTAs x 512"] = t
                if exp.barrier_error():
                    out[f"  !! p2p barrier error at {weak}/{unroll}/{blocks}"] = exp.barrier_error()
                if bestp is None or t < bestp[0]:
                    bestp = (t, blocks, unroll, weak)
    exp.unroll, exp.weak = bestp[2], bool(bestp[3])
    exp.blocks = bestp[1]
    out[f"p2p two-shot ({bestp[1]} CTAs): cast + kernel + small, total"] = timeit(exp)
    ref = (g * exp.pre_scale).clamp(-65504, 65504).to(torch.float16)
    dist.all_reduce(ref)
    exp()
    torch.cuda.synchronize()
    out["p2p result vs NCCL: max rel err"] = float((exp.payload.float() - ref.float()).abs().max() / ref.float().abs().max())
    out["best_p2p"] = {"blocks": bestp[1], "unroll": bestp[2], "weak": bestp[3], "kernel_us": bestp[0]}
except Exception as e:  # noqa: BLE001
    out["p2p"] = f"unavailable: {e!r}"[:200]
if rank == 0:
    for k, v in out.items():
        print(f"{k:72s}: {v:8.1f} us" if isinstance(v, float) else f"{k:72s}: {v}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"exchange_probe_N{world}.json"), "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
