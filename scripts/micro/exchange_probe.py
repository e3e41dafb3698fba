"""2+-GPU timing of the table-gradient exchange variants (pvd_b200.dist.TableGradExchange) in isolation."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "aaai2023-pvd_b200"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from pvd_b200.dist import TableGradExchange
import ctypes as C
from pvd_b200 import _native as nv
g = torch.randn(5303704 * 2, device=dev); small = torch.randn(163840, device=dev)


def timeit(fn, n=40, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3


out = {}
for mode in ("nccl", "auto"):
    ex = TableGradExchange(g, small, mode=mode)
    out[f"exchange[{ex.kind}] total"] = timeit(ex)
    if ex.kind == "multimem":
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        out["  cast only"] = timeit(lambda: nv.check(nv.lib().pvd_cast_f32_to_f16(nv.ptr(g), nv.ptr(ex.payload), C.c_uint64(ex.n), st)))
        out["  barrier only"] = timeit(lambda: ex._hdl.barrier(channel=0))
        out["  multimem kernel only"] = timeit(lambda: nv.check(nv.lib().pvd_multimem_allreduce_f16(C.c_void_p(ex._mc), C.c_uint64(ex._off), C.c_uint64(ex._cnt), st)))
        out["  small nccl only"] = timeit(lambda: dist.all_reduce(small))
if rank == 0:
    for k, v in out.items(): print(f"{k:32s}: {v:8.1f} us")
dist.barrier(); dist.destroy_process_group()
