"""2+-GPU probe: what does the gradient exchange cost?  NCCL all-reduce of the table gradient (fp32 / fp16 payload) vs
torch symmetric-memory collectives, and whether CUDA-IPC peer mappings work between torchrun ranks on this box.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/micro/allreduce_probe.py"""
import os
import time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 5303704 * 2


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3


g32 = torch.randn(N, device=dev)
g16 = torch.empty(N, dtype=torch.float16, device=dev)
small = torch.randn(163840, device=dev)
res = {}
res["nccl fp32 42MB"] = timeit(lambda: dist.all_reduce(g32))
res["nccl fp16 21MB"] = timeit(lambda: dist.all_reduce(g16))
res["cast fp32->fp16"] = timeit(lambda: g16.copy_(g32))
res["nccl fp32 640KB"] = timeit(lambda: dist.all_reduce(small))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(N, dtype=torch.float16, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.normal_()
    for name in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_", "multimem_one_shot_all_reduce"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            res["symm " + name] = timeit(lambda: op(t, "sum", dist.group.WORLD.group_name))
        except Exception as ex:  # noqa: BLE001
            res["symm " + name] = "ERR " + repr(ex)[:120]
    res["symm multicast"] = str(getattr(hdl, "multicast_ptr", None))
except Exception as ex:  # noqa: BLE001
    res["symm"] = "ERR " + repr(ex)[:200]
# CUDA IPC between ranks
try:
    buf = torch.full((1024,), float(rank + 1), device=dev)
    h = buf.untyped_storage()._share_cuda_()
    objs = [None] * world
    dist.all_gather_object(objs, h)
    peer = (rank + 1) % world
    ph = objs[peer]
    st = torch.UntypedStorage._new_shared_cuda(*ph)
    pt = torch.empty(0, dtype=torch.float32, device=st.device).set_(st)
    torch.cuda.synchronize(); dist.barrier()
    res["ipc peer value"] = float(pt[0].item())
    res["ipc peer device"] = str(pt.device)
    res["can_access_peer"] = torch.cuda.can_device_access_peer(local, peer)
except Exception as ex:  # noqa: BLE001
    res["ipc"] = "ERR " + repr(ex)[:200]
if rank == 0:
    for k, v in res.items():
        print(f"{k:34s}: {v if isinstance(v, str) else round(v, 1) if isinstance(v, float) else v}  (us for timings)")
dist.barrier()
dist.destroy_process_group()
