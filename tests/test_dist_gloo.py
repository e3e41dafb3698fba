"""world_size-2 gloo tests (CPU) of the ray-sharding host logic: shard gradients summed over ranks equal the single-process
gradients for the MSE-mean loss and for the non-decomposable norm loss of the distillation trainer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy_field(w, rays):
    # a differentiable stand-in for render(): per-ray colour from shared parameters
    return torch.sigmoid(rays @ w)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pvd_b200 import dist as pd
    torch.manual_seed(0)
    N = 1001  # not divisible by the world size
    rays, gt, tea = torch.randn(N, 3), torch.rand(N, 3), torch.rand(N, 3)
    w = torch.randn(3, 3, requires_grad=True)
    big = torch.zeros(70000, requires_grad=True)  # exercises the un-coalesced branch
    lo, hi = pd.shard_bounds(N, rank, world)
    r, g, t = pd.shard_rays(rays, rays, rank, world, gt, tea)[0::1][0], None, None
    r, _, g, t = pd.shard_rays(rays, rays, rank, world, gt, tea)
    assert r.shape[0] == hi - lo
    # (1) MSE-mean with the GLOBAL normaliser, gradients summed
    pred = _toy_field(w, r)
    loss = ((pred - g) ** 2).sum() / (3 * N) + (big[lo:hi] ** 2).sum() * 0 + big[lo:hi].sum() / N
    loss.backward()
    g_mse = [w.grad.clone(), big.grad.clone()]
    pd.allreduce_sum_(g_mse)
    # (2) norm loss through the sharded norm
    w.grad = None
    n = pd.sharded_norm_l2(_toy_field(w, r) - t)
    n.backward()
    g_norm = [w.grad.clone()]
    pd.allreduce_sum_(g_norm)
    if rank == 0:
        q.put((g_mse[0], g_mse[1], float(n), g_norm[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_gradients_equal_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    torch.manual_seed(0)
    N = 1001
    rays, gt, tea = torch.randn(N, 3), torch.rand(N, 3), torch.rand(N, 3)
    w = torch.randn(3, 3, requires_grad=True)
    big = torch.zeros(70000, requires_grad=True)
    loss = ((_toy_field(w, rays) - gt) ** 2).mean() + big[:N].sum() / N
    loss.backward()
    torch.testing.assert_close(got[0], w.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got[1], big.grad, rtol=1e-6, atol=1e-7)
    w.grad = None
    n = torch.norm(_toy_field(w, rays) - tea)
    n.backward()
    assert abs(got[2] - float(n)) < 1e-5 * float(n)
    torch.testing.assert_close(got[3], w.grad, rtol=1e-5, atol=1e-6)


def test_shard_bounds_cover_everything():
    from pvd_b200.dist import shard_bounds
    for n in (0, 1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
