"""The graph-free training-step engine (what bench.py times) against the autograd path built from the drop-in operators, and
against the CPU oracle's training step; plus CUDA-graph replay equivalence."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _setup(scene, n_rays=1024, seed=0):
    from pvd_b200.engine import HashTrainEngine
    from pvd_b200.fused import HashNeRFField
    torch.manual_seed(seed)
    net = HashNeRFField(num_levels=14, desired_resolution=2048).cuda()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    eng = HashTrainEngine(net, torch.from_numpy(scene["bitfield"]), n_rays, loss_scale=512.0)
    eng.stage()
    ro, rd = scene["batches"][0]
    ro, rd = ro[:n_rays].contiguous(), rd[:n_rays].contiguous()
    gt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(3))
    eng.rays_o.copy_(ro); eng.rays_d.copy_(rd); eng.gt.copy_(gt)
    return net, eng, ro, rd, gt


def test_engine_step_matches_autograd_path_and_oracle(scene):
    import raymarching
    from oracle import field
    net, eng, ro, rd, gt = _setup(scene)
    eng.step(warmup=True)   # sizes the sample buffers from the counter, like the reference's first iterations
    eng.finish_warmup()
    eng.step()              # steady state: M = ceil128(mean_count)
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0
    loss_e = float(eng.loss[0].item())
    gw_e = [g.clone() for g in eng.grad_weights()]
    gt_e = eng.grad_table.clone()
    pred_e, depth_e = eng.final_image()
    # the same step through the drop-in operators + autograd, with the same M (mean_count) so the same rays are dropped
    net.train()
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device="cuda")
    nears, fars = raymarching.near_far_from_aabb(ro.cuda(), rd.cuda(), aabb, 0.2)
    assert torch.equal(nears, eng.nears) and torch.equal(fars, eng.fars)   # fused near/far is the same arithmetic
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro.cuda(), rd.cuda(), 1.0, eng.bitfield, 1, 128, nears, fars, counter,
                                                            eng.mean_count, True, 128, False, 0.0, 1024)
    assert xyzs.shape[0] == eng.M and torch.equal(rays, eng.rays)
    sigma, color = net(xyzs, dirs)
    ws, depth, image = raymarching.composite_rays_train(sigma, color, deltas, rays)
    pred = image + (1 - ws).unsqueeze(-1)
    loss = torch.mean((pred - gt.cuda()) ** 2)
    (loss * 512.0).backward()
    assert abs(loss_e - float(loss)) < 1e-5 * max(1.0, float(loss))
    torch.testing.assert_close(pred_e, pred.detach(), rtol=1e-5, atol=1e-6)
    assert _rel_l2(gt_e, net.encoder.embeddings.grad) < 1e-3      # same kernels; float atomics reorder
    for g, m in zip(gw_e, list(net.sigma_net) + list(net.color_net)):
        assert _rel_l2(g, m.weight.grad) < 1e-3
    # and the oracle's CPU step (same drop rule through M)
    e = net.encoder
    ws_o = [m.weight.detach().cpu() for m in list(net.sigma_net) + list(net.color_net)]
    fn = lambda x, d: field.hash_field_forward(x, d, e.embeddings.detach().cpu(), e.offsets.cpu().numpy(), float(e.per_level_scale),
                                               e.base_resolution, ws_o, quantize_fp16=True)[:2]
    o = field.render_train_step(ro, rd, scene["bitfield"], gt, fn, M=eng.M)
    assert abs(loss_e - float(o["loss"])) < 1e-2 * float(o["loss"])


def test_engine_graph_replay_is_equivalent(scene):
    net, eng, ro, rd, gt = _setup(scene, n_rays=512, seed=1)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    ref_loss, ref_img = float(eng.loss[0]), eng.image.clone()
    ref_gt = eng.grad_table.clone()
    eng.capture()
    eng.loss_slots.zero_(); eng.image.zero_(); eng.grad_table.zero_()
    eng.replay()
    torch.cuda.synchronize()
    assert abs(float(eng.loss[0]) - ref_loss) < 1e-6 * max(1.0, ref_loss)
    torch.testing.assert_close(eng.image, ref_img, rtol=1e-6, atol=1e-7)
    assert _rel_l2(eng.grad_table, ref_gt) < 1e-3
    # new rays through the same graph
    ro2, rd2 = scene["batches"][1]
    eng.rays_o.copy_(ro2[:512]); eng.rays_d.copy_(rd2[:512])
    eng.replay()
    torch.cuda.synchronize()
    assert float(eng.loss[0]) > 0 and int(eng.status.item()) == 0


def test_pipelined_graphs_match_the_serial_step(scene):
    """Two-graph pipelined steady state (march of batch i+1 beside the backward of batch i) gives, batch by batch, what the
    serial step gives: same rays / offsets, same loss, same gradients up to atomic ordering."""
    net, eng, ro0, rd0, gt0 = _setup(scene, n_rays=1024)
    eng.step(warmup=True)
    eng.finish_warmup()
    batches = []
    for b in range(4):
        ro, rd = scene["batches"][b % 3]
        sl = slice(1024 * (b // 3), 1024 * (b // 3) + 1024)
        gt = torch.rand(1024, 3, generator=torch.Generator().manual_seed(10 + b))
        batches.append((ro[sl].contiguous().cuda(), rd[sl].contiguous().cuda(), gt.cuda()))
    # serial reference
    ref = []
    eng.cur = 0
    for ro, rd, gt in batches:
        eng.rays_o.copy_(ro); eng.rays_d.copy_(rd); eng.gt.copy_(gt)
        eng.step()
        torch.cuda.synchronize()
        ref.append((eng.loss.clone(), eng.grad_table.clone(), eng.gw_ws.clone().view(16, -1).sum(0), eng.rays.clone(), eng.image.clone()))
    # pipelined
    eng.capture_pipelined()
    s0 = eng.sets[0]
    s0.rays_o.copy_(batches[0][0]); s0.rays_d.copy_(batches[0][1]); s0.gt.copy_(batches[0][2])
    eng.march(0)
    for i in range(4):
        nxt = eng.sets[(i + 1) & 1]
        ro, rd, gt = batches[(i + 1) % 4]
        nxt.rays_o.copy_(ro); nxt.rays_d.copy_(rd); nxt.gt.copy_(gt)
        eng.replay_pipelined(i)
        torch.cuda.synchronize()
        loss_r, gt_r, gw_r, rays_r, img_r = ref[i]
        assert torch.equal(eng.rays, rays_r), f"step {i}: ray offsets"
        torch.testing.assert_close(eng.image, img_r, rtol=0, atol=0)
        torch.testing.assert_close(eng.loss, loss_r, rtol=1e-5, atol=1e-7)
        assert _rel_l2(eng.grad_table, gt_r) < 1e-4
        assert _rel_l2(eng.gw_ws.view(16, -1).sum(0), gw_r) < 1e-4
    assert int(eng.status.item()) == 0


def test_host_fed_pipelined_graphs(scene):
    """capture_pipelined(host_io=True): batches written into the pinned staging twins reach the device through the H2D node of
    the step's graph, the loss words come back through its D2H node; results equal the device-fed serial step."""
    net, eng, ro0, rd0, gt0 = _setup(scene, n_rays=1024)
    eng.step(warmup=True)
    eng.finish_warmup()
    batches = []
    for b in range(4):
        ro, rd = scene["batches"][b % 3]
        sl = slice(1024 * (b // 3), 1024 * (b // 3) + 1024)
        gt = torch.rand(1024, 3, generator=torch.Generator().manual_seed(20 + b))
        batches.append(torch.stack([ro[sl], rd[sl], gt]).contiguous())   # [3, N, 3] on the host
    ref = []
    eng.cur = 0
    for pk in batches:
        eng.sets[0].inputs.copy_(pk)
        eng.step()
        torch.cuda.synchronize()
        ref.append((eng.loss.clone(), eng.grad_table.clone(), eng.rays.clone()))
    eng.capture_pipelined(host_io=True)
    eng.sets[0].inputs.copy_(batches[0])
    eng.march(0)
    for i in range(4):
        eng.sets[(i + 1) & 1].host.copy_(batches[(i + 1) % 4])
        eng.replay_pipelined(i, host_io=True)
        torch.cuda.synchronize()
        loss_r, gt_r, rays_r = ref[i]
        assert torch.equal(eng.rays, rays_r), f"step {i}: ray offsets"
        host_loss = eng.host_loss.view(-1, 2).sum(0)
        torch.testing.assert_close(host_loss, loss_r.cpu(), rtol=1e-5, atol=1e-7)
        assert _rel_l2(eng.grad_table, gt_r) < 1e-4
    assert int(eng.status.item()) == 0


def test_set_bitfield_reaches_a_captured_graph(scene):
    """A density-grid update between replays (renderer.py:647-773 rewrites density_bitfield every 16 steps): set_bitfield() copies
    into the buffer the captured graph marches against and rebuilds the coarse rejection mask, so the next REPLAY of the existing
    graph gives what a fresh eager step on the new grid gives."""
    import numpy as np
    from pvd_b200 import synthetic as syn
    net, eng, ro, rd, gt = _setup(scene, n_rays=1024, seed=2)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.capture()
    eng.replay()
    torch.cuda.synchronize()
    rays_old = eng.rays.clone()
    # a different scene: the same boxes shifted / one removed -> other cells occupied, other rays hit
    boxes = syn.LEGO_BOXES.copy()
    boxes[:, [0, 3]] += 0.12
    boxes = boxes[:-1]
    grid2 = syn.lego_density_grid(boxes=boxes)
    bf2 = syn.pack_bitfield(grid2)
    assert bf2.shape == scene["bitfield"].shape and not np.array_equal(bf2, scene["bitfield"])
    eng.set_bitfield(torch.from_numpy(bf2))
    eng.replay()
    torch.cuda.synchronize()
    rays_graph, img_graph, loss_graph = eng.rays.clone(), eng.image.clone(), float(eng.loss[0])
    assert not torch.equal(rays_graph, rays_old), "the replay still marched against the old grid"
    # the same step on a FRESH engine built on the new grid (same M so the same rays are dropped, if any)
    from pvd_b200.engine import HashTrainEngine
    eng2 = HashTrainEngine(net, torch.from_numpy(bf2), 1024, loss_scale=512.0)
    eng2.stage()
    eng2.rays_o.copy_(ro); eng2.rays_d.copy_(rd); eng2.gt.copy_(gt)
    eng2.set_mean_count(eng.mean_count)
    eng2.step()
    torch.cuda.synchronize()
    assert torch.equal(eng2.rays, rays_graph), "sample counts / offsets after set_bitfield differ from a fresh engine's"
    torch.testing.assert_close(img_graph, eng2.image, rtol=0, atol=0)
    assert abs(loss_graph - float(eng2.loss[0])) < 1e-6 * max(1.0, loss_graph)
    # and the oracle's march on the new grid agrees on the counts
    from oracle import cpu
    on, of = cpu.near_far_from_aabb(ro.numpy(), rd.numpy(), np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2)
    o = cpu.march_rays_train(ro.numpy(), rd.numpy(), 1.0, bf2, 1, 128, on, of, M=eng.M, perturb=True)
    assert np.array_equal(rays_graph.cpu().numpy(), o[3])


def test_coarse_mask_is_not_used_beyond_bound_one(scene):
    """A direct caller with C == 1 and bound > 1 (the mask maps cells with 1/bound, the marcher with min(1, bound)): the fused count
    entry must give the plain two-phase march's counts (samples were pruned)."""
    import ctypes as C
    from pvd_b200 import _native as nv
    ro, rd = scene["batches"][0]
    ro, rd = ro[:512].contiguous().cuda(), rd[:512].contiguous().cuda()
    bf = torch.from_numpy(scene["bitfield"]).cuda()
    bound = 2.0
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device="cuda")
    N = 512
    l = nv.lib()
    ws = torch.empty(int(l.pvd_march_rays_train_workspace_words(N, 1024)), dtype=torch.int32, device="cuda")
    nears, fars = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    st = nv.stream_of(ro)
    nv.check(l.pvd_march_rays_train_count_aabb(nv.ptr(ro), nv.ptr(rd), nv.ptr(bf), nv.ptr(aabb), C.c_float(0.2), C.c_float(bound), C.c_float(0.0),
                                               C.c_uint32(1024), C.c_uint32(N), C.c_uint32(1), C.c_uint32(128), nv.ptr(nears), nv.ptr(fars),
                                               nv.ptr(rays), nv.ptr(counter), C.c_uint32(1), C.c_uint32(0), nv.ptr(ws), st))
    import raymarching
    n2, f2 = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    rays2 = torch.empty(N, 3, dtype=torch.int32, device="cuda")
    counter2 = torch.zeros(2, dtype=torch.int32, device="cuda")
    # the reference semantics from the CPU oracle (no mask anywhere)
    from oracle import cpu
    import numpy as np
    o = cpu.march_rays_train(ro.cpu().numpy(), rd.cpu().numpy(), bound, scene["bitfield"], 1, 128, n2.cpu().numpy(), f2.cpu().numpy(),
                             M=N * 1024, perturb=True)
    torch.cuda.synchronize()
    assert np.array_equal(rays.cpu().numpy(), o[3]), "C == 1 with bound > 1: counts differ from the oracle (coarse mask misapplied)"
