"""Size-independent properties of the engine steps at BASELINE's FULL size (4096 rays x 1024 max steps, L=14 / VM-48 at 300^3),
where the CPU oracle would take minutes: bookkeeping identities, the loss recomputed from the engine's own outputs, linearity in
the loss scale, repeatability, and a student that equals its teacher."""
import pytest
import torch

pytestmark = pytest.mark.gpu

N = 4096


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _hash(seed, teacher=False):
    from pvd_b200.fused import HashNeRFField, _Args
    torch.manual_seed(seed)
    net = HashNeRFField(num_levels=14, desired_resolution=2048, is_teacher=teacher, args=_Args()).cuda()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    return net


def _feed(eng, scene, b=0, gt=None):
    ro, rd = scene["batches"][b]
    for rs in eng.sets:
        rs.rays_o.copy_(ro); rs.rays_d.copy_(rd)
        if gt is not None:
            rs.gt.copy_(gt)


def test_hash_step_full_size_properties(scene):
    from pvd_b200.engine import HashTrainEngine
    net = _hash(0)
    gt = torch.rand(N, 3, generator=torch.Generator().manual_seed(1)).cuda()
    eng = HashTrainEngine(net, torch.from_numpy(scene["bitfield"]), N, loss_scale=1024.0)
    eng.stage()
    _feed(eng, scene, 0, gt)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0
    rays = eng.rays.cpu().long()
    total = int(eng.counter[0])
    # bookkeeping: ray ids in order, offsets = exclusive scan of the counts, counter = (sum, N), everything fits M
    assert torch.equal(rays[:, 0], torch.arange(N)) and int(rays[:, 2].sum()) == total and int(eng.counter[1]) == N
    assert torch.equal(rays[:, 1], torch.cumsum(rays[:, 2], 0) - rays[:, 2])
    assert 50_000 < total < eng.M and eng.M % 128 == 0 and int(rays[:, 2].max()) <= 1024
    # samples stay inside the scene cube, steps are positive
    x = eng.xyzs[:total]
    assert float(x.abs().max()) <= 1.0 and float(eng.deltas[:total, 0].min()) > 0
    # compositing: weights_sum in [0, 1], finite pixels; the loss is the MSE of the engine's own final image
    ws = eng.weights_sum
    assert float(ws.min()) >= 0 and float(ws.max()) <= 1 + 1e-5 and bool(torch.isfinite(eng.image).all())
    pred, depth = eng.final_image()
    loss = float(eng.loss[0])
    assert abs(loss - float(((pred - gt) ** 2).mean())) < 1e-5 * loss
    assert float(depth.min()) >= 0
    assert int(eng.loss[1]) == int((rays[:, 2] > 0).sum())          # rays that carried samples
    g1, w1, l1, r1 = eng.grad_table.clone(), eng.gw_ws.clone().view(16, -1).sum(0), eng.loss.clone(), eng.rays.clone()
    # repeatability: the same batch again -> identical rays, loss to the last bits of the slot sums, gradients to atomic order
    eng.step()
    torch.cuda.synchronize()
    assert torch.equal(eng.rays, r1) and abs(float(eng.loss[0]) - float(l1[0])) < 1e-5 * float(l1[0])
    assert _rel_l2(eng.grad_table, g1) < 1e-4
    # linearity in the loss scale (GradScaler): every gradient doubles, the loss does not move
    eng.loss_scale = 2048.0
    eng.step()
    torch.cuda.synchronize()
    assert abs(float(eng.loss[0]) - float(l1[0])) < 1e-5 * float(l1[0])
    assert _rel_l2(eng.grad_table, 2 * g1) < 2e-3 and _rel_l2(eng.gw_ws.view(16, -1).sum(0), 2 * w1) < 2e-3
    # only table entries of cells the samples touch receive gradient: far fewer than the table has, but many
    nz = int((eng.grad_table.abs().sum(1) > 0).sum())
    assert 100_000 < nz < eng.grad_table.shape[0]


def test_pair_step_full_size_properties(scene):
    from pvd_b200.engine import PairDistillEngine
    from pvd_b200.fused import _Args
    from pvd_b200.fused_vm import VMNeRFField
    rates = (1.0, 0.002, 0.002, 0.002)
    tea = _hash(2, teacher=True)
    torch.manual_seed(3)
    stu = VMNeRFField(resolution0=300, args=_Args()).cuda()
    eng = PairDistillEngine(tea, stu, torch.from_numpy(scene["bitfield"]), N, rates=rates, stage=3, l1_reg_weight=0.0, loss_scale=128.0)
    eng.stage()
    _feed(eng, scene, 1)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0
    t = eng.loss_terms()
    total = float(eng.loss[0])
    assert abs(total - sum(r * t[k] for r, k in zip(rates, ("rgb", "fea", "color", "sigma")))) < 1e-5 * total
    pred_s, pred_t = eng.final_images()
    assert abs(t["rgb"] - float(torch.norm(pred_t - pred_s))) < 1e-4 * t["rgb"]
    M, cnt = eng.M, int(eng.counter[0])
    assert abs(t["fea"] - float(torch.norm(eng.feat[:M] - eng.feat_tea[:M]))) < 1e-4 * t["fea"]
    assert abs(t["color"] - float(torch.norm(eng.rgbs[:M] - eng.rgbs_tea[:M]))) < 1e-4 * t["color"]
    # padding rows are zeros for both networks' inputs, and carry per-sample gradients only
    assert float(eng.xyzs[cnt:M].abs().max()) == 0.0 and float(eng.grad_sigmas[cnt:M].abs().max()) == 0.0
    assert float(eng.grad_feat[cnt:M].abs().max()) > 0
    g = eng.grads()
    assert all(bool(torch.isfinite(v).all()) for v in g.values()) and float(g["color_mat.0"].abs().sum()) > 0


def test_student_equal_to_teacher_has_zero_loss_and_gradient(scene):
    from pvd_b200.engine import PairDistillEngine
    tea, stu = _hash(5, teacher=True), _hash(5)
    eng = PairDistillEngine(tea, stu, torch.from_numpy(scene["bitfield"]), N, stage=3, loss_scale=128.0)
    eng.stage()
    _feed(eng, scene, 2)
    eng.step(warmup=True)
    eng.finish_warmup()
    eng.step()
    torch.cuda.synchronize()
    assert float(eng.loss[0]) == 0.0 and all(v == 0.0 for v in eng.loss_terms().values())
    assert float(eng.grad_table.abs().max()) == 0.0 and float(eng.gw_ws.abs().max()) == 0.0
