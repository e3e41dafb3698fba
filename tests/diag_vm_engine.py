"""Diagnostic (test infrastructure; run by hand on a GPU box: `python tests/diag_vm_engine.py`): VMTrainEngine gradients vs (a) the autograd module path on the GPU, (b) the CPU oracle.  Prints rel-L2 per parameter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "aaai2023-pvd_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import field
from pvd_b200 import synthetic as syn
from pvd_b200.engine import VMTrainEngine
import test_gpu_pair_engine as T

rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))
_, bitfield, _ = syn.lego_bitfield()
ro, rd = syn.make_ray_batches(3, 4096, seed=0)[0]
ro, rd = ro[:640].contiguous(), rd[:640].contiguous()
gt = torch.rand(640, 3, generator=torch.Generator().manual_seed(3))
LS = 65536.0
for l1 in (0.0, 1e-2):
    net = T._vm_net(21)
    net.density_bitfield.copy_(torch.from_numpy(bitfield))
    eng = VMTrainEngine(net, torch.from_numpy(bitfield), 640, loss_scale=LS, l1_reg_weight=l1)
    T._run_engine(eng, ro.cuda(), rd.cuda(), gt.cuda())
    got = {k: v.clone() / LS for k, v in eng.grads().items()}
    # (a) autograd module path
    net.train()
    out = net.render(ro.cuda().unsqueeze(0), rd.cuda().unsqueeze(0), bg_color=1, perturb=True)
    loss = torch.mean((out["image"][0] - gt.cuda()) ** 2) + net.density_loss() * l1
    (loss * LS).backward()
    auto = {n: p.grad / LS for n, p in net.named_parameters() if p.grad is not None}
    # (b) oracle
    f_s, named, P = T._vm_oracle(net)
    o = field.render_train_step(ro, rd, bitfield, gt, lambda x, d: f_s(x, d)[:2], M=eng.M)
    (o["loss"] + l1 * field.vm_density_loss(P["sm"], P["sv"])).backward()
    print(f"--- l1 {l1}: loss engine {float(eng.loss[0]):.6f} autograd {float(loss):.6f} oracle {float(o['loss']):.6f}")
    for k in got:
        print(f"{k:22s} engine-vs-oracle {rel(got[k], named[k].grad):.4f}  autograd-vs-oracle {rel(auto[k], named[k].grad):.4f}  engine-vs-autograd {rel(got[k], auto[k]):.4f}   |ref| {float(named[k].grad.norm()):.3e}")
