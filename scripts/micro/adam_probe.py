"""Which fp32 arithmetic do torch's CUDA kernels behind torch.optim.AdamW (single-tensor path) actually perform?

For every ATen op of `_single_tensor_adam` (lerp_, mul_, addcmul_, sqrt, div by a Python scalar, add_, addcdiv_) this runs the op on
random CUDA tensors and compares the result BIT FOR BIT with candidate formulas evaluated in float64 (a product of two fp32 values
is exact in fp64, so `float32(float64(a) * b + c)` reproduces a fused multiply-add up to double rounding, and plain fp32 torch ops
reproduce the unfused forms).  The candidates that match are what csrc/optim.cu must spell out with __fmaf_rn / __fmul_rn.

    python scripts/micro/adam_probe.py        (on the GPU box; prints one JSON line)
"""
import json

import torch

torch.manual_seed(0)
dev = "cuda"
n = 1 << 20
f64 = lambda t: t.double()
f32 = lambda t: t.float()


def fma(a, b, c):
    return f32(f64(a) * f64(b) + f64(c))


def frac_equal(a, b):
    return float((a == b).float().mean())


out = {}
m = torch.randn(n, device=dev) * 0.1
g = torch.randn(n, device=dev)
v = torch.rand(n, device=dev) * 0.5 + 1e-4
p = torch.randn(n, device=dev) * 0.3
beta1, beta2, eps, lr, wd = 0.9, 0.99, 1e-15, 1e-2, 0.01
w1, w2 = 1 - beta1, 1 - beta2
w1f = torch.tensor(w1, device=dev, dtype=torch.float32)
w2f = torch.tensor(w2, device=dev, dtype=torch.float32)
b2f = torch.tensor(beta2, device=dev, dtype=torch.float32)

# ---- lerp_
ref = m.clone().lerp_(g, w1)
out["lerp"] = {"fma(w, g - m, m)": frac_equal(ref, fma(w1f, g - m, m)),
               "m + fl(w * (g - m))": frac_equal(ref, m + w1f * (g - m)),
               "g - fl((g - m) * (1 - w))": frac_equal(ref, g - (g - m) * (1 - w1f)),
               "fma(-(g-m), (1-w), g)": frac_equal(ref, fma(-(g - m), (1 - w1f), g))}
# ---- mul_ by a Python scalar
ref = v.clone().mul_(beta2)
out["mul_scalar"] = {"fl(v * fl32(beta2))": frac_equal(ref, v * b2f), "fl32(f64(v) * beta2)": frac_equal(ref, f32(f64(v) * beta2))}
vb = v * b2f
# ---- addcmul_
ref = vb.clone().addcmul_(g, g, value=w2)
out["addcmul"] = {"fma(fl(w2 * g), g, vb)": frac_equal(ref, fma(w2f * g, g, vb)),
                  "fma(w2, fl(g * g), vb)": frac_equal(ref, fma(w2f, g * g, vb)),
                  "vb + fl(fl(w2 * g) * g)": frac_equal(ref, vb + (w2f * g) * g),
                  "vb + fl(w2 * fl(g * g))": frac_equal(ref, vb + w2f * (g * g)),
                  "fl32(f64: vb + w2 * g * g)": frac_equal(ref, f32(f64(vb) + f64(w2f) * f64(g) * f64(g)))}
# ---- sqrt, division by a Python scalar, add_(eps)
step = 7
bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
bc2s = bc2 ** 0.5
s = v.sqrt()
out["sqrt"] = {"correctly rounded (vs f64 sqrt)": frac_equal(s, f32(f64(v).sqrt()))}
ref = s / bc2s
inv = torch.tensor(1.0, device=dev) / torch.tensor(bc2s, device=dev, dtype=torch.float32)
out["div_scalar"] = {"fl(s * fl32(1 / fl32(b)))": frac_equal(ref, s * inv), "true division s / fl32(b)": frac_equal(ref, s / torch.tensor(bc2s, device=dev, dtype=torch.float32)),
                     "fl32(f64(s) / b)": frac_equal(ref, f32(f64(s) / bc2s)), "fl(s * fl32(1 / b))": frac_equal(ref, s * torch.tensor(1.0 / bc2s, device=dev, dtype=torch.float32))}
d = (s / bc2s).add_(eps)
out["add_eps"] = {"fl(x + fl32(eps))": frac_equal(d, (s / bc2s) + torch.tensor(eps, device=dev, dtype=torch.float32))}
for stp in (2, 3, 5, 11):                  # the two reciprocal candidates coincide at some steps and differ at others
    b = (1 - beta2 ** stp) ** 0.5
    ref = s / b
    out[f"div_scalar_step{stp}"] = {"fl(s * fl32(1 / fl32(b)))": frac_equal(ref, s * (torch.tensor(1.0, device=dev) / torch.tensor(b, device=dev, dtype=torch.float32))),
                                    "fl(s * fl32(1 / b))": frac_equal(ref, s * torch.tensor(1.0 / b, device=dev, dtype=torch.float32))}
# ---- addcdiv_
step_size = lr / bc1
ref = p.clone().addcdiv_(m, d, value=-step_size)
alpha = torch.tensor(-step_size, device=dev, dtype=torch.float32)
q = m / d
out["addcdiv"] = {"fma(alpha, fl(m / d), p)": frac_equal(ref, fma(alpha, q, p)),
                  "p + fl(alpha * fl(m / d))": frac_equal(ref, p + alpha * q),
                  "fl32(f64: p + alpha * m / d)": frac_equal(ref, f32(f64(p) + f64(alpha) * f64(m) / f64(d))),
                  "fma(fl(alpha * m), 1/d ...) i.e. p + fl(alpha*m)/d": frac_equal(ref, p + (alpha * m) / d),
                  "fl32(f64(p) + f64(fl(alpha * m)) / f64(d))": frac_equal(ref, f32(f64(p) + f64(alpha * m) / f64(d)))}
# ---- decay
ref = p.clone().mul_(1 - lr * wd)
out["decay"] = {"fl(p * fl32(1 - lr wd))": frac_equal(ref, p * torch.tensor(1 - lr * wd, device=dev, dtype=torch.float32))}
# ---- the whole step, one tensor, against torch.optim.AdamW
P = torch.nn.Parameter(p.clone())
opt = torch.optim.AdamW([P], lr=lr, betas=(beta1, beta2), eps=eps, weight_decay=wd, foreach=False, fused=False)
P.grad = g.clone()
opt.step()
pm = p * torch.tensor(1 - lr * wd, device=dev, dtype=torch.float32)
m1 = fma(w1f, g - 0.0, torch.zeros_like(g))
v1 = fma(w2f * g, g, torch.zeros_like(g))
bc1, bc2 = 1 - beta1, 1 - beta2
inv1 = torch.tensor(1.0, device=dev) / torch.tensor(bc2 ** 0.5, device=dev, dtype=torch.float32)
d1 = v1.sqrt() * inv1 + torch.tensor(eps, device=dev, dtype=torch.float32)
p1 = fma(torch.tensor(-(lr / bc1), device=dev, dtype=torch.float32), m1 / d1, pm)
st = opt.state[P]
out["first_step_vs_formula"] = {"param": frac_equal(P.detach(), p1), "exp_avg": frac_equal(st["exp_avg"], m1), "exp_avg_sq": frac_equal(st["exp_avg_sq"], v1)}
print(json.dumps(out))

# ---- three whole steps: torch.optim.AdamW vs (a) the formulas above with HOST-computed scalars, (b) the fused kernel; and the fused
# kernel's DEVICE-computed scalars against the host's
import ctypes as C
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "aaai2023-pvd_b200"))
from pvd_b200.optim import FusedAdamW, PvdAdamSlot  # noqa: E402

hexf = lambda x: struct.pack(">f", float(x)).hex()
n2 = 1 << 18
p0 = torch.randn(n2, device=dev) * 0.3
P = torch.nn.Parameter(p0.clone())
topt = torch.optim.AdamW([P], lr=lr, betas=(beta1, beta2), eps=eps, weight_decay=wd, foreach=False, fused=False)
pe, me, ve = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)          # (a) emulation
pk, gk = p0.clone(), torch.zeros_like(p0)                                      # (b) kernel
fopt = FusedAdamW([dict(param=pk, grad=gk, lr=lr, zero_grad=True)], betas=(beta1, beta2), eps=eps, weight_decay=wd, loss_scale=1.0, check_finite=False)
T = lambda x: torch.tensor(x, device=dev, dtype=torch.float32)
rows = []
for step in range(1, 4):
    gg = torch.randn(n2, device=dev)
    P.grad = gg.clone()
    topt.step()
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    step_size, bc2s = lr / bc1, bc2 ** 0.5
    pe = pe * T(1 - lr * wd)
    me = fma(T(w1), gg - me, me)
    ve = fma(T(w2), gg * gg, ve * T(beta2))
    inv = T(1.0 / bc2s)                      # reciprocal in double, then fp32 (what ATen's div-by-scalar does)
    de = ve.sqrt() * inv + T(eps)
    pe = fma(T(-step_size), me / de, pe)
    gk.copy_(gg)
    fopt.step()
    torch.cuda.synchronize()
    st = fopt.read_state()
    slot = PvdAdamSlot.from_buffer_copy(bytes(fopt.slots.cpu().numpy().tobytes())[:C.sizeof(PvdAdamSlot)])
    tst = topt.state[P]
    rows.append({"step": step,
                 "emulation_vs_torch": {"param": frac_equal(pe, P.detach()), "m": frac_equal(me, tst["exp_avg"]), "v": frac_equal(ve, tst["exp_avg_sq"])},
                 "kernel_vs_torch": {"param": frac_equal(pk, P.detach()), "m": frac_equal(fopt.exp_avg[0], tst["exp_avg"]), "v": frac_equal(fopt.exp_avg_sq[0], tst["exp_avg_sq"])},
                 "kernel_vs_emulation_param": frac_equal(pk, pe),
                 "scalars_device_vs_host(hex)": {"neg_step_size": (hexf(slot.neg_step_size), hexf(-step_size)), "decay": (hexf(slot.decay), hexf(1 - lr * wd)),
                                                 "inv_bc2_sqrt": (hexf(st.inv_bc2_sqrt), hexf(float(inv))), "w1": (hexf(st.w1), hexf(w1)), "w2": (hexf(st.w2), hexf(w2)),
                                                 "beta2": (hexf(st.beta2_f), hexf(beta2)), "eps": (hexf(st.eps), hexf(eps)), "grad_scale": st.grad_scale,
                                                 "device_step": st.step}})
print(json.dumps({"three_steps": rows}))
