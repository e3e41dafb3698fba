"""`trunc_exp`: exp() forward, gradient evaluated at the input clamped to [-12, 12] (reference: tools/activation.py:6-21)."""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd


class _trunc_exp(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-12, 12))


trunc_exp = _trunc_exp.apply
