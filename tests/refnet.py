"""Import the REFERENCE's own `distill_mutual/network.py` + `renderer.py` over THIS repo's drop-in packages.  TEST INFRASTRUCTURE.

`/root/reference` exists only in the build container (never on the GPU box), so everything here is CPU-side: the reference's
`NeRFNetwork` / `NeRFRenderer` classes are imported with `raymarching`, `gridencoder`, `shencoder`, `tools` resolving to
`aaai2023-pvd_b200/` (the drop-in boundary under test) and the third-party packages the reference's `utils.py` wants but this
image lacks replaced by empty stubs (none is touched by constructing a network or calling `forward`).

The drop-in operators themselves have no CPU path (by design), so for CPU evaluation `cpu_standins(net)` swaps the two encoder
MODULES for stand-ins that compute the same functions with the C oracle -- the reference's `forward` (its layer stacks, clamp,
trunc_exp, concat order, side-channel attributes) is what runs.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "aaai2023-pvd_b200")

# what distill_mutual/utils.py (pulled in by renderer.py:11) and renderer.py import at module scope
_THIRD_PARTY = ["lpips", "tensorboardX", "imageio", "cv2", "matplotlib", "matplotlib.pyplot", "trimesh", "mcubes", "torch_ema", "IPython",
                "pandas", "tqdm", "rich", "rich.console", "packaging", "packaging.version"]


class _Any:
    """Absorbs whatever module-scope code does with a stubbed package (`lpips.LPIPS(net="alex").eval().cuda()`, utils.py:312)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Any,), {})


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "distill_mutual"))


def stubbed() -> list:
    return [m for m in _THIRD_PARTY if isinstance(sys.modules.get(m), _Stub)]


def load():
    """(network module, renderer module) of the reference, imported over the repo's drop-in packages."""
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in _THIRD_PARTY:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001  (absent, or present but broken without its native deps)
            sys.modules[name] = _Stub(name)
    if REF not in sys.path:
        sys.path.append(REF)          # AFTER the repo: only `distill_mutual` / `just_train_tea` resolve into the reference
    import gridencoder, raymarching, shencoder, tools  # noqa: F401
    for m in (gridencoder, raymarching, shencoder, tools):
        assert os.path.abspath(m.__file__).startswith(PKG), f"{m.__name__} resolved to {m.__file__}, not to the drop-in package"
    from distill_mutual import network, renderer
    assert os.path.abspath(network.__file__).startswith(REF)
    return network, renderer


def make_args(**over):
    """The argparse namespace fields NeRFNetwork reads (defaults: main_distill_mutual.py:182-219)."""
    a = types.SimpleNamespace(sigma_clip_min=-2.0, sigma_clip_max=7.0, global_step=10 ** 9, stage_iters={"stage1": -1, "stage2": -1},
                              render_stu_first=True, plenoxel_degree=3, plenoxel_res="[128,128,128]", PE=10, skip=3, nerf_layer_num=8,
                              nerf_layer_wide=256, resolution0=300, enable_edit_plenoxel=False)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def construct(cls, *a, **k):
    """Construct a reference module on a box without a GPU: `init_plenoxel_volume` calls `.cuda()` on its ParameterList
    (network.py:191); that one call is made a no-op for the duration of the constructor."""
    orig = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, device=None: self
    try:
        return cls(*a, **k)
    finally:
        torch.nn.Module.cuda = orig


class _CpuGrid(torch.nn.Module):
    """GridEncoder.forward(x, bound) on the CPU through the C oracle (differentiable w.r.t. the table)."""

    def __init__(self, enc):
        super().__init__()
        self.enc = enc

    def forward(self, x, bound=1):
        from oracle.field import _GridEncode
        e = self.enc
        x01 = (x + bound) / (2 * bound)     # gridencoder/grid.py:211
        return _GridEncode.apply(x01.view(-1, 3), e.embeddings, e.offsets.cpu().numpy(), float(e.per_level_scale), int(e.base_resolution))


class _CpuSH(torch.nn.Module):
    def __init__(self, degree):
        super().__init__()
        self.degree = degree

    def forward(self, d, **kw):
        from oracle import cpu
        return torch.from_numpy(cpu.sh_encode_forward(np.ascontiguousarray(d.detach().numpy().reshape(-1, 3)), self.degree))


def cpu_standins(net):
    """Swap the CUDA-only encoder modules of a reference NeRFNetwork for CPU stand-ins (same parameters, same functions)."""
    if getattr(net, "encoder", None) is not None:
        net.encoder = _CpuGrid(net.encoder)
    net.encoder_dir = _CpuSH(net.encoder_dir.degree)
    return net
