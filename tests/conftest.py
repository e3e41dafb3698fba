import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "aaai2023-pvd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def scene():
    """Lego-shaped occupancy bitfield + a few deterministic 4096-ray batches (CPU tensors)."""
    from pvd_b200 import synthetic as syn
    grid, bitfield, sha = syn.lego_bitfield()
    return {"grid": grid, "bitfield": bitfield, "sha": sha, "batches": syn.make_ray_batches(3, 4096, seed=0)}


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own CUDA extensions (oracle/_ref = /root/reference compiled unmodified by oracle/build_ref.py).

    None ONLY when oracle/_ref was never built (a checkout without the reference).  When the directory holds the built modules
    and they do not import, that is a hard failure: the comparisons with the reference must not be skipped silently."""
    d = os.path.join(ROOT, "oracle", "_ref")
    built = os.path.isdir(d) and any(f.endswith(".so") for f in os.listdir(d))
    if not built:
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import torch  # noqa: F401  (libtorch must be loaded first)
        import _raymarching, _gridencoder, _shencoder
    except Exception as ex:  # noqa: BLE001
        pytest.fail(f"oracle/_ref exists but the reference extensions do not import ({ex!r}): the reference comparisons would be skipped")
    return {"raymarching": _raymarching, "gridencoder": _gridencoder, "shencoder": _shencoder}
