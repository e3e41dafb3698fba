"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares,
and the Python operator packages expose the reference's names.  No compute calls (no GPU here)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(pvd_[a-zA-Z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    from pvd_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_abi_version_and_error_strings():
    from pvd_b200 import _native
    l = _native.lib()
    assert l.pvd_abi_version() == _native.ABI_VERSION
    assert b"invalid argument" in l.pvd_error_string(-1)
    assert b"unsupported" in l.pvd_error_string(-2)
    assert l.pvd_march_rays_train_workspace_words(4096, 1024) == 16384 + 2 * 4096 + 2 * 4096 * 1024


def test_argument_validation_needs_no_gpu():
    """NULL pointers are rejected before any launch (the reference would dereference them)."""
    from pvd_b200 import _native
    l = _native.lib()
    rc = l.pvd_near_far_from_aabb(None, None, None, ctypes.c_uint32(4), ctypes.c_float(0.2), None, None, None)
    assert rc == -1
    rc = l.pvd_sh_encode_forward(None, None, ctypes.c_uint32(0), ctypes.c_uint32(3), ctypes.c_uint32(4), 0, None, None)
    assert rc == 0  # empty batch is a no-op


def test_operator_packages_mirror_reference_names():
    import gridencoder
    import raymarching
    import shencoder
    from tools.activation import trunc_exp
    from tools.encoding import get_encoder
    for name in ("near_far_from_aabb", "polar_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
                 "composite_rays_train", "march_rays", "composite_rays", "compact_rays"):
        assert callable(getattr(raymarching, name))
    enc, dim = get_encoder("hashgrid", desired_resolution=2048, num_levels=14)
    assert dim == 28 and tuple(enc.embeddings.shape) == (5303704, 2) and enc.offsets.dtype.is_floating_point is False
    enc16, dim16 = get_encoder("hashgrid", desired_resolution=2048, num_levels=16)
    assert dim16 == 32 and tuple(enc16.embeddings.shape) == (6119864, 2)
    sh, d = get_encoder("sphere_harmonics")
    assert d == 16 and isinstance(sh, shencoder.SHEncoder)
    fr, d = get_encoder("frequency", multires=10)
    assert d == 63
    assert callable(trunc_exp) and callable(gridencoder.grid_encode)


def test_no_cpu_fallback():
    """The operators refuse CPU tensors that the reference would also reject; nothing routes to the oracle."""
    import pytest
    import torch
    from gridencoder import grid_encode
    emb = torch.zeros(64, 2)
    offs = torch.tensor([0, 64], dtype=torch.int32)
    with pytest.raises(RuntimeError):
        grid_encode(torch.rand(4, 3), emb, offs, 2.0, 16)
    for top in ("aaai2023-pvd_b200", "scripts"):   # the product and its tooling: only tests/, smoke() and bench.py's baseline legs may
        for path in glob.glob(os.path.join(ROOT, top, "**", "*.py"), recursive=True):
            src = open(path).read()
            assert "import oracle" not in src and "from oracle" not in src, f"{path} imports the oracle"
    # bench.py: the oracle is reachable only from the CPU baseline and the reference arm, never from run_ours / build_engine
    src = open(os.path.join(ROOT, "bench.py")).read()
    ours = src[src.index("def build_engine("):src.index("# ------------------------------------------------------------------------------------------------ reference arm")]
    ours = ours.replace('cpu_baseline(args.workload', "")   # the one call of the baseline leg at the end of run_ours
    assert "oracle" not in ours


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Empty work is a no-op, NULL pointers are rejected before any launch (pair-loss, upkeep and vm two-kernel entry points)."""
    from pvd_b200 import _native
    l = _native.lib()
    u32, f32, u64 = ctypes.c_uint32, ctypes.c_float, ctypes.c_uint64
    assert l.pvd_pair_sample_sq(None, None, None, None, u32(0), None, None) == 0
    assert l.pvd_pair_sample_sq(None, None, None, None, u32(128), None, None) == -1
    assert l.pvd_pair_composite(None, None, None, None, None, None, None, u32(128), u32(0), None, None, None, None, None, None, None, None) == 0
    assert l.pvd_pair_composite(None, None, None, None, None, None, None, u32(128), u32(4), None, None, None, None, None, None, None, None) == -1
    assert l.pvd_pair_combine(None, None, None, None, None, None, f32(1.0), u32(0), None, None, None, None, None, None) == 0
    assert l.pvd_pair_combine(None, None, None, None, None, None, f32(1.0), u32(8), None, None, None, None, None, None) == -1
    assert l.pvd_zero_sample_tail(None, None, u32(4), u32(0), None, None, None, None) == 0
    assert l.pvd_zero_sample_tail(None, None, u32(4), u32(128), None, None, None, None) == -1
    assert l.pvd_l1_mean_reg(None, u64(0), f32(1e-4), f32(1.0), None, None, None) == 0
    assert l.pvd_l1_mean_reg(None, u64(16), f32(0.0), f32(1.0), None, None, None) == 0      # zero weight: nothing to do
    assert l.pvd_l1_mean_reg(None, u64(16), f32(1e-4), f32(1.0), None, None, None) == -1
    assert l.pvd_density_grid_points(None, None, u32(0), u32(128), f32(1.0), None, None) == 0
    assert l.pvd_density_grid_points(None, None, u32(8), u32(128), f32(1.0), None, None) == -1
    assert l.pvd_density_grid_update(None, None, None, None, u32(0), u32(0), f32(1.0), f32(0.95), None, None) == 0
    assert l.pvd_density_grid_update(None, None, None, None, u32(8), u32(8), f32(1.0), f32(0.95), None, None) == -1
    assert l.pvd_packbits_mean(None, u32(0), None, u32(8), f32(0.01), None, None, None) == 0
    assert l.pvd_packbits_mean(None, u32(1), None, u32(8), f32(0.01), None, None, None) == -1
    l.pvd_vm_backward_workspace_bytes.restype = ctypes.c_uint64
    assert l.pvd_vm_backward_workspace_bytes(u32(129)) == 2 * 36864 + 2 * 128 * 4
    assert l.pvd_vm_field_backward_ws(None, None, None, None, None, None, None, u32(128), None, None, None, None, None) == -1
    assert l.pvd_vm_field_backward(None, None, None, None, None, None, None, u32(0), None, None, None, None) == 0
    assert l.pvd_mlp_field_forward(None, None, None, u32(0), None, None, None, None, None) == 0
    assert l.pvd_hash_field_backward(None, None, None, None, None, None, None, u32(0), None, None, None, None, None, None) == 0


def test_python_constants_match_the_header():
    """Sizes the host side allocates by (pvd_b200/fused*.py) are the header's #defines, not copies that can drift."""
    import re
    hdr = open(os.path.join(ROOT, "include", "pvd_b200_fused.h")).read()

    def define(name):
        m = re.search(rf"#define\s+{name}\s+(.+)", hdr)
        assert m, name
        expr = re.sub(r"(\d+)u\b", r"\1", m.group(1).split("/*")[0].strip())
        return int(eval(expr, {"__builtins__": {}}, {}))

    from pvd_b200 import fused, fused_mlp
    assert fused.WBLOB_BYTES == define("PVD_FIELD_WBLOB_BYTES")
    assert fused.GW_WS_FLOATS == define("PVD_FIELD_GW_FLOATS") * define("PVD_FIELD_GW_COPIES")
    assert fused.LOSS_SLOTS == define("PVD_LOSS_SLOTS")
    assert fused_mlp.MLP_WBLOB_BYTES == define("PVD_MLP_WBLOB_BYTES")
    assert fused_mlp.MLP_WBLOB_T_BYTES == define("PVD_MLP_WBLOB_T_BYTES")
    assert fused_mlp.MLP_SAVE_TILE_BYTES == define("PVD_MLP_SAVE_TILE_BYTES")
    assert fused_mlp.MLP_GRAD_TILE_BYTES == define("PVD_MLP_GRAD_TILE_BYTES")
    assert fused_mlp.MLP_GW_FLOATS == define("PVD_MLP_GW_FLOATS")
