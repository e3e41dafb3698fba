"""Staged copies of parameters (fp16 table shadow, packed weight tiles) must follow in-place writes that do NOT bump
`param._version` -- `param.data.copy_()` is what torch_ema's copy_to() / restore() (distill_mutual/utils.py:1210-1212, 1364,
1463-1469) and reset_parameters do: evaluation after an EMA swap ran on stale weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pts(n=2048, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, 3, generator=g) * 1.9 - 0.95).cuda()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    return x, d


def test_grid_encoder_autocast_follows_data_copy():
    from gridencoder import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(num_levels=14, desired_resolution=2048).cuda()
    enc.embeddings.data.uniform_(-0.5, 0.5)
    x, _ = _pts()
    with torch.autocast("cuda", dtype=torch.float16):
        a = enc(x, bound=1).clone()
    v0 = enc.embeddings._version
    new = torch.empty_like(enc.embeddings).uniform_(-0.5, 0.5)
    enc.embeddings.data.copy_(new)                       # EMA-style swap: the version counter does not move
    assert enc.embeddings._version == v0
    with torch.autocast("cuda", dtype=torch.float16):
        b = enc(x, bound=1).clone()
    fresh = GridEncoder(num_levels=14, desired_resolution=2048).cuda()
    fresh.embeddings.data.copy_(new)
    with torch.autocast("cuda", dtype=torch.float16):
        c = fresh(x, bound=1)
    assert not torch.equal(a, b) and torch.equal(b, c), "forward under autocast used a stale fp16 copy of the table"


@pytest.mark.parametrize("kind", ["hash", "vm", "mlp"])
def test_fused_fields_follow_data_copy(kind):
    from pvd_b200.fused import HashNeRFField
    from pvd_b200.fused_mlp import MLPNeRFField
    from pvd_b200.fused_vm import VMNeRFField
    make = {"hash": lambda: HashNeRFField(num_levels=14, desired_resolution=2048), "vm": lambda: VMNeRFField(resolution0=32, scale=0.4),
            "mlp": lambda: MLPNeRFField(is_teacher=False)}[kind]
    torch.manual_seed(1)
    net = make().cuda().eval()
    if kind == "hash":
        net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    x, d = _pts(seed=1)
    with torch.no_grad():
        s0, c0 = net(x, d)
        s0, c0 = s0.clone(), c0.clone()
        torch.manual_seed(2)
        other = make().cuda()
        if kind == "hash":
            other.encoder.embeddings.data.uniform_(-0.5, 0.5)
        for p, q in zip(net.parameters(), other.parameters()):
            p.data.copy_(q.data)                          # torch_ema.copy_to()
        s1, c1 = net(x, d)
        s2, c2 = other.eval()(x, d)
    assert not torch.equal(c0, c1)
    assert torch.equal(s1, s2) and torch.equal(c1, c2), f"{kind}: forward after param.data.copy_() ran on stale staged weights"


def test_frozen_teacher_staging_is_cached_and_invalidates():
    from pvd_b200.fused import HashNeRFField
    torch.manual_seed(3)
    net = HashNeRFField(num_levels=14, desired_resolution=2048, is_teacher=True).cuda().eval()
    for p in net.parameters():
        p.requires_grad_(False)
    e = net.encoder.embeddings
    t1 = net._staged.table_for(e, True)
    k1 = net._staged._table_key
    t2 = net._staged.table_for(e, True)
    assert k1 is not None and net._staged._table_key == k1 and t1.data_ptr() == t2.data_ptr()
    with torch.no_grad():
        e.copy_(torch.zeros_like(e))                     # load_state_dict-style write: bumps the version -> re-staged
    t3 = net._staged.table_for(e, True)
    assert net._staged._table_key != k1 and float(t3.abs().max()) == 0.0
