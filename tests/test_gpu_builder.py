"""GPU parity: build_sparse_matrices through the C ABI -- COO indices and values bit-exact against the
reference vectors, gradients within tolerance."""
import numpy as np
import pytest
import torch

from oracle import complex_builder_oracle as cbo
from oracle import rectifier_oracle as ro
from tests.helpers import NAMES, OPS, assert_close, golden_cases, load_golden, sha

pytestmark = pytest.mark.gpu


def _ops(built):
    return [built.adjacencies[f"rank_{r}"] for r in range(4)] + [built.incidences[f"rank_{r}"] for r in (1, 2, 3)]


@pytest.mark.parametrize("case", golden_cases())
def test_operators_bit_exact_against_reference_vectors(case):
    import topo_audio_autoencoder_b200 as T
    fx = load_golden(case)
    n = int(fx["n_vertices"])
    mats = T.ConstraintMatrices.create(n)
    pl = [torch.from_numpy(fx[f"out_{k}"]).cuda().requires_grad_(True) for k in NAMES]
    probs = T.RectifiedProbs(*pl, torch.cat(pl))
    act = {k: torch.from_numpy(fx[f"active_{k}"]).cuda() for k in NAMES}
    built = T.build_sparse_matrices(probs, mats, act)
    assert (built is None) == bool(fx["empty"])       # empty complex -> None, never an exception
    if built is None:
        return
    ops = _ops(built)
    for name, o in zip(OPS, ops):
        assert o.is_coalesced() and o.indices().dtype == torch.int64 and o.values().dtype == torch.float32
        assert list(o.shape) == list(fx[f"{name}_shape"]), name
        assert o._nnz() == int(fx[f"{name}_nnz"]), name
        assert sha(o.indices()) == str(fx[f"{name}_idx_sha"]), f"{name}: COO indices are not bit-exact"
        assert sha(o.values()) == str(fx[f"{name}_val_sha"]), f"{name}: COO values are not bit-exact"
        assert o.values().requires_grad
    if f"{OPS[0]}_up" in fx.files:
        loss = sum((o.values() * torch.from_numpy(fx[f"{name}_up"]).cuda()).sum() for name, o in zip(OPS, ops))
        grads = torch.autograd.grad(loss, pl, allow_unused=True)
        for k, g, l in zip(NAMES, grads, pl):
            g = torch.zeros_like(l) if g is None else g
            assert_close(f"operators-grad/{case}/{k}", g, torch.from_numpy(fx[f"opgrad_{k}"]), rtol=1e-5, atol=2e-6)


def test_default_size_gradients_against_oracle():
    import topo_audio_autoencoder_b200 as T
    fx = load_golden("hc20")
    n = 20
    mats, tab = T.ConstraintMatrices.create(n), ro.make_tables(n)
    g = torch.Generator().manual_seed(9)
    pc = [torch.from_numpy(fx[f"out_{k}"]).clone().requires_grad_(True) for k in NAMES]
    act_c = {k: p.nonzero().squeeze(-1) for k, p in zip(NAMES, pc)}
    adj, inc = cbo.build_sparse_matrices(pc, tab, act_c)
    ops_c = [adj[f"rank_{r}"] for r in range(4)] + [inc[f"rank_{r}"] for r in (1, 2, 3)]
    ups = [torch.randn(o._nnz(), generator=g) for o in ops_c]
    gc = torch.autograd.grad(sum((o.values() * w).sum() for o, w in zip(ops_c, ups)), pc, allow_unused=True)

    pg = [p.detach().cuda().requires_grad_(True) for p in pc]
    built = T.build_sparse_matrices(T.RectifiedProbs(*pg, torch.cat(pg)), mats, {k: v.cuda() for k, v in act_c.items()})
    ops_g = _ops(built)
    for a, b in zip(ops_g, ops_c):
        assert torch.equal(a.indices().cpu(), b.indices()) and torch.equal(a.values().detach().cpu(), b.values().detach())
    gg = torch.autograd.grad(sum((o.values() * w.cuda()).sum() for o, w in zip(ops_g, ups)), pg, allow_unused=True)
    for k, a, b, l in zip(NAMES, gg, gc, pc):
        a = torch.zeros_like(l) if a is None else a
        b = torch.zeros_like(l) if b is None else b
        assert_close(f"operators-grad/hc20-live/{k}", a, b, rtol=1e-5, atol=1e-5)


def test_inconsistent_active_sets_follow_the_reference():
    """The builder takes the caller's index lists at face value (complex_builder.py:47, 57-59): an active
    list that drops simplices with non-zero probability must drop their rows/columns and everything that
    factors through them."""
    import topo_audio_autoencoder_b200 as T
    n = 7
    mats, tab = T.ConstraintMatrices.create(n), ro.make_tables(n)
    g = torch.Generator().manual_seed(11)
    probs = [torch.rand(s, generator=g) * 0.9 + 0.05 for s in tab.sizes]
    act = {k: torch.arange(len(p)) for k, p in zip(NAMES, probs)}
    act["triangles"] = act["triangles"][::2]
    act["vertices"] = act["vertices"][1:]
    adj, inc = cbo.build_sparse_matrices(probs, tab, act)
    pg = [p.cuda() for p in probs]
    built = T.build_sparse_matrices(T.RectifiedProbs(*pg, torch.cat(pg)), mats, {k: v.cuda() for k, v in act.items()})
    ops_c = [adj[f"rank_{r}"] for r in range(4)] + [inc[f"rank_{r}"] for r in (1, 2, 3)]
    for name, a, b in zip(OPS, _ops(built), ops_c):
        assert a.shape == b.shape, name
        assert torch.equal(a.indices().cpu(), b.indices()), name
        assert torch.equal(a.values().cpu(), b.values()), name
