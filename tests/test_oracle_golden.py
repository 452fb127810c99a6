"""CPU: the oracle restatements reproduce the committed reference vectors bit for bit.

tests/golden/*.npz were written by oracle/make_golden.py from the unmodified reference
rectifier.py / complex_builder.py (which also asserted equality at generation time); this test
keeps the oracle pinned wherever the suite runs, without /root/reference."""
import numpy as np
import pytest
import torch

from oracle import complex_builder_oracle as cbo
from oracle import rectifier_oracle as ro
from tests.helpers import NAMES, OPS, bits_equal, golden_cases, load_golden, sha


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_vectors(case):
    fx = load_golden(case)
    n = int(fx["n_vertices"])
    tab = ro.make_tables(n)
    if "edges" in fx.files:
        assert np.array_equal(tab.edges.numpy(), fx["edges"].astype(np.int64).reshape(-1, 2))
        assert np.array_equal(tab.triangles.numpy(), fx["triangles"].astype(np.int64).reshape(-1, 3))
        assert np.array_equal(tab.tetra.numpy(), fx["tetra"].astype(np.int64).reshape(-1, 4))
    else:
        want = fx["tables_sha"]
        assert [sha(tab.edges), sha(tab.triangles), sha(tab.tetra)] == list(want)

    leaves = [torch.from_numpy(fx[f"in_{k}"]).clone().requires_grad_(True) for k in NAMES]
    outs = list(ro.enforce_constraints(*leaves, tab))
    ups = [torch.from_numpy(fx[f"up_{k}"]) for k in NAMES]
    grads = torch.autograd.grad(outs, leaves, ups, allow_unused=True)
    for k, o, g, l in zip(NAMES, outs, grads, leaves):
        assert bits_equal(o, torch.from_numpy(fx[f"out_{k}"])), f"rectified {k}"
        g = torch.zeros_like(l) if g is None else g
        assert bits_equal(g, torch.from_numpy(fx[f"grad_{k}"])), f"gradient {k}"
        assert np.array_equal(o.detach().nonzero().squeeze(-1).numpy(), fx[f"active_{k}"]), f"active set {k}"

    pl = [o.detach().clone().requires_grad_(True) for o in outs]
    act = {k: p.nonzero().squeeze(-1) for k, p in zip(NAMES, pl)}
    built = cbo.build_sparse_matrices(pl, tab, act)
    assert (built is None) == bool(fx["empty"])
    if built is None:
        return
    adj, inc = built
    ops = [adj[f"rank_{r}"] for r in range(4)] + [inc[f"rank_{r}"] for r in (1, 2, 3)]
    for name, o in zip(OPS, ops):
        assert list(o.shape) == list(fx[f"{name}_shape"])
        assert o._nnz() == int(fx[f"{name}_nnz"])
        assert sha(o.indices()) == str(fx[f"{name}_idx_sha"])
        assert sha(o.values()) == str(fx[f"{name}_val_sha"])
    if f"{OPS[0]}_up" in fx.files:
        loss = sum((o.values() * torch.from_numpy(fx[f"{name}_up"])).sum() for name, o in zip(OPS, ops))
        g = torch.autograd.grad(loss, pl, allow_unused=True)
        for k, gi, l in zip(NAMES, g, pl):
            gi = torch.zeros_like(l) if gi is None else gi
            assert bits_equal(gi, torch.from_numpy(fx[f"opgrad_{k}"])), f"operator gradient {k}"


def test_full_complex_nnz_closed_form():
    """nnz of the full complex follows n(n-1), 2(n-2)C(n,2), 3(n-3)C(n,3), 4(n-4)C(n,4)."""
    from math import comb
    fx = load_golden("full20")
    n = 20
    want = [n * (n - 1), 2 * (n - 2) * comb(n, 2), 3 * (n - 3) * comb(n, 3), 4 * (n - 4) * comb(n, 4),
            2 * comb(n, 2), 3 * comb(n, 3), 4 * comb(n, 4)]
    assert [int(fx[f"{o}_nnz"]) for o in OPS] == want
