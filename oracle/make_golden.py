"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle against it.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (the only place /root/reference
exists):

    python oracle/make_golden.py

It imports /root/reference/rectifier.py and /root/reference/complex_builder.py as they are (both
are pure torch and import cleanly), runs them on seeded inputs, asserts that the oracle
restatements reproduce every output BIT FOR BIT (values, gradients, COO indices), and writes
the vectors that travel to the GPU box.  Nothing at test time reads /root/reference.
"""
from __future__ import annotations

import hashlib
import io
import os
import sys
import contextlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import rectifier as ref_rect            # noqa: E402  (reference, unmodified)
import complex_builder as ref_cb        # noqa: E402  (reference, unmodified)

from oracle import rectifier_oracle as ro          # noqa: E402
from oracle import complex_builder_oracle as cbo   # noqa: E402

torch.autograd.set_detect_anomaly(False)           # the reference switches it on at import
OUT = os.path.join(ROOT, "tests", "golden")
NAMES = cbo.RANK_NAMES


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hard_concrete_like(n, gen, p_zero=0.2, p_one=0.15):
    """Probabilities with exact zeros and exact ones, like a clamped gate produces."""
    x = torch.rand(n, generator=gen)
    r = torch.rand(n, generator=gen)
    x = torch.where(r < p_zero, torch.zeros_like(x), x)
    x = torch.where(r > 1 - p_one, torch.ones_like(x), x)
    return x


def cases():
    g = torch.Generator().manual_seed(511990)      # the reference's seed (encoder.py:91)
    sizes = lambda n: ro.make_tables(n).sizes      # noqa: E731

    # 1. the reference's own demo vector: rectifier.py:168-187
    torch.manual_seed(42)
    v = torch.rand(7); v[1] = 0; v[5] = 0
    e = torch.rand(21); e[10] = 0
    t = torch.rand(35); tt = torch.rand(35)
    yield "demo7", 7, (v, e, t, tt)

    # 2. zeros planted at every rank (SURVEY.md section 8a "structure facts")
    n = 9
    nv, ne, nt, nq = sizes(n)
    v = torch.rand(nv, generator=g); v[3] = 0
    e = torch.rand(ne, generator=g); e[[0, 7, 20]] = 0
    t = torch.rand(nt, generator=g); t[[1, 2, 40, 41, 80]] = 0
    tt = torch.rand(nq, generator=g); tt[[0, 5, 6, 100]] = 0
    yield "planted9", n, (v, e, t, tt)

    # 3. clamp-gate-like inputs, with ties 0 == 0 and 1 vs 1
    for n in (5, 6, 12):
        yield f"hc{n}", n, tuple(hard_concrete_like(s, g) for s in sizes(n))

    # 4. all-active (nothing is zero), vertices shifted like encoder.py:333 does
    n = 8
    ps = [torch.rand(s, generator=g) * 0.98 + 0.01 for s in sizes(n)]
    ps[0] = ps[0] + 2.0
    yield "full8_bias", n, tuple(ps)

    # 5. empty complex: every vertex zero
    n = 5
    ps = [hard_concrete_like(s, g) for s in sizes(n)]
    ps[0] = torch.zeros_like(ps[0])
    yield "empty5", n, tuple(ps)

    # 6. tiny complexes without tetrahedra / triangles
    for n in (2, 3, 4):
        yield f"tiny{n}", n, tuple(hard_concrete_like(s, g, 0.1, 0.1) + 0.0 for s in sizes(n))

    # 7. the default size, hashes only for the operators
    n = 20
    yield "hc20", n, tuple(hard_concrete_like(s, g) for s in sizes(n))
    yield "full20", n, tuple(torch.rand(s, generator=g) * 0.98 + 0.01 for s in sizes(n))


def run_reference(n, probs, gen):
    """Reference rectifier + builder, with fixed upstream gradients."""
    mats = ref_rect.ConstraintMatrices.create(n)
    leaves = [p.clone().requires_grad_(True) for p in probs]
    rect = ref_rect.enforce_constraints(*leaves, mats)
    outs = [rect.vertices, rect.edges, rect.triangles, rect.tetra]
    ups = [torch.randn(o.shape, generator=gen) for o in outs]
    grads = torch.autograd.grad(outs, leaves, ups, allow_unused=True)
    grads = [torch.zeros_like(l) if g_ is None else g_ for l, g_ in zip(leaves, grads)]

    # operators: the builder is fed detached rectified probabilities as fresh leaves so that its
    # gradient is pinned on its own
    pl = [o.detach().clone().requires_grad_(True) for o in outs]
    rp = ref_rect.RectifiedProbs(*pl, torch.cat(pl))
    act = {k: p.nonzero().squeeze(-1) for k, p in zip(NAMES, pl)}
    with contextlib.redirect_stdout(io.StringIO()):
        built = ref_cb.build_sparse_matrices(rp, mats, act)
    ops, op_ups, op_grads = None, None, None
    if built is not None:
        ops = [built.adjacencies[f"rank_{r}"] for r in range(4)] + [built.incidences[f"rank_{r}"] for r in (1, 2, 3)]
        op_ups = [torch.randn(o._nnz(), generator=gen) for o in ops]
        loss = sum((o.values() * w).sum() for o, w in zip(ops, op_ups))
        op_grads = torch.autograd.grad(loss, pl, allow_unused=True)
        op_grads = [torch.zeros_like(l) if g_ is None else g_ for l, g_ in zip(pl, op_grads)]
    return mats, outs, ups, grads, act, ops, op_ups, op_grads


def run_oracle(n, probs, ups, op_ups):
    tab = ro.make_tables(n)
    leaves = [p.clone().requires_grad_(True) for p in probs]
    outs = list(ro.enforce_constraints(*leaves, tab))
    grads = torch.autograd.grad(outs, leaves, ups, allow_unused=True)
    grads = [torch.zeros_like(l) if g_ is None else g_ for l, g_ in zip(leaves, grads)]
    pl = [o.detach().clone().requires_grad_(True) for o in outs]
    act = {k: p.nonzero().squeeze(-1) for k, p in zip(NAMES, pl)}
    built = cbo.build_sparse_matrices(pl, tab, act)
    ops, op_grads = None, None
    if built is not None:
        adj, inc = built
        ops = [adj[f"rank_{r}"] for r in range(4)] + [inc[f"rank_{r}"] for r in (1, 2, 3)]
        loss = sum((o.values() * w).sum() for o, w in zip(ops, op_ups))
        op_grads = torch.autograd.grad(loss, pl, allow_unused=True)
        op_grads = [torch.zeros_like(l) if g_ is None else g_ for l, g_ in zip(pl, op_grads)]
    return tab, outs, grads, ops, op_grads


def bits_equal(a, b):
    return a.shape == b.shape and torch.equal(a.contiguous().view(torch.int32) if a.dtype == torch.float32 else a,
                                               b.contiguous().view(torch.int32) if b.dtype == torch.float32 else b)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen = torch.Generator().manual_seed(20260118)
    report = []
    for name, n, probs in cases():
        mats, outs, ups, grads, act, ops, op_ups, op_grads = run_reference(n, probs, gen)
        tab, o_outs, o_grads, o_ops, o_op_grads = run_oracle(n, probs, ups, op_ups)

        # --- pin the oracle: bit-for-bit against the reference ---
        assert torch.equal(tab.edges.reshape(-1, 2), mats.indices.edges.reshape(-1, 2)), name
        if len(tab.triangles):
            assert torch.equal(tab.triangles, mats.indices.triangles.reshape(-1, 3)), name
        if len(tab.tetra):
            assert torch.equal(tab.tetra, mats.indices.tetra.reshape(-1, 4)), name
        assert bits_equal(tab.v2e, mats.vertex_to_edge) and bits_equal(tab.e2t, mats.edge_to_triangle) \
            and bits_equal(tab.t2tt, mats.triangle_to_tetra), name
        for a, b in zip(outs, o_outs):
            assert bits_equal(a.detach(), b.detach()), f"{name}: rectified values differ"
        for a, b in zip(grads, o_grads):
            assert bits_equal(a, b), f"{name}: rectifier gradients differ"
        assert (ops is None) == (o_ops is None), name
        if ops is not None:
            for a, b in zip(ops, o_ops):
                assert a.shape == b.shape and torch.equal(a.indices(), b.indices()), f"{name}: COO indices differ"
                assert bits_equal(a.values().detach(), b.values().detach()), f"{name}: COO values differ"
            for a, b in zip(op_grads, o_op_grads):
                assert bits_equal(a, b), f"{name}: operator gradients differ"

        # --- write the fixture ---
        fx = {"n_vertices": np.int64(n), "empty": np.bool_(ops is None)}
        for k, p, o, u, g_ in zip(NAMES, probs, outs, ups, grads):
            fx[f"in_{k}"] = p.numpy()
            fx[f"out_{k}"] = o.detach().numpy()
            fx[f"up_{k}"] = u.numpy()
            fx[f"grad_{k}"] = g_.numpy()
            fx[f"active_{k}"] = act[k].numpy()
        if n <= 9:
            fx["edges"] = mats.indices.edges.numpy().astype(np.int16)
            fx["triangles"] = mats.indices.triangles.numpy().astype(np.int16)
            fx["tetra"] = mats.indices.tetra.numpy().astype(np.int16)
        else:
            fx["tables_sha"] = np.array([sha(mats.indices.edges.numpy()), sha(mats.indices.triangles.numpy()),
                                         sha(mats.indices.tetra.numpy())])
        if ops is not None:
            op_names = [f"adj{r}" for r in range(4)] + [f"inc{r}" for r in (1, 2, 3)]
            for on, o, w in zip(op_names, ops, op_ups):
                fx[f"{on}_shape"] = np.array(o.shape, dtype=np.int64)
                fx[f"{on}_nnz"] = np.int64(o._nnz())
                fx[f"{on}_idx_sha"] = np.array(sha(o.indices().numpy()))
                fx[f"{on}_val_sha"] = np.array(sha(o.values().detach().numpy()))
                if n <= 12:
                    fx[f"{on}_idx"] = o.indices().numpy().astype(np.int32)
                    fx[f"{on}_val"] = o.values().detach().numpy()
                    fx[f"{on}_up"] = w.numpy()
                else:
                    # the upstream weights are regenerated from the stored seed material instead
                    fx[f"{on}_up_sum"] = np.float64(w.double().sum().item())
            if n <= 12:
                for k, g_ in zip(NAMES, op_grads):
                    fx[f"opgrad_{k}"] = g_.numpy()
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **fx)
        report.append((name, n, None if ops is None else [o._nnz() for o in ops], os.path.getsize(path)))

    for r in report:
        print(r)
    print("oracle pinned bit-for-bit against reference rectifier.py / complex_builder.py on", len(report), "cases")


if __name__ == "__main__":
    main()
