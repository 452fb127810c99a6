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


# ------------------------------------------------------------------------------------------------
# fixtures written from the reference's own encoder.py / precompute_distances.py / custom_sccn.py / decoder.py
# (oracle/make_golden_glue.py, third-party imports stubbed): the restatements stay pinned without /root/reference
# ------------------------------------------------------------------------------------------------
def test_binary_gumbel_oracle_matches_reference():
    from oracle import gate_oracle as go
    fx = load_golden("ref_gumbel")
    for i in range(int(fx["n_cases"])):
        lo = torch.from_numpy(fx[f"c{i}_logits"]).requires_grad_(True)
        out = go.binary_gumbel_train(lo, torch.from_numpy(fx[f"c{i}_gumbels"]), float(fx[f"c{i}_temp"]))
        (g,) = torch.autograd.grad(out, lo, torch.from_numpy(fx[f"c{i}_up"]))
        assert bits_equal(out, torch.from_numpy(fx[f"c{i}_out"])) and bits_equal(g, torch.from_numpy(fx[f"c{i}_grad"]))


@pytest.mark.parametrize("case", ["ref_glue_n6", "ref_glue_n9"])
def test_glue_oracle_matches_reference(case):
    from oracle import glue_oracle as glo
    from oracle.param_fill import value_for
    fx = load_golden(case)
    n, ch, seed = int(fx["n_vertices"]), int(fx["channels"]), int(fx["seed"])
    bias = torch.tensor([float(fx["vertex_bias"])])
    parts = glo.split_simplices(torch.from_numpy(fx["split_in"]), n, bias)
    for k, p in zip(NAMES, parts):
        assert bits_equal(p, torch.from_numpy(fx[f"split_{k}"]))
    probs = [torch.from_numpy(fx[f"prob_{k}"]).requires_grad_(True) for k in NAMES]
    triples = []
    for tbl, size in zip(("vertex_embeddings", "edge_embeddings", "triangle_embeddings", "tetra_embeddings"), glo.rank_sizes(n)):
        triples.append(tuple(value_for(seed, f"{tbl}.{leaf}", torch.empty(shape)).requires_grad_(True)
                             for leaf, shape in (("0.weight", (size, ch)), ("1.weight", (ch,)), ("1.bias", (ch,)))))
    emb = glo.active_embeddings(triples, probs)
    ups = [torch.from_numpy(fx[f"emb_up_{r}"]) for r in range(4)]
    flat = [t for tr in triples for t in tr]
    grads = torch.autograd.grad([emb[f"rank_{r}"] for r in range(4)], probs + flat, ups, allow_unused=True)
    names = [f"prob_{k}" for k in NAMES] + [f"{t}_{w}" for t in ("vtab", "etab", "ttab", "qtab") for w in ("weight", "ln_w", "ln_b")]
    for r, k in enumerate(NAMES):
        assert np.array_equal(emb["active_indices"][k].numpy(), fx[f"active_{k}"])
        assert bits_equal(emb[f"rank_{r}"], torch.from_numpy(fx[f"emb_{r}"]))
    for nm, g, leaf in zip(names, grads, probs + flat):
        g = torch.zeros_like(leaf) if g is None else g
        assert bits_equal(g, torch.from_numpy(fx[f"embgrad_{nm}"])), nm
    for v, want, wg in zip(fx["vp_in"], fx["vp_out"], fx["vp_grad"]):
        vl = torch.from_numpy(v).requires_grad_(True)
        p = glo.vertex_penalty(vl, int(fx["min_active"]), int(fx["max_active"]))
        (g,) = torch.autograd.grad(p, vl, allow_unused=True)
        g = torch.zeros_like(vl) if g is None else g
        assert bits_equal(p, torch.tensor(want)) and bits_equal(g, torch.from_numpy(wg))
    pl = [torch.from_numpy(fx[f"prob_{k}"]).requires_grad_(True) for k in NAMES]
    ent = glo.entropy_loss(*pl)
    assert bits_equal(ent, torch.tensor(fx["entropy"]))
    for k, g in zip(NAMES, torch.autograd.grad(ent, pl)):
        assert bits_equal(g, torch.from_numpy(fx[f"entgrad_{k}"]))


def test_distance_oracle_matches_reference():
    from oracle import distance_oracle as do
    fx = load_golden("ref_distance")
    a, b = torch.from_numpy(fx["bmd_a"]), torch.from_numpy(fx["bmd_b"])
    for norm in ("L1", "L2"):
        for rel in (0, 1):
            assert bits_equal(do.batch_mean_difference(a, b, norm=norm, relative=bool(rel)), torch.from_numpy(fx[f"bmd_{norm}_{rel}"]))
    got = do.batch_audio_distance(torch.from_numpy(fx["bad_x"]), torch.from_numpy(fx["bad_y"]))
    assert bits_equal(got, torch.from_numpy(fx["bad_out"]))
    m = do.pairwise_matrix(torch.from_numpy(fx["cd_audio"]), batch_size=4)
    assert bits_equal(m, torch.from_numpy(fx["cd_matrix"]))
    vals, idx = do.neighbour_order(m)
    assert np.array_equal(idx.numpy(), fx["cd_sorted_idx"]) and bits_equal(vals, torch.from_numpy(fx["cd_sorted_vals"]))


def test_sccn_oracle_matches_reference():
    from oracle.param_fill import fill_by_name
    from oracle.sccn_oracle import OracleSCCN
    from tests.helpers import load_sccn_case, ref_sccn_cases, run_sccn_case
    assert len(ref_sccn_cases()) >= 5
    for case in ref_sccn_cases():
        fx = load_golden(case)
        model = fill_by_name(OracleSCCN(int(fx["channels"]), int(fx["max_rank"]), int(fx["n_layers"])), int(fx["seed"]))
        model.train(bool(fx["train"]))
        feats, inc, adj, ups = load_sccn_case(fx)
        out, gf, gm, gp = run_sccn_case(model, feats, inc, adj, ups)
        for k, v in out.items():
            assert (v is None) == bool(fx[f"outnone_{k}"]), (case, k)
            if v is not None:
                assert bits_equal(v, torch.from_numpy(fx[f"out_{k}"])), (case, k)
        for k, g in gf.items():
            if g is not None:
                assert bits_equal(g, torch.from_numpy(fx[f"gx_{k}"])), (case, "gx", k)
        for (kind, k), g in gm.items():
            if g is not None:
                assert bits_equal(g, torch.from_numpy(fx[f"g{kind}_{k}"])), (case, kind, k)
        for k, g in gp.items():
            if g is not None:
                assert bits_equal(g, torch.from_numpy(fx[f"gp_{k}"])), (case, "gp", k)
            else:
                assert f"gp_{k}" not in fx.files, (case, k)


def test_decoder_tail_oracle_matches_reference():
    from oracle.decoder_oracle import OracleDecoderTail
    from oracle.param_fill import fill_by_name
    fx = load_golden("ref_decoder_tail")
    tail = fill_by_name(OracleDecoderTail(64, 250, 16), int(fx["seed"])).train()
    xs = {f"rank_{r}": torch.from_numpy(fx[f"x_rank_{r}"]).requires_grad_(True) for r in range(4)}
    y = tail(xs)
    assert bits_equal(y, torch.from_numpy(fx["out"]))
    grads = torch.autograd.grad(y, list(xs.values()), torch.from_numpy(fx["up"]))
    for r, g in enumerate(grads):
        assert torch.allclose(g, torch.from_numpy(fx[f"gx_rank_{r}"]), rtol=1e-6, atol=1e-7)
