"""GPU parity: SCCN message passing through the C ABI vs the oracle restatement of custom_sccn.py.

PARITY UNPINNED for ``Conv`` (TopoModelX is absent): both sides use neighborhood @ (x @ W)."""
import pytest
import torch

from oracle import complex_builder_oracle as cbo
from oracle import glue_oracle as glo
from oracle import rectifier_oracle as ro
from oracle.sccn_oracle import OracleSCCN
from tests.helpers import NAMES, assert_close, hard_concrete_like

pytestmark = pytest.mark.gpu

# Deep chains (6 layers x LayerNorm) and batch-summed parameter gradients accumulate fp32 rounding of
# both implementations; activations of a single layer are held to the north-star tolerance.
DEEP = dict(rtol=1e-4, atol=1e-5)


def _pair(channels, max_rank, n_layers, seed=0):
    import topo_audio_autoencoder_b200 as T
    torch.manual_seed(seed)
    ref = OracleSCCN(channels, max_rank, n_layers)
    with torch.no_grad():      # move every parameter off its init so nothing is trivially 1 or 0
        for p in ref.parameters():
            p.add_(0.05 * torch.randn_like(p))
    ours = T.GradientSCCN(channels, max_rank, n_layers).cuda()
    missing = ours.load_state_dict(ref.state_dict(), strict=True)
    return ref, ours


def _compare_params(tag, ours, ref, **tol):
    ref_grads = dict(ref.named_parameters())
    for name, p in ours.named_parameters():
        rg = ref_grads[name].grad
        if rg is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, name
            continue
        assert p.grad is not None, f"{name}: no gradient"
        assert_close(f"{tag}/d{name}", p.grad, rg, **tol)


def test_reference_smoke_script_shapes():
    """test_sccn.py:4-65: random rank-0/1 features, 5-nnz uncoalesced random operators, 4 layers; the
    reference only prints -- here outputs and every gradient are compared with the oracle."""
    g = torch.Generator().manual_seed(7)
    nv, ne, C = 20, 40, 64
    x0, x1 = torch.randn(nv, C, generator=g), torch.randn(ne, C, generator=g)

    def rand_sparse(r, c):
        idx = torch.stack([torch.randint(0, r, (5,), generator=g), torch.randint(0, c, (5,), generator=g)])
        return idx, torch.rand(5, generator=g) + 0.5

    specs = [rand_sparse(nv, nv), rand_sparse(ne, ne), rand_sparse(nv, ne)]
    shapes = [(nv, nv), (ne, ne), (nv, ne)]
    ref, ours = _pair(C, 1, 4)
    ref.train(); ours.train()

    def run(model, dev):
        xs = [x0.to(dev).requires_grad_(True), x1.to(dev).requires_grad_(True)]
        mats = [torch.sparse_coo_tensor(i.to(dev), v.to(dev), s).requires_grad_() for (i, v), s in zip(specs, shapes)]
        out = model({"rank_0": xs[0], "rank_1": xs[1]}, {"rank_1": mats[2]}, {"rank_0": mats[0], "rank_1": mats[1]})
        loss = sum(o.sum() for o in out.values())
        loss.backward()
        return out, xs, mats

    out_r, xs_r, mats_r = run(ref, "cpu")
    out_g, xs_g, mats_g = run(ours, "cuda")
    for k in out_r:
        assert_close(f"sccn-generic/smoke/{k}", out_g[k], out_r[k], **DEEP)
    for i in range(2):
        assert_close(f"sccn-generic/smoke/dx{i}", xs_g[i].grad, xs_r[i].grad, **DEEP)
    for i in range(3):
        assert_close(f"sccn-generic/smoke/dA{i}", mats_g[i].grad.to_dense(), mats_r[i].grad.to_dense(), **DEEP)
    _compare_params("sccn-generic/smoke", ours, ref, **DEEP)


def _complex_inputs(n, batch, regime, seed):
    g = torch.Generator().manual_seed(seed)
    tab = ro.make_tables(n)
    N = sum(tab.sizes)
    if regime == "full":
        probs = torch.rand(batch, N, generator=g) * 0.98 + 0.01
    else:
        probs = hard_concrete_like((batch, N), g, p_zero=0.12, p_one=0.15)
    return tab, probs


def _oracle_stage(ref, tab, probs, emb_params, weights, training=True):
    """Per-sample oracle chain: rectify -> active sets -> embeddings -> operators -> SCCN."""
    ref.train(training)
    outs, total = [], 0.0
    for b in range(probs.shape[0]):
        parts = torch.split(probs[b], list(tab.sizes))
        rect = ro.enforce_constraints(*parts, tab)
        emb = glo.active_embeddings(emb_params, rect)
        adj, inc = cbo.build_sparse_matrices(rect, tab, emb["active_indices"])
        feats = {f"rank_{r}": emb[f"rank_{r}"] for r in range(4)}
        out = ref(feats, inc, adj)
        outs.append(out)
        for r in range(4):
            idx = emb["active_indices"][NAMES[r]]
            total = total + (out[f"rank_{r}"] * weights[r][b][idx]).sum()
    return outs, total


@pytest.mark.parametrize("n,batch,regime,layers,channels", [
    (8, 3, "hc", 1, 64), (8, 2, "full", 2, 64), (20, 2, "hc", 6, 64), (20, 1, "full", 6, 64), (9, 2, "hc", 2, 32)])
def test_matrix_free_stage_matches_oracle(n, batch, regime, layers, channels):
    import topo_audio_autoencoder_b200 as T
    tab, probs = _complex_inputs(n, batch, regime, seed=n * 31 + batch)
    ref, ours = _pair(channels, 3, layers, seed=n)
    head = T.ComplexHead(n, embedding_dim=channels).cuda()
    g = torch.Generator().manual_seed(17)
    weights = [torch.randn(batch, s, channels, generator=g) for s in tab.sizes]

    # oracle
    emb_leaves = []
    for name in head._embedding_names:
        emb, ln = getattr(head, name)
        emb_leaves.append(tuple(t.detach().cpu().clone().requires_grad_(True) for t in (emb.weight, ln.weight, ln.bias)))
    pc = probs.clone().requires_grad_(True)
    outs_c, loss_c = _oracle_stage(ref, tab, pc, emb_leaves, weights)
    loss_c.backward()

    # ours: rectify -> active sets -> embeddings -> matrix-free SCCN, whole batch at once
    pg = probs.cuda().requires_grad_(True)
    ours.train()
    rect = T.rectify_batch(pg, head.constraints)
    cx = head.batched_complex(rect, sync=True)
    xs = ours.forward_complex(cx, head.embed(cx))
    loss_g = 0.0
    hc = cx.host_counts
    o = head._tables.offsets
    for r in range(4):
        rows = torch.split(xs[r], hc[:, r].tolist())
        for b in range(batch):
            idx = cx.act_idx[b, o[r]:o[r] + int(hc[b, r])].long()
            assert_close(f"sccn-stage/n={n}/{regime}/L={layers}/C={channels}/b={b}/rank_{r}", rows[b],
                         outs_c[b][f"rank_{r}"], **(DEEP if layers > 1 else {}))
            loss_g = loss_g + (rows[b] * weights[r][b].cuda()[idx]).sum()
    loss_g.backward()
    tag = f"sccn-stage/n={n}/{regime}/L={layers}/C={channels}"
    assert_close(f"{tag}/loss", loss_g.detach().reshape(1), loss_c.detach().reshape(1), rtol=1e-4, atol=1e-4)
    assert_close(f"{tag}/dprobs", pg.grad, pc.grad, **DEEP)
    _compare_params(tag, ours, ref, rtol=1e-4, atol=1e-4)
    for name, leaves in zip(head._embedding_names, emb_leaves):
        emb, ln = getattr(head, name)
        assert_close(f"{tag}/d{name}.table", emb.weight.grad, leaves[0].grad, rtol=1e-4, atol=1e-4)
        assert_close(f"{tag}/d{name}.ln_w", ln.weight.grad, leaves[1].grad, rtol=1e-4, atol=1e-4)
        assert_close(f"{tag}/d{name}.ln_b", ln.bias.grad, leaves[2].grad, rtol=1e-4, atol=1e-4)


def test_generic_operators_path_equals_matrix_free_path():
    """The reference-signature forward (explicit sparse operators from build_sparse_matrices) and the
    matrix-free batch path are two evaluations of the same layer."""
    import topo_audio_autoencoder_b200 as T
    n, C = 10, 64
    tab, probs = _complex_inputs(n, 1, "hc", seed=5)
    _, ours = _pair(C, 3, 2, seed=1)
    head = T.ComplexHead(n, embedding_dim=C).cuda()
    for training in (True, False):
        ours.train(training)
        rect = T.rectify_batch(probs.cuda(), head.constraints)
        cx = head.batched_complex(rect, sync=True)
        xs = head.embed(cx)
        a = ours.forward_complex(cx, xs)
        parts = [p[0] for p in torch.split(rect, head._tables.counts, dim=1)]
        rp = T.RectifiedProbs(*parts, rect[0])
        act = head.get_active_simplex_embeddings(*parts)
        mats = T.build_sparse_matrices(rp, head.constraints, act["active_indices"])
        b = ours({f"rank_{r}": act[f"rank_{r}"] for r in range(4)}, mats.incidences, mats.adjacencies)
        for r in range(4):
            assert_close(f"sccn-paths/train={training}/rank_{r}", a[r], b[f"rank_{r}"], rtol=1e-4, atol=1e-5)


def test_missing_ranks_are_skipped():
    """custom_sccn.py:69-71, 123-125."""
    import topo_audio_autoencoder_b200 as T
    ref, ours = _pair(64, 2, 2)
    ref.eval(); ours.eval()
    x0 = torch.randn(5, 64)
    idx = torch.tensor([[0, 1, 2], [1, 2, 3]])
    a0 = torch.sparse_coo_tensor(idx, torch.ones(3), (5, 5)).coalesce()
    out_r = ref({"rank_0": x0, "rank_1": None}, {}, {"rank_0": a0})
    out_g = ours({"rank_0": x0.cuda(), "rank_1": None}, {}, {"rank_0": a0.cuda()})
    assert out_g["rank_1"] is None and out_g["rank_2"] is None
    assert_close("sccn-generic/missing/rank_0", out_g["rank_0"], out_r["rank_0"], rtol=1e-5, atol=2e-6)
    # no operator at all: features pass through unchanged
    out_g = ours({"rank_0": x0.cuda()}, {}, {})
    assert torch.equal(out_g["rank_0"].cpu(), x0)
