"""GPU parity: SCCN message passing through the C ABI vs the oracle restatement of custom_sccn.py.

PARITY UNPINNED for ``Conv`` (TopoModelX is absent): both sides use neighborhood @ (x @ W).

Tolerance: single ops are held to rtol 1e-5 / atol 1e-6.  Multi-layer outputs and gradients are held to
the same element-wise tolerance OR, where two fp32 evaluations cannot agree that closely, to the
fp64-anchored criterion of helpers.assert_fp32_equivalent (as close to the fp64 answer as the oracle's own
fp32 run)."""
import copy

import pytest
import torch

from oracle import complex_builder_oracle as cbo
from oracle import glue_oracle as glo
from oracle import rectifier_oracle as ro
from oracle.sccn_oracle import OracleSCCN
from tests.helpers import NAMES, assert_close, assert_fp32_equivalent, hard_concrete_like

pytestmark = pytest.mark.gpu


def _pair(channels, max_rank, n_layers, seed=0):
    import topo_audio_autoencoder_b200 as T
    torch.manual_seed(seed)
    ref = OracleSCCN(channels, max_rank, n_layers)
    with torch.no_grad():      # move every parameter off its init so nothing is trivially 1 or 0
        for p in ref.parameters():
            p.add_(0.05 * torch.randn_like(p))
    ours = T.GradientSCCN(channels, max_rank, n_layers).cuda()
    ours.load_state_dict(ref.state_dict(), strict=True)
    return ref, copy.deepcopy(ref).double(), ours


def _compare_params(tag, ours, ref, ref64):
    g32, g64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    # noise floor: 1e-6 of the largest parameter-gradient magnitude of the model (sums over all rows)
    floor = 5e-6 * max(p.grad.abs().max().item() for p in ref64.parameters() if p.grad is not None)
    for name, p in ours.named_parameters():
        if g32[name].grad is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, name
            continue
        assert p.grad is not None, f"{name}: no gradient"
        assert_fp32_equivalent(f"{tag}/d{name}", p.grad, g32[name].grad, g64[name].grad, floor=floor)


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
    ref, ref64, ours = _pair(C, 1, 4)

    def run(model, dev, dtype):
        model.train()
        xs = [x0.clone().to(dev, dtype).requires_grad_(True), x1.clone().to(dev, dtype).requires_grad_(True)]
        mats = [torch.sparse_coo_tensor(i.to(dev), v.to(dev, dtype), s).requires_grad_() for (i, v), s in zip(specs, shapes)]
        out = model({"rank_0": xs[0], "rank_1": xs[1]}, {"rank_1": mats[2]}, {"rank_0": mats[0], "rank_1": mats[1]})
        sum(o.sum() for o in out.values()).backward()
        return out, xs, mats

    out_r, xs_r, mats_r = run(ref, "cpu", torch.float32)
    out_d, xs_d, mats_d = run(ref64, "cpu", torch.float64)
    out_g, xs_g, mats_g = run(ours, "cuda", torch.float32)
    for k in out_r:
        assert_fp32_equivalent(f"sccn-generic/smoke/{k}", out_g[k], out_r[k], out_d[k])
    for i in range(2):
        assert_fp32_equivalent(f"sccn-generic/smoke/dx{i}", xs_g[i].grad, xs_r[i].grad, xs_d[i].grad)
    for i in range(3):
        # torch.mm(sparse, dense) hands the sparse leaf a DENSE gradient (the full outer product); only the
        # entries on the operator's pattern are gradients of anything, and those are what the SDDMM emits
        on = torch.sparse_coo_tensor(specs[i][0], torch.ones(5), shapes[i]).to_dense() != 0
        ours_dense = mats_g[i].grad.to_dense().cpu()
        assert (ours_dense[~on] == 0).all()
        assert_fp32_equivalent(f"sccn-generic/smoke/dA{i}", ours_dense[on], mats_r[i].grad.to_dense()[on],
                               mats_d[i].grad.to_dense()[on])
    _compare_params("sccn-generic/smoke", ours, ref, ref64)


def _complex_inputs(n, batch, regime, seed):
    g = torch.Generator().manual_seed(seed)
    tab = ro.make_tables(n)
    N = sum(tab.sizes)
    if regime == "full":
        probs = torch.rand(batch, N, generator=g) * 0.98 + 0.01
    elif regime == "mixed":
        # fully active and partly active ranks side by side (the aggregation skips the index lookups per sample and
        # rank when the whole rank is active): sample 0 everything, sample 1 loses triangles (and with them their
        # tetrahedra), sample 2 only tetrahedra
        probs = torch.rand(batch, N, generator=g) * 0.98 + 0.01
        off = [0, tab.sizes[0], tab.sizes[0] + tab.sizes[1], tab.sizes[0] + tab.sizes[1] + tab.sizes[2]]
        if batch > 1:
            kill = torch.rand(tab.sizes[2], generator=g) < 0.3
            probs[1, off[2]:off[3]][kill] = 0.0
        if batch > 2:
            kill = torch.rand(tab.sizes[3], generator=g) < 0.4
            probs[2, off[3]:][kill] = 0.0
    else:
        probs = hard_concrete_like((batch, N), g, p_zero=0.12, p_one=0.15)
    return tab, probs


def _oracle_stage(ref, tab, probs, emb_params, weights):
    """Per-sample oracle chain: rectify -> active sets -> embeddings -> operators -> SCCN (train mode).
    dtype follows `probs`."""
    ref.train()
    dt = probs.dtype
    tab = copy.copy(tab)
    tab.v2e, tab.e2t, tab.t2tt = tab.v2e.to(dt), tab.e2t.to(dt), tab.t2tt.to(dt)
    outs, total = [], 0.0
    for b in range(probs.shape[0]):
        parts = torch.split(probs[b], list(tab.sizes))
        rect = ro.enforce_constraints(*parts, tab)
        emb = glo.active_embeddings(emb_params, rect)
        adj, inc = cbo.build_sparse_matrices(rect, tab, emb["active_indices"])
        out = ref({f"rank_{r}": emb[f"rank_{r}"] for r in range(4)}, inc, adj)
        outs.append(out)
        for r in range(4):
            idx = emb["active_indices"][NAMES[r]]
            total = total + (out[f"rank_{r}"] * weights[r][b][idx].to(dt)).sum()
    return outs, total


@pytest.mark.parametrize("n,batch,regime,layers,channels", [
    (8, 3, "hc", 1, 64), (8, 2, "full", 2, 64), (8, 3, "mixed", 2, 64), (20, 2, "hc", 6, 64), (20, 1, "full", 6, 64), (9, 2, "hc", 2, 32)])
def test_matrix_free_stage_matches_oracle(n, batch, regime, layers, channels):
    import topo_audio_autoencoder_b200 as T
    tab, probs = _complex_inputs(n, batch, regime, seed=n * 31 + batch)
    ref, ref64, ours = _pair(channels, 3, layers, seed=n)
    head = T.ComplexHead(n, embedding_dim=channels).cuda()
    g = torch.Generator().manual_seed(17)
    weights = [torch.randn(batch, s, channels, generator=g) for s in tab.sizes]

    def emb_leaves(dtype):
        out = []
        for name in head._embedding_names:
            emb, ln = getattr(head, name)
            out.append(tuple(t.detach().cpu().to(dtype).clone().requires_grad_(True) for t in (emb.weight, ln.weight, ln.bias)))
        return out

    runs = {}
    for dtype, model in ((torch.float32, ref), (torch.float64, ref64)):
        leaves = emb_leaves(dtype)
        pc = probs.to(dtype).clone().requires_grad_(True)
        outs, loss = _oracle_stage(model, tab, pc, leaves, weights)
        loss.backward()
        runs[dtype] = (outs, loss, pc, leaves)
    outs_c, loss_c, pc, leaves_c = runs[torch.float32]
    outs_d, loss_d, pd, leaves_d = runs[torch.float64]

    # ours: rectify -> active sets -> embeddings -> matrix-free SCCN, whole batch at once
    pg = probs.cuda().requires_grad_(True)
    ours.train()
    rect = T.rectify_batch(pg, head.constraints)
    cx = head.batched_complex(rect, sync=True)
    xs = ours.forward_complex(cx, head.embed(cx))
    loss_g = 0.0
    hc, o = cx.host_counts, head._tables.offsets
    tag = f"sccn-stage/n={n}/{regime}/L={layers}/C={channels}"
    for r in range(4):
        rows = torch.split(xs[r], hc[:, r].tolist())
        for b in range(batch):
            idx = cx.act_idx[b, o[r]:o[r] + int(hc[b, r])].long()
            assert_fp32_equivalent(f"{tag}/b={b}/rank_{r}", rows[b], outs_c[b][f"rank_{r}"], outs_d[b][f"rank_{r}"])
            loss_g = loss_g + (rows[b] * weights[r][b].cuda()[idx]).sum()
    loss_g.backward()
    assert_fp32_equivalent(f"{tag}/loss", loss_g.detach().reshape(1), loss_c.detach().reshape(1), loss_d.detach().reshape(1))
    assert_fp32_equivalent(f"{tag}/dprobs", pg.grad, pc.grad, pd.grad)
    _compare_params(tag, ours, ref, ref64)
    for name, lc, ld in zip(head._embedding_names, leaves_c, leaves_d):
        emb, ln = getattr(head, name)
        assert_fp32_equivalent(f"{tag}/d{name}.table", emb.weight.grad, lc[0].grad, ld[0].grad)
        # LayerNorm parameter gradients are sums over every simplex of the rank and the batch (thousands of
        # terms of either sign): their fp32 summation noise is a few 1e-6 of the sum's scale in any order
        for which, i in (("ln_w", 1), ("ln_b", 2)):
            p = ln.weight if i == 1 else ln.bias
            assert_fp32_equivalent(f"{tag}/d{name}.{which}", p.grad, lc[i].grad, ld[i].grad,
                                   floor=5e-6 * ld[i].grad.abs().max().item())


def test_single_layer_forward_meets_the_strict_tolerance():
    """One SCCN layer, eval mode (no LayerNorm): rtol 1e-5 / atol 1e-6 element-wise."""
    import topo_audio_autoencoder_b200 as T
    n, C = 8, 64
    tab, probs = _complex_inputs(n, 2, "hc", seed=77)
    ref, _, ours = _pair(C, 3, 1, seed=3)
    head = T.ComplexHead(n, embedding_dim=C).cuda()
    emb = [tuple(t.detach().cpu() for t in (getattr(head, nm)[0].weight, getattr(head, nm)[1].weight, getattr(head, nm)[1].bias))
           for nm in head._embedding_names]
    ours.eval(); ref.eval()
    rect = T.rectify_batch(probs.cuda(), head.constraints)
    cx = head.batched_complex(rect, sync=True)
    xs = ours.forward_complex(cx, head.embed(cx))
    for b in range(2):
        parts = torch.split(probs[b], list(tab.sizes))
        r_ = ro.enforce_constraints(*parts, tab)
        e = glo.active_embeddings(emb, r_)
        adj, inc = cbo.build_sparse_matrices(r_, tab, e["active_indices"])
        out = ref({f"rank_{r}": e[f"rank_{r}"] for r in range(4)}, inc, adj)
        for r in range(4):
            rows = torch.split(xs[r], cx.host_counts[:, r].tolist())[b]
            assert_close(f"sccn-single-layer/b={b}/rank_{r}", rows, out[f"rank_{r}"])


def test_generic_operators_path_equals_matrix_free_path():
    """The reference-signature forward (explicit sparse operators from build_sparse_matrices) and the
    matrix-free batch path are two evaluations of the same layer."""
    import topo_audio_autoencoder_b200 as T
    n, C = 10, 64
    tab, probs = _complex_inputs(n, 1, "hc", seed=5)
    _, _, ours = _pair(C, 3, 2, seed=1)
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
            assert_close(f"sccn-paths/train={training}/rank_{r}", a[r], b[f"rank_{r}"], rtol=1e-5, atol=2e-5)


def test_missing_ranks_are_skipped():
    """custom_sccn.py:69-71, 123-125."""
    import topo_audio_autoencoder_b200 as T
    ref, _, ours = _pair(64, 2, 2)
    ref.eval(); ours.eval()
    x0 = torch.randn(5, 64)
    idx = torch.tensor([[0, 1, 2], [1, 2, 3]])
    a0 = torch.sparse_coo_tensor(idx, torch.ones(3), (5, 5)).coalesce()
    out_r = ref({"rank_0": x0, "rank_1": None}, {}, {"rank_0": a0})
    out_g = ours({"rank_0": x0.cuda(), "rank_1": None}, {}, {"rank_0": a0.cuda()})
    assert out_g["rank_1"] is None and out_g["rank_2"] is None
    assert_close("sccn-generic/missing/rank_0", out_g["rank_0"], out_r["rank_0"], rtol=1e-5, atol=5e-6)
    # no operator at all: features pass through unchanged
    out_g = ours({"rank_0": x0.cuda()}, {}, {})
    assert torch.equal(out_g["rank_0"].cpu(), x0)
