"""GPU parity: the whole complex stage (gate -> rectifier -> active sets -> embeddings -> operators ->
SCCN -> penalties) against the per-sample oracle chain, and the reference-shaped generate_complex."""
import pytest
import torch

from oracle import complex_builder_oracle as cbo
from oracle import gate_oracle as go
from oracle import glue_oracle as glo
from oracle import rectifier_oracle as ro
from oracle.sccn_oracle import OracleSCCN
import copy

from tests.helpers import NAMES, assert_close, assert_fp32_equivalent

pytestmark = pytest.mark.gpu
DEEP = dict(rtol=1e-4, atol=1e-5)


def _emb_params(head):
    out = []
    for name in head._embedding_names:
        emb, ln = getattr(head, name)
        out.append(tuple(t.detach().cpu() for t in (emb.weight, ln.weight, ln.bias)))
    return out


def _dump_worst(tag, got, r32, r64, rect, off, k=8):
    from tests.helpers import REPORT
    g, a, d = got.detach().double().cpu(), r32.detach().double().cpu(), r64.detach().double().cpu()
    err = (g - d).abs()
    top = torch.topk(err.flatten(), k).indices
    n = g.shape[1]
    with open(REPORT, "a") as f:
        for t in top.tolist():
            b, i = divmod(t, n)
            rank = sum(i >= o for o in off[1:4])
            f.write(f"   worst {tag}: b={b} idx={i} rank={rank} ours={g[b, i]:.6e} o32={a[b, i]:.6e} o64={d[b, i]:.6e} "
                    f"rect={rect[b, i].item():.6e}\n")


def _oracle_chain(stage, ref, tab, off, logits, u, bias_on, dtype):
    """gate -> glue -> SCCN -> penalties per sample, in `dtype`, with sum-of-squares + penalty loss."""
    head, n = stage.head, stage.head.num_vertices
    tab = copy.copy(tab)
    tab.v2e, tab.e2t, tab.t2tt = tab.v2e.to(dtype), tab.e2t.to(dtype), tab.t2tt.to(dtype)
    emb = [tuple(t.to(dtype) for t in triple) for triple in _emb_params(head)]
    lc = logits.to(dtype).clone().requires_grad_(True)
    loc = torch.relu(torch.cat([p.detach().cpu() for p in (head.vertex_bias, head.edge_bias, head.triangle_bias,
                                                           head.tetra_bias)])).to(dtype)
    if bias_on == "probs":
        loc = torch.zeros(4, dtype=dtype)
    z = go.hard_concrete(lc, u.to(dtype), head.sampler.current_temp, head.sampler.gamma.item(), head.sampler.zeta.item(), loc, off)
    per_sample, loss = [], 0.0
    for b in range(logits.shape[0]):
        res = glo.complex_from_probs(z[b], n, head.vertex_bias.detach().cpu().to(dtype), tab, emb, bias_on == "probs")
        assert res is not None
        e, (adj, inc), rect = res
        o = ref({f"rank_{r}": e[f"rank_{r}"] for r in range(4)}, inc, adj)
        vp = glo.vertex_penalty(rect[0], head.min_active_vertices, head.max_active_vertices)
        ent = glo.entropy_loss(*rect)
        per_sample.append((e, o, rect, vp, ent))
        loss = loss + sum(o[f"rank_{r}"].pow(2).sum() for r in range(4)) + 0.3 * vp + 0.7 * ent
    loss.backward()
    return per_sample, loss, lc


@pytest.mark.parametrize("bias_on", ["logits", "probs"])
def test_complex_stage_end_to_end(bias_on):
    import topo_audio_autoencoder_b200 as T
    n, B, C, L = 9, 3, 64, 2
    torch.manual_seed(2)
    stage = T.ComplexStage(n, channels=C, n_layers=L, bias_on=bias_on).cuda().train()
    tab = ro.make_tables(n)
    off = glo.rank_offsets(n)
    g = torch.Generator().manual_seed(511990)
    logits = torch.randn(B, off[4], generator=g)
    u = torch.rand(B, off[4], generator=g).clamp_(1e-6, 1 - 1e-6)

    lg = logits.cuda().requires_grad_(True)
    out = stage(lg, u.cuda(), sync=True)
    cx = out["complex"]

    ref = OracleSCCN(C, 3, L).train()
    ref.load_state_dict({k: v.detach().cpu() for k, v in stage.sccn.state_dict().items()})
    ref64 = copy.deepcopy(ref).double()
    s32, loss_c, lc = _oracle_chain(stage, ref, tab, off, logits, u, bias_on, torch.float32)
    s64, loss_d, ld = _oracle_chain(stage, ref64, tab, off, logits, u, bias_on, torch.float64)

    hc = cx.host_counts
    for b in range(B):
        (emb, o, rect, vp, ent), (_, o64, _, _, _) = s32[b], s64[b]
        assert_close(f"stage/{bias_on}/b={b}/vertex_penalty", out["vertex_penalty"][b].reshape(1), vp.detach().reshape(1))
        assert_close(f"stage/{bias_on}/b={b}/entropy_loss", out["entropy_loss"][b].reshape(1), ent.detach().reshape(1))
        assert_close(f"stage/{bias_on}/b={b}/rectified", out["rectified"][b], torch.cat(rect))
        for r in range(4):
            idx = emb["active_indices"][NAMES[r]]
            got_idx = cx.act_idx[b, off[r]:off[r] + int(hc[b, r])].long().cpu()
            assert torch.equal(got_idx, idx), "active index sets must be bit-exact"
            rows = T.ComplexStage.split_per_sample(cx, out[f"rank_{r}"], r)[b]
            assert_fp32_equivalent(f"stage/{bias_on}/b={b}/rank_{r}", rows, o[f"rank_{r}"], o64[f"rank_{r}"])
    loss_g = sum(out[f"rank_{r}"].pow(2).sum() for r in range(4)) + 0.3 * out["vertex_penalty"].sum() \
        + 0.7 * out["entropy_loss"].sum()
    loss_g.backward()
    assert_fp32_equivalent(f"stage/{bias_on}/loss", loss_g.detach().reshape(1), loss_c.detach().reshape(1), loss_d.detach().reshape(1))
    _dump_worst(f"stage/{bias_on}/dlogits", lg.grad, lc.grad, ld.grad, out["rectified"], off)
    assert_fp32_equivalent(f"stage/{bias_on}/dlogits", lg.grad, lc.grad, ld.grad)


def test_generate_complex_reference_shape_and_empty_convention():
    import topo_audio_autoencoder_b200 as T
    n, C = 7, 64
    head = T.ComplexHead(n, embedding_dim=C, bias_on="logits").cuda().train()
    off = glo.rank_offsets(n)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(off[4], generator=g).cuda()
    u = torch.rand(off[4], generator=g).clamp_(1e-6, 1 - 1e-6).cuda()
    emb, mats = head.generate_complex(logits, u)
    assert set(emb) == {"rank_0", "rank_1", "rank_2", "rank_3"}
    assert set(mats.adjacencies) == {"rank_0", "rank_1", "rank_2", "rank_3"} and set(mats.incidences) == {"rank_1", "rank_2", "rank_3"}
    act = head.active_simplices
    for r, k in enumerate(NAMES):
        assert act[k].dtype == torch.int64 and emb[f"rank_{r}"].shape == (len(act[k]), C)
        assert mats.adjacencies[f"rank_{r}"].shape == (len(act[k]), len(act[k]))
    # an all-closed gate gives an empty complex: (None, None, None), never an exception (encoder.py:365-366)
    res = head.generate_complex(torch.full((off[4],), -50.0, device="cuda"), u)
    assert res == (None, None, None)


def test_gate_rectifier_penalty_chain_gradients():
    """The part of the stage upstream of the SCCN, isolated: logits -> Hard Concrete -> rectifier -> linear
    functional + penalties, gradients w.r.t. the logits and the gate / bias parameters."""
    import topo_audio_autoencoder_b200 as T
    n, B = 9, 4
    torch.manual_seed(4)
    head = T.ComplexHead(n, embedding_dim=64).cuda().train()
    tab, off = ro.make_tables(n), glo.rank_offsets(n)
    g = torch.Generator().manual_seed(99)
    logits = torch.randn(B, off[4], generator=g)
    u = torch.rand(B, off[4], generator=g).clamp_(1e-6, 1 - 1e-6)
    w = torch.randn(B, off[4], generator=g)

    lg = logits.cuda().requires_grad_(True)
    rect = head.rectified_batch(lg, u.cuda())
    vp = head.compute_vertex_penalty(rect[:, :off[1]])
    ent = head.compute_entropy_loss(*torch.split(rect, tab.sizes, dim=1))
    ((rect * w.cuda()).sum() + 0.3 * vp.sum() + 0.7 * ent.sum()).backward()

    runs = {}
    for dt in (torch.float32, torch.float64):
        t2 = copy.copy(tab)
        t2.v2e, t2.e2t, t2.t2tt = tab.v2e.to(dt), tab.e2t.to(dt), tab.t2tt.to(dt)
        lc = logits.to(dt).clone().requires_grad_(True)
        bias = [p.detach().cpu().to(dt).clone().requires_grad_(True) for p in
                (head.vertex_bias, head.edge_bias, head.triangle_bias, head.tetra_bias)]
        z = go.hard_concrete(lc, u.to(dt), head.sampler.current_temp, -0.1, 1.1, torch.relu(torch.cat(bias)), off)
        loss = 0.0
        rects = []
        for b in range(B):
            r = ro.enforce_constraints(*torch.split(z[b], tab.sizes), t2)
            rects.append(torch.cat(r))
            loss = loss + (torch.cat(r) * w[b].to(dt)).sum() + 0.3 * glo.vertex_penalty(r[0], 8, 16) + 0.7 * glo.entropy_loss(*r)
        loss.backward()
        runs[dt] = (lc, bias, torch.stack(rects))
    (lc, bc, rc), (ld, bd, rd) = runs[torch.float32], runs[torch.float64]
    assert_close("chain/rectified", rect, rc)
    _dump_worst("chain/dlogits", lg.grad, lc.grad, ld.grad, rect, off)
    assert_fp32_equivalent("chain/dlogits", lg.grad, lc.grad, ld.grad)
    for name, p, a, d in zip(("vertex", "edge", "triangle", "tetra"),
                             (head.vertex_bias, head.edge_bias, head.triangle_bias, head.tetra_bias), bc, bd):
        assert_fp32_equivalent(f"chain/d{name}_bias", p.grad, a.grad, d.grad, floor=1e-5)
