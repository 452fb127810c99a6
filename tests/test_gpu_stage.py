"""GPU parity: the whole complex stage (gate -> rectifier -> active sets -> embeddings -> operators ->
SCCN -> penalties) against the per-sample oracle chain, and the reference-shaped generate_complex."""
import pytest
import torch

from oracle import complex_builder_oracle as cbo
from oracle import gate_oracle as go
from oracle import glue_oracle as glo
from oracle import rectifier_oracle as ro
from oracle.sccn_oracle import OracleSCCN
from tests.helpers import NAMES, assert_close

pytestmark = pytest.mark.gpu
DEEP = dict(rtol=1e-4, atol=1e-5)


def _emb_params(head):
    out = []
    for name in head._embedding_names:
        emb, ln = getattr(head, name)
        out.append(tuple(t.detach().cpu() for t in (emb.weight, ln.weight, ln.bias)))
    return out


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
    head = stage.head
    lc = logits.clone().requires_grad_(True)
    loc = torch.relu(torch.cat([p.detach().cpu() for p in (head.vertex_bias, head.edge_bias, head.triangle_bias, head.tetra_bias)]))
    if bias_on == "probs":
        loc = torch.zeros(4)
    beta = head.sampler.current_temp
    z = go.hard_concrete(lc, u, beta, head.sampler.gamma.item(), head.sampler.zeta.item(), loc, off)
    loss_c, loss_g = 0.0, 0.0
    hc = cx.host_counts
    for b in range(B):
        res = glo.complex_from_probs(z[b], n, head.vertex_bias.detach().cpu(), tab, _emb_params(head), bias_on == "probs")
        assert res is not None
        emb, (adj, inc), rect = res
        o = ref({f"rank_{r}": emb[f"rank_{r}"] for r in range(4)}, inc, adj)
        vp = glo.vertex_penalty(rect[0], head.min_active_vertices, head.max_active_vertices)
        ent = glo.entropy_loss(*rect)
        assert_close(f"stage/{bias_on}/b={b}/vertex_penalty", out["vertex_penalty"][b].reshape(1), vp.detach().reshape(1))
        assert_close(f"stage/{bias_on}/b={b}/entropy_loss", out["entropy_loss"][b].reshape(1), ent.detach().reshape(1))
        assert_close(f"stage/{bias_on}/b={b}/rectified", out["rectified"][b], torch.cat(rect))
        for r in range(4):
            idx = emb["active_indices"][NAMES[r]]
            got_idx = cx.act_idx[b, off[r]:off[r] + int(hc[b, r])].long().cpu()
            assert torch.equal(got_idx, idx), "active index sets must be bit-exact"
            rows = T.ComplexStage.split_per_sample(cx, out[f"rank_{r}"], r)[b]
            assert_close(f"stage/{bias_on}/b={b}/rank_{r}", rows, o[f"rank_{r}"], **DEEP)
            loss_c = loss_c + o[f"rank_{r}"].pow(2).sum()
        loss_c = loss_c + 0.3 * vp + 0.7 * ent
    for r in range(4):
        loss_g = loss_g + out[f"rank_{r}"].pow(2).sum()
    loss_g = loss_g + 0.3 * out["vertex_penalty"].sum() + 0.7 * out["entropy_loss"].sum()
    loss_c.backward()
    loss_g.backward()
    assert_close(f"stage/{bias_on}/loss", loss_g.detach().reshape(1), loss_c.detach().reshape(1), rtol=1e-4, atol=1e-3)
    assert_close(f"stage/{bias_on}/dlogits", lg.grad, lc.grad, rtol=1e-4, atol=1e-4)


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
