"""GPU: regressions for the round-1 advisor findings: a rank with no live row in the whole batch under
sync=True, gate temperatures under CUDA-graph replay, tables that follow the module to another device."""
import pytest
import torch

from oracle import glue_oracle as glo
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _stage(n=9, C=64, L=2, **kw):
    import topo_audio_autoencoder_b200 as T
    torch.manual_seed(3)
    return T.ComplexStage(n, channels=C, n_layers=L, **kw).cuda().train()


@pytest.mark.parametrize("closed", [(3,), (2, 3), (1, 2, 3)])
def test_rank_without_live_rows_under_sync(closed):
    """All gates of the highest ranks closed (e.g. no tetrahedron survives): forward(sync=True) and backward run, the
    dead ranks come back empty, the others equal the run in which those ranks are merely masked downstream."""
    import topo_audio_autoencoder_b200 as T
    stage = _stage(gate="hard_concrete", bias_on="logits")
    off = glo.rank_offsets(9)
    g = torch.Generator().manual_seed(1)
    B = 3
    logits = torch.randn(B, off[4], generator=g)
    for r in closed:
        logits[:, off[r]:off[r + 1]] = -60.0            # far below the clamp: z == 0 exactly
    u = torch.rand(B, off[4], generator=g).clamp_(1e-6, 1 - 1e-6)
    lg = logits.cuda().requires_grad_(True)
    out = stage(lg, u.cuda(), sync=True)
    cx = out["complex"]
    for r in closed:
        assert int(cx.host_counts[:, r].sum()) == 0
        assert all(t.shape[0] == 0 for t in T.ComplexStage.split_per_sample(cx, out[f"rank_{r}"], r))
    loss = sum(out[f"rank_{r}"].pow(2).sum() for r in range(4)) + out["vertex_penalty"].sum()
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(lg.grad).all()
    for r in closed:
        assert (lg.grad[:, off[r]:off[r + 1]] == 0).all(), "a clamped gate passes no gradient"
    # the same batch without the host synchronisation (buffers sized by the bound) must give the same live rows
    lg2 = logits.cuda().requires_grad_(True)
    out2 = stage(lg2, u.cuda(), sync=False)
    live = out2["complex"].row_off[:, B].tolist()
    for r in range(4):
        assert live[r] == int(cx.host_counts[:, r].sum())
        assert torch.equal(out2[f"rank_{r}"][:live[r]], out[f"rank_{r}"][:live[r]])
    loss2 = sum(out2[f"rank_{r}"].pow(2).sum() for r in range(4)) + out2["vertex_penalty"].sum()
    loss2.backward()
    assert_close("advice/empty-rank/dlogits", lg.grad, lg2.grad, rtol=1e-5, atol=1e-6 * max(1.0, lg2.grad.abs().max().item()))


@pytest.mark.parametrize("gate", ["hard_concrete", "binary_gumbel"])
def test_graph_replay_follows_the_annealed_temperature(gate):
    """trainer.py:266 rewrites sampler.current_temp every epoch: a captured step must see the new value."""
    from topo_audio_autoencoder_b200.graph import GraphedStep
    kw = dict(gate="hard_concrete", bias_on="logits") if gate == "hard_concrete" else dict(gate="binary_gumbel", bias_on="probs")
    stage = _stage(**kw)
    head = stage.head
    N, B, C = head.total_simplices, 2, 64
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(B, N, generator=g).cuda()
    if gate == "hard_concrete":
        noise = torch.rand(B, N, generator=g).clamp_(1e-6, 1 - 1e-6).cuda()
    else:
        noise = (-torch.empty(2, B, N).exponential_(generator=g).log()).cuda()
    ups = [torch.randn(B * c, C, generator=g).cuda() for c in head._tables.counts] + [torch.ones(B).cuda(), torch.ones(B).cuda()]
    graphed = GraphedStep(stage, logits, noise, ups)
    sampler = head.sampler if gate == "hard_concrete" else head.gumbel
    first = graphed.replay(logits, noise)["rectified"].clone()
    for temp in (0.3, 1.7):
        if gate == "hard_concrete":
            sampler.current_temp = temp                  # the assignment the reference's trainer makes
        else:
            sampler.set_temperature(temp)
        got = graphed.replay(logits, noise)["rectified"].clone()
        want = stage(logits, noise)["rectified"]
        assert torch.equal(got, want), f"replay ignored the temperature {temp}"
        assert not torch.equal(got, first)
    if gate == "hard_concrete":
        assert graphed.captures == 1, "Hard Concrete reads its temperature from device memory: no re-capture"
    else:
        assert graphed.captures == 3, "BinaryGumbel's temperature is a by-value kernel argument: re-captured per change"


def test_tables_follow_the_module_to_another_device():
    import topo_audio_autoencoder_b200 as T
    from topo_audio_autoencoder_b200._lib import TopoError
    if torch.cuda.device_count() < 2:
        # one visible device: the device check itself is still exercised
        head = T.ComplexHead(6, embedding_dim=64).cuda()
        assert head._tables.device == torch.device("cuda", torch.cuda.current_device())
        pytest.skip("needs two CUDA devices")
    head = T.ComplexHead(6, embedding_dim=64).to("cuda:0")
    t0 = head._tables
    head = head.to("cuda:1")
    assert head._tables.device == torch.device("cuda:1") and head._tables is not t0
    with torch.cuda.device(1):
        z = torch.rand(2, head.total_simplices, device="cuda:1")
        r = T.rectify_batch(z, head.constraints)
        assert torch.isfinite(r).all()
        stale = T.ConstraintMatrices(None, None, None, head.constraints.indices, _tables=t0)
        with pytest.raises(TopoError, match="another device"):
            T.rectify_batch(z, stale)


def test_hand_built_constraint_matrices_are_validated():
    import topo_audio_autoencoder_b200 as T
    ref = T.ConstraintMatrices.create(5)
    ok = T.ConstraintMatrices(ref.vertex_to_edge, ref.edge_to_triangle, ref.triangle_to_tetra, ref.indices)
    assert ok.n_vertices == 5
    bad = ref.edge_to_triangle.clone()
    bad[0, 0] = 1 - bad[0, 0]
    with pytest.raises(ValueError, match="canonical"):
        T.ConstraintMatrices(ref.vertex_to_edge, bad, ref.triangle_to_tetra, ref.indices)
    idx = T.SimplexIndices(edges=ref.indices.edges.flip(0), triangles=ref.indices.triangles, tetra=ref.indices.tetra)
    with pytest.raises(ValueError, match="itertools"):
        T.ConstraintMatrices(ref.vertex_to_edge, ref.edge_to_triangle, ref.triangle_to_tetra, idx)
