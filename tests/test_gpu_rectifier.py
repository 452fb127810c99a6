"""GPU parity: tables, rectifier and active sets through the C ABI vs the reference vectors / oracle."""
import numpy as np
import pytest
import torch

from oracle import rectifier_oracle as ro
from tests.helpers import NAMES, assert_close, golden_cases, hard_concrete_like, load_golden, sha

pytestmark = pytest.mark.gpu


def _cuda_leaves(fx):
    return [torch.from_numpy(fx[f"in_{k}"]).cuda().requires_grad_(True) for k in NAMES]


@pytest.mark.parametrize("case", golden_cases())
def test_enforce_constraints_matches_reference_vectors(case):
    import topo_audio_autoencoder_b200 as T
    fx = load_golden(case)
    n = int(fx["n_vertices"])
    mats = T.ConstraintMatrices.create(n)
    # tables: bit-exact index sets
    if "edges" in fx.files:
        assert np.array_equal(mats.indices.edges.cpu().numpy(), fx["edges"].astype(np.int64).reshape(-1, 2))
        assert np.array_equal(mats.indices.triangles.cpu().numpy(), fx["triangles"].astype(np.int64).reshape(-1, 3))
        assert np.array_equal(mats.indices.tetra.cpu().numpy(), fx["tetra"].astype(np.int64).reshape(-1, 4))
    else:
        assert [sha(mats.indices.edges), sha(mats.indices.triangles), sha(mats.indices.tetra)] == list(fx["tables_sha"])
    leaves = _cuda_leaves(fx)
    rect = T.enforce_constraints(*leaves, mats)
    outs = [rect.vertices, rect.edges, rect.triangles, rect.tetra]
    assert torch.equal(rect.all_simplices, torch.cat(outs))
    ups = [torch.from_numpy(fx[f"up_{k}"]).cuda() for k in NAMES]
    grads = torch.autograd.grad(outs, leaves, ups, allow_unused=True)
    for k, o, g in zip(NAMES, outs, grads):
        want = torch.from_numpy(fx[f"out_{k}"])
        assert_close(f"rectify/{case}/{k}", o, want)
        # the sparsity mask is bit-exact: zeros are exact zeros in both
        assert torch.equal(o.detach().cpu() == 0, want == 0), f"{case}: zero set of {k} differs"
        assert np.array_equal(o.detach().nonzero().squeeze(-1).cpu().numpy(), fx[f"active_{k}"])
        assert_close(f"rectify-grad/{case}/{k}", g, torch.from_numpy(fx[f"grad_{k}"]))


def test_face_matrices_match_oracle():
    import topo_audio_autoencoder_b200 as T
    for n in (4, 7):
        mats, tab = T.ConstraintMatrices.create(n), ro.make_tables(n)
        assert torch.equal(mats.vertex_to_edge.cpu(), tab.v2e)
        assert torch.equal(mats.edge_to_triangle.cpu(), tab.e2t)
        assert torch.equal(mats.triangle_to_tetra.cpu(), tab.t2tt)


def test_tie_rule_and_masked_branch():
    """own == constraint (0 == 0) halves the gradient; own = 0 < constraint keeps it all;
    the masked branch passes nothing to the faces."""
    import topo_audio_autoencoder_b200 as T
    n = 4
    mats, tab = T.ConstraintMatrices.create(n), ro.make_tables(n)
    v = torch.tensor([0.0, 0.5, 0.7, 0.9])
    e = torch.tensor([0.0, 0.3, 0.0, 0.0, 0.8, 0.6])   # edge 0 ties 0 == 0, edge 2: own 0, faces nonzero... 
    t = torch.tensor([0.2, 0.0, 0.4, 0.9])
    tt = torch.tensor([0.5])
    cpu = [x.clone().requires_grad_(True) for x in (v, e, t, tt)]
    gpu = [x.clone().cuda().requires_grad_(True) for x in (v, e, t, tt)]
    o_cpu = ro.enforce_constraints(*cpu, tab)
    r = T.enforce_constraints(*gpu, mats)
    o_gpu = [r.vertices, r.edges, r.triangles, r.tetra]
    ups = [torch.ones_like(x) for x in o_cpu]
    g_cpu = torch.autograd.grad(list(o_cpu), cpu, ups)
    g_gpu = torch.autograd.grad(o_gpu, gpu, [u.cuda() for u in ups])
    for k, a, b in zip(NAMES, g_gpu, g_cpu):
        assert_close(f"tie-rule/grad/{k}", a, b)
    assert g_gpu[1][0].item() == 0.5      # 0 == 0 tie


def test_batched_equals_per_sample_and_is_idempotent():
    import topo_audio_autoencoder_b200 as T
    n, B = 20, 5
    mats = T.ConstraintMatrices.create(n)
    g = torch.Generator().manual_seed(3)
    probs = hard_concrete_like((B, mats._tables.total), g).cuda()
    out = T.rectify_batch(probs, mats)
    c = mats._tables.counts
    for b in range(B):
        parts = torch.split(probs[b], c)
        single = T.enforce_constraints(*parts, mats)
        assert torch.equal(single.all_simplices, out[b])
    again = T.rectify_batch(out, mats)
    assert torch.equal(again, out), "rectification is idempotent"
    # a simplex with a zero face is exactly zero; nothing exceeds its own probability
    assert (out <= probs).all()
    tab = ro.make_tables(n)
    o = mats._tables.offsets
    e = out[:, o[1]:o[2]].cpu()
    vz = (out[:, :o[1]].cpu()[:, tab.edges] == 0).any(-1)
    assert (e[vz] == 0).all()


def test_large_complex_properties():
    """Config 3 sizes: beyond what the dense reference can hold, checked through invariants."""
    import topo_audio_autoencoder_b200 as T
    n, B = 40, 2
    mats = T.ConstraintMatrices.create(n)
    g = torch.Generator().manual_seed(5)
    probs = hard_concrete_like((B, mats._tables.total), g).cuda().requires_grad_(True)
    out = T.rectify_batch(probs, mats)
    assert (out <= probs).all() and torch.equal(T.rectify_batch(out.detach(), mats), out.detach())
    out.sum().backward()
    assert torch.isfinite(probs.grad).all()
    pos, act, counts, row_off = T.active_sets(out.detach(), mats._tables)
    o = mats._tables.offsets
    for r in range(4):
        seg = out[:, o[r]:o[r + 1]].detach()
        assert torch.equal(counts[:, r].cpu().long(), (seg != 0).sum(1).cpu())
        for b in range(B):
            want = seg[b].nonzero().squeeze(-1)
            got = act[b, o[r]:o[r] + len(want)].long()
            assert torch.equal(got, want)
            assert torch.equal(pos[b, o[r]:o[r + 1]][want].long(), torch.arange(len(want), device="cuda"))
            assert (pos[b, o[r]:o[r + 1]][seg[b] == 0] == -1).all()
    assert torch.equal(row_off[:, 1:].cpu().long(), counts.cpu().long().cumsum(0).t())
