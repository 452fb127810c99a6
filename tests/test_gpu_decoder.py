"""GPU: the decoder consumer's cross-attention over compact rows (csrc/attention.cu) against stock PyTorch --
an fp64 scaled dot-product per sample, nn.MultiheadAttention on a padded memory, and the per-sample reference path of the
whole decoder tail (decoder.py:131-175), forward and backward."""
import pytest
import torch
import torch.nn.functional as F

from oracle.param_fill import fill_by_name
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_convolutions():
    """cuDNN convolutions default to TF32 on this GPU (torch.backends.cudnn.allow_tf32): inputs that differ in the last fp32
    bit can round to different TF32 values and come out 1e-3 apart, which would drown what these tests compare."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _segments(counts):
    """counts [B, 3] -> (seg [B, 3, 2] int32, total rows): rank after rank, samples in order inside a rank"""
    tot = counts.sum(0)
    base = torch.tensor([0, int(tot[0]), int(tot[0] + tot[1])])
    starts = torch.cumsum(counts, 0) - counts
    seg = torch.stack([base.unsqueeze(0) + starts, counts], dim=2).to(torch.int32)
    return seg, int(tot.sum())


def _reference(q, k, v, seg, heads):
    b, q_len, c = q.shape
    out = torch.zeros_like(q)
    for s in range(b):
        rows = torch.cat([torch.arange(int(a), int(a + n)) for a, n in seg[s].tolist()]).to(q.device)
        ks, vs = k[rows], v[rows]
        for h in range(heads):
            sl = slice(16 * h, 16 * h + 16)
            p = torch.softmax(q[s, :, sl] @ ks[:, sl].t() / 4.0, dim=-1)
            out[s, :, sl] = p @ vs[:, sl]
    return out


@pytest.mark.parametrize("counts,q_len", [
    ([[5, 0, 3], [1, 130, 0], [129, 257, 1]], 37),          # empty runs, runs around the 128 / 256 tile sizes, one-row runs
    ([[190, 1140, 4845], [17, 600, 3000]], 250),            # the full 20-vertex complex next to a sparse one, the decoder's L
    ([[3, 2, 1]], 300),                                     # more queries than two CTAs' worth, a tiny memory
])
def test_kernel_against_fp64_attention(counts, q_len):
    from topo_audio_autoencoder_b200.attention import segment_cross_attention
    counts = torch.tensor(counts)
    seg, rows = _segments(counts)
    g = torch.Generator().manual_seed(int(counts.sum()) + q_len)
    b, heads = counts.shape[0], 4
    q = torch.randn(b, q_len, 64, generator=g).cuda().requires_grad_(True)
    k = (torch.randn(rows, 64, generator=g) * 1.5).cuda().requires_grad_(True)
    v = torch.randn(rows, 64, generator=g).cuda().requires_grad_(True)
    up = torch.randn(b, q_len, 64, generator=g).cuda()
    out = segment_cross_attention(q, k, v, seg.cuda(), heads, int(counts.max()))
    gq, gk, gv = torch.autograd.grad(out, (q, k, v), up)
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    ref = _reference(q64, k64, v64, seg, heads)
    rq, rk, rv = torch.autograd.grad(ref, (q64, k64, v64), up.double())
    tag = f"attention/B={b}/L={q_len}/rows={rows}"
    assert_close(tag + "/out", out, ref.float(), rtol=1e-5, atol=1e-6)
    # gradients are fp32 sums of up to 300 (dk, dv) or 6,175 (dq) signed terms: absolute tolerance 1e-6 of the largest entry
    for name, got, want in (("dq", gq, rq), ("dk", gk, rk), ("dv", gv, rv)):
        assert_close(f"{tag}/{name}", got, want.float(), rtol=1e-5, atol=1e-6 * max(1.0, want.abs().max().item()))
    # deterministic: no atomics anywhere
    out2 = segment_cross_attention(q, k, v, seg.cuda(), heads, int(counts.max()))
    g2 = torch.autograd.grad(out2, (q, k, v), up)
    assert torch.equal(out, out2) and all(torch.equal(a, c) for a, c in zip((gq, gk, gv), g2))


def test_rows_outside_every_run_are_never_touched():
    from topo_audio_autoencoder_b200.attention import segment_cross_attention
    seg = torch.tensor([[[2, 3], [9, 0], [12, 5]]], dtype=torch.int32)           # rows 0-1, 5-11 and 17+ belong to nobody
    g = torch.Generator().manual_seed(4)
    q = torch.randn(1, 9, 64, generator=g).cuda().requires_grad_(True)
    k = torch.randn(20, 64, generator=g)
    v = torch.randn(20, 64, generator=g)
    live = torch.zeros(20, dtype=torch.bool)
    live[2:5] = True
    live[12:17] = True
    k[~live] = float("nan")
    v[~live] = float("nan")
    k, v = k.cuda().requires_grad_(True), v.cuda().requires_grad_(True)
    out = segment_cross_attention(q, k, v, seg.cuda(), 4, 5)
    assert torch.isfinite(out).all()
    gq, gk, gv = torch.autograd.grad(out.sum(), (q, k, v))
    assert torch.isfinite(gq).all() and torch.isfinite(gk[live.cuda()]).all()
    assert (gk[~live.cuda()] == 0).all() and (gv[~live.cuda()] == 0).all()


@pytest.mark.parametrize("full", [False, True])
def test_batched_consumer_on_compact_rows_equals_per_sample_reference_path(full):
    """DecoderTail.forward_batched (kernel path) against the reference's per-sample order of operations with the stock
    nn.MultiheadAttention (decoder.py:131-175), both on the GPU; dead rows of the compact buffers hold NaN."""
    from topo_audio_autoencoder_b200.decoder import DecoderTail
    tail = fill_by_name(DecoderTail(64, 250, 16), 5).train().cuda()
    g = torch.Generator().manual_seed(3)
    if full:
        counts = torch.tensor([[20, 190, 1140, 4845]] * 3)
    else:
        counts = torch.tensor([[4, 3, 0, 0], [6, 15, 20, 15], [1, 0, 2, 1], [20, 190, 1140, 4845], [9, 30, 7, 0]])
    tot = counts.sum(0).tolist()
    xs = [torch.cat([torch.randn(tot[r], 64, generator=g), torch.full((5, 64), float("nan"))]).cuda() for r in range(4)]
    leaves = [x.clone().requires_grad_(True) for x in xs]
    out = tail.forward_batched(leaves, counts)
    up = torch.randn(out.shape, generator=g).cuda()
    params = [p for _, p in tail.named_parameters()]
    gb = torch.autograd.grad(out, leaves + params, up, allow_unused=True)
    starts = torch.cumsum(counts, 0) - counts
    leaves2 = [x.clone().requires_grad_(True) for x in xs]
    outs = []
    for b in range(counts.shape[0]):
        sample = {f"rank_{r}": (leaves2[r][starts[b, r]:starts[b, r] + counts[b, r]] if counts[b, r] else None) for r in range(4)}
        outs.append(tail(sample))
    ref = torch.cat(outs)
    gs = torch.autograd.grad(ref, leaves2 + params, up, allow_unused=True)
    assert torch.isfinite(out).all()
    # (1) kernel path against the SAME batched pre-processing with the stock nn.MultiheadAttention on a padded memory:
    #     only the attention core differs -> fp32 rounding, forward and backward
    with torch.no_grad():
        att = tail.attend_batched(xs, counts)
    stock = fill_by_name(DecoderTail(64, 250, 16), 5).train().cuda()
    stock._attend_compact = lambda q, k, v, c, base, st: stock._attend_padded(q, k, v, c, base, st, c[:, 1:].sum(dim=1))
    leaves3 = [x.clone().requires_grad_(True) for x in xs]
    with torch.no_grad():
        att_stock = stock.attend_batched(xs, counts)
    assert_close(f"decoder/kernel-vs-stock-mha/attend/full={full}", att, att_stock, rtol=1e-5, atol=2e-6 * att_stock.abs().max().item())
    out_stock = stock.forward_batched(leaves3, counts)
    g3 = torch.autograd.grad(out_stock, leaves3, up)
    assert_close(f"decoder/kernel-vs-stock-mha/out/full={full}", out, out_stock, rtol=1e-4, atol=2e-5 * out_stock.abs().max().item())
    for r in range(4):
        assert_close(f"decoder/kernel-vs-stock-mha/dx{r}/full={full}", gb[r][:tot[r]], g3[r][:tot[r]], rtol=1e-3,
                     atol=2e-5 * g3[r][:tot[r]].abs().max().item())
    # (2) against the reference's per-sample order of operations: batched vs single-sample cuDNN / cuBLAS calls round
    #     differently before the LayerNorms / GroupNorms, which divide by the deviation of a sample's signal (tiny for a
    #     sample made of one or two vertices, whose 250 queries are nearly identical)
    with torch.no_grad():
        att_ref = torch.cat([tail.attend({f"rank_{r}": (xs[r][starts[b, r]:starts[b, r] + counts[b, r]] if counts[b, r] else None)
                                          for r in range(4)}) for b in range(counts.shape[0])])
    assert_close(f"decoder/batched-kernel/attend/full={full}", att, att_ref, rtol=1e-4, atol=2e-5 * att_ref.abs().max().item())
    assert_close(f"decoder/batched-kernel/full={full}", out, ref, rtol=1e-3, atol=2e-4 * ref.abs().max().item())
    for r in range(4):
        assert (gb[r][tot[r]:] == 0).all(), "dead rows received gradient"
        scale = gs[r][:tot[r]].abs().max().item()
        assert_close(f"decoder/batched-kernel/dx{r}/full={full}", gb[r][:tot[r]], gs[r][:tot[r]], rtol=1e-3, atol=5e-4 * scale)
    g_max = max(c.abs().max().item() for c in gs[4:] if c is not None)
    for (name, _), a, c in zip(tail.named_parameters(), gb[4:], gs[4:]):
        if a is None or c is None:
            assert a is None and c is None, name
            continue
        scale = c.abs().max().item() + 1e-12
        # (a bias in front of a LayerNorm, and the key bias -- it shifts every score of a query alike -- have mathematically
        # zero gradients: both sides hold the rounding noise of sums over all rows, hence the floor relative to the largest one)
        assert (a - c).abs().max().item() <= 5e-4 * scale + 1e-5 * g_max, (name, (a - c).abs().max().item(), scale, g_max)


def test_kernel_equals_stock_multihead_attention_on_padded_memory():
    """the whole module (in-projection, heads, out-projection): _attend_compact vs nn.MultiheadAttention with a mask"""
    from topo_audio_autoencoder_b200.decoder import DecoderTail
    tail = fill_by_name(DecoderTail(64, 250, 16), 9).cuda()
    counts = torch.tensor([[2, 40, 300, 700], [2, 7, 0, 55]])
    g = torch.Generator().manual_seed(8)
    q = torch.randn(2, 250, 64, generator=g).cuda()
    tot = counts.sum(0).tolist()
    keys = torch.randn(sum(tot[1:]), 64, generator=g).cuda()
    values = torch.randn(sum(tot[1:]), 64, generator=g).cuda()
    base = [0, tot[1], tot[1] + tot[2]]
    starts = [torch.cumsum(counts[:, r], 0) - counts[:, r] for r in (1, 2, 3)]
    a = tail._attend_compact(q, keys, values, counts, base, starts)
    b = tail._attend_padded(q, keys, values, counts, base, starts, counts[:, 1:].sum(1))
    assert_close("decoder/compact-vs-padded-mha", a, b, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("c", [32, 64, 128])
def test_row_layernorm_kernel_against_torch(c):
    """topo_layernorm_fwd/bwd at the consumer's sizes (row counts that do not divide the rows-per-warp grouping), 2-D and 3-D inputs"""
    from topo_audio_autoencoder_b200.decoder import RowLayerNorm
    g = torch.Generator().manual_seed(c)
    ln = RowLayerNorm(c).cuda()
    with torch.no_grad():
        ln.weight.copy_(1 + 0.1 * torch.randn(c, generator=g))
        ln.bias.copy_(0.1 * torch.randn(c, generator=g))
    for shape in [(100003, c), (7, 251, c), (1, c)]:
        x = (torch.randn(shape, generator=g) * 3 + 0.5).cuda().requires_grad_(True)
        up = torch.randn(shape, generator=g).cuda()
        y = ln(x)
        gx, gw, gb = torch.autograd.grad(y, (x, ln.weight, ln.bias), up)
        x64 = x.detach().double().requires_grad_(True)
        w64, b64 = ln.weight.detach().double().requires_grad_(True), ln.bias.detach().double().requires_grad_(True)
        ref = F.layer_norm(x64, (c,), w64, b64, ln.eps)
        rx, rw, rb = torch.autograd.grad(ref, (x64, w64, b64), up.double())
        tag = f"row-layernorm/C={c}/{'x'.join(map(str, shape))}"
        assert_close(tag + "/y", y, ref.float(), rtol=1e-5, atol=1e-6)
        assert_close(tag + "/dx", gx, rx.float(), rtol=1e-5, atol=2e-6)
        assert_close(tag + "/dgamma", gw, rw.float(), rtol=1e-5, atol=1e-5 * rw.abs().max().item())
        assert_close(tag + "/dbeta", gb, rb.float(), rtol=1e-5, atol=1e-5 * rb.abs().max().item())
