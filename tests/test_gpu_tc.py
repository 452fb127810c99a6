"""GPU: the tcgen05 (3xTF32) GEMM primitive against an fp64 matmul."""
import pytest
import torch

from tests.helpers import report

pytestmark = pytest.mark.gpu

_DEBUG = []


def _debug_lib():
    """libtopo_b200_debug.so: the unit-test GEMM entry points are not part of the product library"""
    if not _DEBUG:
        from topo_audio_autoencoder_b200 import _lib
        _DEBUG.append(_lib.load_debug())
    return _DEBUG[0]


@pytest.mark.parametrize("rows", [1, 128, 1000, 20000])
def test_tf32x3_gemm_is_fp32_accurate(rows):
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream
    g = torch.Generator().manual_seed(rows)
    a = (torch.randn(rows, 64, generator=g) * torch.logspace(-3, 3, 64)).cuda()     # wide dynamic range per column
    w = torch.randn(64, 64, generator=g).cuda()
    out = torch.full((rows, 64), float("nan"), device="cuda")
    check(_debug_lib().topo_debug_gemm_tf32x3(ptr(a), ptr(w), rows, 0, ptr(out), stream()))
    torch.cuda.synchronize()
    want64 = a.double() @ w.double()
    fp32 = (a @ w).double()                       # cuBLAS fp32 (no TF32) as the accuracy yardstick
    scale = (a.double().abs() @ w.double().abs())  # per-element condition scale sum|a||w|
    ours = ((out.double() - want64).abs() / scale).max().item()
    theirs = ((fp32 - want64).abs() / scale).max().item()
    report(f"tc/gemm3xtf32/rows={rows}", out, want64.float())
    assert torch.isfinite(out).all()
    assert ours < 2e-6, f"relative-to-condition error {ours:.3e} (cuBLAS fp32: {theirs:.3e})"


@pytest.mark.parametrize("n_msgs,apply_ln,rows", [(3, True, 1000), (2, False, 130), (1, True, 64), (3, True, 40000)])
def test_tensor_core_combine_forward_matches_the_fp32_kernel(n_msgs, apply_ln, rows):
    """The tcgen05 combine and the FFMA combine are two implementations of one contract."""
    import ctypes as C
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream
    from topo_audio_autoencoder_b200.custom_sccn import _make_params
    from tests.helpers import assert_close
    g = torch.Generator().manual_seed(rows + n_msgs)
    ch = 64
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()     # noqa: E731
    aggs = [rnd(rows, ch) * 2 for _ in range(n_msgs)]
    ws = [rnd(ch, ch) * 0.2 for _ in range(n_msgs)]
    scales = [torch.tensor([0.7 + 0.2 * k]).cuda() for k in range(n_msgs)]
    x = rnd(rows, ch)
    tensors = [rnd(ch, ch) * 0.2, rnd(ch) * 0.1, rnd(ch) * 0.3, rnd(1), 1 + 0.1 * rnd(ch), 0.1 * rnd(ch)]
    params = _make_params(ch, n_msgs, aggs, ws, scales, x, tensors, 1e-5, apply_ln)
    live = torch.tensor([rows - 3], dtype=torch.int32).cuda() if rows > 100 else None
    outs = []
    for fn in (lib.topo_sccn_combine_fwd, lib.topo_sccn_combine_fwd_tc):
        out = torch.zeros(rows, ch, device="cuda")
        check(fn(C.byref(params), rows, ptr(live, torch.int32), ptr(out), stream()))
        torch.cuda.synchronize()
        outs.append(out)
    # fp64 reference of the same formula
    m = [s.double() * (a.double() @ w.double()) + x.double() for a, w, s in zip(aggs, ws, scales)]
    w1, b1, w2, b2, gam, bet = (t.double() for t in tensors)
    sc = torch.stack([torch.nn.functional.gelu(mk @ w1.t() + b1) @ w2 + b2 for mk in m])
    att = torch.softmax(sc, dim=0)
    ref = sum(att[k].unsqueeze(1) * m[k] for k in range(n_msgs))
    if apply_ln:
        ref = torch.nn.functional.layer_norm(ref, (ch,), gam, bet, 1e-5)
    n_live = rows - 3 if live is not None else rows
    ref[n_live:] = 0
    e_simt = (outs[0].double() - ref).abs().max().item()
    e_tc = (outs[1].double() - ref).abs().max().item()
    report(f"tc/combine-fwd/msgs={n_msgs}/ln={apply_ln}/rows={rows}", outs[1], outs[0])
    assert e_tc <= 4 * e_simt + 1e-6, f"tensor-core error {e_tc:.3e} vs FFMA error {e_simt:.3e} (both against fp64)"
    assert (outs[1][n_live:] == 0).all(), "rows past the live count must not be touched"


def test_a_operand_from_tensor_memory():
    """mode 3: the A operand of the MMA lives in tensor memory (written with tcgen05.st by the row threads)."""
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream
    g = torch.Generator().manual_seed(9)
    a = torch.randn(500, 64, generator=g).cuda()
    w = torch.randn(64, 64, generator=g).cuda()
    out = torch.full((500, 64), float("nan"), device="cuda")
    check(_debug_lib().topo_debug_gemm_tf32x3(ptr(a), ptr(w), 500, 3, ptr(out), stream()))
    torch.cuda.synchronize()
    want = a.double() @ w.double()
    scale = a.double().abs() @ w.double().abs()
    err = ((out.double() - want).abs() / scale).max().item()
    report("tc/gemm3xtf32/a-in-tmem", out, want.float())
    assert err < 2e-6, err


@pytest.mark.parametrize("n_msgs,rows", [(3, 1000), (2, 130), (1, 64), (3, 40000)])
def test_tensor_core_conv_backward_matches_the_fp32_kernel(n_msgs, rows):
    """dL/dagg_k and the weight-gradient product agg_k^T dL/dm_k: tcgen05 kernel vs FFMA kernel vs fp64."""
    import ctypes as C
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream, CombineGrads
    from topo_audio_autoencoder_b200.custom_sccn import _make_params
    g = torch.Generator().manual_seed(rows * 7 + n_msgs)
    ch = 64
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()     # noqa: E731
    aggs = [rnd(rows, ch) for _ in range(n_msgs)]
    ws = [rnd(ch, ch) * 0.2 for _ in range(n_msgs)]
    scales = [torch.tensor([0.7 + 0.2 * k]).cuda() for k in range(n_msgs)]
    tensors = [rnd(ch, ch), rnd(ch), rnd(ch), rnd(1), rnd(ch), rnd(ch)]
    params = _make_params(ch, n_msgs, aggs, ws, scales, None, tensors, 1e-5, False)
    dm = rnd(n_msgs, rows, ch).contiguous()
    live_n = rows - 3 if rows > 100 else rows
    live = torch.tensor([live_n], dtype=torch.int32).cuda()
    results = []
    for fn in (lib.topo_sccn_combine_bwd_conv, lib.topo_sccn_combine_bwd_conv_tc):
        g_aggs = [torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)]
        wprod = [torch.zeros(ch, ch, device="cuda") for _ in range(n_msgs)]
        grads = CombineGrads()
        for k in range(n_msgs):
            grads.g_agg[k], grads.g_wprod[k] = ptr(g_aggs[k]), ptr(wprod[k])
        check(fn(C.byref(params), rows, ptr(live, torch.int32), C.byref(grads), ptr(dm), stream()))
        torch.cuda.synchronize()
        results.append((g_aggs, wprod))
    for k in range(n_msgs):
        d = dm[k, :live_n].double()
        want_g = scales[k].double() * (d @ ws[k].double().t())
        cond_g = scales[k].double().abs() * (d.abs() @ ws[k].double().abs().t())
        want_p = aggs[k][:live_n].double().t() @ d
        cond_p = aggs[k][:live_n].double().abs().t() @ d.abs()
        for name, idx, want, cond, sl in (("g_agg", 0, want_g, cond_g, slice(0, live_n)), ("wprod", 1, want_p, cond_p, slice(None))):
            tcv = results[1][idx][k][sl].double()
            err = ((tcv - want).abs() / cond).max().item()        # relative to sum |a||b|, like the GEMM primitive
            report(f"tc/conv-bwd/{name}/msgs={n_msgs}/rows={rows}/k={k}", tcv.float(), want.float())
            assert err < 2e-6, (name, k, err)
        assert (results[1][0][k][live_n:] == 0).all()


@pytest.mark.parametrize("mode,shape", [(0, None), (1, None), (2, (64, 64))])
@pytest.mark.parametrize("rows", [1, 128, 300, 20000])
def test_bf16x3_gemm_in_both_operand_majors(mode, shape, rows):
    """One 16-bit SWIZZLE_128B image read K-major (mode 1) and MN-major (modes 0, 2: y = x W with W stored
    [in][out], and the row contraction x^T y) -- the descriptor fields pinned here are the ones tc16.cuh uses."""
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream
    g = torch.Generator().manual_seed(rows + mode)
    a = (torch.randn(rows, 64, generator=g) * torch.logspace(-3, 3, 64)).cuda()
    second = (torch.randn(rows, 64, generator=g) if mode == 2 else torch.randn(64, 64, generator=g)).cuda()
    out = torch.zeros(*(shape or (rows, 64)), device="cuda")
    check(_debug_lib().topo_debug_gemm_bf16x3(ptr(a), ptr(second), rows, mode, 16384, 1024, 2048, ptr(out), stream()))
    torch.cuda.synchronize()
    a64, s64 = a.double(), second.double()
    if mode == 0:
        want, cond = a64 @ s64, a64.abs() @ s64.abs()
    elif mode == 1:
        want, cond = a64 @ s64.t(), a64.abs() @ s64.abs().t()
    else:
        want, cond = a64.t() @ s64, a64.abs().t() @ s64.abs()
    err = ((out.double() - want).abs() / cond).max().item()
    report(f"tc/gemm-bf16x3/mode={mode}/rows={rows}", out, want.float())
    assert torch.isfinite(out).all()
    assert err < 5e-7, f"relative-to-condition error {err:.3e}"


@pytest.mark.parametrize("n_msgs,apply_ln,rows,residual", [(3, True, 1000, True), (2, False, 130, True), (1, True, 64, False),
                                                            (2, True, 40000, True), (3, False, 5000, True)])
def test_fused_tensor_core_backward_matches_fp64_autograd(n_msgs, apply_ln, rows, residual):
    """topo_sccn_combine_bwd_tc (one bf16x3 kernel) against fp64 autograd of the same formula, with the FFMA +
    3xTF32 two-kernel backward as the accuracy yardstick."""
    import ctypes as C
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream, CombineGrads
    from topo_audio_autoencoder_b200.custom_sccn import _make_params
    g = torch.Generator().manual_seed(rows * 3 + n_msgs)
    ch = 64
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()     # noqa: E731
    aggs = [rnd(rows, ch) * 1.5 for _ in range(n_msgs)]
    ws = [rnd(ch, ch) * 0.2 for _ in range(n_msgs)]
    scales = [torch.tensor([0.7 + 0.2 * k]).cuda() for k in range(n_msgs)]
    x = rnd(rows, ch) if residual else None
    tensors = [rnd(ch, ch) * 0.2, rnd(ch) * 0.1, rnd(ch) * 0.3, rnd(1), 1 + 0.1 * rnd(ch), 0.1 * rnd(ch)]
    g_out = rnd(rows, ch)
    live_n = rows - 3 if rows > 100 else rows
    live = torch.tensor([live_n], dtype=torch.int32).cuda()
    saved = ([torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)],
             [torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)], torch.zeros(3, rows, device="cuda"))
    params = _make_params(ch, n_msgs, aggs, ws, scales, x, tensors, 1e-5, apply_ln, saved)
    out = torch.zeros(rows, ch, device="cuda")
    check(lib.topo_sccn_combine_fwd_tc(C.byref(params), rows, ptr(live, torch.int32), ptr(out), stream()))

    def run(fused):
        res = {"g_agg": [torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)],
               "wprod": [torch.zeros(ch, ch, device="cuda") for _ in range(n_msgs)],
               "g_x": torch.zeros(rows, ch, device="cuda") if residual else None,
               "w1": torch.zeros(ch, ch, device="cuda"), "b1": torch.zeros(ch, device="cuda"),
               "w2": torch.zeros(ch, device="cuda"), "b2": torch.zeros(1, device="cuda"),
               "gamma": torch.zeros(ch, device="cuda"), "beta": torch.zeros(ch, device="cuda")}
        grads = CombineGrads()
        for k in range(n_msgs):
            grads.g_agg[k], grads.g_wprod[k] = ptr(res["g_agg"][k]), ptr(res["wprod"][k])
        grads.g_x = ptr(res["g_x"])
        grads.g_att_w1, grads.g_att_b1, grads.g_att_w2, grads.g_att_b2 = ptr(res["w1"]), ptr(res["b1"]), ptr(res["w2"]), ptr(res["b2"])
        grads.g_ln_gamma, grads.g_ln_beta = ptr(res["gamma"]), ptr(res["beta"])
        if fused:
            check(lib.topo_sccn_combine_bwd_tc(C.byref(params), rows, ptr(live, torch.int32), ptr(g_out), C.byref(grads), stream()))
        else:
            ws_buf = torch.zeros(n_msgs * rows * ch, device="cuda")
            check(lib.topo_sccn_combine_bwd_attention(C.byref(params), rows, ptr(live, torch.int32), ptr(g_out), C.byref(grads),
                                                      ptr(ws_buf), stream()))
            check(lib.topo_sccn_combine_bwd_conv_tc(C.byref(params), rows, ptr(live, torch.int32), C.byref(grads), ptr(ws_buf), stream()))
        torch.cuda.synchronize()
        return res

    fused, two_kernel = run(True), run(False)

    # fp64 autograd of the same formula on the live rows
    leaf = lambda t: t[:live_n].double().clone().requires_grad_(True) if t.dim() == 2 and t.shape[0] == rows else t.double().clone().requires_grad_(True)  # noqa: E731
    a64 = [leaf(a) for a in aggs]
    w64 = [w.double().clone().requires_grad_(True) for w in ws]
    x64 = leaf(x) if residual else None
    w1, b1, w2, b2, gam, bet = (t.double().clone().requires_grad_(True) for t in tensors)
    prods = [a @ w for a, w in zip(a64, w64)]
    for p_ in prods:
        p_.retain_grad()
    m = [s.double() * p_ + (x64 if residual else 0) for p_, s in zip(prods, scales)]
    sc = torch.stack([torch.nn.functional.gelu(mk @ w1.t() + b1) @ w2 + b2 for mk in m])
    att = torch.softmax(sc, dim=0)
    y = sum(att[k].unsqueeze(1) * m[k] for k in range(n_msgs))
    if apply_ln:
        y = torch.nn.functional.layer_norm(y, (ch,), gam, bet, 1e-5)
    y.backward(g_out[:live_n].double())

    def check_one(name, got_f, got_2, want):
        scale = want.abs().max().item() + 1e-30
        e_f = (got_f.double() - want).abs().max().item() / scale
        e_2 = (got_2.double() - want).abs().max().item() / scale
        report(f"tc/bwd-fused/{name}/msgs={n_msgs}/ln={apply_ln}/rows={rows}", got_f.float(), want.float())
        assert torch.isfinite(got_f).all(), name
        assert e_f <= 4 * e_2 + 2e-6, f"{name}: fused {e_f:.3e} vs two-kernel {e_2:.3e} (relative to max |want|, both against fp64)"

    for k in range(n_msgs):
        check_one(f"g_agg{k}", fused["g_agg"][k][:live_n], two_kernel["g_agg"][k][:live_n], a64[k].grad)
        # g_wprod[k] = agg_k^T dL/dm_k  =  d/d(agg_k W_k) scaled back by 1 / scale_k
        want_p = a64[k].detach().t() @ (prods[k].grad / scales[k].double())
        check_one(f"wprod{k}", fused["wprod"][k], two_kernel["wprod"][k], want_p)
        assert (fused["g_agg"][k][live_n:] == 0).all()
    if residual:
        check_one("g_x", fused["g_x"][:live_n], two_kernel["g_x"][:live_n], x64.grad)
    check_one("w1", fused["w1"], two_kernel["w1"], w1.grad)
    check_one("b1", fused["b1"], two_kernel["b1"], b1.grad)
    check_one("w2", fused["w2"], two_kernel["w2"], w2.grad)
    check_one("b2", fused["b2"], two_kernel["b2"], b2.grad)
    if apply_ln:
        check_one("gamma", fused["gamma"], two_kernel["gamma"], gam.grad)
        check_one("beta", fused["beta"], two_kernel["beta"], bet.grad)


@pytest.mark.parametrize("n_msgs,apply_ln,rows,residual,max_ctas", [
    (3, True, 1000, True, 0), (2, False, 130, True, 0), (1, True, 64, False, 0), (2, True, 40000, True, 0), (3, False, 5000, True, 0),
    (2, True, 257, False, 0),
    # few CTAs, many tiles each: the forward's staging stream runs ahead across tile borders (three accumulators, two H0
    # buffers) -- every message count with and without the residual unit
    (3, True, 3000, False, 2), (1, False, 2000, False, 3), (2, True, 3000, True, 1), (1, True, 1500, True, 2), (3, True, 2500, True, 1),
    (2, False, 2000, False, 1)])
def test_second_generation_forward_and_its_backward(n_msgs, apply_ln, rows, residual, max_ctas):
    """topo_sccn_combine_fwd_tc2 (bf16x3, one product per message, tile-fragment saves) against fp64, with the
    FFMA forward as the accuracy yardstick; then the fused backward on its tile-fragment saves against the same
    backward on the first-generation forward's row-major saves (two layouts, one result)."""
    import ctypes as C
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream, CombineGrads
    from topo_audio_autoencoder_b200.custom_sccn import _make_params
    g = torch.Generator().manual_seed(rows * 5 + n_msgs)
    ch = 64
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()     # noqa: E731
    aggs = [rnd(rows, ch) * 2 for _ in range(n_msgs)]
    ws = [rnd(ch, ch) * 0.2 for _ in range(n_msgs)]
    scales = [torch.tensor([0.7 + 0.2 * k]).cuda() for k in range(n_msgs)]
    x = rnd(rows, ch) if residual else None
    tensors = [rnd(ch, ch) * 0.2, rnd(ch) * 0.1, rnd(ch) * 0.3, rnd(1), 1 + 0.1 * rnd(ch), 0.1 * rnd(ch)]
    g_out = rnd(rows, ch)
    live_n = rows - 3 if rows > 100 else rows
    live = torch.tensor([live_n], dtype=torch.int32).cuda()
    rows_pad = -(-rows // 128) * 128

    def forward(gen):
        pad = rows_pad if gen == 2 else rows
        saved = ([torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)],
                 [torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)], torch.zeros(3, rows, device="cuda"))
        params = _make_params(ch, n_msgs, aggs, ws, scales, x, tensors, 1e-5, apply_ln, saved, gen == 2, None, max_ctas)
        out = torch.zeros(rows, ch, device="cuda")
        fn = {0: lib.topo_sccn_combine_fwd, 1: lib.topo_sccn_combine_fwd_tc, 2: lib.topo_sccn_combine_fwd_tc2}[gen]
        check(fn(C.byref(params), rows, ptr(live, torch.int32), ptr(out), stream()))
        torch.cuda.synchronize()
        return out, params, saved

    out0, _, _ = forward(0)
    out1, params1, saved1 = forward(1)
    out2, params2, saved2 = forward(2)
    m = [s.double() * (a.double() @ w.double()) + (x.double() if residual else 0) for a, w, s in zip(aggs, ws, scales)]
    w1, b1, w2, b2, gam, bet = (t.double() for t in tensors)
    pre = [mk @ w1.t() + b1 for mk in m]
    sc = torch.stack([torch.nn.functional.gelu(p_) @ w2 + b2 for p_ in pre])
    att = torch.softmax(sc, dim=0)
    ref = sum(att[k].unsqueeze(1) * m[k] for k in range(n_msgs))
    if apply_ln:
        ref = torch.nn.functional.layer_norm(ref, (ch,), gam, bet, 1e-5)
    ref[live_n:] = 0
    e0 = (out0.double() - ref).abs().max().item()
    e2 = (out2.double() - ref).abs().max().item()
    report(f"tc/fwd2/out/msgs={n_msgs}/ln={apply_ln}/rows={rows}", out2, ref.float())
    assert torch.isfinite(out2).all()
    assert e2 <= 4 * e0 + 1e-6, f"second-generation forward error {e2:.3e} vs FFMA error {e0:.3e} (both against fp64)"
    assert (out2[live_n:] == 0).all(), "rows past the live count must not be touched"
    # the saved tensors, un-permuted: element (row, col) sits at tile*8192 + ((col//16*4 + col%16//4)*128 + row%128)*4 + col%4
    rr = torch.arange(rows_pad, device="cuda").unsqueeze(1)
    cc = torch.arange(ch, device="cuda").unsqueeze(0)
    pos = (rr // 128) * 8192 + ((cc // 16 * 4 + cc % 16 // 4) * 128 + rr % 128) * 4 + cc % 4
    for k in range(n_msgs):
        got_m = saved2[0][k].reshape(-1)[pos][:live_n]
        got_p = saved2[1][k].reshape(-1)[pos][:live_n]
        report(f"tc/fwd2/saved_m{k}/msgs={n_msgs}/rows={rows}", got_m, m[k][:live_n].float())
        assert (got_m.double() - m[k][:live_n]).abs().max().item() <= 2e-6 * m[k].abs().max().item() + 1e-6
        assert (got_p.double() - pre[k][:live_n]).abs().max().item() <= 2e-6 * pre[k].abs().max().item() + 2e-6
        assert (saved2[2][k][:live_n].double() - sc[k][:live_n]).abs().max().item() <= 4e-6 * sc.abs().max().item() + 2e-6

    def backward(params):
        res = {"g_agg": [torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)],
               "wprod": [torch.zeros(ch, ch, device="cuda") for _ in range(n_msgs)],
               "g_x": torch.zeros(rows, ch, device="cuda") if residual else None,
               "w1": torch.zeros(ch, ch, device="cuda"), "b1": torch.zeros(ch, device="cuda"),
               "w2": torch.zeros(ch, device="cuda"), "b2": torch.zeros(1, device="cuda"),
               "gamma": torch.zeros(ch, device="cuda"), "beta": torch.zeros(ch, device="cuda")}
        grads = CombineGrads()
        for k in range(n_msgs):
            grads.g_agg[k], grads.g_wprod[k] = ptr(res["g_agg"][k]), ptr(res["wprod"][k])
        grads.g_x = ptr(res["g_x"])
        grads.g_att_w1, grads.g_att_b1, grads.g_att_w2, grads.g_att_b2 = ptr(res["w1"]), ptr(res["b1"]), ptr(res["w2"]), ptr(res["b2"])
        grads.g_ln_gamma, grads.g_ln_beta = ptr(res["gamma"]), ptr(res["beta"])
        check(lib.topo_sccn_combine_bwd_tc(C.byref(params), rows, ptr(live, torch.int32), ptr(g_out), C.byref(grads), stream()))
        torch.cuda.synchronize()
        return res

    b1_, b2_ = backward(params1), backward(params2)
    for name in ("g_agg", "wprod"):
        for k in range(n_msgs):
            a_, b_ = b1_[name][k], b2_[name][k]
            tol = 2e-5 * a_.abs().max().item() + 1e-6
            assert (a_ - b_).abs().max().item() <= tol, (name, k, (a_ - b_).abs().max().item(), tol)
    for name in ("g_x", "w1", "b1", "w2", "b2", "gamma", "beta"):
        a_, b_ = b1_[name], b2_[name]
        if a_ is None or (name in ("gamma", "beta") and not apply_ln):
            continue
        tol = 2e-5 * a_.abs().max().item() + 2e-6
        if name == "b2":      # softmax is shift invariant: the exact gradient is 0 and both values are pure round-off
            tol = 1e-4
        assert (a_ - b_).abs().max().item() <= tol, (name, (a_ - b_).abs().max().item(), tol)


@pytest.mark.parametrize("rows", [128, 256, 1000, 20000])
def test_cta_pair_gemm_shares_the_weight_image(rows):
    """mode 4: tcgen05.mma.cta_group::2 issued by clusters of two CTAs, each holding its own 128-row operand tile and
    HALF of the weight image (M = 256 product, 128 accumulator rows per CTA): remote mbarrier arrival, multicast commit,
    cta_group::2 tensor-memory allocation -- the primitive for halving the resident weight images of the combine kernels."""
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream
    g = torch.Generator().manual_seed(rows + 4)
    a = (torch.randn(rows, 64, generator=g) * torch.logspace(-3, 3, 64)).cuda()
    w = torch.randn(64, 64, generator=g).cuda()
    out = torch.full((rows, 64), float("nan"), device="cuda")
    check(_debug_lib().topo_debug_gemm_bf16x3(ptr(a), ptr(w), rows, 4, 0, 0, 0, ptr(out), stream()))
    torch.cuda.synchronize()
    want = a.double() @ w.double().t()
    cond = a.double().abs() @ w.double().abs().t()
    err = ((out.double() - want).abs() / cond).max().item()
    report(f"tc/gemm-bf16x3/cta-pair/rows={rows}", out, want.float())
    assert torch.isfinite(out).all()
    assert err < 5e-7, f"relative-to-condition error {err:.3e}"

@pytest.mark.gpu
@pytest.mark.parametrize("n_msgs,residual,max_ctas", [(2, True, 0), (2, True, 3), (3, True, 5), (3, False, 2), (1, True, 4)])
def test_forward_and_row_gradients_are_bit_reproducible(n_msgs, residual, max_ctas):
    """The kernels hand shared-memory slots, operand slots and tensor-memory accumulators from one tile to the next
    without a closing barrier; a hazard there would show up as run-to-run differences.  Ten runs of the forward (output,
    saved activations, scores) and of the fused backward's row gradients must agree bit for bit (the parameter
    gradients meet in global atomics across CTAs and are excluded)."""
    import ctypes as C
    from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream, CombineGrads
    from topo_audio_autoencoder_b200.custom_sccn import _make_params
    rows, ch = 20000, 64
    g = torch.Generator().manual_seed(77 + n_msgs)
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()     # noqa: E731
    aggs = [rnd(rows, ch) * 2 for _ in range(n_msgs)]
    ws = [rnd(ch, ch) * 0.2 for _ in range(n_msgs)]
    scales = [torch.tensor([0.7 + 0.2 * k]).cuda() for k in range(n_msgs)]
    x = rnd(rows, ch) if residual else None
    tensors = [rnd(ch, ch) * 0.2, rnd(ch) * 0.1, rnd(ch) * 0.3, rnd(1), 1 + 0.1 * rnd(ch), 0.1 * rnd(ch)]
    g_out = rnd(rows, ch)
    pad = -(-rows // 128) * 128

    def run():
        saved = ([torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)],
                 [torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)], torch.zeros(3, rows, device="cuda"))
        params = _make_params(ch, n_msgs, aggs, ws, scales, x, tensors, 1e-5, True, saved, True, None, max_ctas)
        out = torch.zeros(rows, ch, device="cuda")
        check(lib.topo_sccn_combine_fwd_tc2(C.byref(params), rows, None, ptr(out), stream()))
        res = {"g_agg": [torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)],
               "wprod": [torch.zeros(ch, ch, device="cuda") for _ in range(n_msgs)],
               "g_x": torch.zeros(rows, ch, device="cuda") if residual else None,
               "w1": torch.zeros(ch, ch, device="cuda"), "b1": torch.zeros(ch, device="cuda"),
               "w2": torch.zeros(ch, device="cuda"), "b2": torch.zeros(1, device="cuda"),
               "gamma": torch.zeros(ch, device="cuda"), "beta": torch.zeros(ch, device="cuda")}
        grads = CombineGrads()
        for k in range(n_msgs):
            grads.g_agg[k], grads.g_wprod[k] = ptr(res["g_agg"][k]), ptr(res["wprod"][k])
        grads.g_x = ptr(res["g_x"])
        grads.g_att_w1, grads.g_att_b1, grads.g_att_w2, grads.g_att_b2 = ptr(res["w1"]), ptr(res["b1"]), ptr(res["w2"]), ptr(res["b2"])
        grads.g_ln_gamma, grads.g_ln_beta = ptr(res["gamma"]), ptr(res["beta"])
        check(lib.topo_sccn_combine_bwd_tc(C.byref(params), rows, None, ptr(g_out), C.byref(grads), stream()))
        torch.cuda.synchronize()
        return [out, *saved[0], *saved[1], saved[2], *res["g_agg"]] + ([res["g_x"]] if residual else [])

    first = run()
    assert all(torch.isfinite(t).all() for t in first)
    for it in range(9):
        again = run()
        for i, (a, b) in enumerate(zip(first, again)):
            assert torch.equal(a, b), f"run {it + 1}: tensor {i} differs in {(a != b).sum().item()} elements"
