"""GPU: the cta_group::2 debug GEMM (mode 4) against fp64."""
import sys
import torch
from topo_audio_autoencoder_b200 import _lib
sys.path.insert(0, ".")
from topo_audio_autoencoder_b200._lib import lib, ptr, stream  # noqa: E402

for rows in (256, 128, 1000, 20000):
    g = torch.Generator().manual_seed(rows)
    a = (torch.randn(rows, 64, generator=g) * torch.logspace(-2, 2, 64)).cuda()
    w = torch.randn(64, 64, generator=g).cuda()
    out = torch.full((rows, 64), float("nan"), device="cuda")
    rc = _lib.load_debug().topo_debug_gemm_bf16x3(ptr(a), ptr(w), rows, 4, 0, 0, 0, ptr(out), stream())
    torch.cuda.synchronize()
    want = a.double() @ w.double().t()
    cond = a.double().abs() @ w.double().abs().t()
    err = ((out.double() - want).abs() / cond)
    print(f"rows {rows}: rc={rc} finite={bool(torch.isfinite(out).all())} max err/cond={err.max().item():.3e}", flush=True)
    if not torch.isfinite(out).all() or err.max().item() > 1e-6:
        bad = (~torch.isfinite(out)) | (err > 1e-6)
        print("   bad rows (first 10):", bad.any(dim=1).nonzero().flatten()[:10].tolist(), " bad cols of first bad row:",
              bad[bad.any(dim=1).nonzero().flatten()[0]].nonzero().flatten().tolist()[:16])
