"""Shared helpers of the parity tests."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.txt")
NAMES = ("vertices", "edges", "triangles", "tetra")
OPS = [f"adj{r}" for r in range(4)] + [f"inc{r}" for r in (1, 2, 3)]

# north-star tolerance for fp32 activations, losses and gradients
RTOL, ATOL = 1e-5, 1e-6


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sha(a) -> str:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == torch.float32:
        return torch.equal(a.view(torch.int32), b.view(torch.int32))
    return torch.equal(a, b)


def report(tag: str, got: torch.Tensor, want: torch.Tensor) -> str:
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    if got.numel() == 0:
        line = f"{tag:60s} empty"
    else:
        err = (got - want).abs()
        scale = want.abs().max().item()
        rel = (err / (want.abs() + 1e-30)).max().item()
        viol = (err - (ATOL + RTOL * want.abs())).max().item()
        line = (f"{tag:60s} n={got.numel():8d} max|err|={err.max().item():.3e} max|ref|={scale:.3e} "
                f"maxrel={rel:.3e} worst(err-tol)={viol:+.3e}")
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
    return line


def assert_close(tag, got, want, rtol=RTOL, atol=ATOL):
    line = report(tag, got, want)
    assert got.shape == want.shape, f"{tag}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    ok = torch.allclose(got.detach().cpu().float(), want.detach().cpu().float(), rtol=rtol, atol=atol, equal_nan=True)
    assert ok, f"{line} (rtol={rtol}, atol={atol})"


def hard_concrete_like(shape, gen, p_zero=0.2, p_one=0.15):
    x = torch.rand(shape, generator=gen)
    r = torch.rand(shape, generator=gen)
    x = torch.where(r < p_zero, torch.zeros_like(x), x)
    return torch.where(r > 1 - p_one, torch.ones_like(x), x)
