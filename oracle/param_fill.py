"""Deterministic, name-keyed parameter values.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference modules, the oracle restatements and the CUDA-backed modules share state-dict names but
construct their parameters in different orders, so a global seed cannot reproduce one module's
initialisation in another.  ``fill_by_name`` gives every parameter a value that depends only on
(seed, its state-dict name, its shape): a fixture then stores the seed instead of the weights.
"""
from __future__ import annotations

import hashlib

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return torch.Generator().manual_seed(int.from_bytes(h[:7], "little"))


def tensor_by_name(seed: int, name: str, shape, kind: str = "normal", scale: float = 1.0, shift: float = 0.0):
    g = _gen(seed, name)
    if kind == "uniform":
        return (torch.rand(tuple(shape), generator=g) * 2 - 1) * scale + shift
    return torch.randn(tuple(shape), generator=g) * scale + shift


def value_for(seed: int, name: str, p: torch.Tensor) -> torch.Tensor:
    """Weights ~ U(-a, a) with a Xavier-like bound, LayerNorm / scale parameters near 1, biases small."""
    leaf = name.rsplit(".", 1)[-1]
    if p.dim() >= 2:                                   # Linear / Conv / Embedding weights
        fan = sum(p.shape[-2:]) if "embedding" not in name else 2.0
        return tensor_by_name(seed, name, p.shape, "uniform", (6.0 / fan) ** 0.5 if "embedding" not in name else 1.0)
    if leaf == "weight" or "message_scales" in name or "attention_scale" in name:
        return tensor_by_name(seed, name, p.shape, "normal", 0.1, 1.0)     # 1-D weights are normalisation gains
    return tensor_by_name(seed, name, p.shape, "normal", 0.1)


@torch.no_grad()
def fill_by_name(module: torch.nn.Module, seed: int, prefix: str = "") -> torch.nn.Module:
    for name, p in module.named_parameters():
        p.copy_(value_for(seed, prefix + name, p).to(p.dtype))
    return module
