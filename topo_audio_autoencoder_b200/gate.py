"""Stochastic gates of the complex stage (csrc/gate.cu).

``HardConcrete`` is the gate the reference's README.md:15-18 describes (sample, stretch
gamma/zeta, clamp, STE, learned temperature / stretch / location bias) and that
``trainer.py:266`` expects at ``model.encoder.sampler`` with a writable ``current_temp``.  The
reference ships no code for it, so the formula is this repo's (DESIGN.md "Hard Concrete spec").

``BinaryGumbel`` mirrors the shipped gate, reference encoder.py:26-53.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import lib, check, ptr, stream, i64_array


def _aligned(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The gate kernels move 128 bits per access: a row sliced out of a [B, N] batch with N % 4 != 0 starts at an address
    that is not 16-byte aligned and is copied once."""
    if t is None:
        return None
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


class _HardConcreteFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, u, params, offsets, training, ste):
        logits, u = _aligned(logits), _aligned(u)
        params = params.contiguous()
        z = torch.empty_like(logits)
        batch = logits.numel() // offsets[4]
        off = i64_array(offsets)
        check(lib.topo_hard_concrete_fwd(ptr(logits), ptr(u), ptr(params), off, batch, int(training), int(ste),
                                         ptr(z), stream()))
        ctx.save_for_backward(logits, u, params)
        ctx.offsets, ctx.training, ctx.batch = list(offsets), bool(training), batch
        return z

    @staticmethod
    def backward(ctx, grad_z):
        logits, u, params = ctx.saved_tensors
        grad_z = _aligned(grad_z)
        grad_logits = torch.empty_like(logits)
        grad_params = torch.empty(7, dtype=torch.float32, device=logits.device)
        offsets = i64_array(ctx.offsets)
        # per-CTA sums, added in CTA order: the seven parameter gradients are bit-reproducible
        workspace = torch.empty(int(lib.topo_hard_concrete_bwd_workspace_floats(offsets, ctx.batch)), dtype=torch.float32,
                                device=logits.device)
        check(lib.topo_hard_concrete_bwd(ptr(logits), ptr(u), ptr(params), offsets, ctx.batch, int(ctx.training), ptr(grad_z),
                                         ptr(grad_logits), ptr(grad_params), ptr(workspace), stream()))
        return grad_logits, None, grad_params, None, None, None


def hard_concrete(logits: torch.Tensor, u: Optional[torch.Tensor], params: torch.Tensor, offsets: Sequence[int],
                  training: bool = True, ste: bool = False) -> torch.Tensor:
    """logits [..., N]; u like logits (uniform noise, required when training); params float[7] on
    the device = (beta, gamma, zeta, loc_0..loc_3); offsets = the five rank boundaries of the axis."""
    if logits.shape[-1] != offsets[4]:
        raise ValueError(f"last axis must be {offsets[4]}, got {logits.shape[-1]}")
    if training and u is None:
        raise ValueError("training mode needs the uniform noise tensor u")
    return _HardConcreteFn.apply(logits, u, params, tuple(int(o) for o in offsets), training, ste)


class HardConcrete(nn.Module):
    """Hard Concrete sampler.

    Learned: ``log_temp_scale`` (temperature = current_temp * exp(log_temp_scale)), ``gamma``,
    ``zeta``; the per-rank location bias is passed in by the owner (the reference keeps the four rank
    biases on the encoder, encoder.py:167-170).  ``current_temp`` is the annealed temperature the
    trainer writes (trainer.py:266); ``set_temperature`` mirrors BinaryGumbel's (encoder.py:49-53).
    """

    def __init__(self, offsets: Sequence[int], start_temp: float = 2.0 / 3.0, min_temp: float = 0.01,
                 gamma: float = -0.1, zeta: float = 1.1, ste: bool = False):
        super().__init__()
        self.offsets = [int(o) for o in offsets]
        self.start_temp, self.min_temp = start_temp, min_temp
        # the annealed temperature lives in DEVICE memory (a non-persistent buffer that follows .to()): a CUDA graph that
        # captured the gate then reads the value of the moment of each replay, not of the capture
        self.register_buffer("_temp_buf", torch.tensor([float(start_temp)]), persistent=False)
        self._temp_host = float(start_temp)
        self.ste = ste
        self.log_temp_scale = nn.Parameter(torch.zeros(1))
        self.gamma = nn.Parameter(torch.tensor([float(gamma)]))
        self.zeta = nn.Parameter(torch.tensor([float(zeta)]))

    @property
    def current_temp(self) -> float:
        return self._temp_host

    @current_temp.setter
    def current_temp(self, value: float) -> None:
        """trainer.py:266 assigns this attribute every epoch: the write lands in the device buffer (one fill kernel)."""
        self._temp_host = float(value)
        self._temp_buf.fill_(float(value))

    def set_temperature(self, temp: float) -> None:
        self.current_temp = max(float(temp), self.min_temp)

    def pack_params(self, loc: Optional[torch.Tensor]) -> torch.Tensor:
        beta = self._temp_buf * torch.exp(self.log_temp_scale)
        if loc is None:
            loc = torch.zeros(4, dtype=torch.float32, device=beta.device)
        return torch.cat([beta, self.gamma, self.zeta, loc.reshape(4)]).to(torch.float32)

    def forward(self, logits: torch.Tensor, u: Optional[torch.Tensor] = None, loc: Optional[torch.Tensor] = None):
        if self.training and u is None:
            # open interval (0, 1): log(u) and log(1 - u) stay finite
            u = torch.rand_like(logits).clamp_(1e-6, 1.0 - 1e-6)
        return hard_concrete(logits, u if self.training else None, self.pack_params(loc), self.offsets,
                             training=self.training, ste=self.ste)


class _BinaryGumbelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, gumbels, temp):
        logits, gumbels = _aligned(logits), _aligned(gumbels)
        probs = torch.empty_like(logits)
        check(lib.topo_binary_gumbel_fwd(ptr(logits), ptr(gumbels), float(temp), logits.numel(), ptr(probs), stream()))
        ctx.save_for_backward(logits, gumbels)
        ctx.temp = float(temp)
        return probs

    @staticmethod
    def backward(ctx, grad_probs):
        logits, gumbels = ctx.saved_tensors
        grad = torch.empty_like(logits)
        grad_probs = _aligned(grad_probs)
        check(lib.topo_binary_gumbel_bwd(ptr(logits), ptr(gumbels), ctx.temp, logits.numel(),
                                         ptr(grad_probs), ptr(grad), stream()))
        return grad, None, None


class BinaryGumbel(nn.Module):
    """reference encoder.py:26-53.  Training: softmax(([l, 1-l] + Gumbel) / temp, dim 0)[0], with the
    noise drawn as the reference does (encoder.py:36) unless ``gumbels`` [2, *l.shape] is injected.
    Eval: the reference's branch reduces a 1-D input to a scalar; the evident
    intent, the hard decision softmax([l, 1-l] / temp)[0] > 0.5, is what is implemented."""

    def __init__(self, start_temp: float = 1.0, min_temp: float = 0.01):
        super().__init__()
        self.start_temp, self.min_temp = start_temp, min_temp
        self.current_temp = start_temp

    def forward(self, logits: torch.Tensor, gumbels: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.training:
            if gumbels is None:
                gumbels = -torch.empty((2,) + tuple(logits.shape), device=logits.device,
                                       dtype=logits.dtype).exponential_().log()
            return _BinaryGumbelFn.apply(logits, gumbels, self.current_temp)
        return ((logits - (1 - logits)) > 0).to(logits.dtype)

    def set_temperature(self, temp: float) -> None:       # encoder.py:49-53
        self.current_temp = self.min_temp if temp < self.min_temp else temp
