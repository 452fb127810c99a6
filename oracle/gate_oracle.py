"""Oracle: the stochastic gates.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* ``binary_gumbel_train`` follows encoder.py:33-41 (BinaryGumbel.forward, training branch) with
  the Gumbel noise passed in instead of drawn (encoder.py:36 draws ``-Exp(1).log()``).
* ``hard_concrete`` -- PARITY UNPINNED.  The reference has no Hard Concrete code; README.md:15-18
  only names the ingredients (sample, stretch gamma/zeta, STE, learned temperature and location
  bias).  The spec below is the builder's, after Louizos, Welling & Kingma 2018, "Learning
  Sparse Neural Networks through L0 Regularization", eqs. (10)-(12).
"""
from __future__ import annotations

import torch

RANK_SLICES_DOC = "rank r owns the slice [offsets[r], offsets[r+1]) of the simplex axis"


def binary_gumbel_train(logits: torch.Tensor, gumbels: torch.Tensor, temp: float) -> torch.Tensor:
    """encoder.py:34-41.  ``gumbels`` has shape [2, *logits.shape]."""
    pair = torch.stack([logits, 1 - logits])
    return torch.softmax((pair + gumbels) / temp, dim=0)[0]


def hard_concrete(logits, u, beta, gamma, zeta, loc, offsets, training=True, ste=False):
    """Builder's Hard Concrete gate.

    logits : [..., N] log-alpha
    u      : [..., N] uniform noise in (0, 1), injected for determinism
    beta   : temperature (python float or 0-d tensor)
    gamma, zeta : stretch limits (gamma < 0 < 1 < zeta)
    loc    : [4] per-rank location bias, already rectified by the caller (the reference applies
             ``relu`` to its rank bias, encoder.py:292)
    offsets: 5 ints, rank boundaries on the simplex axis

    training:  s = sigmoid((log u - log(1-u) + logits + loc_r) / beta)
    eval:      s = sigmoid(logits + loc_r)
    both:      z = clamp(s * (zeta - gamma) + gamma, 0, 1)
    ste:       value (z > 0.5), gradient of z   -- the in-repo STE idiom, encoder.py:354-357
    """
    loc_full = torch.cat([loc[r].expand(offsets[r + 1] - offsets[r]) for r in range(4)])
    x = logits + loc_full
    if training:
        x = (torch.log(u) - torch.log(1 - u) + x) / beta
    s = torch.sigmoid(x)
    z = torch.clamp(s * (zeta - gamma) + gamma, 0.0, 1.0)
    if ste:
        z = z + ((z > 0.5).to(z.dtype) - z).detach()
    return z
