"""Oracle: SCCN message passing with message attention.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED for ``Conv``.

Follows custom_sccn.py:7-162 (GradientSCCNLayer / GradientSCCN).  The base classes
``SCCNLayer`` / ``SCCN`` and the ``Conv`` they instantiate come from TopoModelX
(pyt-team/TopoModelX; path-imported at custom_sccn.py:3-4 from a git-ignored clone, no version
pinned, not on this machine).  What is restated here from the published algorithm:

  Conv(x_source, neighborhood) = neighborhood @ (x_source @ weight)
      weight [in, out], Xavier-uniform with gain 1.414, no bias, no activation
      (SCCNLayer builds its convs with update_func=None)

and from the reference's own call sites: argument order (x_source, neighborhood)
(custom_sccn.py:78-81, 95-98, 113-116), ``.weight`` (custom_sccn.py:46-58), the ModuleDict keys
``rank_r`` with same-rank for r in 0..R, low_to_high for r in 1..R, high_to_low for r in 0..R-1.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class StandInConv(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(channels, channels))
        nn.init.xavier_uniform_(self.weight, gain=1.414)

    def forward(self, x_source, neighborhood):
        xw = torch.mm(x_source, self.weight)
        if neighborhood.is_sparse:          # gradient to the operator's values stays on its pattern
            return torch.sparse.mm(neighborhood, xw)
        return torch.mm(neighborhood, xw)


def _present(d, key):
    return key in d and d[key] is not None


class OracleSCCNLayer(nn.Module):
    """custom_sccn.py:7-138."""

    def __init__(self, channels, max_rank, residual=True, is_final_layer=False):
        super().__init__()
        self.max_rank = max_rank
        self.residual = residual
        self.is_final_layer = is_final_layer
        ranks = range(max_rank + 1)
        self.convs_same_rank = nn.ModuleDict({f"rank_{r}": StandInConv(channels) for r in ranks})
        self.convs_low_to_high = nn.ModuleDict({f"rank_{r}": StandInConv(channels) for r in ranks if r > 0})
        self.convs_high_to_low = nn.ModuleDict({f"rank_{r}": StandInConv(channels) for r in ranks if r < max_rank})
        self.layer_norms = nn.ModuleDict({f"rank_{r}": nn.LayerNorm(channels) for r in ranks})          # :15-18
        self.message_scales = nn.ParameterDict({k: nn.Parameter(torch.ones(1))                           # :21-25
                                                for k in ("same_rank", "low_to_high", "high_to_low")})
        self.message_attention = nn.ModuleDict({                                                         # :28-34
            f"rank_{r}": nn.Sequential(nn.Linear(channels, channels), nn.GELU(), nn.Linear(channels, 1))
            for r in ranks})

    def _messages(self, r, features, incidences, adjacencies):
        """custom_sccn.py:73-120: the list of (message [+ residual]) tensors for rank r."""
        key, x = f"rank_{r}", features[f"rank_{r}"]
        msgs = []

        def push(m, unconditional_residual):
            # same-rank adds the residual whenever self.residual (:82-85); the cross-rank
            # messages also require matching shapes (:99-102, :117-120)
            if self.residual and (unconditional_residual or m.shape == x.shape):
                m = m + x
            msgs.append(m)

        if _present(adjacencies, key):
            push(self.convs_same_rank[key](x, adjacencies[key]) * self.message_scales["same_rank"], True)
        up = f"rank_{r + 1}"
        if r < self.max_rank and _present(features, up) and _present(incidences, up):
            push(self.convs_high_to_low[key](features[up], incidences[up]) * self.message_scales["high_to_low"], False)
        down = f"rank_{r - 1}"
        if r > 0 and _present(features, down) and _present(incidences, key):
            push(self.convs_low_to_high[key](features[down], incidences[key].transpose(1, 0))
                 * self.message_scales["low_to_high"], False)
        return msgs

    def forward(self, features, incidences, adjacencies):
        out = {}
        for r in range(self.max_rank + 1):
            key = f"rank_{r}"
            if not _present(features, key):                       # :69-71
                out[key] = None
                continue
            msgs = self._messages(r, features, incidences, adjacencies)
            if not msgs:                                          # :123-125
                out[key] = features[key]
                continue
            stacked = torch.stack(msgs)                           # :128
            attn = F.softmax(self.message_attention[key](stacked), dim=0)   # :129-130
            y = (stacked * attn).sum(dim=0)                       # :132
            if self.training and not self.is_final_layer:         # :133-134
                y = self.layer_norms[key](y)
            out[key] = y
        return out


class OracleSCCN(nn.Module):
    """custom_sccn.py:140-162.  ``residual`` is accepted and ignored by the reference (layers are
    built with the default residual=True, :147-155); ``update_func`` never reaches the forward."""

    def __init__(self, channels, max_rank, n_layers=2, update_func="sigmoid", residual=False):
        super().__init__()
        self.max_rank = max_rank
        self.layers = nn.ModuleList([
            OracleSCCNLayer(channels, max_rank, is_final_layer=(i == n_layers - 1)) for i in range(n_layers)])

    def forward(self, features, incidences, adjacencies):
        for layer in self.layers:
            features = layer(features, incidences, adjacencies)
        return features
