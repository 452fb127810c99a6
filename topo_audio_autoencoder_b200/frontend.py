"""The multi-band convolutional front-end of the reference encoder -- STOCK PyTorch, out of the hot path's scope.

Reference: encoder.py:104-165 (modules) and :390-426 (forward up to the logits).  The north star keeps this part stock;
it is here only so that the full training step (BASELINE.json config 4) can run: band signals [B, 16, 4000] -> logits
[B, total_simplices], which is what the complex stage consumes.  Same module names and shapes, so a reference state dict
loads unchanged.

One deviation in HOW, not WHAT: the reference runs its 16 band stacks in a Python loop (encoder.py:396-401).  The 16
stacks are independent and identically shaped, so the forward here evaluates them as ONE grouped convolution per layer
(groups = 16) with the per-band weights concatenated on the fly -- 9 launches instead of 144.  GroupNorm(2, 8) per band
is GroupNorm(32, 128) over the band-major concatenation, and so on.  tests/test_frontend.py checks the result against
the reference's own modules.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class ConvFrontEnd(nn.Module):
    def __init__(self, num_vertices: int, num_bands: int = 16, dropout: float = 0.1):
        super().__init__()
        self.num_bands = num_bands
        self.total_simplices = sum(math.comb(num_vertices, k) for k in (1, 2, 3, 4))
        self.band_processors = nn.ModuleList([                                                     # encoder.py:104-120
            nn.Sequential(
                nn.Conv1d(1, 8, kernel_size=15, stride=2, padding=7), nn.GroupNorm(2, 8), nn.GELU(),
                nn.Conv1d(8, 16, kernel_size=7, stride=2, padding=3), nn.GroupNorm(4, 16), nn.GELU(),
                nn.Conv1d(16, 16, kernel_size=5, stride=2, padding=2), nn.GroupNorm(4, 16), nn.GELU(),
            ) for _ in range(num_bands)])
        self.skip_maxpool = nn.MaxPool1d(kernel_size=2, stride=2)                                  # :123-124
        self.skip_weight = nn.Parameter(torch.tensor(0.1))
        self.cross_band = nn.Sequential(                                                           # :127-136
            nn.Conv1d(num_bands * 16, 192, kernel_size=5, padding=2, groups=4), nn.GroupNorm(12, 192), nn.GELU(),
            nn.Conv1d(192, 128, kernel_size=7, padding=3), nn.GroupNorm(8, 128), nn.GELU())
        self.temporal_reduction = nn.Sequential(                                                   # :139-150
            nn.Conv1d(128, 128, kernel_size=7, stride=4, padding=3, groups=8), nn.GroupNorm(8, 128), nn.GELU(),
            nn.Conv1d(128, 128, kernel_size=7, stride=2, padding=3, groups=8), nn.GroupNorm(8, 128), nn.GELU(),
            nn.Conv1d(128, 128, kernel_size=3, stride=2, padding=1), nn.GroupNorm(8, 128), nn.GELU())
        self.to_simplices = nn.Sequential(                                                         # :153-165
            nn.Linear(4096, 2048), nn.LayerNorm(2048), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(2048, 1024), nn.LayerNorm(1024), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(1024, self.total_simplices))

    def _bands(self, x: torch.Tensor) -> torch.Tensor:
        """encoder.py:396-404 for all bands at once: [B, bands, T] -> [B, bands * 16, T / 8]."""
        nb = self.num_bands
        for conv_i, norm_i in ((0, 1), (3, 4), (6, 7)):
            convs = [bp[conv_i] for bp in self.band_processors]
            norms = [bp[norm_i] for bp in self.band_processors]
            c0 = convs[0]
            w = torch.cat([c.weight for c in convs], dim=0)
            b = torch.cat([c.bias for c in convs], dim=0)
            x = F.conv1d(x, w, b, stride=c0.stride, padding=c0.padding, groups=nb)
            g = torch.cat([n.weight for n in norms], dim=0)
            h = torch.cat([n.bias for n in norms], dim=0)
            x = F.gelu(F.group_norm(x, norms[0].num_groups * nb, g, h, norms[0].eps))
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._bands(x)
        skip = self.skip_maxpool(x.transpose(1, 2)).transpose(1, 2)                                # :407
        x = self.cross_band(x) + self.skip_weight * skip                                           # :411-415
        x = self.temporal_reduction(x)                                                             # :419
        return self.to_simplices(x.flatten(1))                                                     # :423-426 (batch kept)
