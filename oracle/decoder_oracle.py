"""Oracle: the decoder tail that CONSUMES the SCCN output (reference decoder.py:19-175 minus the SCCN itself).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned: oracle/make_golden_glue.py runs the reference's
own ``AudioDecoder.forward`` (decoder.py:120-175) on a planted SCCN output and asserts this restatement
reproduces it (output bit for bit, gradients to 1e-6).

  constructor   decoder.py:31-108 (same attribute names -> same state-dict keys)
  forward       decoder.py:131-175: x0.1 scaling (:132, :149), vertex->query MLP (:133), temporal conv over the
                vertex axis (:136-137), linear interpolation to the initial sequence length (:140), rank 1-3
                rows concatenated as attention memory (:145-153; absent / None ranks are skipped), shared
                pre-attention LayerNorm (:156-157), bottleneck key / value projections (:158-159), 4-head
                cross-attention x attention_scale (:162-163), GELU residual + post-norm (:166-167), four
                upsampling blocks (:173-174).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Scale(nn.Module):
    def __init__(self, s):
        super().__init__()
        self.s = s

    def forward(self, x):
        return x * self.s


class OracleDecoderTail(nn.Module):
    def __init__(self, hidden=64, initial_sequence_length=250, output_channels=16):
        super().__init__()
        self.initial_sequence_length = initial_sequence_length
        self.vertex_to_query = nn.Sequential(nn.Linear(hidden, hidden * 2), nn.LayerNorm(hidden * 2), nn.GELU(),
                                             nn.Linear(hidden * 2, hidden), nn.LayerNorm(hidden), nn.GELU())
        self.temporal_conv = nn.Sequential(nn.Conv1d(hidden, hidden, 3, padding=1, groups=8), nn.GroupNorm(8, hidden), nn.GELU(),
                                           nn.Conv1d(hidden, hidden, 3, padding=1, groups=8), nn.GroupNorm(8, hidden), nn.GELU())
        self.pre_attention_norm = nn.LayerNorm(hidden)
        self.post_attention_norm = nn.LayerNorm(hidden)
        self.cross_attention = nn.MultiheadAttention(embed_dim=hidden, num_heads=4, batch_first=True, dropout=0.0)
        self.attention_scale = nn.Parameter(torch.ones(1) * 0.5)
        mid = hidden // 2
        self.key_proj = nn.Sequential(nn.Linear(hidden, mid), nn.LayerNorm(mid), nn.GELU(), nn.Linear(mid, hidden), nn.LayerNorm(hidden))
        self.value_proj = nn.Sequential(nn.Linear(hidden, mid), nn.LayerNorm(mid), nn.GELU(), nn.Linear(mid, hidden), nn.LayerNorm(hidden))
        chans = [hidden, hidden // 2, hidden // 4, output_channels]
        self.upsample_blocks = nn.ModuleList()
        for i in range(4):
            cin, cout = chans[i], chans[min(i + 1, 3)]
            self.upsample_blocks.append(nn.Sequential(
                nn.Upsample(scale_factor=2, mode="linear", align_corners=False),
                nn.Conv1d(cin, cin, 3, padding=1, groups=cin), nn.Conv1d(cin, cout, 1),
                nn.GroupNorm(min(8, cout), cout), nn.GELU(), _Scale(1.0 / (2 ** (i + 1)))))

    def attention_inputs(self, output):
        """decoder.py:131-159 -> (query [1, L, C], keys [1, M, C], values [1, M, C])."""
        v = self.vertex_to_query(output["rank_0"] * 0.1)
        q = self.temporal_conv(v.transpose(0, 1).unsqueeze(0))
        q = F.interpolate(q, size=self.initial_sequence_length, mode="linear", align_corners=False).transpose(1, 2)
        mem = [output[f"rank_{r}"] * 0.1 for r in range(1, 4) if output.get(f"rank_{r}") is not None]
        mem = self.pre_attention_norm(torch.cat(mem, dim=0).unsqueeze(0))
        q = self.pre_attention_norm(q)
        return q, self.key_proj(mem), self.value_proj(mem)

    def attend(self, output):
        """decoder.py:131-167 -> [1, L, C] (the consumer of the hot path's compact rows)."""
        q, k, v = self.attention_inputs(output)
        a, _ = self.cross_attention(query=q, key=k, value=v)
        return self.post_attention_norm(q + F.gelu(a * self.attention_scale))

    def forward(self, output):
        x = self.attend(output).transpose(1, 2)
        for blk in self.upsample_blocks:
            x = blk(x)
        return x
