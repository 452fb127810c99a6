"""Drop-in for the reference's ``decoder.py``: the consumer of the complex stage's output.

Reference: decoder.py:19-175.  ``AudioDecoder`` keeps the constructor, attribute names (state-dict compatible) and the
``forward(feature_embeddings, complex_matrices, desired_length)`` signature; its ``sccn`` is this package's
``GradientSCCN`` (csrc kernels).  The tail itself -- vertex->query MLP, temporal conv, cross-attention, upsampling --
stays stock PyTorch, as the reference has it and the north star asks.

What is this repo's own is the BATCHED consumer, ``DecoderTail.forward_batched``: it reads the compact per-rank rows the
stage emits for a whole batch (inactive simplices are absent, samples concatenated, ``row_off`` marks the sample borders)
and reproduces decoder.py:131-167 per sample without a Python loop over samples:
  * row-wise layers (x0.1, vertex_to_query, pre-attention LayerNorm, key / value projections) run once over all rows;
  * the temporal conv + interpolation along the vertex axis depend on a sample's active-vertex count, so samples are
    grouped by that count (at most n_vertices + 1 groups; one group when every vertex is active);
  * the rank 1-3 rows of a sample are its attention memory WHERE THEY LIE: the scaled dot-product core of the
    cross-attention runs in csrc/attention.cu over per-sample (first row, row count) runs of the compact layout -- no
    padding to the longest sample, no key-padding mask, no scatter (on CPU tensors, for host-side tests, the stock module
    runs on a padded copy instead).
An inactive row must never reach the memory: the compaction contract is checked by tests/test_decoder_tail.py (CPU) and
tests/test_gpu_decoder.py (the CUDA path).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .custom_sccn import GradientSCCN


class ScaleLayer(nn.Module):
    """reference decoder.py:177-183."""

    def __init__(self, scale_factor):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x):
        return x * self.scale_factor


def interpolate_linear(x: torch.Tensor, size=None, scale_factor=None) -> torch.Tensor:
    """``F.interpolate(x, mode="linear", align_corners=False)`` of a [B, C, T] signal, evaluated by the bilinear 2-D operator on
    ONE image whose rows are the B * C signals (height scale 1: the second row's weight is exactly 0): the same weights and the
    same arithmetic -- bit-identical on the CPU, forward and backward -- but PyTorch's CUDA kernels give one thread per output
    PIXEL the loop over batch and channels, so the 1-D operator runs 4,000 threads for a 16 MB tensor (0.5 ms on a B200, 22 ms
    per training step over the decoder's five interpolations and their backwards) where this form runs one thread per element."""
    b, c, t = x.shape
    img = x.reshape(1, 1, b * c, t)
    kw = dict(size=(b * c, size)) if size is not None else dict(scale_factor=(1.0, float(scale_factor)))
    return F.interpolate(img, mode="bilinear", align_corners=False, **kw).reshape(b, c, -1)


class RowLayerNorm(nn.LayerNorm):
    """``nn.LayerNorm`` over the last axis (same parameters, same state-dict keys).  CUDA fp32 inputs with 32, 64 or 128 channels go
    through ``topo_layernorm_fwd/bwd`` (csrc/combine.cu): the consumer normalises the 395,200 memory rows of a 64-clip batch
    five times per step, and PyTorch's kernel moves such short rows at a tenth of the HBM bandwidth."""

    def forward(self, x):
        c = x.shape[-1]
        if (x.is_cuda and x.dtype == torch.float32 and self.elementwise_affine and self.bias is not None
                and tuple(self.normalized_shape) == (c,) and c in (32, 64, 128) and x.numel() > 0):
            from .encoder_complex import _LayerNormFn
            return _LayerNormFn.apply(x.reshape(-1, c), self.weight, self.bias, self.eps).view(x.shape)
        return super().forward(x)


class LinearUpsample(nn.Upsample):
    """``nn.Upsample(scale_factor=k, mode="linear", align_corners=False)`` (decoder.py:93) through ``interpolate_linear``."""

    def forward(self, x):
        return interpolate_linear(x, scale_factor=self.scale_factor)


class DecoderTail(nn.Module):
    """Everything of reference ``AudioDecoder`` except ``self.sccn`` (decoder.py:31-108, 131-175)."""

    def __init__(self, sccn_hidden_dim: int = 64, initial_sequence_length: int = 250, output_channels: int = 16):
        super().__init__()
        h = sccn_hidden_dim
        self.hidden = h
        self.initial_sequence_length = initial_sequence_length
        self.vertex_to_query = nn.Sequential(nn.Linear(h, h * 2), RowLayerNorm(h * 2), nn.GELU(),              # :34-41
                                             nn.Linear(h * 2, h), RowLayerNorm(h), nn.GELU())
        self.temporal_conv = nn.Sequential(nn.Conv1d(h, h, kernel_size=3, padding=1, groups=8), nn.GroupNorm(8, h), nn.GELU(),   # :44-51
                                           nn.Conv1d(h, h, kernel_size=3, padding=1, groups=8), nn.GroupNorm(8, h), nn.GELU())
        self.pre_attention_norm = RowLayerNorm(h)                                                              # :54-55
        self.post_attention_norm = RowLayerNorm(h)
        self.cross_attention = nn.MultiheadAttention(embed_dim=h, num_heads=4, batch_first=True, dropout=0.0)  # :58-63
        self.attention_scale = nn.Parameter(torch.ones(1) * 0.5)                                               # :66
        mid = h // 2
        self.key_proj = nn.Sequential(nn.Linear(h, mid), RowLayerNorm(mid), nn.GELU(), nn.Linear(mid, h), RowLayerNorm(h))   # :70-83
        self.value_proj = nn.Sequential(nn.Linear(h, mid), RowLayerNorm(mid), nn.GELU(), nn.Linear(mid, h), RowLayerNorm(h))
        channels = [h, h // 2, h // 4, output_channels]                                                        # :86-105
        self.upsample_blocks = nn.ModuleList()
        for i in range(4):
            cin, cout = channels[i], channels[min(i + 1, len(channels) - 1)]
            self.upsample_blocks.append(nn.Sequential(
                LinearUpsample(scale_factor=2, mode="linear", align_corners=False),
                nn.Conv1d(cin, cin, kernel_size=3, padding=1, groups=cin), nn.Conv1d(cin, cout, kernel_size=1),
                nn.GroupNorm(min(8, cout), cout), nn.GELU(), ScaleLayer(1.0 / (2 ** (i + 1)))))
        self.apply(self._init_weights)                                                                         # :108

    @staticmethod
    def _init_weights(m):                                                                                      # :110-118
        if isinstance(m, (nn.Linear, nn.Conv1d)):
            nn.init.kaiming_normal_(m.weight, mode="fan_in", nonlinearity="linear")
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    # ---- one sample, the reference's own order of operations (decoder.py:131-175) ----------------------------
    def attend(self, output: Dict[str, Optional[torch.Tensor]]) -> torch.Tensor:
        v = self.vertex_to_query(output["rank_0"] * 0.1)                                                       # :132-133
        q = self.temporal_conv(v.transpose(0, 1).unsqueeze(0))                                                 # :136-137
        q = interpolate_linear(q, size=self.initial_sequence_length).transpose(1, 2)                          # :140-141
        mem = [output[f"rank_{r}"] * 0.1 for r in range(1, 4) if output.get(f"rank_{r}") is not None]          # :144-150
        mem = self.pre_attention_norm(torch.cat(mem, dim=0).unsqueeze(0))                                      # :153-156
        q = self.pre_attention_norm(q)
        a, _ = self.cross_attention(query=q, key=self.key_proj(mem), value=self.value_proj(mem))               # :158-162
        return self.post_attention_norm(q + F.gelu(a * self.attention_scale))                                  # :163-167

    def upsample(self, x: torch.Tensor) -> torch.Tensor:
        x = x.transpose(1, 2)                                                                                  # :170
        for block in self.upsample_blocks:
            x = block(x)
        return x

    def forward(self, output: Dict[str, Optional[torch.Tensor]]) -> torch.Tensor:
        return self.upsample(self.attend(output))

    # ---- a whole batch of compact rows -----------------------------------------------------------------------
    def attend_batched(self, xs: Sequence[torch.Tensor], counts: torch.Tensor) -> torch.Tensor:
        """xs[r]: [sum_b counts[b, r], C] compact rows of rank r, samples concatenated (rows past the total are
        ignored); counts: [B, 4] HOST tensor of active rows per sample and rank.  -> [B, L, C]."""
        counts = counts.to("cpu", torch.int64)
        b, dev, h = counts.shape[0], xs[0].device, self.hidden
        tot = counts.sum(dim=0).tolist()
        if int(counts[:, 0].min()) == 0:
            raise ValueError("a sample without active vertices has no decoder input (the reference returns None, "
                             "audio2complex.py:47-48); filter empty complexes before the decoder")
        mem_len = counts[:, 1:].sum(dim=1)
        if int(mem_len.min()) == 0:
            raise ValueError("a sample without any active edge, triangle or tetrahedron has no attention memory "
                             "(decoder.py:153 would concatenate an empty list)")
        # queries: row-wise MLP over all vertex rows, then per group of equal vertex count
        v = self.vertex_to_query(xs[0][:tot[0]] * 0.1)
        starts0 = torch.cumsum(counts[:, 0], 0) - counts[:, 0]
        q = torch.empty(b, self.initial_sequence_length, h, dtype=v.dtype, device=dev)
        for n0 in torch.unique(counts[:, 0]).tolist():
            members = torch.nonzero(counts[:, 0] == n0).squeeze(1)
            idx = (starts0[members].unsqueeze(1) + torch.arange(n0).unsqueeze(0)).to(dev)                    # [g, n0]
            grp = v[idx].transpose(1, 2)                                                                       # [g, C, n0]
            grp = interpolate_linear(self.temporal_conv(grp), size=self.initial_sequence_length)
            q[members.to(dev)] = grp.transpose(1, 2)
        q = self.pre_attention_norm(q)
        # memory: the rank 1..3 rows of every sample, rank after rank, exactly as the stage emits them
        rows = torch.cat([xs[r][:tot[r]] for r in (1, 2, 3)], dim=0) * 0.1
        rows = self.pre_attention_norm(rows)
        keys, values = self.key_proj(rows), self.value_proj(rows)
        base = [0, tot[1], tot[1] + tot[2]]
        starts = [torch.cumsum(counts[:, r], 0) - counts[:, r] for r in (1, 2, 3)]
        if keys.is_cuda:
            a = self._attend_compact(q, keys, values, counts, base, starts)
        else:
            a = self._attend_padded(q, keys, values, counts, base, starts, mem_len)
        return self.post_attention_norm(q + F.gelu(a * self.attention_scale))

    def _attend_compact(self, q, keys, values, counts, base, starts):
        """CUDA: nn.MultiheadAttention's arithmetic (F.multi_head_attention_forward: packed in-projection, scaled dot-product
        per head, out-projection) with the scaled dot-product core on the compact rows (csrc/attention.cu) -- the memory is
        neither padded nor scattered, each sample attends over its own three runs of rows."""
        from .attention import segment_cross_attention
        mha, h = self.cross_attention, self.hidden
        w, bias = mha.in_proj_weight, mha.in_proj_bias
        seg = torch.stack([torch.stack([base[j] + starts[j], counts[:, r]], dim=1) for j, r in enumerate((1, 2, 3))], dim=1)
        seg = seg.to(torch.int32).contiguous().to(q.device, non_blocking=True)
        qp = F.linear(q, w[:h], bias[:h])
        kp = F.linear(keys, w[h:2 * h], bias[h:2 * h])
        vp = F.linear(values, w[2 * h:], bias[2 * h:])
        a = segment_cross_attention(qp, kp, vp, seg, mha.num_heads, int(counts[:, 1:].max()))
        return F.linear(a, mha.out_proj.weight, mha.out_proj.bias)

    def _attend_padded(self, q, keys, values, counts, base, starts, mem_len):
        """CPU tensors (host-side tests): the stock module on a memory padded to the longest sample, with a key-padding mask."""
        b, dev, h = counts.shape[0], q.device, self.hidden
        m_max = int(mem_len.max())
        sample_of, slot_of, src = [], [], []
        offset_in_sample = torch.zeros(b, dtype=torch.int64)
        for j, r in enumerate((1, 2, 3)):
            c = counts[:, r]
            sid = torch.repeat_interleave(torch.arange(b), c)
            within = torch.arange(int(c.sum())) - torch.repeat_interleave(starts[j], c)
            sample_of.append(sid)
            slot_of.append(offset_in_sample[sid] + within)
            src.append(base[j] + torch.arange(int(c.sum())))
            offset_in_sample = offset_in_sample + c
        sample_of, slot_of, src = (torch.cat(t).to(dev) for t in (sample_of, slot_of, src))
        k_pad = keys.new_zeros(b, m_max, h)
        v_pad = values.new_zeros(b, m_max, h)
        k_pad[sample_of, slot_of] = keys[src]
        v_pad[sample_of, slot_of] = values[src]
        pad_mask = (torch.arange(m_max).unsqueeze(0) >= mem_len.unsqueeze(1)).to(dev)
        a, _ = self.cross_attention(query=q, key=k_pad, value=v_pad, key_padding_mask=pad_mask, need_weights=False)
        return a

    def forward_batched(self, xs: Sequence[torch.Tensor], counts: torch.Tensor) -> torch.Tensor:
        """-> [B, output_channels, 16 L] band signals (decoder.py:170-175)."""
        return self.upsample(self.attend_batched(xs, counts))


class AudioDecoder(DecoderTail):
    """reference decoder.py:19-175, same signature.  ``forward`` takes one sample's embeddings and sparse operators."""

    def __init__(self, sccn_hidden_dim: int = 64, initial_sequence_length: int = 250, output_channels: int = 16):
        super().__init__(sccn_hidden_dim, initial_sequence_length, output_channels)
        self.sccn = GradientSCCN(channels=sccn_hidden_dim, max_rank=3, n_layers=6, update_func="gelu")      # :25-30
        self.sccn.apply(self._init_weights)          # the reference's self.apply (:108) also reaches the SCCN's Linear layers

    def forward(self, feature_embeddings, complex_matrices, desired_length=None):
        output = self.sccn(feature_embeddings, complex_matrices.incidences, complex_matrices.adjacencies)    # :129
        return DecoderTail.forward(self, output)
