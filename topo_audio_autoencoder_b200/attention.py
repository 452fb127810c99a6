"""Cross-attention of the decoder consumer over the stage's compact rows (csrc/attention.cu).

Reference: decoder.py:58-63, 144-162 -- ``nn.MultiheadAttention(embed_dim=64, num_heads=4)`` with one sample's vertices-made
queries and that sample's active edges / triangles / tetrahedra as memory.  ``segment_cross_attention`` is the scaled
dot-product core of that module for a whole batch whose memory rows stay in the concatenated compact layout: no padding to
the longest sample, no key-padding mask, no scatter.  The projections around it (in_proj, out_proj) stay ``F.linear``.
"""
from __future__ import annotations

import torch

from ._lib import check, lib, ptr, stream

HEAD_DIM = 16


class _SegmentCrossAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, seg, heads, max_run_len):
        if not (q.is_cuda and k.is_cuda and v.is_cuda and seg.is_cuda):
            raise RuntimeError("segment_cross_attention runs on CUDA tensors only (csrc/attention.cu)")
        q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        b, q_len, c = q.shape
        if c != heads * HEAD_DIM or k.shape != v.shape or k.shape[1] != c:
            raise ValueError("q [B, L, 16 * heads], k / v [rows, 16 * heads] expected")
        if seg.dtype != torch.int32 or tuple(seg.shape) != (b, 3, 2) or not seg.is_contiguous():
            raise ValueError("seg must be a contiguous int32 [B, 3, 2] tensor of (first row, row count)")
        out = torch.empty_like(q)
        lse2 = torch.empty(b, heads, q_len, dtype=torch.float32, device=q.device)
        check(lib.topo_cross_attention_fwd(ptr(q), ptr(k), ptr(v), seg.data_ptr(), b, q_len, heads, ptr(out), ptr(lse2), stream()))
        ctx.save_for_backward(q, k, v, seg, out, lse2)
        ctx.heads, ctx.max_run_len = heads, int(max_run_len)
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, seg, out, lse2 = ctx.saved_tensors
        d_out = d_out.contiguous().float()
        b, q_len, _ = q.shape
        dq = torch.empty_like(q)
        dk, dv = torch.zeros_like(k), torch.zeros_like(v)
        d_row = torch.empty_like(lse2)
        check(lib.topo_cross_attention_bwd(ptr(q), ptr(k), ptr(v), seg.data_ptr(), ptr(out), ptr(lse2), ptr(d_out), b, q_len,
                                           ctx.heads, ctx.max_run_len, ptr(d_row), ptr(dq), ptr(dk), ptr(dv), stream()))
        return dq, dk, dv, None, None, None


def segment_cross_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, seg: torch.Tensor, heads: int,
                            max_run_len: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(16)) v per (sample, head).  q [B, L, 16 * heads] and k, v [rows, 16 * heads] are already projected;
    ``seg[b, r] = (first row, row count)`` of sample b's r-th run of memory rows; ``max_run_len`` = the longest run."""
    return _SegmentCrossAttention.apply(q, k, v, seg, heads, max_run_len)
