"""Drop-in for the reference's ``complex_builder.py``.

    build_sparse_matrices(probs, matrices, active_indices) -> SparseSimplicialMatrices | None
                                     reference complex_builder.py:23-115

Same dict keys, shapes, int64 row-major-sorted COO indices and fp32 values as the reference's
dense-product-then-nonzero construction, built by csrc/operators.cu without any dense matrix.
Values carry autograd history back to the four probability vectors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from ._lib import lib, check, ptr, ptr_array, stream
from .rectifier import ConstraintMatrices, RectifiedProbs, _Tables

RANK_KEYS = ("vertices", "edges", "triangles", "tetra")
# operator order of the C ABI: adjacency rank_0..3, incidence rank_1..3
OP_RANK = (0, 1, 2, 3, 0, 1, 2)


@dataclass
class SparseSimplicialMatrices:          # reference complex_builder.py:9-15
    adjacencies: Dict[str, torch.Tensor]
    incidences: Dict[str, torch.Tensor]


class _OperatorValuesFn(torch.autograd.Function):
    """probs [N] -> the seven value vectors; pattern buffers ride along as non-differentiable."""

    @staticmethod
    def forward(ctx, probs, tables: _Tables, pos, act_idx, counts, max_rows):
        probs = probs.contiguous()
        dev = probs.device
        row_ptr = torch.empty(7, max_rows + 1, dtype=torch.int32, device=dev)
        check(lib.topo_operators_count(tables.handle, ptr(probs), ptr(pos, torch.int32), ptr(act_idx, torch.int32),
                                       ptr(counts, torch.int32), max_rows, ptr(row_ptr, torch.int32), stream()))
        nnz = row_ptr[:, max_rows].tolist()          # the one host sync: sparse tensors need their sizes
        rows = [torch.empty(n, dtype=torch.int64, device=dev) for n in nnz]
        cols = [torch.empty(n, dtype=torch.int64, device=dev) for n in nnz]
        vals = [torch.empty(n, dtype=torch.float32, device=dev) for n in nnz]
        check(lib.topo_operators_fill(tables.handle, ptr(probs), ptr(pos, torch.int32), ptr(act_idx, torch.int32),
                                      ptr(counts, torch.int32), max_rows, ptr(row_ptr, torch.int32),
                                      ptr_array(rows, 7, torch.int64), ptr_array(cols, 7, torch.int64),
                                      ptr_array(vals, 7), stream()))
        ctx.save_for_backward(probs, pos, act_idx, counts, row_ptr)
        ctx.tables, ctx.max_rows = tables, max_rows
        ctx.mark_non_differentiable(*rows, *cols)
        return (*vals, *rows, *cols)

    @staticmethod
    def backward(ctx, *grads):
        probs, pos, act_idx, counts, row_ptr = ctx.saved_tensors
        g_vals = [g.contiguous() if g is not None and g.numel() else None for g in grads[:7]]
        grad_probs = torch.zeros_like(probs)
        check(lib.topo_operators_bwd(ctx.tables.handle, ptr(probs), ptr(pos, torch.int32), ptr(act_idx, torch.int32),
                                     ptr(counts, torch.int32), ctx.max_rows, ptr(row_ptr, torch.int32),
                                     ptr_array(g_vals, 7), ptr(grad_probs), stream()))
        return grad_probs, None, None, None, None, None


def index_sets_to_device_layout(active_indices: Dict[str, torch.Tensor], tables: _Tables, device):
    """User-supplied ascending index lists -> (pos [N], act_idx [N], counts [4]) int32 device arrays
    in the layout topo_active_sets produces."""
    n_total = tables.total
    pos = torch.full((n_total,), -1, dtype=torch.int32, device=device)
    act = torch.full((n_total,), -1, dtype=torch.int32, device=device)
    counts = []
    for r, key in enumerate(RANK_KEYS):
        idx = active_indices[key].to(device=device, dtype=torch.int64).reshape(-1)
        n = idx.numel()
        counts.append(n)
        if n:
            pos[tables.offsets[r] + idx] = torch.arange(n, dtype=torch.int32, device=device)
            act[tables.offsets[r]: tables.offsets[r] + n] = idx.to(torch.int32)
    return pos, act, torch.tensor(counts, dtype=torch.int32, device=device), counts


def build_sparse_matrices(probs: RectifiedProbs, matrices: ConstraintMatrices,
                          active_indices: Dict[str, torch.Tensor]) -> Optional[SparseSimplicialMatrices]:
    """reference complex_builder.py:23-115 (same signature, same None-for-empty convention)."""
    if len(active_indices["vertices"]) == 0:           # complex_builder.py:30-32
        return None
    tables = matrices._tables
    flat = torch.cat([probs.vertices, probs.edges, probs.triangles, probs.tetra])
    pos, act, counts_dev, counts = index_sets_to_device_layout(active_indices, tables, flat.device)
    max_rows = max(max(counts), 1)
    outs = _OperatorValuesFn.apply(flat, tables, pos, act, counts_dev, max_rows)
    vals, rows, cols = outs[:7], outs[7:14], outs[14:21]

    def coo(op: int, n_rows: int, n_cols: int) -> torch.Tensor:
        idx = torch.stack([rows[op], cols[op]])
        return torch.sparse_coo_tensor(idx, vals[op], (n_rows, n_cols), is_coalesced=True)

    adj = {f"rank_{r}": coo(r, counts[r], counts[r]) for r in range(4)}
    inc = {f"rank_{r}": coo(3 + r, counts[r - 1], counts[r]) for r in (1, 2, 3)}
    return SparseSimplicialMatrices(adjacencies=adj, incidences=inc)
