"""Oracle: weighted incidence / adjacency operators of the active sub-complex.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows complex_builder.py:23-115 (build_sparse_matrices).  The reference's operators are
unsigned and probability-weighted (complex_builder.py:52-54), and their sparsity pattern is
whatever ``torch.nonzero`` finds in the dense fp32 result (complex_builder.py:82-85).
"""
from __future__ import annotations

import torch

from .rectifier_oracle import OracleTables

RANK_NAMES = ("vertices", "edges", "triangles", "tetra")


def dense_operators(probs, tables: OracleTables, active):
    """Dense versions of the seven operators, before sparsification.

    probs  : (v, e, t, tt) rectified probability vectors
    active : dict name -> int64 index vector (ascending)
    Returns (adjacencies[0..3], incidences[1..3]) as dense tensors, or None for an empty
    complex (complex_builder.py:30-32).
    """
    v, e, t, tt = probs
    av, ae, at, aq = (active[k] for k in RANK_NAMES)
    if len(av) == 0:
        return None

    # vertex adjacency: the edge probability, symmetric (complex_builder.py:35-47)
    a0 = torch.zeros(len(v), len(v), dtype=e.dtype)
    i, j = tables.edges[:, 0], tables.edges[:, 1]
    a0 = a0.index_put((i, j), e).index_put((j, i), e)
    a0 = a0[av][:, av]

    # incidences: 0/1 face matrix transposed, columns scaled by the coface probability
    # (complex_builder.py:52-59)
    inc1 = (tables.v2e.T * e.unsqueeze(0))[av][:, ae]
    inc2 = (tables.e2t.T * t.unsqueeze(0))[ae][:, at]
    inc3 = (tables.t2tt.T * tt.unsqueeze(0))[at][:, aq]

    # higher adjacencies as dense products, diagonal removed (complex_builder.py:62-70)
    def off_diag(m):
        return m * (1 - torch.eye(m.shape[0]))

    a1 = off_diag(inc2 @ inc2.T)
    a2 = off_diag(inc3 @ inc3.T)
    a3 = off_diag(inc3.T @ inc3)
    return [a0, a1, a2, a3], [inc1, inc2, inc3]


def to_coo(dense: torch.Tensor) -> torch.Tensor:
    """complex_builder.py:82-85: nonzero -> COO -> coalesce (row-major sorted int64 indices)."""
    idx = torch.nonzero(dense).t()
    return torch.sparse_coo_tensor(idx, dense[idx[0], idx[1]], dense.size()).coalesce()


def build_sparse_matrices(probs, tables: OracleTables, active):
    """complex_builder.py:23-115.  Returns (adjacencies, incidences) dicts keyed like the
    reference ('rank_0'..'rank_3', 'rank_1'..'rank_3'), or None."""
    dense = dense_operators(probs, tables, active)
    if dense is None:
        return None
    adj, inc = dense
    return ({f"rank_{r}": to_coo(m) for r, m in enumerate(adj)},
            {f"rank_{r + 1}": to_coo(m) for r, m in enumerate(inc)})
