"""Oracle: constraint tables and the geometric-mean face rectifier.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference:
  * tables        -- rectifier.py:24-64  (ConstraintMatrices.create)
  * rectification -- rectifier.py:75-127 (enforce_constraints)

The tables here are built with ``itertools.combinations`` and dictionary look-ups,
independently of the closed-form rank/unrank used by the CUDA library, so the two
can be compared.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import torch


@dataclass
class OracleTables:
    n_vertices: int
    edges: torch.Tensor        # [C(n,2), 2] int64, lexicographic
    triangles: torch.Tensor    # [C(n,3), 3]
    tetra: torch.Tensor        # [C(n,4), 4]
    tri_edges: torch.Tensor    # [C(n,3), 3] edge ids of each triangle's faces
    tet_tris: torch.Tensor     # [C(n,4), 4] triangle ids of each tetrahedron's faces
    v2e: torch.Tensor          # dense 0/1 [n_e, n_v]   (rectifier.py:33-36)
    e2t: torch.Tensor          # dense 0/1 [n_t, n_e]   (rectifier.py:39-45)
    t2tt: torch.Tensor         # dense 0/1 [n_tt, n_t]  (rectifier.py:48-55)

    @property
    def sizes(self):
        return (self.n_vertices, len(self.edges), len(self.triangles), len(self.tetra))


def make_tables(n: int) -> OracleTables:
    """rectifier.py:24-64, with the per-face linear searches replaced by dict look-ups."""
    verts = range(n)
    e = list(itertools.combinations(verts, 2))
    t = list(itertools.combinations(verts, 3))
    q = list(itertools.combinations(verts, 4))
    eid = {c: i for i, c in enumerate(e)}
    tid = {c: i for i, c in enumerate(t)}

    tri_edges = [[eid[(a, b)], eid[(a, c)], eid[(b, c)]] for (a, b, c) in t]
    tet_tris = [[tid[(a, b, c)], tid[(a, b, d)], tid[(a, c, d)], tid[(b, c, d)]] for (a, b, c, d) in q]

    def as_idx(rows, width):
        return torch.tensor(rows, dtype=torch.int64).reshape(-1, width)

    edges, triangles, tetra = as_idx(e, 2), as_idx(t, 3), as_idx(q, 4)
    tri_edges, tet_tris = as_idx(tri_edges, 3), as_idx(tet_tris, 4)

    def onehot_rows(faces, n_cols):
        m = torch.zeros(len(faces), n_cols)
        if len(faces):
            m.scatter_(1, faces, 1.0)
        return m

    return OracleTables(
        n_vertices=n, edges=edges, triangles=triangles, tetra=tetra,
        tri_edges=tri_edges, tet_tris=tet_tris,
        v2e=onehot_rows(edges, n), e2t=onehot_rows(tri_edges, len(e)), t2tt=onehot_rows(tet_tris, len(t)),
    )


def _level(own: torch.Tensor, face_probs_log_sum: torch.Tensor, any_face_zero: torch.Tensor, arity: int):
    """One rectification level: constraint = exp(sum(log(face+eps))/arity), forced to an exact
    zero (with a zero gradient path, ``x - x``) where any face is zero, then minimum with the own
    probability.  rectifier.py:90-97 / 102-108 / 113-119."""
    gm = torch.exp(face_probs_log_sum / arity)
    gm = torch.where(any_face_zero, gm - gm, gm)
    return torch.minimum(own, gm)


def enforce_constraints(v, e, t, tt, tables: OracleTables, eps: float = 1e-10):
    """rectifier.py:75-127.  Returns (vertices, edges, triangles, tetra) rectified.

    Edge level gathers vertex pairs by index (rectifier.py:88-92); triangle and tetra levels use
    the dense 0/1 matmul exactly as the reference does (rectifier.py:101, 104, 112, 115), so the
    summation goes through the same torch op.
    """
    pairs = v[tables.edges]
    e_r = _level(e, torch.log(pairs + eps).sum(dim=1), (pairs == 0).any(dim=1), 2)

    def dense_level(own, below, face_matrix, arity):
        log_sum = torch.matmul(face_matrix, torch.log(below + eps))
        zero = (face_matrix @ (below == 0).to(face_matrix.dtype)).bool()
        return _level(own, log_sum, zero, arity)

    t_r = dense_level(t, e_r, tables.e2t, 3)
    tt_r = dense_level(tt, t_r, tables.t2tt, 4)
    return v, e_r, t_r, tt_r
