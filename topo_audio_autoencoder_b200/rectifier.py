"""Drop-in for the reference's ``rectifier.py``: same names, argument order and return types.

    ConstraintMatrices.create(n)     reference rectifier.py:24-64
    enforce_constraints(v, e, t, tt, matrices, eps=1e-10) -> RectifiedProbs
                                     reference rectifier.py:75-127

The arithmetic runs in libtopo_b200 (csrc/tables.cu, csrc/rectifier.cu).  Besides the reference's
per-sample 1-D vectors, every probability argument may carry a leading batch dimension
([B, n_r]); the whole batch is then rectified by the same launches.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import lib, check, ptr, stream


@dataclass
class SimplexIndices:            # reference rectifier.py:7-11
    edges: torch.Tensor          # [C(n,2), 2] int64
    triangles: torch.Tensor      # [C(n,3), 3]
    tetra: torch.Tensor          # [C(n,4), 4]


class _Tables:
    """Owner of one ``topo_tables`` handle (static tables of one vertex count on one device)."""

    def __init__(self, n_vertices: int, upload: bool = True):
        handle = C.c_void_p()
        check(lib.topo_tables_create_ex(int(n_vertices), 1 if upload else 0, C.byref(handle)))
        self.handle = handle
        self.n_vertices = int(n_vertices)
        self.on_device = upload
        self.device = torch.device("cuda", torch.cuda.current_device()) if upload else torch.device("cpu")
        counts, offsets = (C.c_int64 * 4)(), (C.c_int64 * 5)()
        check(lib.topo_tables_sizes(handle, counts, offsets))
        self.counts = [int(c) for c in counts]
        self.offsets = [int(o) for o in offsets]
        self.offsets_c = offsets
        self.total = self.offsets[4]

    def simplex_vertices(self, rank: int) -> torch.Tensor:
        out = torch.empty(self.counts[rank], rank + 1, dtype=torch.int64)
        if out.numel():
            check(lib.topo_tables_simplex_vertices(self.handle, rank, out.data_ptr()))
        return out

    def faces(self, rank: int) -> torch.Tensor:
        out = torch.empty(self.counts[rank], rank + 1, dtype=torch.int32)
        if out.numel():
            check(lib.topo_tables_faces(self.handle, rank, out.data_ptr()))
        return out

    def cofaces(self, rank: int) -> torch.Tensor:
        out = torch.empty(self.counts[rank], max(0, self.n_vertices - 1 - rank), dtype=torch.int32)
        if out.numel():
            check(lib.topo_tables_cofaces(self.handle, rank, out.data_ptr()))
        return out

    def face_matrix(self, rank: int) -> torch.Tensor:
        out = torch.zeros(self.counts[rank], self.counts[rank - 1], dtype=torch.float32, device=self.device)
        if out.numel():
            check(lib.topo_tables_face_matrix(self.handle, rank, ptr(out), stream()))
        return out

    def __del__(self):
        try:
            lib.topo_tables_destroy(self.handle)
        except Exception:
            pass


class ConstraintMatrices:
    """reference rectifier.py:13-64.

    ``vertex_to_edge`` / ``edge_to_triangle`` / ``triangle_to_tetra`` are the reference's dense 0/1
    face matrices; here they are materialised on first access only (the kernels never read them,
    they walk int32 face tables).
    """

    def __init__(self, v2e: Optional[torch.Tensor], e2t: Optional[torch.Tensor], t2tt: Optional[torch.Tensor],
                 indices: SimplexIndices, _tables: Optional[_Tables] = None):
        self._dense = {1: v2e, 2: e2t, 3: t2tt}
        self.indices = indices
        if _tables is None:       # built by hand, as the reference constructor allows
            if v2e is None:
                raise ValueError("ConstraintMatrices needs either dense matrices or ConstraintMatrices.create(n)")
            _tables = _Tables(v2e.shape[1])
            # the kernels walk the canonical face tables of the complete 3-skeleton: a hand-built instance is accepted
            # only if it describes that same complex (anything else would be silently ignored otherwise)
            canon = (None, _tables.simplex_vertices(1), _tables.simplex_vertices(2), _tables.simplex_vertices(3))
            for rank, (dense, idx) in enumerate(((v2e, indices.edges), (e2t, indices.triangles), (t2tt, indices.tetra)), start=1):
                if idx is not None and not torch.equal(idx.reshape(-1, rank + 1).cpu().to(torch.int64), canon[rank]):
                    raise ValueError(f"ConstraintMatrices: rank-{rank} index list is not the complete skeleton in "
                                     "itertools.combinations order; only ConstraintMatrices.create(n)'s complex is supported")
                if dense is not None and not torch.equal(dense.to(_tables.device, torch.float32), _tables.face_matrix(rank)):
                    raise ValueError(f"ConstraintMatrices: rank-{rank} face matrix differs from the canonical 0/1 face "
                                     "matrix; only ConstraintMatrices.create(n)'s complex is supported")
        self._tables = _tables

    @classmethod
    def create(cls, n_vertices: int, device=None) -> "ConstraintMatrices":
        """``device``: the CUDA device the static tables are uploaded to (default: the current one).  The tables do not
        follow ``nn.Module.to()``; owners re-create them when they move (ComplexHead._apply)."""
        if device is not None:
            with torch.cuda.device(device):
                return cls.create(n_vertices)
        tables = _Tables(n_vertices)
        dev = tables.device
        indices = SimplexIndices(edges=tables.simplex_vertices(1).to(dev),
                                 triangles=tables.simplex_vertices(2).to(dev),
                                 tetra=tables.simplex_vertices(3).to(dev))
        return cls(None, None, None, indices, _tables=tables)

    def _face_matrix(self, rank: int) -> torch.Tensor:
        if self._dense[rank] is None:
            self._dense[rank] = self._tables.face_matrix(rank)
        return self._dense[rank]

    @property
    def vertex_to_edge(self) -> torch.Tensor:
        return self._face_matrix(1)

    @property
    def edge_to_triangle(self) -> torch.Tensor:
        return self._face_matrix(2)

    @property
    def triangle_to_tetra(self) -> torch.Tensor:
        return self._face_matrix(3)

    @property
    def n_vertices(self) -> int:
        return self._tables.n_vertices


@dataclass
class RectifiedProbs:            # reference rectifier.py:67-73
    vertices: torch.Tensor
    edges: torch.Tensor
    triangles: torch.Tensor
    tetra: torch.Tensor
    all_simplices: torch.Tensor


class _RectifyFn(torch.autograd.Function):
    """probs [B, N] -> rectified [B, N] (csrc/rectifier.cu)."""

    @staticmethod
    def forward(ctx, probs: torch.Tensor, tables: _Tables, eps: float):
        probs = probs.contiguous()
        out = torch.empty_like(probs)
        check(lib.topo_rectify_fwd(tables.handle, ptr(probs), float(eps), probs.shape[0], ptr(out), stream()))
        ctx.save_for_backward(probs, out)
        ctx.tables, ctx.eps = tables, float(eps)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        probs, out = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_in = torch.empty_like(probs)
        workspace = torch.empty_like(probs)
        check(lib.topo_rectify_bwd(ctx.tables.handle, ptr(probs), ptr(out), ptr(grad_out), ctx.eps, probs.shape[0],
                                   ptr(grad_in), ptr(workspace), stream()))
        return grad_in, None, None


def rectify_batch(probs: torch.Tensor, matrices: ConstraintMatrices, eps: float = 1e-10) -> torch.Tensor:
    """Batched entry: probs [B, N] on the simplex axis -> rectified [B, N]."""
    if probs.dim() != 2 or probs.shape[1] != matrices._tables.total:
        raise ValueError(f"expected [B, {matrices._tables.total}], got {tuple(probs.shape)}")
    return _RectifyFn.apply(probs, matrices._tables, eps)


def enforce_constraints(vertex_probs: torch.Tensor, edge_probs: torch.Tensor, triangle_probs: torch.Tensor,
                        tetra_probs: torch.Tensor, matrices: ConstraintMatrices, eps: float = 1e-10) -> RectifiedProbs:
    """reference rectifier.py:75-127 (same signature).  1-D inputs as in the reference, or [B, n_r]."""
    parts = (vertex_probs, edge_probs, triangle_probs, tetra_probs)
    counts = matrices._tables.counts
    batched = vertex_probs.dim() == 2
    for p, c in zip(parts, counts):
        if p.shape[-1] != c or p.dim() != vertex_probs.dim():
            raise ValueError(f"probability vectors must have sizes {counts} on the last axis")
    flat = torch.cat([p if batched else p.unsqueeze(0) for p in parts], dim=1)
    out = _RectifyFn.apply(flat, matrices._tables, eps)
    v, e, t, tt = torch.split(out, counts, dim=1)
    if not batched:
        v, e, t, tt = v[0], e[0], t[0], tt[0]
        return RectifiedProbs(vertices=v, edges=e, triangles=t, tetra=tt, all_simplices=out[0])
    return RectifiedProbs(vertices=v, edges=e, triangles=t, tetra=tt, all_simplices=out)
