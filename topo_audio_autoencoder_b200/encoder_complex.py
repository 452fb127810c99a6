"""The complex-generation half of the reference's ``AudioEncoder`` and the batched complex stage.

``ComplexHead`` carries exactly the encoder attributes and methods the hot path touches
(reference encoder.py):
    parameters / tables          :86-98, 167-197
    compute_vertex_penalty       :199-203
    compute_entropy_loss         :205-225
    get_active_simplex_embeddings:227-263
    split_simplices              :291-297
    generate_complex             :324-388
so the stock convolutional front-end (encoder.py:104-165, 390-433) can own one and call
``generate_complex(logits)``.  ``ComplexStage`` adds the decoder's SCCN (decoder.py:25-30, 129) and
runs gate -> rectify -> active sets -> embeddings -> 6 SCCN layers for a whole batch with no host
synchronisation; it is what bench.py times.

Glue that is this repo's own (the reference's generate_complex raises):
see DESIGN.md "Glue".
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import lib, check, ptr, ptr_array, stream
from .complex_builder import RANK_KEYS, SparseSimplicialMatrices, build_sparse_matrices
from . import custom_sccn as _sccn
from .custom_sccn import BatchedComplex, GradientSCCN
from .gate import BinaryGumbel, HardConcrete
from .rectifier import ConstraintMatrices, RectifiedProbs, _Tables, rectify_batch


# --------------------------------------------------------------------------------------------
# autograd wrappers
# --------------------------------------------------------------------------------------------
def active_sets(probs: torch.Tensor, tables: _Tables):
    """[B, N] probabilities -> (pos, act_idx, counts, row_off) int32 device arrays; no host sync."""
    probs = probs.detach().contiguous()
    b, dev = probs.shape[0], probs.device
    pos = torch.empty(b, tables.total, dtype=torch.int32, device=dev)
    act = torch.empty(b, tables.total, dtype=torch.int32, device=dev)
    counts = torch.empty(b, 4, dtype=torch.int32, device=dev)
    row_off = torch.empty(4, b + 1, dtype=torch.int32, device=dev)
    check(lib.topo_active_sets(tables.handle, ptr(probs), b, ptr(pos, torch.int32), ptr(act, torch.int32),
                               ptr(counts, torch.int32), ptr(row_off, torch.int32), stream()))
    return pos, act, counts, row_off


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x, gamma, beta = x.contiguous(), gamma.contiguous(), beta.contiguous()
        y = torch.empty_like(x)
        check(lib.topo_layernorm_fwd(x.shape[0], x.shape[1], ptr(x), ptr(gamma), ptr(beta), float(eps), ptr(y), stream()))
        ctx.save_for_backward(x, gamma)
        ctx.eps = float(eps)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, gamma = ctx.saved_tensors
        g_x = torch.empty_like(x)
        g_gamma, g_beta = torch.zeros_like(gamma), torch.zeros_like(gamma)
        g_y = g_y.contiguous()         # keep every buffer a launch reads alive in a named variable
        # per-CTA column sums, added in CTA order: the affine gradients are bit-reproducible
        workspace = torch.empty(int(lib.topo_layernorm_bwd_workspace_floats(x.shape[0], x.shape[1])), dtype=torch.float32, device=x.device)
        check(lib.topo_layernorm_bwd(x.shape[0], x.shape[1], ptr(x), ptr(gamma), ctx.eps, ptr(g_y),
                                     ptr(g_x), ptr(g_gamma), ptr(g_beta), ptr(workspace), stream()))
        return g_x, g_gamma, g_beta, None


class _EmbedFn(torch.autograd.Function):
    """X_r[row] = lne_r[id] * p[id] for the four ranks (encoder.py:242-247)."""

    @staticmethod
    def forward(ctx, cx: BatchedComplex, probs, l0, l1, l2, l3):
        probs = probs.contiguous()
        lnes = [t.contiguous() for t in (l0, l1, l2, l3)]
        ch = lnes[0].shape[1]
        view = cx.view(probs)
        outs = [torch.empty(cx.rows_max[r], ch, dtype=torch.float32, device=probs.device) for r in range(4)]   # live rows are all written
        # the four ranks are independent: largest first on the current stream, the others on side streams
        forked = _sccn._Forked(probs.device, None) if _sccn.CONCURRENT_RANKS else None
        for position, r in enumerate((3, 2, 1, 0)):
            if cx.rows_max[r]:
                st = forked.stream_for(position) if forked is not None else stream()
                check(lib.topo_embed_fwd(cx.tables.handle, C.byref(view), r, ch, ptr(lnes[r]), ptr(outs[r]), st))
        if forked is not None:
            forked.join()
        ctx.save_for_backward(probs, *lnes)
        ctx.cx, ctx.ch = cx, ch
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_xs):
        probs, lnes = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        cx, ch = ctx.cx, ctx.ch
        view = cx.view(probs)
        g_probs = torch.zeros_like(probs)
        g_lnes = [torch.zeros_like(lnes[r]) for r in range(4)]
        keep = [g.contiguous() if g is not None else None for g in g_xs]      # bound to names until the launches are queued
        # rank r only touches its own slice of g_probs: the four launches run side by side
        forked = _sccn._Forked(probs.device, None) if _sccn.CONCURRENT_RANKS else None
        for position, r in enumerate((3, 2, 1, 0)):
            if keep[r] is not None and cx.rows_max[r]:
                st = forked.stream_for(position) if forked is not None else stream()
                check(lib.topo_embed_bwd(cx.tables.handle, C.byref(view), r, ch, ptr(lnes[r]), ptr(keep[r]),
                                         ptr(g_lnes[r]), ptr(g_probs), st))
        if forked is not None:
            forked.join()
        return (None, g_probs, *g_lnes)


class _PenaltiesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, tables: _Tables, min_active, max_active):
        probs = probs.contiguous()
        b = probs.shape[0]
        vp = torch.empty(b, dtype=torch.float32, device=probs.device)
        ent = torch.empty(b, dtype=torch.float32, device=probs.device)
        check(lib.topo_penalties_fwd(tables.handle, ptr(probs), b, float(min_active), float(max_active), ptr(vp), ptr(ent), stream()))
        ctx.save_for_backward(probs)
        ctx.cfg = (tables, float(min_active), float(max_active))
        return vp, ent

    @staticmethod
    def backward(ctx, g_vp, g_ent):
        (probs,) = ctx.saved_tensors
        tables, lo, hi = ctx.cfg
        g = torch.empty_like(probs)
        # both upstream gradients usually arrive as expanded (stride-0) tensors; their contiguous copies
        # must outlive the launch, so they are bound to names (two inline temporaries would alias)
        g_vp = g_vp.contiguous() if g_vp is not None else None
        g_ent = g_ent.contiguous() if g_ent is not None else None
        check(lib.topo_penalties_bwd(tables.handle, ptr(probs), probs.shape[0], lo, hi, ptr(g_vp), ptr(g_ent), ptr(g),
                                     stream()))
        return g, None, None, None


# --------------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------------
class ComplexHead(nn.Module):
    """Complex-generation state and methods of the reference AudioEncoder (see module docstring)."""

    def __init__(self, num_vertices: int, embedding_dim: int = 64, min_active_vertices: int = 8,
                 max_active_vertices: int = 16, gate: str = "hard_concrete", bias_on: str = "logits",
                 start_temp: float = 2.0 / 3.0, ste: bool = False):
        super().__init__()
        if gate not in ("hard_concrete", "binary_gumbel"):
            raise ValueError("gate must be 'hard_concrete' or 'binary_gumbel'")
        if bias_on not in ("logits", "probs"):
            raise ValueError("bias_on must be 'logits' or 'probs'")
        self.num_vertices = num_vertices                                   # encoder.py:86-98
        self.num_edges = math.comb(num_vertices, 2)
        self.num_triangles = math.comb(num_vertices, 3)
        self.num_tetra = math.comb(num_vertices, 4)
        self.total_simplices = self.num_vertices + self.num_edges + self.num_triangles + self.num_tetra
        self.seed = 511990
        self.embedding_dim = embedding_dim
        self.min_active_vertices, self.max_active_vertices = min_active_vertices, max_active_vertices
        self.gate_kind, self.bias_on = gate, bias_on
        self.constraints = ConstraintMatrices.create(num_vertices)
        self.active_simplices = None

        self.vertex_bias = nn.Parameter(torch.ones(1) * 2.0)              # encoder.py:167-170
        self.edge_bias = nn.Parameter(torch.ones(1))
        self.triangle_bias = nn.Parameter(torch.ones(1))
        self.tetra_bias = nn.Parameter(torch.ones(1) * 1.5)

        self.gumbel = BinaryGumbel()                                       # encoder.py:172
        self.sampler = HardConcrete(self.constraints._tables.offsets, start_temp=start_temp, ste=ste)   # trainer.py:266

        sizes = (self.num_vertices, self.num_edges, self.num_triangles, self.num_tetra)
        names = ("vertex_embeddings", "edge_embeddings", "triangle_embeddings", "tetra_embeddings")
        for name, n in zip(names, sizes):                                  # encoder.py:177-195
            setattr(self, name, nn.Sequential(nn.Embedding(max(n, 1), embedding_dim), nn.LayerNorm(embedding_dim)))
        self._embedding_names = names

    def _apply(self, fn, *args, **kwargs):
        """nn.Module.to() / .cuda(): the static tables are device memory owned by libtopo_b200, not parameters or
        buffers, so they are re-created on the device the parameters moved to."""
        out = super()._apply(fn, *args, **kwargs)
        dev = self.vertex_bias.device
        if dev.type == "cuda" and dev != self.constraints._tables.device:
            self.constraints = ConstraintMatrices.create(self.num_vertices, device=dev)
        return out

    # ---- reference methods -------------------------------------------------------------------
    @property
    def _tables(self) -> _Tables:
        return self.constraints._tables

    def split_simplices(self, logits):
        """encoder.py:291-297 (adds relu(vertex_bias) to the vertex slice of whatever it is given)."""
        o = self._tables.offsets
        return (logits[..., o[0]:o[1]] + F.relu(self.vertex_bias), logits[..., o[1]:o[2]],
                logits[..., o[2]:o[3]], logits[..., o[3]:o[4]])

    def _flat(self, parts) -> Tuple[torch.Tensor, bool]:
        batched = parts[0].dim() == 2
        return torch.cat([p if batched else p.unsqueeze(0) for p in parts], dim=1), batched

    def compute_vertex_penalty(self, vertex_probs):
        """encoder.py:199-203.  [n_v] -> scalar, or [B, n_v] -> [B]."""
        batched = vertex_probs.dim() == 2
        v = vertex_probs if batched else vertex_probs.unsqueeze(0)
        pad = torch.zeros(v.shape[0], self.total_simplices - self.num_vertices, dtype=v.dtype, device=v.device)
        vp, _ = _PenaltiesFn.apply(torch.cat([v, pad], dim=1), self._tables, self.min_active_vertices,
                                   self.max_active_vertices)
        return vp if batched else vp[0]

    def compute_entropy_loss(self, vertex_probs, edge_probs, triangle_probs, tetra_probs):
        """encoder.py:205-221 (line 223 raises in the reference and is omitted)."""
        flat, batched = self._flat((vertex_probs, edge_probs, triangle_probs, tetra_probs))
        _, ent = _PenaltiesFn.apply(flat, self._tables, self.min_active_vertices, self.max_active_vertices)
        return ent if batched else ent[0]

    def _normalised_tables(self) -> List[torch.Tensor]:
        out = []
        for name in self._embedding_names:
            emb, ln = getattr(self, name)
            out.append(_LayerNormFn.apply(emb.weight, ln.weight, ln.bias, ln.eps))
        return out

    def batched_complex(self, rectified: torch.Tensor, sync: bool = False) -> BatchedComplex:
        """Active sets of a rectified [B, N] batch.  sync=True fetches the per-sample counts so that
        the compact feature tensors can be allocated exactly and split per sample."""
        t = self._tables
        pos, act, counts, row_off = active_sets(rectified, t)
        b = rectified.shape[0]
        rows_max = [b * c for c in t.counts]
        host_counts = None
        if sync:
            host_counts = counts.cpu()
            # exact totals, but never a zero-size buffer: a rank without a single live row in the whole batch (say no
            # tetrahedron survived the gate) keeps one dead row so that every kernel still receives a valid pointer;
            # the device-side live count of that rank is 0 and no kernel touches the row
            rows_max = [max(int(v), 1) for v in host_counts.sum(dim=0).tolist()]
        return BatchedComplex(tables=t, probs=rectified, pos=pos, act_idx=act, counts=counts, row_off=row_off,
                              rows_max=rows_max, host_counts=host_counts)

    def embed(self, cx: BatchedComplex) -> List[torch.Tensor]:
        return list(_EmbedFn.apply(cx, cx.probs, *self._normalised_tables()))

    def get_active_simplex_embeddings(self, vertices, edges, triangles, tetra, device=None):
        """encoder.py:227-263 for one sample: per-rank [n_r_active, C] rows plus the int64 index sets."""
        flat, _ = self._flat((vertices, edges, triangles, tetra))
        cx = self.batched_complex(flat, sync=True)
        xs = self.embed(cx)
        o, n = self._tables.offsets, [int(v) for v in cx.host_counts[0].tolist()]
        out = {f"rank_{r}": xs[r][:n[r]] for r in range(4)}        # an empty rank is [0, C], as the reference returns it
        out["active_indices"] = {key: cx.act_idx[0, o[r]:o[r] + n[r]].to(torch.int64) for r, key in enumerate(RANK_KEYS)}
        return out

    # ---- gate + rectifier for a batch ---------------------------------------------------------
    def rank_location_bias(self) -> torch.Tensor:
        return F.relu(torch.cat([self.vertex_bias, self.edge_bias, self.triangle_bias, self.tetra_bias]))

    def gate(self, logits: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[.., N] logits -> gate output on the simplex axis (before rectification)."""
        if self.gate_kind == "binary_gumbel":
            z = self.gumbel(logits, noise)
        else:
            loc = self.rank_location_bias() if self.bias_on == "logits" else None
            z = self.sampler(logits, noise, loc)
        if self.bias_on == "probs":      # encoder.py:333 applied literally: the bias lands on probabilities
            o = self._tables.offsets
            z = torch.cat([z[..., :o[1]] + F.relu(self.vertex_bias), z[..., o[1]:]], dim=-1)
        return z

    def rectified_batch(self, logits: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        return rectify_batch(self.gate(logits, noise), self.constraints)

    def generate_complex(self, logits: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """encoder.py:324-388 for one sample: logits [N] -> (embeddings, complex_matrices), or
        (None, None, None) for an empty complex (:365-366).  Side effect: ``self.active_simplices``."""
        rect = self.rectified_batch(logits.reshape(1, -1), None if noise is None else noise.reshape(
            (2, 1, -1) if self.gate_kind == "binary_gumbel" else (1, -1)))
        counts = self._tables.counts
        v, e, t, tt = (p[0] for p in torch.split(rect, counts, dim=1))
        if torch.sum(v) == 0:
            return None, None, None
        probs = RectifiedProbs(vertices=v, edges=e, triangles=t, tetra=tt, all_simplices=rect[0])
        active = self.get_active_simplex_embeddings(v, e, t, tt, logits.device)
        self.active_simplices = active["active_indices"]
        embeddings = {f"rank_{r}": active[f"rank_{r}"] for r in range(4)}
        complex_matrices = build_sparse_matrices(probs, self.constraints, active["active_indices"])
        return embeddings, complex_matrices


class ComplexStage(nn.Module):
    """gate -> rectify -> active sets -> embeddings -> SCCN x n_layers for a batch of logits.

    forward(logits [B, N], noise [B, N]) -> dict with the compact per-rank features ('rank_r'), the
    BatchedComplex ('complex') and the per-sample penalties.  No host synchronisation unless
    ``sync=True`` (needed only to split the compact rows per sample for the decoder tail).
    """

    def __init__(self, num_vertices: int = 20, channels: int = 64, n_layers: int = 6, **head_kwargs):
        super().__init__()
        self.head = ComplexHead(num_vertices, embedding_dim=channels, **head_kwargs)
        self.sccn = GradientSCCN(channels=channels, max_rank=3, n_layers=n_layers, update_func="gelu")   # decoder.py:25-30

    def forward(self, logits: torch.Tensor, noise: Optional[torch.Tensor] = None, sync: bool = False):
        head = self.head
        rect = head.rectified_batch(logits, noise)
        cx = head.batched_complex(rect, sync=sync)
        xs = self.sccn.forward_complex(cx, head.embed(cx))
        vp, ent = _PenaltiesFn.apply(rect, head._tables, head.min_active_vertices, head.max_active_vertices)
        out = {f"rank_{r}": xs[r] for r in range(4)}
        out.update(complex=cx, rectified=rect, vertex_penalty=vp, entropy_loss=ent)
        return out

    @torch.no_grad()
    def calibrate(self, logits: torch.Tensor, noise: Optional[torch.Tensor] = None) -> List[float]:
        """Measure the live fraction of every rank on a representative batch (one host synchronisation) and store it as
        the SM-split estimate of the concurrent rank launches (custom_sccn.ROW_FRACTION_HINT)."""
        rect = self.head.rectified_batch(logits, noise)
        cx = self.head.batched_complex(rect, sync=True)
        totals = cx.host_counts.sum(dim=0).tolist()
        bound = [logits.shape[0] * c for c in self.head._tables.counts]
        _sccn.ROW_FRACTION_HINT[:] = [min(1.0, max(t / b, 1.0 / max(b, 1))) if b else 1.0 for t, b in zip(totals, bound)]
        return list(_sccn.ROW_FRACTION_HINT)

    @staticmethod
    def split_per_sample(cx: BatchedComplex, x: torch.Tensor, rank: int) -> List[torch.Tensor]:
        if cx.host_counts is None:
            raise ValueError("split_per_sample needs forward(..., sync=True)")
        sizes = cx.host_counts[:, rank].tolist()
        return list(torch.split(x[:sum(sizes)], sizes, dim=0))        # rows past the live total are allocation slack
