"""Oracle: the encoder-side glue of the complex stage and the structural penalties.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restated from encoder.py (the module itself cannot be imported: it needs ``toponetx``, which
it never uses):
  * split_simplices                 encoder.py:291-297
  * get_active_simplex_embeddings   encoder.py:227-263 (tables: encoder.py:177-195)
  * compute_vertex_penalty          encoder.py:199-203
  * compute_entropy_loss            encoder.py:205-225, minus line 223 which raises on ragged input
  * generate_complex                encoder.py:324-388 -- the reference version is broken
    (line 325 truncates the logits to the vertex slice).  The chain restated
    here is the builder's repair, documented in DESIGN.md "Glue": gate -> split -> rectify ->
    active sets -> embeddings -> operators.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import complex_builder_oracle as cbo
from . import rectifier_oracle as ro


def rank_sizes(n):
    return [math.comb(n, k) for k in (1, 2, 3, 4)]


def rank_offsets(n):
    off = [0]
    for s in rank_sizes(n):
        off.append(off[-1] + s)
    return off


def split_simplices(x, n_vertices, vertex_bias):
    """encoder.py:291-297: four slices; relu(vertex_bias) is added to the vertex slice of
    whatever is passed in (the reference passes probabilities, encoder.py:333)."""
    o = rank_offsets(n_vertices)
    return (x[o[0]:o[1]] + F.relu(vertex_bias), x[o[1]:o[2]], x[o[2]:o[3]], x[o[3]:o[4]])


def active_indices(v, e, t, tt):
    """encoder.py:230-233: ascending int64 positions of the non-zero entries."""
    names = cbo.RANK_NAMES
    return {k: p.nonzero().squeeze(-1) for k, p in zip(names, (v, e, t, tt))}


def active_embeddings(tables_and_norms, probs):
    """encoder.py:242-254: LayerNorm(Embedding(idx)) * p[idx] per rank.

    tables_and_norms: four (embedding_weight [n_r, C], ln_weight [C], ln_bias [C]) triples.
    """
    act = active_indices(*probs)
    out = {}
    for r, ((emb, g, b), p, name) in enumerate(zip(tables_and_norms, probs, cbo.RANK_NAMES)):
        idx = act[name]
        rows = F.layer_norm(emb[idx], (emb.shape[1],), g, b)
        out[f"rank_{r}"] = rows * p[idx].unsqueeze(-1)
    out["active_indices"] = act
    return out


def vertex_penalty(vertex_probs, min_active, max_active):
    """encoder.py:199-203."""
    count = vertex_probs.sum()
    return F.relu(min_active - count) + F.relu(count - max_active)


def entropy_loss(v, e, t, tt):
    """encoder.py:205-221 (line 223, a ragged torch.stack that raises, is omitted)."""
    act = torch.stack([v.mean(), e.mean(), t.mean(), tt.mean()])
    q = act / (act.sum() + 1e-10)
    return -0.1 * (-(q * torch.log(q + 1e-10)).sum())


def complex_from_probs(z, n_vertices, vertex_bias, tables: ro.OracleTables, emb_params, bias_on_probs):
    """Builder's repaired generate_complex, downstream of the gate: z [N] gate output.

    bias_on_probs=True reproduces encoder.py:333 literally (split_simplices applied to
    probabilities, so vertices live in [relu(b), 1 + relu(b)]); False slices without the bias
    (the location bias then belongs to the gate).  Returns (embeddings, (adj, inc), rectified) or
    None for an empty complex (encoder.py:365-366)."""
    if bias_on_probs:
        v, e, t, tt = split_simplices(z, n_vertices, vertex_bias)
    else:
        o = rank_offsets(n_vertices)
        v, e, t, tt = (z[o[i]:o[i + 1]] for i in range(4))
    rect = ro.enforce_constraints(v, e, t, tt, tables)
    if torch.sum(rect[0]) == 0:
        return None
    emb = active_embeddings(emb_params, rect)
    mats = cbo.build_sparse_matrices(rect, tables, emb["active_indices"])
    return emb, mats, rect
