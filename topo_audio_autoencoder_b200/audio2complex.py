"""Drop-in for the reference's model shell (audio2complex.py:18-72) and encoder (encoder.py:70-197, 390-433), batched.

``AudioEncoder`` = the stock convolutional front-end (frontend.py) + the complex-generation state (``ComplexHead``), with
the reference's attribute names on ONE module, so a reference ``encoder`` state dict loads unchanged and
``model.encoder.sampler.current_temp`` (trainer.py:266) exists.  ``AudioAutoencoder`` = encoder + ``AudioDecoder``.

Deviations, both forced:
  * ``rave.pqmf.PQMF`` is not on this machine: ``forward`` takes the 16 PQMF band signals [B, 16, 4000] directly and returns
    band signals; the analysis / synthesis filter bank stays outside.
  * the reference's ``encoder.forward`` raises before it reaches the decoder; the glue here is DESIGN.md "Glue", run for a
    whole batch: front-end -> gate -> rectifier -> active sets -> embeddings -> SCCN -> decoder tail.  A clip whose
    complex is empty (no active vertex, or nothing above rank 0 to attend to) is reported in ``valid`` instead of
    returning ``None`` for the whole call; the trainer charges it the reference's invalid-state penalty (trainer.py:278-279).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .decoder import AudioDecoder
from .encoder_complex import ComplexHead, _PenaltiesFn
from .frontend import ConvFrontEnd


class AudioEncoder(ComplexHead):
    def __init__(self, num_vertices, num_bands=16, embedding_dim=128, dropout=0.1, min_active_vertices=8,
                 max_active_vertices=16, gate: str = "hard_concrete", bias_on: str = "logits", **head_kwargs):
        super().__init__(num_vertices, embedding_dim=embedding_dim, min_active_vertices=min_active_vertices,
                         max_active_vertices=max_active_vertices, gate=gate, bias_on=bias_on, **head_kwargs)
        fe = ConvFrontEnd(num_vertices, num_bands, dropout)
        self.num_bands = num_bands
        for name in ("band_processors", "skip_maxpool", "cross_band", "temporal_reduction", "to_simplices"):
            setattr(self, name, getattr(fe, name))
        self.skip_weight = fe.skip_weight
        self._front = [fe]                        # not a submodule: its parameters are registered above, once

    def logits(self, x: torch.Tensor) -> torch.Tensor:
        """encoder.py:390-426: band signals [B, bands, T] -> [B, total_simplices]."""
        return self._front[0](x)


class AudioAutoencoder(nn.Module):
    def __init__(self, num_vertices, num_bands=16, sccn_hidden_dim=64, min_active_vertices=8, max_active_vertices=20,
                 gate: str = "hard_concrete", bias_on: str = "logits"):
        super().__init__()
        self.encoder = AudioEncoder(num_vertices=num_vertices, num_bands=num_bands, embedding_dim=sccn_hidden_dim,
                                    min_active_vertices=min_active_vertices, max_active_vertices=max_active_vertices,
                                    gate=gate, bias_on=bias_on)
        self.decoder = AudioDecoder(sccn_hidden_dim=sccn_hidden_dim, initial_sequence_length=250, output_channels=num_bands)
        self.seed = 511990

    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    def forward(self, bands: torch.Tensor, noise: Optional[torch.Tensor] = None
                ) -> Tuple[Optional[torch.Tensor], Dict[str, torch.Tensor], torch.Tensor]:
        """bands [B, 16, T] -> (decoded bands of the valid clips [B_valid, 16, 16 L] or None, {'binary_entropy': [B],
        'diversity': [B]}, valid [B] bool on the host)."""
        enc, dec = self.encoder, self.decoder
        logits = enc.logits(bands)
        rect = enc.rectified_batch(logits, noise)
        all_active = enc.gate_kind == "binary_gumbel" and enc.training      # the shipped gate never emits an exact zero
        cx = enc.batched_complex(rect, sync=not all_active)
        xs = dec.sccn.forward_complex(cx, enc.embed(cx))
        vp, ent = _PenaltiesFn.apply(rect, enc._tables, enc.min_active_vertices, enc.max_active_vertices)
        diversity = {"binary_entropy": ent, "diversity": vp}
        b = bands.shape[0]
        if all_active:
            counts = torch.tensor([enc._tables.counts] * b)
        else:
            counts = cx.host_counts.to(torch.int64)
        valid = (counts[:, 0] > 0) & (counts[:, 1:].sum(dim=1) > 0)
        if not bool(valid.any()):
            return None, diversity, valid
        if not bool(valid.all()):            # drop the rows of invalid clips from the compact buffers
            keep_rows = []
            for r in range(4):
                keep = torch.repeat_interleave(valid, counts[:, r])
                keep_rows.append(xs[r][:keep.numel()][keep.to(xs[r].device)])
            xs, counts = keep_rows, counts[valid]
        return dec.forward_batched(xs, counts), diversity, valid
