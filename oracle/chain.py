"""Oracle: the whole complex stage for a batch, per clip, as the reference processes data (batch size 1, trainer.py:93).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  gate -> split -> rectifier -> active sets -> embeddings -> dense
operator build -> SCCN x L -> penalties, in fp32 or fp64, with the upstream gradients of a compact batched run mapped
back onto each clip's rows.  Every piece is one of the pinned restatements (rectifier_oracle, complex_builder_oracle,
glue_oracle, gate_oracle, sccn_oracle); the glue between them is DESIGN.md "Glue" (the reference's generate_complex
raises).  Used by tests/test_gpu_bench_parity.py and __graft_entry__.smoke().
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence

import torch

from . import complex_builder_oracle as cbo
from . import gate_oracle as go
from . import glue_oracle as glo
from . import rectifier_oracle as ro
from .sccn_oracle import OracleSCCN

NAMES = cbo.RANK_NAMES


class StageOracle:
    """CPU twin of ``ComplexStage`` built from the state dict of a GPU stage."""

    def __init__(self, state: Dict[str, torch.Tensor], n_vertices: int, channels: int, n_layers: int, gate: str, bias_on: str,
                 min_active: float, max_active: float, temp: float, gamma: float = -0.1, zeta: float = 1.1,
                 dtype=torch.float32):
        self.n, self.C, self.L = n_vertices, channels, n_layers
        self.gate, self.bias_on, self.dtype = gate, bias_on, dtype
        self.min_active, self.max_active, self.temp, self.gamma, self.zeta = min_active, max_active, temp, gamma, zeta
        self.off = glo.rank_offsets(n_vertices)
        tab = ro.make_tables(n_vertices)
        tab.v2e, tab.e2t, tab.t2tt = tab.v2e.to(dtype), tab.e2t.to(dtype), tab.t2tt.to(dtype)
        self.tab = tab
        state = {k: v.detach().cpu() for k, v in state.items()}
        self.sccn = OracleSCCN(channels, 3, n_layers).train()
        self.sccn.load_state_dict({k[len("sccn."):]: v for k, v in state.items() if k.startswith("sccn.")})
        self.sccn = self.sccn.to(dtype)
        self.leaves: Dict[str, torch.Tensor] = {}
        for k, v in state.items():
            if k.startswith("head.") and v.dtype.is_floating_point and not k.endswith("_temp_buf"):
                self.leaves[k] = v.to(dtype).clone().requires_grad_(True)

    def emb_params(self):
        out = []
        for nm in ("vertex_embeddings", "edge_embeddings", "triangle_embeddings", "tetra_embeddings"):
            out.append((self.leaves[f"head.{nm}.0.weight"], self.leaves[f"head.{nm}.1.weight"], self.leaves[f"head.{nm}.1.bias"]))
        return out

    def named_grads(self) -> Dict[str, Optional[torch.Tensor]]:
        out = {f"sccn.{k}": p.grad for k, p in self.sccn.named_parameters()}
        out.update({k: v.grad for k, v in self.leaves.items()})
        return out

    def run(self, logits: torch.Tensor, noise: torch.Tensor, ups: Sequence[torch.Tensor], up_vp: torch.Tensor,
            up_ent: torch.Tensor):
        """logits [B, N]; noise [2, B, N] (binary_gumbel) or [B, N] (hard_concrete); ups[r] [sum_b n_r(b), C] upstream
        gradient of the compact rank-r rows; up_vp / up_ent [B].  Runs forward + backward and returns per-clip records
        and d loss / d logits."""
        dt = self.dtype
        lc = logits.to(dt).clone().requires_grad_(True)
        lv = self.leaves
        bias = [lv["head.vertex_bias"], lv["head.edge_bias"], lv["head.triangle_bias"], lv["head.tetra_bias"]]
        if self.gate == "binary_gumbel":
            z = go.binary_gumbel_train(lc, noise.to(dt), self.temp)
        else:
            loc = torch.relu(torch.cat(bias)) if self.bias_on == "logits" else torch.zeros(4, dtype=dt)
            beta = self.temp * torch.exp(lv["head.sampler.log_temp_scale"])
            z = go.hard_concrete(lc, noise.to(dt), beta, lv["head.sampler.gamma"], lv["head.sampler.zeta"], loc, self.off)
        records, loss = [], 0.0
        cursor = [0, 0, 0, 0]
        for b in range(logits.shape[0]):
            res = glo.complex_from_probs(z[b], self.n, bias[0], self.tab, self.emb_params(), self.bias_on == "probs")
            if res is None:
                raise ValueError(f"clip {b}: empty complex")
            emb, (adj, inc), rect = res
            out = self.sccn({f"rank_{r}": emb[f"rank_{r}"] for r in range(4)}, inc, adj)
            vp = glo.vertex_penalty(rect[0], self.min_active, self.max_active)
            ent = glo.entropy_loss(*rect)
            for r in range(4):
                n_r = out[f"rank_{r}"].shape[0]
                loss = loss + (out[f"rank_{r}"] * ups[r][cursor[r]:cursor[r] + n_r].to(dt)).sum()
                cursor[r] += n_r
            loss = loss + vp * up_vp[b].to(dt) + ent * up_ent[b].to(dt)
            records.append(dict(active=emb["active_indices"], out=[out[f"rank_{r}"].detach() for r in range(4)],
                                rect=torch.cat(rect).detach(), vp=vp.detach(), ent=ent.detach()))
        loss.backward()
        return records, lc.grad, loss.detach()


def stage_oracles(stage, gate: str, bias_on: str):
    """(fp32 twin, fp64 twin) of a ``ComplexStage`` living on the GPU."""
    head = stage.head
    kw = dict(state=stage.state_dict(), n_vertices=head.num_vertices, channels=stage.sccn.channels,
              n_layers=len(stage.sccn.layers), gate=gate, bias_on=bias_on, min_active=head.min_active_vertices,
              max_active=head.max_active_vertices,
              temp=float(head.gumbel.current_temp if gate == "binary_gumbel" else head.sampler.current_temp))
    return StageOracle(dtype=torch.float32, **kw), StageOracle(dtype=torch.float64, **kw)
