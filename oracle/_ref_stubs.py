"""Stub modules that let the UNMODIFIED reference files import in the authoring container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by oracle/make_golden_glue.py, nowhere else.

The reference imports three third-party packages that are not on this machine (no network, no
pinned versions anywhere in the reference):

  * ``toponetx.classes.SimplicialComplex``   encoder.py:5 -- imported, never used.  Empty class.
  * ``rave.core`` / ``rave.pqmf``            precompute_distances.py:9, loss.py:1, audio2complex.py:6.
    ``AudioDistanceV1(multiscale_stft_factory, log_epsilon)`` keeps ``self.multiscale_stft`` and
    ``self.log_epsilon`` -- the two attributes the reference's own subclass reads
    (precompute_distances.py:36-41).  ``MultiScaleSTFT`` is the builder's restatement of acids-rave
    (Hann window, hop = scale // 4, centred, reflect padding, magnitude) -- STFT parity stays UNPINNED;
    what the stub pins is everything the reference does AFTER the transform.
  * ``TopoModelX.topomodelx.nn.simplicial.{sccn,sccn_layer}``   custom_sccn.py:3-4, decoder.py:4.
    ``SCCNLayer(channels, max_rank, aggr_func, update_func)`` builds the three ``ModuleDict``s of
    ``Conv`` that the reference's forward indexes (custom_sccn.py:46-58, 78-116); ``SCCN`` builds
    ``.layers`` (overwritten at custom_sccn.py:147).  ``Conv`` is the stand-in
    ``neighborhood @ (x_source @ weight)`` -- Conv arithmetic stays UNPINNED; what the stub pins is
    the reference's own forward body (message order, residual rules, scales, attention, LayerNorm rule).
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn


def _module(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class Conv(nn.Module):
    """Stand-in for TopoModelX ``Conv`` as SCCNLayer instantiates it: no bias, no activation."""

    def __init__(self, in_channels, out_channels, update_func=None, **_):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        nn.init.xavier_uniform_(self.weight, gain=1.414)

    def forward(self, x_source, neighborhood):
        xw = torch.mm(x_source, self.weight)
        if neighborhood.is_sparse:
            return torch.sparse.mm(neighborhood, xw)
        return torch.mm(neighborhood, xw)


class SCCNLayer(nn.Module):
    def __init__(self, channels, max_rank, aggr_func="sum", update_func="sigmoid"):
        super().__init__()
        self.channels, self.max_rank = channels, max_rank
        self.aggr_func, self.update_func = aggr_func, update_func
        ranks = range(max_rank + 1)
        self.convs_same_rank = nn.ModuleDict({f"rank_{r}": Conv(channels, channels) for r in ranks})
        self.convs_low_to_high = nn.ModuleDict({f"rank_{r}": Conv(channels, channels) for r in range(1, max_rank + 1)})
        self.convs_high_to_low = nn.ModuleDict({f"rank_{r}": Conv(channels, channels) for r in range(max_rank)})


class SCCN(nn.Module):
    def __init__(self, channels, max_rank, n_layers=2, update_func="sigmoid"):
        super().__init__()
        self.layers = nn.ModuleList(SCCNLayer(channels, max_rank, update_func=update_func) for _ in range(n_layers))


class MultiScaleSTFT(nn.Module):
    def __init__(self, scales, sample_rate, magnitude=True, **_):
        super().__init__()
        self.scales, self.magnitude = list(scales), magnitude

    def forward(self, x):
        x = x.reshape(-1, x.shape[-1])
        out = []
        for s in self.scales:
            win = torch.hann_window(s, dtype=x.dtype, device=x.device)
            spec = torch.stft(x, n_fft=s, hop_length=s // 4, win_length=s, window=win, center=True,
                              pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
            out.append(spec.abs() if self.magnitude else spec)
        return out


class AudioDistanceV1(nn.Module):
    def __init__(self, multiscale_stft, log_epsilon):
        super().__init__()
        self.multiscale_stft = multiscale_stft()
        self.log_epsilon = log_epsilon


def install() -> None:
    """Register the stubs in sys.modules (idempotent)."""
    if "TopoModelX" in sys.modules and getattr(sys.modules["TopoModelX"], "_topo_stub", False):
        return
    tnx = _module("toponetx")
    tnx_c = _module("toponetx.classes")
    tnx_c.SimplicialComplex = type("SimplicialComplex", (), {})
    tnx.classes = tnx_c

    rave = _module("rave")
    core = _module("rave.core")
    core.AudioDistanceV1, core.MultiScaleSTFT = AudioDistanceV1, MultiScaleSTFT
    try:
        from einops import rearrange
        core.rearrange = rearrange                      # loss.py:1 imports it (unused)
    except ImportError:                                 # pragma: no cover
        core.rearrange = None
    pq = _module("rave.pqmf")
    pq.PQMF = type("PQMF", (nn.Module,), {"__init__": lambda self, *a, **k: nn.Module.__init__(self)})
    rave.core, rave.pqmf = core, pq

    names = ["TopoModelX", "TopoModelX.topomodelx", "TopoModelX.topomodelx.nn", "TopoModelX.topomodelx.nn.simplicial",
             "TopoModelX.topomodelx.nn.simplicial.sccn", "TopoModelX.topomodelx.nn.simplicial.sccn_layer"]
    mods = [_module(n) for n in names]
    for parent, child, n in zip(mods[:-2], mods[1:-1], names[1:-1]):
        setattr(parent, n.rsplit(".", 1)[1], child)
    mods[3].sccn, mods[3].sccn_layer = mods[4], mods[5]
    mods[4].SCCN = SCCN
    mods[5].SCCNLayer = SCCNLayer
    mods[0]._topo_stub = True
