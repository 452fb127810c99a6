"""CUDA-graph capture of one complex-stage training step.

The stage is ~130 short launches per step (4 ranks x 6 layers x forward/backward kernels plus the gate,
rectifier and bookkeeping); at B200 speeds the Python / autograd / ctypes launch path costs as much as
the kernels themselves.  ``GraphedStep`` captures forward + backward once (static shapes: the buffers are
sized by the batch bound B * n_r and the live row counts stay on the device, so nothing in the step
synchronises or depends on the data) and replays it with one ``cudaGraphLaunch``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


class GraphedStep:
    """forward(logits, noise) + backward(upstream gradients) of a ``ComplexStage`` as one CUDA graph.

    After ``replay(logits, noise)``: ``outputs`` holds the stage's output tensors (static buffers),
    ``logits_grad`` the gradient w.r.t. the logits and every parameter's ``.grad`` is updated in place.

    Python-side state and the graph.  Read from device memory at every replay (never baked in): parameters, inputs,
    upstream gradients, the Hard Concrete temperature (``HardConcrete._temp_buf``).  Baked in at capture, because they are
    by-value kernel arguments or shape decisions: the BinaryGumbel temperature, ``training`` / ``ste`` flags, LayerNorm
    epsilons, the penalty bounds, the batch size.  ``replay`` compares these with their captured values and re-captures
    the graph when one changed (the reference's trainer anneals the temperature once per epoch, trainer.py:266).
    """

    def __init__(self, stage, logits: torch.Tensor, noise: Optional[torch.Tensor], upstream: Sequence[torch.Tensor],
                 warmup: int = 3):
        self.stage = stage
        self.params = [p for p in stage.parameters() if p.requires_grad]
        self.logits = logits.detach().clone().requires_grad_(True)
        self.noise = None if noise is None else noise.detach().clone()
        self.upstream = [u.detach().clone() for u in upstream]
        self.outputs: Dict[str, torch.Tensor] = {}
        self.logits_grad: Optional[torch.Tensor] = None
        self.captures = 0
        self._capture(warmup)

    def _baked_state(self):
        """Everything the captured launches hold by value."""
        head = self.stage.head
        return (float(head.gumbel.current_temp), bool(self.stage.training), bool(head.sampler.ste), head.gate_kind, head.bias_on,
                float(head.min_active_vertices), float(head.max_active_vertices),
                tuple(bool(layer.training) for layer in self.stage.sccn.layers))

    def _capture(self, warmup: int = 1):
        self.graph = torch.cuda.CUDAGraph()
        self._baked = self._baked_state()
        self.captures += 1
        if hasattr(self.stage, "calibrate"):        # live-row estimate for the SM split of the concurrent rank launches
            self.stage.calibrate(self.logits.detach(), self.noise)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the capture: allocator, function attributes
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        for p in self.params:
            p.grad = None
        self.logits.grad = None
        with torch.cuda.graph(self.graph):
            out = self._run()
        self.outputs = {k: v for k, v in out.items() if isinstance(v, torch.Tensor)}
        self.logits_grad = self.logits.grad
        # the graph writes parameter gradients into these static buffers on every replay
        self.param_grads = [p.grad for p in self.params]

    def _run(self):
        for p in self.params:
            p.grad = None
        self.logits.grad = None
        out = self.stage(self.logits, self.noise)
        heads = [out[f"rank_{r}"] for r in range(4)] + [out["vertex_penalty"], out["entropy_loss"]]
        torch.autograd.backward(heads, self.upstream)
        return out

    def replay(self, logits: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        """Copy new inputs into the static buffers (device or pinned host tensors) and launch the graph."""
        if logits is not None:
            self.logits.data.copy_(logits, non_blocking=True)
        if noise is not None and self.noise is not None:
            self.noise.copy_(noise, non_blocking=True)
        if self._baked_state() != self._baked:          # e.g. BinaryGumbel.set_temperature between epochs
            self._capture()
        self.graph.replay()
        for p, g in zip(self.params, self.param_grads):      # re-attach if the caller cleared or replaced .grad
            p.grad = g
        return self.outputs
