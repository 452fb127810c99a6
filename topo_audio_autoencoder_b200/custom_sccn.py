"""Drop-in for the reference's ``custom_sccn.py`` (GradientSCCNLayer / GradientSCCN).

Reference: custom_sccn.py:7-162.  Same constructor signatures, parameter names (state-dict
compatible: ``convs_same_rank.rank_r.weight`` ...), dict-in / dict-out forward, and the same
"missing rank -> skipped" rules (custom_sccn.py:69-71, 123-125).

Two execution paths, both through libtopo_b200:
  * ``forward(features, incidences, adjacencies)``: caller-supplied sparse COO operators (any
    pattern, as in the reference's test_sccn.py) -> CSR SpMM / SDDMM kernels + fused combine.
  * ``forward_complex(batch, features)``: a whole batch of complexes described by their
    rectified probabilities; neighbourhoods are walked matrix-free (csrc/aggregate.cu).

The base classes of the reference (TopoModelX ``SCCNLayer`` / ``SCCN`` / ``Conv``) are not on
this machine; ``Conv`` here is the published contract: ``neighborhood @ (x_source @ weight)``,
weight [in, out], Xavier-uniform gain 1.414.  The kernels evaluate it as
``(neighborhood @ x_source) @ weight``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import (lib, check, ptr, ptr_array, stream, ComplexView, CombineParams, CombineGrads, ImageJob, WgradJob, CTA_PARTIAL_FLOATS,
                   WEIGHT_IMAGE_BYTES)
from .rectifier import _Tables


# --------------------------------------------------------------------------------------------
# generic sparse operators
# --------------------------------------------------------------------------------------------
class CsrPattern:
    """int32 CSR of a coalesced COO pattern, plus the CSR of its transpose and the entry
    permutation between the two (index plumbing, done once per operator per forward)."""

    def __init__(self, indices: torch.Tensor, shape: Tuple[int, int]):
        rows, cols = indices[0], indices[1]
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(rows.numel())
        self.row_ptr = torch._convert_indices_from_coo_to_csr(rows, self.shape[0], out_int32=True)
        self.col = cols.to(torch.int32)
        key = cols * self.shape[0] + rows
        self.perm = torch.argsort(key, stable=True)
        self.row_ptr_t = torch._convert_indices_from_coo_to_csr(cols[self.perm], self.shape[1], out_int32=True)
        self.col_t = rows[self.perm].to(torch.int32)


def _spmm_raw(row_ptr, col, vals, x, n_rows):
    if n_rows == 0 or vals.numel() == 0 or x.shape[0] == 0:       # an empty rank or an operator without entries (the
        return torch.zeros(n_rows, x.shape[1], dtype=torch.float32, device=x.device)   # reference multiplies empty tensors)
    y = torch.empty(n_rows, x.shape[1], dtype=torch.float32, device=x.device)
    check(lib.topo_spmm_csr(n_rows, ptr(row_ptr, torch.int32), ptr(col, torch.int32), ptr(vals), ptr(x),
                            x.shape[1], ptr(y), stream()))
    return y


class _SpmmFn(torch.autograd.Function):
    """y = A x (or A^T x) with gradients to the values (SDDMM on the pattern) and to x."""

    @staticmethod
    def forward(ctx, vals, x, pattern: CsrPattern, transposed: bool):
        vals, x = vals.contiguous(), x.contiguous()
        ctx.save_for_backward(vals, x)
        ctx.pattern, ctx.transposed = pattern, transposed
        if transposed:
            return _spmm_raw(pattern.row_ptr_t, pattern.col_t, vals[pattern.perm].contiguous(), x, pattern.shape[1])
        return _spmm_raw(pattern.row_ptr, pattern.col, vals, x, pattern.shape[0])

    @staticmethod
    def backward(ctx, g_y):
        vals, x = ctx.saved_tensors
        p, g_y = ctx.pattern, g_y.contiguous()
        g_vals = torch.empty_like(vals)
        if vals.numel() == 0 or x.shape[0] == 0 or g_y.shape[0] == 0:
            return g_vals, torch.zeros_like(x), None, None
        if ctx.transposed:      # y = A^T x:  dx = A g_y ;  dA[i,j] = x[i] . g_y[j]
            g_x = _spmm_raw(p.row_ptr, p.col, vals, g_y, p.shape[0])
            check(lib.topo_sddmm_csr(p.shape[0], ptr(p.row_ptr, torch.int32), ptr(p.col, torch.int32), ptr(x),
                                     ptr(g_y), x.shape[1], ptr(g_vals), stream()))
        else:                   # y = A x:    dx = A^T g_y ;  dA[i,j] = g_y[i] . x[j]
            g_x = _spmm_raw(p.row_ptr_t, p.col_t, vals[p.perm].contiguous(), g_y, p.shape[1])
            check(lib.topo_sddmm_csr(p.shape[0], ptr(p.row_ptr, torch.int32), ptr(p.col, torch.int32), ptr(g_y),
                                     ptr(x), x.shape[1], ptr(g_vals), stream()))
        return g_vals, g_x, None, None


# --------------------------------------------------------------------------------------------
# fused message combine
# --------------------------------------------------------------------------------------------
# "tc": tcgen05 tensor-core kernels (3xTF32) where instantiated (channels == 64); "simt": the fp32 FFMA
# kernels everywhere.  Both are hand-written CUDA behind the same C ABI; there is no PyTorch fallback.
COMBINE_IMPL = "tc"
# keep m_k and the pre-GELU attention layer from the forward (2 x n_msgs x rows x C floats per call) so the
# backward does not recompute two of its five GEMMs per message
SAVE_ACTIVATIONS = True
# with the tensor-core forward's saved activations, run the whole combine backward as ONE bf16x3 tensor-core
# kernel (csrc/combine_bwd_tc.cu) instead of the FFMA attention kernel + the 3xTF32 conv kernel
FUSED_BACKWARD = True
# tensor-core forward generation: 1 = 3xTF32, two chained GEMMs per message (csrc/combine_tc.cu);
# 2 = bf16x3, one 128 x 128 x 64 product per message, double-buffered, tile-fragment saves (csrc/combine_fwd16.cu)
FORWARD_TC_GENERATION = 2
# build the weights' operand images once per layer and step (one small launch) instead of inside every
# tensor-core kernel launch
SHARED_WEIGHT_IMAGES = True
# launch the four ranks of a layer side by side (four streams, SMs divided by work) instead of one after another
CONCURRENT_RANKS = os.environ.get("TOPO_CONCURRENT_RANKS", "1") not in ("0", "")      # 0: one rank after another, full grid each (profiling)
# make the neighbourhood aggregation part of the same autograd node as the combine: its backward then works on the
# node's own gradient buffers (no clones of the in-place updated ones, no zero fills, no autograd additions)
FUSED_LAYER_NODE = True
# Expected fraction of live rows per rank, for dividing the SMs between the four concurrent rank launches.  Buffers are
# sized by the bound B * n_r and the live counts stay on the device (no host synchronisation), so the split needs an
# estimate: 1.0 = every candidate simplex active (the shipped BinaryGumbel gate).  ComplexStage.calibrate() measures it
# once on a representative batch (GraphedStep does so before capturing); a wrong estimate costs balance, never correctness.
ROW_FRACTION_HINT = [1.0, 1.0, 1.0, 1.0]


class _CombineFn(torch.autograd.Function):
    """out = [LayerNorm] sum_k softmax_k(att(m_k)) m_k,  m_k = scale_k (agg_k @ W_k) + x."""

    @staticmethod
    def forward(ctx, n_msgs, apply_ln, ln_eps, n_rows_dev, zero_dead_rows, images, x, att_w1, att_b1, att_w2, att_b2, ln_g, ln_b, *rest):
        aggs = [t.contiguous() for t in rest[:n_msgs]]
        ws = [t.contiguous() for t in rest[n_msgs:2 * n_msgs]]
        scales = [t.contiguous() for t in rest[2 * n_msgs:3 * n_msgs]]
        rows, ch = aggs[0].shape
        x_c = x.contiguous() if x is not None else None
        tensors = [att_w1.contiguous(), att_b1.contiguous(), att_w2.contiguous(), att_b2.contiguous(),
                   ln_g.contiguous(), ln_b.contiguous()]
        dev = aggs[0].device
        saved = None
        use_tc = COMBINE_IMPL == "tc" and ch == 64
        tile_fragment = False
        if SAVE_ACTIVATIONS and any(ctx.needs_input_grad):
            # the messages and the pre-GELU attention layer, kept for the backward (skips both recompute GEMMs)
            tile_fragment = use_tc and FORWARD_TC_GENERATION == 2
            rows_alloc = -(-rows // 128) * 128 if tile_fragment else rows     # the tile-fragment layout permutes inside 128-row tiles
            saved = ([torch.empty(rows_alloc, ch, dtype=torch.float32, device=dev) for _ in range(n_msgs)],
                     [torch.empty(rows_alloc, ch, dtype=torch.float32, device=dev) for _ in range(n_msgs)],
                     torch.empty(3, rows, dtype=torch.float32, device=dev))      # attention scores (tensor-core forward)
        params = _make_params(ch, n_msgs, aggs, ws, scales, x_c, tensors, ln_eps, apply_ln, saved, tile_fragment, images)
        # rows past the live count are never read by any kernel; they are zero-filled only where the tensor is
        # handed to the caller (last layer), so that padded buffers are safe to reduce over
        out = (torch.zeros if zero_dead_rows else torch.empty)(rows, ch, dtype=torch.float32, device=dev)
        fwd = (lib.topo_sccn_combine_fwd_tc2 if tile_fragment else lib.topo_sccn_combine_fwd_tc) if use_tc else lib.topo_sccn_combine_fwd
        check(fwd(C.byref(params), rows, ptr(n_rows_dev, torch.int32), ptr(out), stream()))
        ctx.save_for_backward(x_c, n_rows_dev, *tensors, *aggs, *ws, *scales)
        ctx.saved_act = saved
        ctx.images = images
        ctx.cfg = (n_msgs, bool(apply_ln), float(ln_eps), x is not None, use_tc and saved is not None, tile_fragment)
        return out

    @staticmethod
    def backward(ctx, g_out):
        n_msgs, apply_ln, ln_eps, has_x, fused, tile_fragment = ctx.cfg
        saved = ctx.saved_tensors
        x_c, n_rows_dev, tensors = saved[0], saved[1], list(saved[2:8])
        aggs = list(saved[8:8 + n_msgs])
        ws = list(saved[8 + n_msgs:8 + 2 * n_msgs])
        scales = list(saved[8 + 2 * n_msgs:8 + 3 * n_msgs])
        rows, ch = aggs[0].shape
        dev = aggs[0].device
        params = _make_params(ch, n_msgs, aggs, ws, scales, x_c, tensors, ln_eps, apply_ln, ctx.saved_act, tile_fragment, ctx.images)
        g_aggs = [torch.empty(rows, ch, dtype=torch.float32, device=dev) for _ in range(n_msgs)]
        g_x = torch.empty(rows, ch, dtype=torch.float32, device=dev) if has_x else None
        wprod = [torch.zeros(ch, ch, dtype=torch.float32, device=dev) for _ in range(n_msgs)]
        g_w1, g_b1 = torch.zeros_like(tensors[0]), torch.zeros_like(tensors[1])
        g_w2, g_b2 = torch.zeros_like(tensors[2]), torch.zeros_like(tensors[3])
        g_g, g_b = torch.zeros_like(tensors[4]), torch.zeros_like(tensors[5])
        grads = CombineGrads()
        for k in range(3):
            grads.g_agg[k] = ptr(g_aggs[k]) if k < n_msgs else None
            grads.g_wprod[k] = ptr(wprod[k]) if k < n_msgs else None
        grads.g_x = ptr(g_x)
        grads.g_att_w1, grads.g_att_b1 = ptr(g_w1), ptr(g_b1)
        grads.g_att_w2, grads.g_att_b2 = ptr(g_w2), ptr(g_b2)
        grads.g_ln_gamma, grads.g_ln_beta = ptr(g_g), ptr(g_b)
        g_out = g_out.contiguous()     # named: a temporary would be freed before the launch reads it
        if fused and (FUSED_BACKWARD or tile_fragment):     # the FFMA kernels read row-major saves only
            # one tensor-core kernel: attention / LayerNorm backward, input gradients and weight-gradient products
            check(lib.topo_sccn_combine_bwd_tc(C.byref(params), rows, ptr(n_rows_dev, torch.int32), ptr(g_out),
                                               C.byref(grads), stream()))
        else:
            workspace = torch.empty(n_msgs * rows * ch, dtype=torch.float32, device=dev)
            check(lib.topo_sccn_combine_bwd_attention(C.byref(params), rows, ptr(n_rows_dev, torch.int32),
                                                      ptr(g_out), C.byref(grads), ptr(workspace), stream()))
            conv = lib.topo_sccn_combine_bwd_conv_tc if (COMBINE_IMPL == "tc" and ch == 64) else lib.topo_sccn_combine_bwd_conv
            check(conv(C.byref(params), rows, ptr(n_rows_dev, torch.int32), C.byref(grads), ptr(workspace), stream()))
        # finish the conv-weight chain: dW_k = scale_k P_k,  dscale_k = <W_k, P_k>
        g_ws = [wprod[k] * scales[k] for k in range(n_msgs)]
        g_ss = [(wprod[k] * ws[k]).sum().reshape(scales[k].shape) for k in range(n_msgs)]
        return (None, None, None, None, None, None, g_x, g_w1, g_b1, g_w2, g_b2,
                g_g if apply_ln else None, g_b if apply_ln else None, *g_aggs, *g_ws, *g_ss)


# ---------------------------------------------------------------------------------------------------
# one layer, all ranks: four tensor-core launches side by side
# ---------------------------------------------------------------------------------------------------
_SIDE_STREAMS: Dict[int, List[torch.cuda.Stream]] = {}


def _side_streams(device: torch.device) -> List[torch.cuda.Stream]:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = [torch.cuda.Stream(device=idx) for _ in range(3)]
    return _SIDE_STREAMS[idx]


BWD_THREE_MESSAGE_COST = float(os.environ.get("TOPO_BWD_NM3_COST", "1.15"))
BWD_FIXED_COST = float(os.environ.get("TOPO_BWD_FIXED_COST", "1.2"))
FWD_FIXED_COST = float(os.environ.get("TOPO_FWD_FIXED_COST", "1.2"))
FWD_THREE_MESSAGE_COST = float(os.environ.get("TOPO_FWD_NM3_COST", "1.0"))


def _sm_shares(costs: Sequence[float], n_sm: int, units: Optional[Sequence[int]] = None) -> List[int]:
    """Split the SMs between concurrent persistent launches so that they finish together.

    costs[i] = total work of launch i; units[i] = its number of indivisible tiles (a launch with s CTAs takes
    ceil(units / s) tile times).  Greedy: every live launch starts with one SM, each further SM goes to the launch
    that would currently finish last."""
    live = [i for i, c in enumerate(costs) if c > 0]
    shares = [0] * len(costs)
    if not live:
        return shares
    if units is None:
        units = [max(1, int(round(c))) for c in costs]
    per_unit = [costs[i] / max(units[i], 1) for i in range(len(costs))]
    for i in live:
        shares[i] = 1

    def finish(i):
        return -(-units[i] // shares[i]) * per_unit[i]

    for _ in range(max(0, n_sm - len(live))):
        worst = max(live, key=lambda i: (finish(i), costs[i]))
        if shares[worst] >= units[worst]:                 # already one CTA per tile: more SMs cannot help it
            rest = [i for i in live if shares[i] < units[i]]
            if not rest:
                break
            worst = max(rest, key=lambda i: (finish(i), costs[i]))
        shares[worst] += 1
    return shares


class _Forked:
    """Run rank launches concurrently: the largest on the current stream, the others on side streams that fork
    from it and join back.  Torch's current stream never changes, so every allocation belongs to it."""

    def __init__(self, device, order):
        self.main = torch.cuda.current_stream(device)
        self.side = _side_streams(device)
        self.order = order                      # ranks, heaviest first
        self.used = []
        fork = torch.cuda.Event()
        fork.record(self.main)
        self.fork = fork

    def stream_for(self, position: int) -> int:
        if position == 0:
            return self.main.cuda_stream
        s = self.side[position - 1]
        s.wait_event(self.fork)
        self.used.append(s)
        return s.cuda_stream

    def join(self):
        for s in self.used:
            ev = torch.cuda.Event()
            ev.record(s)
            self.main.wait_event(ev)


class _LayerCombineFn(torch.autograd.Function):
    """The message combine of ALL ranks of one layer (tensor-core path, C = 64): what _CombineFn does per rank,
    launched as four concurrent kernels with the SMs divided by work, one set of weight images, one zero fill
    for every accumulated gradient and one launch for the conv-weight chain tail."""

    @staticmethod
    def forward(ctx, cfg, *flat):
        ranks = cfg["ranks"]                    # per rank: dict(n_msgs, apply_ln, ln_eps, n_rows_dev, zero_dead_rows, has_x, images)
        cx = cfg.get("cx")                      # given: the neighbourhood aggregation is part of this node (flat[0] = probs, no aggregates in flat)
        per, pos = [], (1 if cx is not None else 0)
        n_agg = 0 if cx is not None else 1
        for rc in ranks:
            n = rc["n_msgs"]
            xin = flat[pos].contiguous()
            tens = [t.contiguous() for t in flat[pos + 1:pos + 7]]
            aggs = [t.contiguous() for t in flat[pos + 7:pos + 7 + n * n_agg]]
            ws = [t.contiguous() for t in flat[pos + 7 + n * n_agg:pos + 7 + n * (n_agg + 1)]]
            scales = [t.contiguous() for t in flat[pos + 7 + n * (n_agg + 1):pos + 7 + n * (n_agg + 2)]]
            per.append(dict(x=xin if rc["has_x"] else None, xin=xin, tens=tens, aggs=aggs, ws=ws, scales=scales, first=pos))
            pos += 7 + n * (n_agg + 2)
        if cx is not None:
            # down[r] = I_{r+1} X_{r+1}, up[r] = I_r^T X_{r-1}, same[r] = A_r X_r for the whole batch (csrc/aggregate.cu)
            probs = flat[0].contiguous()
            xs = [pr["xin"] for pr in per]
            ch_, dev_ = xs[0].shape[1], probs.device
            counts = cx.tables.counts

            def new(r, source_rank=None):
                empty_source = source_rank is not None and counts[source_rank] == 0
                return (torch.zeros if empty_source else torch.empty)(cx.rows_max[r], ch_, dtype=torch.float32, device=dev_)

            down = [new(0, 1), new(1, 2), new(2, 3), None]
            up = [None, new(1), new(2), new(3)]
            same = [new(r) for r in range(4)]
            view = cx.view(probs)
            check(lib.topo_sccn_aggregate_fwd(cx.tables.handle, C.byref(view), ch_, ptr_array(xs, 4), ptr_array(down, 4),
                                              ptr_array(up, 4), ptr_array(same, 4), stream()))
            for r, pr in enumerate(per):
                pr["aggs"] = [same[r]] + ([down[r]] if r < 3 else []) + ([up[r]] if r > 0 else [])
            ctx.agg = dict(cx=cx, probs=probs, xs=xs, down2=down[2], up2=up[2], up3=up[3])
        else:
            ctx.agg = None
        dev = per[0]["aggs"][0].device
        ch = per[0]["aggs"][0].shape[1]
        need_grad = SAVE_ACTIVATIONS and any(ctx.needs_input_grad)
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        tiles = [max(1, int(-(-pr["aggs"][0].shape[0] * ROW_FRACTION_HINT[r] // 128))) if pr["aggs"][0].shape[0] else 0
                 for r, pr in enumerate(per)]
        # per tile: a fixed part + one part per message
        costs = [t * (rc["n_msgs"] + FWD_FIXED_COST) * (FWD_THREE_MESSAGE_COST if rc["n_msgs"] == 3 else 1.0) for t, rc in zip(tiles, ranks)]
        shares = _sm_shares(costs, n_sm, tiles)
        # the backward of a three-message rank issues both halves of its row contractions from one thread (tensor memory has no
        # room for per-half accumulators there, and an accumulator must have ONE issuing thread to be reproducible): its tiles
        # cost a little more than the forward's ratio says
        ctx.shares_bwd = _sm_shares([t * (rc["n_msgs"] + BWD_FIXED_COST) * (BWD_THREE_MESSAGE_COST if rc["n_msgs"] == 3 else 1.0)
                                     for t, rc in zip(tiles, ranks)], n_sm, tiles)
        outs = []
        for pr, rc in zip(per, ranks):
            rows = pr["aggs"][0].shape[0]
            pr["saved"] = None
            if need_grad:
                rows_alloc = -(-rows // 128) * 128
                pr["saved"] = ([torch.empty(rows_alloc, ch, dtype=torch.float32, device=dev) for _ in range(rc["n_msgs"])],
                               [torch.empty(rows_alloc, ch, dtype=torch.float32, device=dev) for _ in range(rc["n_msgs"])],
                               torch.empty(3, rows, dtype=torch.float32, device=dev))
            outs.append((torch.zeros if rc["zero_dead_rows"] else torch.empty)(rows, ch, dtype=torch.float32, device=dev))
        order = sorted(range(len(ranks)), key=lambda i: -costs[i])
        forked = _Forked(dev, order)
        for position, i in enumerate(order):
            pr, rc = per[i], ranks[i]
            rows = pr["aggs"][0].shape[0]
            if rows == 0:
                continue
            params = _make_params(ch, rc["n_msgs"], pr["aggs"], pr["ws"], pr["scales"], pr["x"], pr["tens"], rc["ln_eps"],
                                  rc["apply_ln"], pr["saved"], need_grad, rc["images"], shares[i])
            fwd = lib.topo_sccn_combine_fwd_tc2 if need_grad else lib.topo_sccn_combine_fwd_tc
            check(fwd(C.byref(params), rows, ptr(rc["n_rows_dev"], torch.int32), ptr(outs[i]), forked.stream_for(position)))
        forked.join()
        ctx.per, ctx.ranks, ctx.shares, ctx.order = per, ranks, shares, order
        ctx.n_flat = len(flat)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_outs):
        per, ranks, shares, order = ctx.per, ctx.ranks, ctx.shares_bwd, ctx.order
        dev = per[0]["aggs"][0].device
        ch = per[0]["aggs"][0].shape[1]
        # parameter gradients: every CTA of a rank's launch stores its partial sums in its own slot of `partials[i]`
        # (plain stores) and ONE launch adds the slots in CTA order afterwards (topo_sccn_finish_weight_grads): no
        # floating-point atomics, bit-reproducible.  `acc` receives the results: per rank [w1 | b1 | w2 | gamma | beta | b2].
        sizes = [ch * ch + 4 * ch + 4 for rc in ranks]
        acc = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        grads_flat = [None] * ctx.n_flat
        forked = _Forked(dev, order)
        keep, jobs = [], []
        off = 0
        views = []
        partials = []
        for i, (pr, rc) in enumerate(zip(per, ranks)):
            base = off
            v = {"w1": acc[base:base + ch * ch].view(ch, ch)}
            base += ch * ch
            v["b1"] = acc[base:base + ch]; base += ch          # b1, w2, gamma, beta, b2: the order of a CTA slot's tail
            v["w2"] = acc[base:base + ch]; base += ch
            v["gamma"] = acc[base:base + ch]; base += ch
            v["beta"] = acc[base:base + ch]; base += ch
            v["b2"] = acc[base:base + 1]
            views.append(v)
            off += sizes[i]
            slots = int(lib.topo_sccn_combine_grid(pr["aggs"][0].shape[0], shares[i]))
            partials.append(torch.empty(max(slots, 1) * CTA_PARTIAL_FLOATS, dtype=torch.float32, device=dev))
        # gradient buffers of this call live in locals (not on ctx: they would stay allocated until the graph is freed)
        g_aggs_of, g_x_of, g_x_total_of = [None] * len(ranks), [None] * len(ranks), [None] * len(ranks)
        for position, i in enumerate(order):
            pr, rc, v = per[i], ranks[i], views[i]
            rows, n = pr["aggs"][0].shape[0], rc["n_msgs"]
            g_out = g_outs[i]
            g_aggs = [torch.empty(rows, ch, dtype=torch.float32, device=dev) for _ in range(n)]
            g_x = torch.empty(rows, ch, dtype=torch.float32, device=dev) if rc["has_x"] else None
            g_aggs_of[i], g_x_of[i] = g_aggs, g_x
            if rows == 0 or g_out is None:
                for t in g_aggs:
                    t.zero_()
                if g_x is not None:
                    g_x.zero_()
                continue
            g_out = g_out.contiguous()
            keep.append(g_out)
            params = _make_params(ch, n, pr["aggs"], pr["ws"], pr["scales"], pr["x"], pr["tens"], rc["ln_eps"], rc["apply_ln"],
                                  pr["saved"], True, rc["images"], shares[i])
            grads = CombineGrads()
            for k in range(3):
                grads.g_agg[k] = ptr(g_aggs[k]) if k < n else None
            grads.g_x = ptr(g_x)
            grads.cta_partials = ptr(partials[i])
            check(lib.topo_sccn_combine_bwd_tc(C.byref(params), rows, ptr(rc["n_rows_dev"], torch.int32), ptr(g_out),
                                               C.byref(grads), forked.stream_for(position)))
        forked.join()
        # the ordered sum over the CTA slots for every parameter of the layer in one launch, with the conv-weight chain's tail
        # on top of it: P_k = sum of slots, dW_k = s_k P_k, ds_k = <W_k, P_k>
        n_total = sum(rc["n_msgs"] for rc in ranks)
        g_w_all = torch.empty(n_total, ch, ch, dtype=torch.float32, device=dev)
        g_s_all = torch.empty(n_total, dtype=torch.float32, device=dev)
        arr = (WgradJob * (n_total + 2 * len(ranks)))()
        q = m = 0          # job number, message number

        def _slots(job, i, offset, count):
            pr, rc = per[i], ranks[i]
            rows = pr["aggs"][0].shape[0]
            launched = rows > 0 and g_outs[i] is not None
            job.partials, job.partial_offset, job.count = ptr(partials[i]), offset, count
            job.n_rows_dev, job.rows, job.max_ctas = ptr(rc["n_rows_dev"], torch.int32), (rows if launched else 0), shares[i]

        for i, (pr, rc) in enumerate(zip(per, ranks)):
            for k in range(rc["n_msgs"]):
                arr[q].w, arr[q].scale = ptr(pr["ws"][k]), ptr(pr["scales"][k])
                arr[q].g_w, arr[q].g_scale = g_w_all[m].data_ptr(), g_s_all[m:m + 1].data_ptr()
                _slots(arr[q], i, (1 + k) * ch * ch, ch * ch)
                q += 1
                m += 1
            arr[q].g_w = ptr(views[i]["w1"])
            _slots(arr[q], i, 0, ch * ch)
            q += 1
            arr[q].g_w = ptr(views[i]["b1"])                     # b1, w2, gamma, beta, b2 are contiguous on both sides
            _slots(arr[q], i, 4 * ch * ch, 4 * ch + 1)
            q += 1
        n_jobs = q
        # ... on a side stream: the aggregation backward below does not depend on it
        tail = _Forked(dev, None)
        check(lib.topo_sccn_finish_weight_grads(arr, n_jobs, ch, tail.stream_for(1) if ctx.agg is not None else stream()))
        agg = ctx.agg
        if agg is not None:
            # the aggregation backward runs on this node's own buffers: it updates g_down[2], g_up[2], g_up[3] in
            # place and ACCUMULATES into the residual gradients the combine kernels just wrote (no clones, no zero
            # fills, no autograd additions)
            cx = agg["cx"]
            g_same = [g_aggs_of[r][0] for r in range(4)]
            g_down = [g_aggs_of[r][1] for r in range(3)] + [None]
            g_up = [None] + [g_aggs_of[r][2 if r < 3 else 1] for r in range(1, 4)]
            g_x = [g_x_of[i] if g_x_of[i] is not None else torch.zeros_like(pr["xin"]) for i, pr in enumerate(per)]
            g_probs = torch.zeros_like(agg["probs"])
            view = cx.view(agg["probs"])
            check(lib.topo_sccn_aggregate_bwd(cx.tables.handle, C.byref(view), ch, ptr_array(agg["xs"], 4),
                                              ptr_array([None, None, agg["down2"], None], 4),
                                              ptr_array([None, None, agg["up2"], agg["up3"]], 4),
                                              ptr_array(g_down, 4), ptr_array(g_up, 4), ptr_array(g_same, 4),
                                              ptr_array(g_x, 4), ptr(g_probs), stream()))
            grads_flat[0] = g_probs
            g_x_total_of = list(g_x)
            tail.join()
        # The three message scales are shared by all ranks of the layer: autograd would add up one 1-element gradient per
        # (rank, message) -- seven tiny launches per layer in the middle of the backward chain.  One product with a 0/1
        # matrix sums them per distinct scale; the total is handed to the scale's first occurrence, the others get None.
        owners, groups = [], {}
        for pr, rc in zip(per, ranks):
            for k in range(rc["n_msgs"]):
                key = pr["scales"][k].data_ptr()
                owners.append(groups.setdefault(key, len(groups)))
        onehot = _scale_groups(tuple(owners), len(groups), dev)
        g_s_tot = g_s_all @ onehot                                   # [distinct scales]
        handed = set()
        n_agg = 0 if agg is not None else 1
        q = 0
        for i, (pr, rc) in enumerate(zip(per, ranks)):
            n, first, v = rc["n_msgs"], pr["first"], views[i]
            if agg is not None:
                grads_flat[first] = g_x_total_of[i]
            elif rc["has_x"]:
                grads_flat[first] = g_x_of[i]
            grads_flat[first + 1:first + 5] = [v["w1"], v["b1"], v["w2"], v["b2"]]
            if rc["apply_ln"]:
                grads_flat[first + 5], grads_flat[first + 6] = v["gamma"], v["beta"]
            for k in range(n):
                if n_agg:
                    grads_flat[first + 7 + k] = g_aggs_of[i][k]
                grads_flat[first + 7 + n * n_agg + k] = g_w_all[q]
                grp = owners[q]
                if grp not in handed:
                    handed.add(grp)
                    grads_flat[first + 7 + n * (n_agg + 1) + k] = g_s_tot[grp:grp + 1].reshape(pr["scales"][k].shape)
                q += 1
        return (None, *grads_flat)


_SCALE_GROUPS: Dict[tuple, torch.Tensor] = {}


def _scale_groups(owners: tuple, n_groups: int, device) -> torch.Tensor:
    """[len(owners), n_groups] 0/1 matrix: entry (q, g) = message q uses distinct scale g (cached per device and pattern)."""
    key = (owners, n_groups, str(device))
    if key not in _SCALE_GROUPS:
        m = torch.zeros(len(owners), n_groups, dtype=torch.float32)
        for q, g in enumerate(owners):
            m[q, g] = 1.0
        _SCALE_GROUPS[key] = m.to(device)
    return _SCALE_GROUPS[key]


def _make_params(ch, n_msgs, aggs, ws, scales, x, tensors, ln_eps, apply_ln, saved=None, tile_fragment=False,
                 images: Optional[torch.Tensor] = None, max_ctas: int = 0) -> CombineParams:
    p = CombineParams()
    p.channels, p.n_msgs = ch, n_msgs
    for k in range(3):
        p.agg[k] = ptr(aggs[k]) if k < n_msgs else None
        p.w[k] = ptr(ws[k]) if k < n_msgs else None
        p.scale[k] = ptr(scales[k]) if k < n_msgs else None
        p.saved_m[k] = ptr(saved[0][k]) if (saved is not None and k < n_msgs) else None
        p.saved_pre[k] = ptr(saved[1][k]) if (saved is not None and k < n_msgs) else None
    p.saved_score = ptr(saved[2]) if (saved is not None and len(saved) > 2) else None
    p.saved_layout = 1 if tile_fragment else 0       # TOPO_SAVED_TILE_FRAGMENT / TOPO_SAVED_ROW_MAJOR
    p.weight_images = ptr(images, torch.uint8) if images is not None else None
    p.max_ctas = int(max_ctas)
    p.x = ptr(x)
    p.att_w1, p.att_b1, p.att_w2, p.att_b2 = (ptr(t) for t in tensors[:4])
    p.ln_gamma, p.ln_beta = ptr(tensors[4]), ptr(tensors[5])
    p.ln_eps, p.apply_ln = float(ln_eps), int(bool(apply_ln))
    return p


# --------------------------------------------------------------------------------------------
# matrix-free aggregation over a batch of complexes
# --------------------------------------------------------------------------------------------
@dataclass
class BatchedComplex:
    """A batch of active sub-complexes on the device (the layout of include/topo_b200.h)."""
    tables: _Tables
    probs: torch.Tensor       # [B, N] rectified, carries autograd history
    pos: torch.Tensor         # [B, N] int32
    act_idx: torch.Tensor     # [B, N] int32
    counts: torch.Tensor      # [B, 4] int32
    row_off: torch.Tensor     # [4, B+1] int32
    rows_max: List[int]       # allocation bound per rank (B * n_r, or exact after a sync)
    host_counts: Optional[torch.Tensor] = None   # [B, 4] on the host when the caller synchronised

    @property
    def batch(self) -> int:
        return int(self.probs.shape[0])

    def view(self, probs: Optional[torch.Tensor] = None) -> ComplexView:
        v = ComplexView()
        v.probs = ptr(self.probs.detach() if probs is None else probs)
        v.pos, v.act_idx = ptr(self.pos, torch.int32), ptr(self.act_idx, torch.int32)
        v.counts, v.row_off = ptr(self.counts, torch.int32), ptr(self.row_off, torch.int32)
        v.batch = self.batch
        return v

    def live_rows(self, rank: int) -> torch.Tensor:
        """device int32 scalar: total active rows of `rank` in the batch"""
        return self.row_off[rank, self.batch:self.batch + 1]


class _AggregateFn(torch.autograd.Function):
    """(probs, x0..x3) -> (down0..2, up1..3, same0..3)   [csrc/aggregate.cu]"""

    @staticmethod
    def forward(ctx, cx: BatchedComplex, probs, x0, x1, x2, x3):
        xs = [t.contiguous() for t in (x0, x1, x2, x3)]
        probs = probs.contiguous()
        ch = xs[0].shape[1]
        dev = probs.device
        counts = cx.tables.counts

        def new(r, source_rank=None):
            # every live row is written by the kernel; an aggregate whose source rank does not exist is all zero
            empty_source = source_rank is not None and counts[source_rank] == 0
            return (torch.zeros if empty_source else torch.empty)(cx.rows_max[r], ch, dtype=torch.float32, device=dev)

        down = [new(0, 1), new(1, 2), new(2, 3), None]
        up = [None, new(1), new(2), new(3)]
        same = [new(r) for r in range(4)]
        view = cx.view(probs)
        check(lib.topo_sccn_aggregate_fwd(cx.tables.handle, C.byref(view), ch, ptr_array(xs, 4), ptr_array(down, 4),
                                          ptr_array(up, 4), ptr_array(same, 4), stream()))
        ctx.save_for_backward(probs, *xs, down[2], up[2], up[3])
        ctx.cx, ctx.ch = cx, ch
        return down[0], down[1], down[2], up[1], up[2], up[3], same[0], same[1], same[2], same[3]

    @staticmethod
    def backward(ctx, gd0, gd1, gd2, gu1, gu2, gu3, gs0, gs1, gs2, gs3):
        cx, ch = ctx.cx, ctx.ch
        saved = ctx.saved_tensors
        probs, xs, down2, up2, up3 = saved[0], list(saved[1:5]), saved[5], saved[6], saved[7]
        dev = probs.device

        def dense(g, r, clone=False):
            if g is None:
                return torch.zeros(cx.rows_max[r], ch, dtype=torch.float32, device=dev)
            return g.clone() if clone else g.contiguous()

        # g_down[2], g_up[2], g_up[3] are updated in place by the kernel: never hand it autograd's buffers
        g_down = [dense(gd0, 0), dense(gd1, 1), dense(gd2, 2, clone=True), None]
        g_up = [None, dense(gu1, 1), dense(gu2, 2, clone=True), dense(gu3, 3, clone=True)]
        g_same = [dense(g, r) for r, g in enumerate((gs0, gs1, gs2, gs3))]
        g_x = [torch.zeros(cx.rows_max[r], ch, dtype=torch.float32, device=dev) for r in range(4)]
        g_probs = torch.zeros_like(probs)
        view = cx.view(probs)
        check(lib.topo_sccn_aggregate_bwd(cx.tables.handle, C.byref(view), ch, ptr_array(xs, 4),
                                          ptr_array([None, None, down2, None], 4), ptr_array([None, None, up2, up3], 4),
                                          ptr_array(g_down, 4), ptr_array(g_up, 4), ptr_array(g_same, 4),
                                          ptr_array(g_x, 4), ptr(g_probs), stream()))
        return (None, g_probs, *g_x)


# --------------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------------
class Conv(nn.Module):
    """Stand-in for TopoModelX ``Conv`` as SCCNLayer instantiates it (no bias, no activation)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        nn.init.xavier_uniform_(self.weight, gain=1.414)


def _present(d, key) -> bool:
    return d is not None and key in d and d[key] is not None


class GradientSCCNLayer(nn.Module):
    """reference custom_sccn.py:7-138."""

    def __init__(self, channels, max_rank, aggr_func="sum", update_func="relu", residual=True, is_final_layer=False):
        super().__init__()
        self.channels, self.max_rank = channels, max_rank
        self.aggr_func, self.update_func = aggr_func, update_func     # accepted, unused by the forward (as in the reference)
        self.residual = residual
        self.is_final_layer = is_final_layer
        ranks = range(max_rank + 1)
        self.convs_same_rank = nn.ModuleDict({f"rank_{r}": Conv(channels, channels) for r in ranks})
        self.convs_low_to_high = nn.ModuleDict({f"rank_{r}": Conv(channels, channels) for r in ranks if r > 0})
        self.convs_high_to_low = nn.ModuleDict({f"rank_{r}": Conv(channels, channels) for r in ranks if r < max_rank})
        self.layer_norms = nn.ModuleDict({f"rank_{r}": nn.LayerNorm(channels) for r in ranks})             # :15-18
        self.message_scales = nn.ParameterDict({k: nn.Parameter(torch.ones(1))                              # :21-25
                                                for k in ("same_rank", "low_to_high", "high_to_low")})
        self.message_attention = nn.ModuleDict({                                                            # :28-34
            f"rank_{r}": nn.Sequential(nn.Linear(channels, channels), nn.GELU(), nn.Linear(channels, 1))
            for r in ranks})

    # -- shared tail: messages -> output rows --------------------------------------------------
    def _combine(self, key: str, x: torch.Tensor, msgs: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                 n_rows_dev: Optional[torch.Tensor] = None, zero_dead_rows: bool = True,
                 images: Optional[torch.Tensor] = None) -> torch.Tensor:
        att, ln = self.message_attention[key], self.layer_norms[key]
        apply_ln = self.training and not self.is_final_layer                                               # :133-134
        aggs, ws, scales = zip(*msgs)
        return _CombineFn.apply(len(msgs), apply_ln, ln.eps, n_rows_dev, zero_dead_rows, images, x if self.residual else None,
                                att[0].weight, att[0].bias, att[2].weight.reshape(-1), att[2].bias,
                                ln.weight, ln.bias, *aggs, *ws, *scales)

    # -- reference signature: caller-supplied sparse operators ---------------------------------
    def forward(self, features, incidences, adjacencies, _patterns: Optional[dict] = None):
        patterns = {} if _patterns is None else _patterns

        def pattern_of(mat: torch.Tensor):
            k = id(mat)
            if k not in patterns:
                m = mat if mat.is_coalesced() else mat.coalesce()
                patterns[k] = (CsrPattern(m.indices(), m.shape), m)
            return patterns[k]

        def aggregate(mat, x_src, transposed):
            pat, m = pattern_of(mat)
            return _SpmmFn.apply(m.values(), x_src, pat, transposed)

        out = {}
        for r in range(self.max_rank + 1):
            key = f"rank_{r}"
            if not _present(features, key):                                                                # :69-71
                out[key] = None
                continue
            x = features[key]
            msgs = []
            if _present(adjacencies, key):                                                                 # :77-85
                msgs.append((aggregate(adjacencies[key], x, False),
                             self.convs_same_rank[key].weight, self.message_scales["same_rank"]))
            up_key = f"rank_{r + 1}"
            if r < self.max_rank and _present(features, up_key) and _present(incidences, up_key):         # :88-102
                msgs.append((aggregate(incidences[up_key], features[up_key], False),
                             self.convs_high_to_low[key].weight, self.message_scales["high_to_low"]))
            low_key = f"rank_{r - 1}"
            if r > 0 and _present(features, low_key) and _present(incidences, key):                       # :105-120
                msgs.append((aggregate(incidences[key], features[low_key], True),
                             self.convs_low_to_high[key].weight, self.message_scales["low_to_high"]))
            if not msgs:                                                                                   # :123-125
                out[key] = x
                continue
            out[key] = self._combine(key, x, msgs) if x.shape[0] else x
        return out

    # -- batch of complexes, matrix-free --------------------------------------------------------
    def forward_complex(self, cx: BatchedComplex, xs: Sequence[torch.Tensor], zero_dead_rows: bool = True,
                        images: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        if self.max_rank != 3:
            raise ValueError("forward_complex is built for the reference's max_rank = 3 complexes")
        sc = self.message_scales
        tc = COMBINE_IMPL == "tc" and self.channels == 64
        layer_node = tc and CONCURRENT_RANKS and FORWARD_TC_GENERATION == 2 and all(cx.rows_max[r] for r in range(4))
        fused_agg = layer_node and FUSED_LAYER_NODE
        if fused_agg:
            same = down = up = (None, None, None, None)        # formed inside the layer node
        else:
            d0, d1, d2, u1, u2, u3, s0, s1, s2, s3 = _AggregateFn.apply(cx, cx.probs, *xs)
            same, down, up = (s0, s1, s2, s3), (d0, d1, d2, None), (None, u1, u2, u3)
        per_rank = []
        for r in range(4):
            key = f"rank_{r}"
            msgs = [(same[r], self.convs_same_rank[key].weight, sc["same_rank"])]
            if r < 3:
                msgs.append((down[r], self.convs_high_to_low[key].weight, sc["high_to_low"]))
            if r > 0:
                msgs.append((up[r], self.convs_low_to_high[key].weight, sc["low_to_high"]))
            per_rank.append((key, msgs))
        if images is None:
            images = self._weight_images(xs[0].device) if (tc and SHARED_WEIGHT_IMAGES) else [None] * 4
        if layer_node:
            apply_ln = self.training and not self.is_final_layer
            ranks, flat = [], ([cx.probs] if fused_agg else [])
            for r, (key, msgs) in enumerate(per_rank):
                att, ln = self.message_attention[key], self.layer_norms[key]
                aggs, ws, scales = zip(*msgs)
                ranks.append(dict(n_msgs=len(msgs), apply_ln=apply_ln, ln_eps=ln.eps, n_rows_dev=cx.live_rows(r),
                                  zero_dead_rows=zero_dead_rows, has_x=self.residual, images=images[r]))
                flat += [xs[r], att[0].weight, att[0].bias, att[2].weight.reshape(-1), att[2].bias, ln.weight, ln.bias,
                         *([] if fused_agg else aggs), *ws, *scales]
            return list(_LayerCombineFn.apply(dict(ranks=ranks, cx=cx if fused_agg else None), *flat))
        out = []
        for r, (key, msgs) in enumerate(per_rank):
            out.append(self._combine(key, xs[r], msgs, cx.live_rows(r), zero_dead_rows, images[r]) if cx.rows_max[r] else xs[r])
        return out

    def _image_jobs(self):
        """[(rank key, [(W_k, s_k), ...])] in the message order of forward_complex: same, from above, from below."""
        sc = self.message_scales
        out = []
        for r in range(4):
            key = f"rank_{r}"
            msgs = [(self.convs_same_rank[key].weight, sc["same_rank"])]
            if r < 3:
                msgs.append((self.convs_high_to_low[key].weight, sc["high_to_low"]))
            if r > 0:
                msgs.append((self.convs_low_to_high[key].weight, sc["low_to_high"]))
            out.append((key, msgs))
        return out

    def _weight_images(self, device) -> List[torch.Tensor]:
        return prepare_weight_images([self], device)[0]


def prepare_weight_images(layers, device) -> List[List[torch.Tensor]]:
    """bf16x3 operand images of the weights of `layers` for all ranks, built by ONE launch (csrc/weight_images.cu)
    and shared by the forward and backward tensor-core kernels: per layer and rank [W1 | k: W_k, V_k = s_k W_k W1^T]."""
    plan = [layer._image_jobs() for layer in layers]
    sizes = [[WEIGHT_IMAGE_BYTES * (1 + 2 * len(msgs)) for _, msgs in per_rank] for per_rank in plan]
    buf = torch.empty(sum(sum(sz) for sz in sizes), dtype=torch.uint8, device=device)
    views, jobs, off = [], [], 0
    for layer, per_rank, sz in zip(layers, plan, sizes):
        views.append([])
        for (key, msgs), size in zip(per_rank, sz):
            views[-1].append(buf[off:off + size])
            w1 = layer.message_attention[key][0].weight.detach().contiguous()
            base = buf.data_ptr() + off
            jobs.append((None, None, w1, base))
            for k, (w, s) in enumerate(msgs):
                jobs.append((w.detach().contiguous(), s.detach().contiguous(), w1, base + WEIGHT_IMAGE_BYTES * (1 + 2 * k)))
            off += size
    for first in range(0, len(jobs), 96):
        chunk = jobs[first:first + 96]
        arr = (ImageJob * len(chunk))()
        for q, (w, s, w1, dst) in enumerate(chunk):
            arr[q].w, arr[q].scale, arr[q].att_w1, arr[q].dst = ptr(w), ptr(s), ptr(w1), dst
        check(lib.topo_sccn_prepare_images(arr, len(chunk), layers[0].channels, stream()))
    return views


class GradientSCCN(nn.Module):
    """reference custom_sccn.py:140-162.  ``residual`` is accepted and ignored (the reference
    builds every layer with the default residual=True, :147-155); ``update_func`` never reaches
    the forward."""

    def __init__(self, channels, max_rank, n_layers=2, update_func="sigmoid", residual=False):
        super().__init__()
        self.channels, self.max_rank = channels, max_rank
        self.layers = nn.ModuleList([
            GradientSCCNLayer(channels=channels, max_rank=max_rank, update_func=update_func,
                              is_final_layer=(i == n_layers - 1))
            for i in range(n_layers)])

    def forward(self, features, incidences, adjacencies):
        patterns: dict = {}          # CSR of each operator is built once and shared by all layers
        for layer in self.layers:
            features = layer(features, incidences, adjacencies, _patterns=patterns)
        return features

    def forward_complex(self, cx: BatchedComplex, xs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        images = [None] * len(self.layers)
        if COMBINE_IMPL == "tc" and self.channels == 64 and SHARED_WEIGHT_IMAGES and self.max_rank == 3:
            images = prepare_weight_images(list(self.layers), xs[0].device)      # every layer's weights in one launch
        for i, layer in enumerate(self.layers):
            xs = layer.forward_complex(cx, xs, zero_dead_rows=(i == len(self.layers) - 1), images=images[i])
        return list(xs)
