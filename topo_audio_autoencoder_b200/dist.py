"""Data-parallel plumbing for the complex stage: one process per GPU, torch.distributed.

The stage shards by batch (independent complexes): parameters are replicated, every
rank runs its own clips, and the only exchange is ONE all-reduce of the flattened parameter gradients
per optimizer step (the reference steps every 4 micro-batches, trainer.py:288-293, so gradients are
accumulated locally in between).  The distance sweep shards by row blocks and needs no collective.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_batch(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `total` clips owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def flatten_grads(params: Iterable[torch.nn.Parameter]) -> Tuple[torch.Tensor, List[torch.nn.Parameter]]:
    """Parameters that received no gradient (unused on this path: identical on every replica) are skipped."""
    plist = [p for p in params if p.requires_grad and p.grad is not None]
    return torch.cat([p.grad.reshape(-1) for p in plist]), plist


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True) -> None:
    """One bucket, one collective: NCCL over NVLink on GPUs, gloo in the CPU tests."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    flat, plist = flatten_grads(params)
    if average and dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)          # the division rides inside the collective
    else:
        dist.all_reduce(flat)
        if average:
            flat.div_(dist.get_world_size())
    # one multi-tensor copy back into the gradient buffers (they stay the CUDA graph's static buffers) instead of
    # one launch per parameter: ~180 parameters made this the costliest part of the exchange
    views, offset = [], 0
    for p in plist:
        n = p.numel()
        views.append(flat[offset:offset + n].view_as(p))
        offset += n
    torch._foreach_copy_([p.grad for p in plist], views)


def clip_grad_norm_after_reduce(params: Iterable[torch.nn.Parameter], max_norm: float = 10.0) -> torch.Tensor:
    """Global-norm clipping (trainer.py:290) must see the reduced gradients."""
    return torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], max_norm)
