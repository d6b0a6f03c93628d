"""Batch-sharded data parallelism for the spectral-mixing layer (one process per GPU, torch.distributed).

Every (batch, channel) column is an independent transform (reference: fft(x, dim=1), spectral_layers.py:88), so
the forward and dL/dx need no communication.  The only cross-batch coupling is the filter/bias gradient sum
(wirtinger_ops.py:77-80) -> ONE all-reduce(sum) per layer per step over [weight_real.grad | weight_imag.grad |
bias.grad].  The reference has no distributed code at all; this is the B200-native plumbing.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_batch(x: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Rank r's slice of the batch axis (contiguous shards, sizes differ by at most one)."""
    B = x.shape[0]
    base, rem = divmod(B, world_size)
    start = rank * base + min(rank, rem)
    return x[start: start + base + (1 if rank < rem else 0)]


def filter_grad_tensors(modules: Iterable[torch.nn.Module]) -> List[torch.Tensor]:
    out = []
    for m in modules:
        for name in ("weight_real", "weight_imag", "bias"):
            p = getattr(m, name, None)
            if p is not None and p.grad is not None:
                out.append(p.grad)
    return out


def allreduce_filter_grads(modules: Iterable[torch.nn.Module], group: Optional[dist.ProcessGroup] = None,
                           average: bool = False, async_op: bool = False):
    """Sum (or average) the filter/bias gradients of the given SpectralMixingLayer modules across ranks with a
    single flat all-reduce.  ``sum`` reproduces a single-process run on the concatenated batch exactly
    (up to fp32 summation order); ``average=True`` gives DDP semantics."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    grads = filter_grad_tensors(modules)
    if not grads:
        return None
    flat = torch.cat([g.reshape(-1) for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if average:
            flat.div_(dist.get_world_size(group))
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off: off + n].view_as(g))
            off += n

    if async_op:
        return work, finish
    finish()
    return None
