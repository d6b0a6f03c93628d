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


def filter_grad_params(modules: Iterable[torch.nn.Module]) -> List[torch.nn.Parameter]:
    out = []
    for m in modules:
        for name in ("weight_real", "weight_imag", "bias"):
            p = getattr(m, name, None)
            if p is not None and p.grad is not None:
                out.append(p)
    return out


def filter_grad_tensors(modules: Iterable[torch.nn.Module]) -> List[torch.Tensor]:
    return [p.grad for p in filter_grad_params(modules)]


def _flat_view(grads: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """If the gradients are contiguous, back-to-back slices of one storage (the layout sml_backward writes), return a
    1-D view covering all of them (no copy); else None."""
    g0 = grads[0]
    try:
        base = g0.untyped_storage().data_ptr()
        off = g0.storage_offset()
        for g in grads:
            if (g.dtype != g0.dtype or g.device != g0.device or not g.is_contiguous()
                    or g.untyped_storage().data_ptr() != base or g.storage_offset() != off):
                return None
            off += g.numel()
        return torch.as_strided(g0, (off - g0.storage_offset(),), (1,), g0.storage_offset())
    except Exception:
        return None


def allreduce_filter_grads(modules: Iterable[torch.nn.Module], group: Optional[dist.ProcessGroup] = None,
                           average: bool = False, async_op: bool = False, as_views: bool = True):
    """Sum (or average) the filter/bias gradients of the given SpectralMixingLayer modules across ranks with a
    single flat all-reduce.  ``sum`` reproduces a single-process run on the concatenated batch exactly
    (up to fp32 summation order); ``average=True`` gives DDP semantics.

    ``as_views=True`` (default) re-points each ``param.grad`` at its slice of the reduced flat buffer (DDP's
    ``gradient_as_bucket_view``), so the step costs one concatenation and one collective and no copy back;
    ``as_views=False`` copies the reduced values into the existing ``.grad`` tensors instead."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    params = filter_grad_params(modules)
    if not params:
        return None
    grads = [p.grad for p in params]
    flat = _flat_view(grads)
    in_place = flat is not None       # the fused backward hands out views of ONE flat [gw_re | gw_im | gb] buffer
    if flat is None:
        flat = torch.cat([g.reshape(-1) for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if average:
            flat.div_(dist.get_world_size(group))
        if in_place:
            return
        off = 0
        for p, g in zip(params, grads):
            n = g.numel()
            piece = flat[off: off + n].view_as(g)
            if as_views:
                p.grad = piece
            else:
                g.copy_(piece)
            off += n

    if async_op:
        return work, finish
    finish()
    return None
