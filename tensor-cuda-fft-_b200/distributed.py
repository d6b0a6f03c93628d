"""Batch-sharded data parallelism for the spectral-mixing layer (one process per GPU, torch.distributed).

Every (batch, channel) column is an independent transform (reference: fft(x, dim=1), spectral_layers.py:88), so
the forward and dL/dx need no communication.  The only cross-batch coupling is the filter/bias gradient sum
(wirtinger_ops.py:77-80) -> ONE all-reduce(sum) per layer per step over [weight_real.grad | weight_imag.grad |
bias.grad].  The reference has no distributed code at all; this is the B200-native plumbing.
"""
from __future__ import annotations

import os
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


LAST_ALLREDUCE_PATH = "none"     # "symm_mem multimem" | "symm_mem one_shot" | "nccl" -- what the last call used (bench.py reports it)


class _SymmetricAllReduce:
    """Sum a flat fp32 gradient buffer across the ranks of one node through NVLink/NVSwitch symmetric memory
    (torch.distributed._symmetric_memory): the buffer is copied into a persistent symmetric allocation, reduced in the
    switch with multimem ld_reduce/st when the fabric supports multicast (NVLS) -- else by a one-shot peer-read kernel --
    and copied back.  For the 2-4 MB filter gradient this is latency-bound and about 3x faster than a ring/tree
    all-reduce launch.  Any failure while setting it up leaves ``ok`` False and the caller uses NCCL."""

    _cache = {}

    def __init__(self, numel: int, device: torch.device, group):
        self.ok = False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.group_name = group.group_name
            self.buf = symm_mem.empty(numel, dtype=torch.float32, device=device)
            self.hdl = symm_mem.rendezvous(self.buf, self.group_name)
            self.multicast = bool(getattr(self.hdl, "multicast_ptr", 0))
            self.ok = True
        except Exception as e:   # pragma: no cover - depends on the fabric
            self.err = repr(e)

    @classmethod
    def get(cls, numel: int, device: torch.device, group):
        key = (numel, device.index, id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(numel, device, group)
        return cls._cache[key]

    def __call__(self, flat: torch.Tensor) -> None:
        self.buf.copy_(flat)
        if self.multicast:
            torch.ops.symm_mem.multimem_all_reduce_(self.buf, "sum", self.group_name)
            flat.copy_(self.buf)
        else:
            flat.copy_(torch.ops.symm_mem.one_shot_all_reduce(self.buf, "sum", self.group_name))


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
    work = None
    symm = None
    # NCCL is the default.  The symmetric-memory path (SML_ALLREDUCE=symm) wins a back-to-back microbenchmark at 8 ranks
    # (tools/allreduce_check.py: 36 vs 64 us for the 2.4 MB cfg-2 gradient) but loses inside the real step, where the host
    # runs ahead of the GPU and NCCL's launch cost is hidden: 4 ranks, cfg-2: 0.457 ms per step with NCCL, 0.470 ms with the
    # copy-in / multimem reduce / copy-out sequence.
    mode = os.environ.get("SML_ALLREDUCE", "nccl")
    use_symm = mode == "symm"
    if flat.is_cuda and not async_op and flat.dtype == torch.float32 and use_symm:
        symm = _SymmetricAllReduce.get(flat.numel(), flat.device, group if group is not None else dist.group.WORLD)
    global LAST_ALLREDUCE_PATH
    done = False
    if symm is not None and symm.ok:
        try:
            symm(flat)                  # NVLink/NVSwitch symmetric-memory reduction (NVLS multimem when available)
            LAST_ALLREDUCE_PATH = "symm_mem multimem" if symm.multicast else "symm_mem one_shot"
            done = True
        except Exception as e:          # pragma: no cover - e.g. an op missing from this torch build: use NCCL from now on
            symm.ok, symm.err = False, repr(e)
    if not done:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        LAST_ALLREDUCE_PATH = "nccl"

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
