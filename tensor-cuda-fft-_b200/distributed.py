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


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: Optional[int] = None, local_rank: Optional[int] = None,
                          local_world_size: Optional[int] = None) -> dict:
    """Pin the calling process to the CPU cores of its GPU's NUMA node (sysfs: /sys/bus/pci/devices/<bdf>/numa_node), so that
    the pinned host buffers it allocates AFTERWARDS (first touch) and the threads that fill them sit next to the GPU's PCIe
    root -- what the host-buffer path (sml_fwd_bwd_host) needs when 8 ranks stream at once.  Where the platform exposes no
    node (virtualised hosts report -1) the available cores are split evenly over the local ranks instead.
    Returns a record of what was done (bench.py puts it into the e2e object)."""
    info = {"numa_node": None, "cpus": None, "how": "unchanged"}
    try:
        import torch
        idx = torch.cuda.current_device() if device_index is None else device_index
        pr = torch.cuda.get_device_properties(idx)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = -1
        try:
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        except Exception:
            pass
        avail = sorted(os.sched_getaffinity(0))
        cpus: List[int] = []
        if node >= 0:
            try:
                cpus = [c for c in _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) if c in avail]
            except Exception:
                cpus = []
            info["how"] = f"sysfs numa_node of {bdf}"
        if not cpus:
            lr = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else local_rank
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))) if local_world_size is None else local_world_size
            if lw > 1 and len(avail) >= lw:
                per = len(avail) // lw
                cpus = avail[lr * per: (lr + 1) * per]
                info["how"] = f"no NUMA node exposed for {bdf}: even split of {len(avail)} cores over {lw} local ranks"
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["numa_node"], info["cpus"] = node, f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"
    except Exception as e:   # never fatal: affinity is an optimisation
        info["how"] = f"failed: {e!r}"[:160]
    return info


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


class SymmetricGradBucket:
    """NVLink/NVSwitch symmetric-memory home of the [weight_real.grad | weight_imag.grad | bias.grad] blocks of a set of
    SpectralMixingLayer modules, back to back.

    ``fused=True`` (needs NVSwitch multicast): the batch-reduction kernel of every module's backward pushes its sums with
    ``multimem.red.add`` into the multicast alias of the bucket (sml_backward_allreduce) -- the switch adds them into every
    rank's copy -- so the cross-rank sum needs NO collective launch and no second pass over the data: one cross-rank barrier
    on the stream (``all_reduce()``) and every rank holds the global sum.  Two buffers alternate step by step; each backward
    clears its slice of the other one for the next step.
    ``fused=False``: each backward writes its local sums into its slice and ``all_reduce()`` runs one in-place multimem
    (or two-shot) all-reduce over the bucket.

    ``param.grad`` tensors are views of the bucket (DDP's gradient_as_bucket_view): consume them (optimizer step) before the
    next backward of the same parity overwrites them, and run exactly ONE backward per all-reduce -- micro-batch gradient
    accumulation across several backward calls needs plain (unattached) gradients."""

    def __init__(self, modules, group=None, allocator=None, fused: Optional[bool] = None):
        self.modules = [m for m in modules if getattr(m, "weight_real", None) is not None]
        if not self.modules:
            raise ValueError("no learnable SpectralMixingLayer modules given")
        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.sizes = [2 * m.weight_real.numel() + m.bias.numel() for m in self.modules]
        device = self.modules[0].weight_real.device
        self.numel = sum(self.sizes)
        self.multicast = False
        self.group_name = None
        self.parity = 0
        self.hdls = []
        if allocator is not None:                      # tests: any tensor factory
            self.bufs = [allocator(self.numel, device), allocator(self.numel, device)]
            self.fused = bool(fused)
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.group_name = self.group.group_name
            self.bufs = [symm_mem.empty(self.numel, dtype=torch.float32, device=device) for _ in range(2)]
            self.hdls = [symm_mem.rendezvous(b, self.group_name) for b in self.bufs]
            self.multicast = all(bool(getattr(h, "multicast_ptr", 0)) for h in self.hdls)
            self.fused = self.multicast if fused is None else (bool(fused) and self.multicast)
            for b in self.bufs:
                b.zero_()
            self.hdls[0].barrier(channel=0)            # nobody pushes into a buffer that a peer has not cleared yet
        self.buf = self.bufs[0]
        off = 0
        self.offsets = []
        for m, n in zip(self.modules, self.sizes):
            self.offsets.append(off)
            m._grad_bucket = (self, len(self.offsets) - 1)
            off += n

    # ---- what the autograd function asks for (spectral_layers._SpectralMixFn.backward) ----
    def slot(self, index: int):
        """(flat local slice to write / return views of, multicast address of that slice or 0, local slice of the other parity)."""
        off, n = self.offsets[index], self.sizes[index]
        cur, nxt = self.bufs[self.parity], self.bufs[self.parity ^ 1]
        mc = 0
        if self.fused and self.hdls:
            mc = int(self.hdls[self.parity].multicast_ptr) + off * 4
        return cur[off: off + n], mc, nxt[off: off + n]

    def covers(self, grads: List[torch.Tensor]) -> bool:
        """True if every gradient is a view into the current buffer of this bucket (one collective / barrier sums all of them)."""
        base = self.bufs[self.parity].untyped_storage().data_ptr()
        return all(g.is_contiguous() and g.untyped_storage().data_ptr() == base for g in grads)

    def all_reduce(self) -> str:
        if self.group_name is None:
            raise RuntimeError("bucket was built with a test allocator: no collective available")
        cur = self.parity
        if self.fused:
            self.hdls[cur].barrier(channel=0)          # every rank's pushes into this parity have landed
            self.parity ^= 1
            return "fused: multimem.red from the batch-reduction kernel + one barrier"
        if self.multicast:
            torch.ops.symm_mem.multimem_all_reduce_(self.bufs[cur], "sum", self.group_name)
            how = "symm_mem multimem in place"
        else:
            torch.ops.symm_mem.two_shot_all_reduce_(self.bufs[cur], "sum", self.group_name)
            how = "symm_mem two_shot in place"
        return how


_BUCKETS: List["SymmetricGradBucket"] = []


def attach_symmetric_grad_buffers(modules: Iterable[torch.nn.Module], group: Optional[dist.ProcessGroup] = None,
                                  allocator=None, fused: Optional[bool] = None) -> SymmetricGradBucket:
    """Give the modules' filter/bias gradients a home in NVLink symmetric memory (see SymmetricGradBucket).  Call once after the
    process group is up; ``allreduce_filter_grads`` then finishes the cross-rank sum (a barrier when fused, else one in-place
    collective)."""
    bucket = SymmetricGradBucket(list(modules), group, allocator, fused)
    _BUCKETS.append(bucket)
    return bucket


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
    global LAST_ALLREDUCE_PATH
    if not async_op:
        for bucket in _BUCKETS:      # gradients written straight into a symmetric-memory bucket: one in-place collective
            if bucket.group_name is not None and bucket.covers(grads):
                try:
                    cur = bucket.bufs[bucket.parity]
                    LAST_ALLREDUCE_PATH = bucket.all_reduce()
                    if average:
                        cur.div_(dist.get_world_size(group))
                    return None
                except Exception as e:   # pragma: no cover - an op missing from this torch build: fall through to NCCL
                    bucket.group_name, bucket.err = None, repr(e)
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
