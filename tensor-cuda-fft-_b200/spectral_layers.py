"""B200-native drop-in for ``fft_tensor.spectral_layers`` (reference: /root/reference/fft_tensor/spectral_layers.py).

Same classes, constructor signatures, parameter names/shapes and autograd behaviour; the forward and backward
of ``SpectralMixingLayer`` run as ONE fused sm_100a kernel each (csrc/sml_fast.cuh) behind the C ABI in
include/spectral_mix_b200.h.  CUDA only: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native

_IO_DTYPES = {torch.float32: _native.DTYPE_F32, torch.bfloat16: _native.DTYPE_BF16}


_SHAPE_CACHE = {}


def _shape_info(B: int, T: int, D: int, Fn: int, io: int):
    """(fast_path, xlow_bytes, workspace_bytes) of a problem, cached: three ctypes round trips per call otherwise."""
    key = (B, T, D, Fn, io)
    info = _SHAPE_CACHE.get(key)
    if info is None:
        lib = _native.lib()
        info = (_native.plan(B, T, D, Fn, io)["path"] == "fast", int(lib.sml_xlow_bytes(B, T, D, Fn)),
                int(lib.sml_workspace_bytes(B, T, D, Fn, io)))
        if len(_SHAPE_CACHE) < 4096:
            _SHAPE_CACHE[key] = info
    return info


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream_handle(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """``torch.cuda.device(dev)`` only when dev is not already current (the context manager costs ~15 us per use,
    which is most of a forward call at small shapes)."""

    def __init__(self, device):
        self.ctx = None if device.index is None or device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.contiguous().float()


class _SpectralMixFn(torch.autograd.Function):
    """y = Re(ifft(lowpass(fft(x) * W))) + bias, with the backward of spectral_layers.py:88-116
    (== WirtingerGradient.backward, wirtinger_ops.py:53-82) computed by sml_backward."""

    @staticmethod
    def forward(ctx, x, w_re, w_im, bias, grad_bucket=None):
        if not x.is_cuda:
            raise RuntimeError("SpectralMixingLayer (B200 build) needs a CUDA tensor; there is no CPU path")
        if x.dtype not in _IO_DTYPES:
            raise RuntimeError(f"Unsupported dtype {x.dtype}")   # reference raises on bf16; here fp32 and bf16 are I/O types
        if x.dim() != 3:
            raise RuntimeError(f"expected (B, T, D) input, got shape {tuple(x.shape)}")
        if w_re.device != x.device or w_im.device != x.device or (bias is not None and bias.device != x.device):
            raise RuntimeError(f"input is on {x.device} but the filter parameters are on {w_re.device}")
        if w_re.shape != w_im.shape or w_re.dim() != 2 or w_re.shape[0] != x.shape[2]:
            raise RuntimeError(f"filter shape {tuple(w_re.shape)} / {tuple(w_im.shape)} does not match embed dim {x.shape[2]}")
        B, T, D = x.shape
        Fn = w_re.shape[1]
        io = _IO_DTYPES[x.dtype]
        lib = _native.lib()
        xc = x.contiguous()
        if xc.data_ptr() % 16:      # a contiguous view at an odd storage offset: the fused kernel's TMA needs 16-byte alignment
            xc = xc.clone()
        wr, wi = _f32c(w_re), _f32c(w_im)
        bs = None if bias is None else _f32c(bias)
        y = torch.empty_like(xc)
        need_filter_grad = any(ctx.needs_input_grad[1:])
        fast_path, nbytes, _ = _shape_info(B, T, D, Fn, io)
        xlow = None
        if need_filter_grad or not fast_path:
            xlow = torch.empty(max(nbytes // 8, 1), dtype=torch.complex64, device=x.device)
        with _on_device(x.device):
            _native.check(lib.sml_forward(_ptr(xc), _ptr(wr), _ptr(wi), _ptr(bs), _ptr(y), _ptr(xlow),
                                          B, T, D, Fn, io, _stream_handle(x.device)))
        ctx.save_for_backward(wr, wi, xlow if need_filter_grad else None)
        ctx.shape = (B, T, D, Fn, io)
        ctx.has_bias = bias is not None
        ctx.fast_path = fast_path
        ctx.param_dtypes = (w_re.dtype, w_im.dtype, None if bias is None else bias.dtype)
        ctx.grad_bucket = grad_bucket      # (distributed.SymmetricGradBucket, slot index) or None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable      # double backward is not implemented (the reference never uses it): fail clearly
    def backward(ctx, g):
        wr, wi, xlow = ctx.saved_tensors
        B, T, D, Fn, io = ctx.shape
        lib = _native.lib()
        gc = g.contiguous()
        if gc.dtype != (torch.float32 if io == _native.DTYPE_F32 else torch.bfloat16):
            gc = gc.to(torch.float32 if io == _native.DTYPE_F32 else torch.bfloat16)
        if gc.data_ptr() % 16:
            gc = gc.clone()
        gx = torch.empty_like(gc)
        want = xlow is not None
        flat = None
        gwr = gwi = gb = None
        if want:
            # one flat buffer [gw_re | gw_im | gb] so a data-parallel job can all-reduce it in a single call.  A caller-provided
            # buffer (distributed.attach_symmetric_grad_buffers: a slice of an NVLink symmetric-memory bucket) makes the
            # reduction kernel's own store the collective's input: no copy in, no copy out.
            mc_ptr, flat_next = 0, None
            if ctx.grad_bucket is not None:
                cand, mc_ptr, flat_next = ctx.grad_bucket[0].slot(ctx.grad_bucket[1])
                if cand.device == gc.device and cand.numel() == 2 * D * Fn + D and cand.dtype == torch.float32:
                    flat = cand
                else:
                    mc_ptr = 0
            if flat is None:
                flat = torch.empty(2 * D * Fn + D, dtype=torch.float32, device=gc.device)
            gwr = flat[: D * Fn].view(D, Fn)
            gwi = flat[D * Fn: 2 * D * Fn].view(D, Fn)
            gb = flat[2 * D * Fn:]
        # fast path: scratch only for the filter-gradient terms; generic path: always (low-band spectrum of g)
        ws_bytes = _shape_info(B, T, D, Fn, io)[2] if (want or not ctx.fast_path) else 0
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gc.device) if ws_bytes else None
        with _on_device(gc.device):
            if want and mc_ptr and ctx.fast_path:
                # the batch-reduction kernel pushes its sums through the NVSwitch multicast alias of the bucket: its store is the
                # all-reduce (distributed.SymmetricGradBucket, fused mode); the barrier follows in allreduce_filter_grads
                _native.check(lib.sml_backward_allreduce(_ptr(gc), _ptr(xlow), _ptr(wr), _ptr(wi), _ptr(gx), _ptr(gwr), _ptr(gwi),
                                                         _ptr(gb), _ptr(ws), ws_bytes, B, T, D, Fn, io, mc_ptr, _ptr(flat_next),
                                                         _stream_handle(gc.device)))
            else:
                _native.check(lib.sml_backward(_ptr(gc), _ptr(xlow), _ptr(wr), _ptr(wi), _ptr(gx), _ptr(gwr), _ptr(gwi),
                                               _ptr(gb), _ptr(ws), ws_bytes, B, T, D, Fn, io, _stream_handle(gc.device)))
                if want and mc_ptr:
                    # a shape the fused kernels do not take, inside a fused bucket (whose all_reduce() is only a barrier): sum this
                    # module's block across ranks right here
                    import torch.distributed as dist
                    dist.all_reduce(flat, group=ctx.grad_bucket[0].group)
        need = ctx.needs_input_grad
        dt = ctx.param_dtypes
        return (gx if need[0] else None,
                gwr.to(dt[0]) if (want and need[1]) else None,
                gwi.to(dt[1]) if (want and need[2]) else None,
                gb.to(dt[2]) if (want and ctx.has_bias and need[3]) else None,
                None)


_EXT_CACHE = {}


def _ext_supported(B: int, T: int, D: int, Fn: int, io: int, T_in: int = 0, T_out: int = 0, nyq: bool = False) -> bool:
    """True if the extended fused kernels (sml_forward_ext) take this problem; cached per shape."""
    key = (B, T, D, Fn, io, T_in, T_out, nyq)
    ok = _EXT_CACHE.get(key)
    if ok is None:
        import ctypes
        ext = _native.make_ext(T_in=T_in, T_out=T_out)
        if nyq:
            ext.w_nyq = 1      # only tested for NULL by sml_ext_supported
        ok = _native.lib().sml_ext_supported(B, T, D, Fn, io, ctypes.byref(ext)) == 0
        if len(_EXT_CACHE) < 4096:
            _EXT_CACHE[key] = ok
    return ok


class _LNSpectralResidualFn(torch.autograd.Function):
    """``x + spectral_mix(LayerNorm(x))`` -- the way SpectralMLPBlock calls the layer (spectral_layers.py:161, :185) -- as
    three launches instead of seven HBM passes: a row-statistics pre-pass (sml_ln_stats), ONE fused kernel that normalises the
    rows while it loads them and adds the residual rows while it stores (sml_forward_ext), and in the backward the fused
    kernel (sml_backward) followed by one LayerNorm-backward + skip-connection kernel (sml_ln_backward).

    The LayerNorm affine part never reaches the kernel: the layer is linear in its input, so gamma scales the filter rows and
    beta is a DC term, ``w_eff = gamma[:, None] * W``, ``bias_eff = bias + beta * W_re[:, 0]`` -- computed by the caller with
    ordinary autograd ops, so the gradients of gamma / beta / W / bias follow from the gradients returned here."""

    @staticmethod
    def forward(ctx, x, w_re, w_im, bias, eps):
        import ctypes
        B, T, D = x.shape
        Fn = w_re.shape[1]
        io = _IO_DTYPES[x.dtype]
        lib = _native.lib()
        xc = x.contiguous()
        if xc.data_ptr() % 16:
            xc = xc.clone()
        wr, wi, bs = _f32c(w_re), _f32c(w_im), _f32c(bias)
        stats = torch.empty(B, T, 2, dtype=torch.float32, device=x.device)
        y = torch.empty_like(xc)
        need_filter_grad = any(ctx.needs_input_grad[1:4])
        xlow = None
        if need_filter_grad:
            xlow = torch.empty(max(_shape_info(B, T, D, Fn, io)[1] // 8, 1), dtype=torch.complex64, device=x.device)
        ext = _native.make_ext(row_stats=stats, residual=xc)
        with _on_device(x.device):
            st = _stream_handle(x.device)
            _native.check(lib.sml_ln_stats(_ptr(xc), _ptr(stats), B, T, T, 0, D, float(eps), io, st))
            _native.check(lib.sml_forward_ext(_ptr(xc), _ptr(wr), _ptr(wi), _ptr(bs), _ptr(y), _ptr(xlow), B, T, D, Fn, io,
                                              ctypes.byref(ext), st))
        ctx.save_for_backward(xc, stats, wr, wi, xlow)
        ctx.shape = (B, T, D, Fn, io)
        ctx.param_dtypes = (w_re.dtype, w_im.dtype, bias.dtype)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        xc, stats, wr, wi, xlow = ctx.saved_tensors
        B, T, D, Fn, io = ctx.shape
        lib = _native.lib()
        gc = g.contiguous()
        if gc.dtype != xc.dtype:
            gc = gc.to(xc.dtype)
        if gc.data_ptr() % 16:
            gc = gc.clone()
        gh = torch.empty_like(gc)      # dL/dx^ (gradient with respect to the normalised rows)
        want = xlow is not None
        gwr = gwi = gb = ws = None
        ws_bytes = 0
        if want:
            flat = torch.empty(2 * D * Fn + D, dtype=torch.float32, device=gc.device)
            gwr, gwi, gb = flat[: D * Fn].view(D, Fn), flat[D * Fn: 2 * D * Fn].view(D, Fn), flat[2 * D * Fn:]
            ws_bytes = _shape_info(B, T, D, Fn, io)[2]
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gc.device)
        gx = torch.empty_like(gc)
        with _on_device(gc.device):
            st = _stream_handle(gc.device)
            _native.check(lib.sml_backward(_ptr(gc), _ptr(xlow), _ptr(wr), _ptr(wi), _ptr(gh), _ptr(gwr), _ptr(gwi), _ptr(gb),
                                           _ptr(ws), ws_bytes, B, T, D, Fn, io, st))
            _native.check(lib.sml_ln_backward(_ptr(gh), _ptr(xc), _ptr(stats), _ptr(gc), None, _ptr(gx), B, T, T, 0, D, io, st))
        need = ctx.needs_input_grad
        dt = ctx.param_dtypes
        return (gx if need[0] else None,
                gwr.to(dt[0]) if (want and need[1]) else None,
                gwi.to(dt[1]) if (want and need[2]) else None,
                gb.to(dt[2]) if (want and need[3]) else None,
                None)


class _SpectralResidualFn(torch.autograd.Function):
    """``x + spectral_mix(x)`` (HybridSpectralAttention feeds ``x + global_context`` to its norm, spectral_layers.py:243-248): the
    skip connection is added while the fused kernel stores its rows (sml_forward_ext, residual = x) instead of by a separate
    three-stream add; the backward is the plain fused backward plus the skip connection's gradient."""

    @staticmethod
    def forward(ctx, x, w_re, w_im, bias):
        import ctypes
        B, T, D = x.shape
        Fn = w_re.shape[1]
        io = _IO_DTYPES[x.dtype]
        lib = _native.lib()
        xc = x.contiguous()
        if xc.data_ptr() % 16:
            xc = xc.clone()
        wr, wi, bs = _f32c(w_re), _f32c(w_im), _f32c(bias)
        y = torch.empty_like(xc)
        need_filter_grad = any(ctx.needs_input_grad[1:])
        xlow = torch.empty(max(_shape_info(B, T, D, Fn, io)[1] // 8, 1), dtype=torch.complex64, device=x.device) if need_filter_grad else None
        ext = _native.make_ext(residual=xc)
        with _on_device(x.device):
            _native.check(lib.sml_forward_ext(_ptr(xc), _ptr(wr), _ptr(wi), _ptr(bs), _ptr(y), _ptr(xlow), B, T, D, Fn, io,
                                              ctypes.byref(ext), _stream_handle(x.device)))
        ctx.save_for_backward(wr, wi, xlow)
        ctx.shape = (B, T, D, Fn, io)
        ctx.dtypes = (x.dtype, w_re.dtype, w_im.dtype, bias.dtype)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        wr, wi, xlow = ctx.saved_tensors
        B, T, D, Fn, io = ctx.shape
        lib = _native.lib()
        gc = g.contiguous().to(ctx.dtypes[0])
        if gc.data_ptr() % 16:
            gc = gc.clone()
        gx = torch.empty_like(gc)
        want = xlow is not None
        gwr = gwi = gb = ws = None
        ws_bytes = 0
        if want:
            flat = torch.empty(2 * D * Fn + D, dtype=torch.float32, device=gc.device)
            gwr, gwi, gb = flat[: D * Fn].view(D, Fn), flat[D * Fn: 2 * D * Fn].view(D, Fn), flat[2 * D * Fn:]
            ws_bytes = _shape_info(B, T, D, Fn, io)[2]
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gc.device)
        with _on_device(gc.device):
            _native.check(lib.sml_backward(_ptr(gc), _ptr(xlow), _ptr(wr), _ptr(wi), _ptr(gx), _ptr(gwr), _ptr(gwi), _ptr(gb), _ptr(ws),
                                           ws_bytes, B, T, D, Fn, io, _stream_handle(gc.device)))
        need = ctx.needs_input_grad
        dt = ctx.dtypes
        return ((gx + gc) if need[0] else None,
                gwr.to(dt[1]) if (want and need[1]) else None,
                gwi.to(dt[2]) if (want and need[2]) else None,
                gb.to(dt[3]) if (want and need[3]) else None)


def spectral_mix_residual(x: torch.Tensor, layer: "SpectralMixingLayer") -> torch.Tensor:
    """``x + layer(x)`` with the skip connection fused into the kernel's store (callers test ``fused_residual_supported``)."""
    return _SpectralResidualFn.apply(x, layer.weight_real, layer.weight_imag, layer.bias)


def fused_residual_supported(x: torch.Tensor, layer: "SpectralMixingLayer") -> bool:
    if not (x.is_cuda and x.dim() == 3 and x.dtype in _IO_DTYPES and x.numel() > 0):
        return False
    if not (layer.learnable and layer.weight_real is not None) or (layer.training and layer.dropout.p > 0.0):
        return False
    if getattr(layer, "_grad_bucket", None) is not None:
        return False
    B, T, D = x.shape
    return _ext_supported(B, T, D, layer.weight_real.shape[1], _IO_DTYPES[x.dtype])


def ln_spectral_mix_residual(x: torch.Tensor, norm: nn.LayerNorm, layer: "SpectralMixingLayer") -> torch.Tensor:
    """``x + layer(norm(x))`` through the fused LayerNorm-on-load / residual-on-store kernel; raises if the shape is not
    one the extended kernels take (callers test ``fused_block_supported`` first)."""
    gamma = norm.weight if norm.weight is not None else torch.ones(x.shape[-1], device=x.device)
    w_re = gamma.float()[:, None] * layer.weight_real.float()
    w_im = gamma.float()[:, None] * layer.weight_imag.float()
    bias = layer.bias.float()
    if norm.bias is not None:
        bias = bias + norm.bias.float() * layer.weight_real.float()[:, 0]
    return _LNSpectralResidualFn.apply(x, w_re, w_im, bias, norm.eps)


def fused_block_supported(x: torch.Tensor, norm: nn.LayerNorm, layer: "SpectralMixingLayer") -> bool:
    if not (x.is_cuda and x.dim() == 3 and x.dtype in _IO_DTYPES and x.numel() > 0):
        return False
    if not (layer.learnable and layer.weight_real is not None) or tuple(norm.normalized_shape) != (x.shape[-1],):
        return False
    if layer.training and layer.dropout.p > 0.0:      # dropout sits between the layer and the skip connection (:118, :185)
        return False
    if getattr(layer, "_grad_bucket", None) is not None:
        # data-parallel job with the fused reduce + all-reduce (distributed.attach_symmetric_grad_buffers): the filter gradients must
        # leave through the layer's own backward (sml_backward_allreduce), which the block-level function does not call
        return False
    B, T, D = x.shape
    return _ext_supported(B, T, D, layer.weight_real.shape[1], _IO_DTYPES[x.dtype])


def spectral_mix(x: torch.Tensor, weight_real: torch.Tensor, weight_imag: torch.Tensor,
                 bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Functional form of the fused layer (no dropout)."""
    return _SpectralMixFn.apply(x, weight_real, weight_imag, bias)


def spectral_mix_fwd_bwd_host(x: torch.Tensor, g: torch.Tensor, weight_real: torch.Tensor, weight_imag: torch.Tensor,
                              bias: Optional[torch.Tensor] = None, chunk_batch: int = 0, out=None, device=None,
                              filter_grads: bool = True):
    """Forward + backward of the layer over HOST tensors through ``sml_fwd_bwd_host`` (include/spectral_mix_b200.h).

    ``x`` and ``g`` are CPU tensors (B, T, D), fp32 or bf16 -- pinned memory lets the chunked host->device copies,
    the kernels and the device->host copies overlap.  Returns CPU tensors ``(y, gx, gw_re, gw_im, gb)`` (the last
    three are None with ``filter_grads=False``); ``out=(y, gx)`` reuses caller-provided (pinned) output buffers.
    Same math as ``SpectralMixingLayer.forward`` + autograd backward (spectral_layers.py:88-116 of the reference)."""
    if x.is_cuda or g.is_cuda:
        raise RuntimeError("spectral_mix_fwd_bwd_host takes CPU tensors; use SpectralMixingLayer for device tensors")
    if x.dtype not in _IO_DTYPES or g.dtype != x.dtype or x.shape != g.shape or x.dim() != 3:
        raise RuntimeError(f"expected matching (B, T, D) fp32/bf16 tensors, got {tuple(x.shape)} {x.dtype} / {tuple(g.shape)} {g.dtype}")
    B, T, D = x.shape
    Fn = weight_real.shape[1]
    x, g = x.contiguous(), g.contiguous()
    wr = weight_real.detach().float().contiguous().cpu()
    wi = weight_imag.detach().float().contiguous().cpu()
    bs = None if bias is None else bias.detach().float().contiguous().cpu()
    if out is None:
        y = torch.empty_like(x, pin_memory=x.is_pinned())
        gx = torch.empty_like(x, pin_memory=x.is_pinned())
    else:
        y, gx = out
    gwr = gwi = gb = None
    if filter_grads:
        gwr, gwi, gb = torch.empty(D, Fn), torch.empty(D, Fn), torch.empty(D)
    lib = _native.lib()
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        _native.check(lib.sml_fwd_bwd_host(_ptr(x), _ptr(g), _ptr(wr), _ptr(wi), _ptr(bs), _ptr(y), _ptr(gx), _ptr(gwr),
                                           _ptr(gwi), _ptr(gb), B, T, D, Fn, _IO_DTYPES[x.dtype], int(chunk_batch)))
    return y, gx, gwr, gwi, gb


class SpectralMixingLayer(nn.Module):
    """Drop-in for ``fft_tensor.spectral_layers.SpectralMixingLayer`` (spectral_layers.py:19-132).

    FFT along the sequence axis, learnable complex low-pass filter on the first
    ``k = min(num_filters, T // 2)`` bins (everything else is zeroed), inverse FFT, real part, + bias, dropout.
    Parameters: ``weight_real``/``weight_imag`` of shape (embed_dim, num_filters), ``bias`` (embed_dim,) -- the
    reference's state_dict loads unchanged.
    """

    def __init__(self, embed_dim: int, num_filters: Optional[int] = None, dropout: float = 0.0,
                 learnable: bool = True):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_filters = num_filters or (embed_dim // 2)
        self.learnable = learnable
        if learnable:
            self.weight_real = nn.Parameter(torch.ones(embed_dim, self.num_filters))
            self.weight_imag = nn.Parameter(torch.zeros(embed_dim, self.num_filters))
            self.bias = nn.Parameter(torch.zeros(embed_dim))
        else:
            self.register_parameter("weight_real", None)
            self.register_parameter("weight_imag", None)
            self.register_parameter("bias", None)
        self.dropout = nn.Dropout(dropout)
        self._verify_gradients = True

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, T, D = x.shape
        assert D == self.embed_dim, f"Expected embed_dim={self.embed_dim}, got {D}"
        if self.learnable and self.weight_real is not None:
            if x.numel() == 0:      # empty batch or sequence: nothing to transform (the reference returns an empty tensor too)
                if not x.is_cuda:
                    raise RuntimeError("SpectralMixingLayer (B200 build) needs a CUDA tensor; there is no CPU path")
                y = x + 0.0 * (self.weight_real.sum() + self.weight_imag.sum() + self.bias.sum()).to(x.dtype)
            else:
                y = _SpectralMixFn.apply(x, self.weight_real, self.weight_imag, self.bias, getattr(self, "_grad_bucket", None))
        else:
            # learnable=False is fft followed by ifft(.).real (spectral_layers.py:88, :112): the identity up to
            # rounding (reference self-test :301-309 measures 1.2e-7); returned exactly, as a new tensor.
            if not x.is_cuda:
                raise RuntimeError("SpectralMixingLayer (B200 build) needs a CUDA tensor; there is no CPU path")
            y = x.clone()
        if self.dropout.p == 0.0 or not self.training:     # nn.Dropout is the identity here: skip the call
            return y
        return self.dropout(y)

    def graphed(self, sample: torch.Tensor, num_warmup_iters: int = 3):
        """A CUDA-graph-captured callable of this layer for inputs of ``sample``'s shape and dtype (forward AND backward are
        replayed as graphs: no per-call Python, ctypes or launch work on the host -- what small, launch-bound shapes such as
        BASELINE configs[0] (8, 512, 256) need).  The warm-up iterations build the per-(device, T) tables, which must exist
        before capture.  Usual CUDA-graph rules: fixed shape, dropout inactive, outputs live in graph-owned memory."""
        if not sample.is_cuda:
            raise RuntimeError("SpectralMixingLayer (B200 build) needs a CUDA tensor; there is no CPU path")
        if self.training and self.dropout.p > 0.0:
            raise RuntimeError("graphed(): dropout draws new random numbers per call; use eval() or dropout=0")
        outer = self

        class _Graphed(nn.Module):      # make_graphed_callables replaces the forward of the module it is given: give it a wrapper,
            def __init__(self):         # so that this layer itself stays callable (eagerly, and inside other captures)
                super().__init__()
                self.layer = outer

            def forward(self, x):
                return self.layer(x)

        return torch.cuda.make_graphed_callables(_Graphed(), (sample.detach().clone().requires_grad_(True),),
                                                 num_warmup_iters=num_warmup_iters)

    def graphed_step(self, x: torch.Tensor, g: torch.Tensor, num_warmup_iters: int = 3):
        """ONE CUDA graph for a whole forward + backward of this layer on static buffers (PyTorch's "whole network capture"
        pattern): ``replay, bufs = layer.graphed_step(x, g)``; fill ``bufs["x"]`` / ``bufs["g"]`` with ``copy_``, call ``replay()``
        and read ``bufs["y"]``, ``bufs["gx"]`` and the parameters' ``.grad`` (overwritten by every replay, not accumulated).
        Nothing runs on the host per step but one graph launch: small, launch-bound shapes such as BASELINE configs[0]
        (8, 512, 256) drop from about 0.17 ms per step (eager autograd) to the sum of the kernels."""
        if not x.is_cuda:
            raise RuntimeError("SpectralMixingLayer (B200 build) needs a CUDA tensor; there is no CPU path")
        if self.training and self.dropout.p > 0.0:
            raise RuntimeError("graphed_step(): dropout draws new random numbers per call; use eval() or dropout=0")
        sx = x.detach().clone().requires_grad_(True)
        sg = g.detach().clone()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            for _ in range(num_warmup_iters):       # builds the per-(device, T) tables and tunes the plan outside the capture
                self.zero_grad(set_to_none=True)
                sx.grad = None
                self(sx).backward(sg)
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.zero_grad(set_to_none=True)
        sx.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sy = self(sx)
            sy.backward(sg)
        bufs = {"x": sx.detach(), "g": sg, "y": sy.detach(), "gx": sx.grad}      # x: an alias that can be written with copy_
        return graph.replay, bufs

    def verify_energy_preservation(self, x: torch.Tensor, y: torch.Tensor) -> float:
        """sum(y^2) / (sum(x^2) + 1e-8), spectral_layers.py:122-132."""
        energy_in = torch.sum(x.float() ** 2).item()
        energy_out = torch.sum(y.float() ** 2).item()
        return energy_out / (energy_in + 1e-8)


class SpectralMLPBlock(nn.Module):
    """x + spectral_mix(norm1(x)); x + mlp(norm2(x))  -- spectral_layers.py:135-190 (the layer's main caller)."""

    def __init__(self, embed_dim: int, mlp_ratio: int = 4, dropout: float = 0.1):
        super().__init__()
        self.spectral_mix = SpectralMixingLayer(embed_dim=embed_dim, dropout=dropout, learnable=True)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        hidden_dim = embed_dim * mlp_ratio
        self.mlp = nn.Sequential(
            nn.Linear(embed_dim, hidden_dim),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, embed_dim),
            nn.Dropout(dropout),
        )
        self.fuse_norm_residual = True      # not a parameter / buffer: the reference's state_dict loads unchanged

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # spectral half: one fused kernel (LayerNorm on load, residual on store) where the extended kernels take the
        # shape and dropout is inactive; otherwise the reference's composition around the fused layer
        if self.fuse_norm_residual and fused_block_supported(x, self.norm1, self.spectral_mix):
            x = ln_spectral_mix_residual(x, self.norm1, self.spectral_mix)
        else:
            x = x + self.spectral_mix(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class HybridSpectralAttention(nn.Module):
    """Spectral mixing + (full) softmax attention over the mixed stream -- spectral_layers.py:193-256."""

    def __init__(self, embed_dim: int, num_heads: int = 8, window_size: int = 64, dropout: float = 0.1):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.spectral = SpectralMixingLayer(embed_dim, dropout=dropout)
        self.qkv = nn.Linear(embed_dim, 3 * embed_dim)
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.dropout = nn.Dropout(dropout)
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, T, D = x.shape
        H = self.num_heads
        # x + global_context (spectral_layers.py:243-248): the skip connection rides on the fused kernel's store where the extended
        # kernels take the shape and dropout is inactive
        mixed = spectral_mix_residual(x, self.spectral) if fused_residual_supported(x, self.spectral) else x + self.spectral(x)
        qkv = self.qkv(self.norm(mixed)).reshape(B, T, 3, H, D // H).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = F.softmax((q @ k.transpose(-2, -1)) / math.sqrt(D // H), dim=-1)
        attn = self.dropout(attn)
        out = (attn @ v).transpose(1, 2).reshape(B, T, D)
        out = self.dropout(self.proj(out))
        return x + out
