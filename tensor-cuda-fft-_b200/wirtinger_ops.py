"""B200-native drop-in for ``fft_tensor.wirtinger_ops`` (reference: /root/reference/fft_tensor/wirtinger_ops.py).

``WirtingerGradient`` / ``WirtingerSpectralFilter`` operate on spectra the caller already holds; both the
multiply and its Wirtinger backward (conjugate multiply + batch reduction, :53-82) are CUDA kernels
(csrc/sml_wirtinger.cuh).  The fused layer in spectral_layers.py uses the same gradient formula inside its
backward kernel, so the two agree exactly as they do in the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _native


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_c64_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensor required; there is no CPU path")
    if t.dtype != torch.complex64:
        raise RuntimeError(f"{name}: complex64 required, got {t.dtype}")


class WirtingerGradient(Function):
    """out = x_freq * weight_complex with weight broadcast over dim 0 (wirtinger_ops.py:20-82)."""

    @staticmethod
    def forward(ctx, x_freq: torch.Tensor, weight_complex: torch.Tensor) -> torch.Tensor:
        _require_c64_cuda(x_freq, "x_freq")
        _require_c64_cuda(weight_complex, "weight_complex")
        B = x_freq.shape[0]
        N = x_freq[0].numel()
        if weight_complex.numel() != N or tuple(weight_complex.shape[-(x_freq.dim() - 1):]) != tuple(x_freq.shape[1:]):
            raise RuntimeError("weight_complex must have shape (1, *x_freq.shape[1:]) (it is reduced over dim 0 in backward)")
        xc = x_freq.contiguous()
        wc = weight_complex.contiguous()
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _native.check(_native.lib().sml_wirtinger_mul_forward(xc.data_ptr(), wc.data_ptr(), out.data_ptr(), B, N,
                                                                  _stream(xc.device)))
        ctx.save_for_backward(xc, wc)
        ctx.wshape = weight_complex.shape
        return out

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        xc, wc = ctx.saved_tensors
        B = xc.shape[0]
        N = xc[0].numel()
        g = grad_output.contiguous()
        gx = torch.empty_like(xc)
        gw = torch.empty(N, dtype=torch.complex64, device=xc.device)
        with torch.cuda.device(xc.device):
            _native.check(_native.lib().sml_wirtinger_mul_backward(g.data_ptr(), xc.data_ptr(), wc.data_ptr(),
                                                                   gx.data_ptr(), gw.data_ptr(), B, N, _stream(xc.device)))
        # reference returns sum(dim=0, keepdim=True): shape (1, *x.shape[1:]); autograd needs the weight's own shape
        return gx, gw.view(ctx.wshape)


class ComplexParameter(nn.Module):
    """Real/imag ``nn.Parameter`` pair, same initialisers as wirtinger_ops.py:85-142."""

    def __init__(self, shape: tuple, init_mode: str = "xavier"):
        super().__init__()
        if init_mode == "xavier":
            bound = np.sqrt(3.0 / (shape[0] + shape[1])) if len(shape) == 2 else np.sqrt(3.0 / shape[0])
            self.real = nn.Parameter(torch.empty(shape).uniform_(-bound, bound))
            self.imag = nn.Parameter(torch.empty(shape).uniform_(-bound, bound))
        elif init_mode == "kaiming":
            std = np.sqrt(2.0 / shape[0])
            self.real = nn.Parameter(torch.randn(shape) * std)
            self.imag = nn.Parameter(torch.randn(shape) * std)
        elif init_mode == "uniform":
            self.real = nn.Parameter(torch.empty(shape).uniform_(-1, 1))
            self.imag = nn.Parameter(torch.empty(shape).uniform_(-1, 1))
            with torch.no_grad():
                mag = torch.sqrt(self.real ** 2 + self.imag ** 2)
                self.real /= mag
                self.imag /= mag
        elif init_mode == "ones":
            self.real = nn.Parameter(torch.ones(shape))
            self.imag = nn.Parameter(torch.zeros(shape))
        else:
            raise ValueError(f"Unknown init_mode: {init_mode}")

    def forward(self) -> torch.Tensor:
        return torch.complex(self.real, self.imag)

    def magnitude(self) -> torch.Tensor:
        return torch.sqrt(self.real ** 2 + self.imag ** 2)

    def phase(self) -> torch.Tensor:
        return torch.atan2(self.imag, self.real)


class _WirtingerFilterFn(Function):
    """Fused low-pass filter of a (B,T,D) spectrum: slice, multiply, zero-scatter in one pass each way."""

    @staticmethod
    def forward(ctx, x_freq, w_re, w_im):
        _require_c64_cuda(x_freq, "x_freq")
        B, T, D = x_freq.shape
        Fn = w_re.shape[1]
        xc = x_freq.contiguous()
        wr = w_re.detach().contiguous().float()
        wi = w_im.detach().contiguous().float()
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _native.check(_native.lib().sml_wirtinger_filter_forward(xc.data_ptr(), wr.data_ptr(), wi.data_ptr(),
                                                                     out.data_ptr(), B, T, D, Fn, _stream(xc.device)))
        ctx.save_for_backward(xc, wr, wi)
        return out

    @staticmethod
    def backward(ctx, g):
        xc, wr, wi = ctx.saved_tensors
        B, T, D = xc.shape
        Fn = wr.shape[1]
        gc = g.contiguous()
        gx = torch.empty_like(xc)
        gwr = torch.empty_like(wr)
        gwi = torch.empty_like(wi)
        with torch.cuda.device(xc.device):
            _native.check(_native.lib().sml_wirtinger_filter_backward(gc.data_ptr(), xc.data_ptr(), wr.data_ptr(),
                                                                      wi.data_ptr(), gx.data_ptr(), gwr.data_ptr(),
                                                                      gwi.data_ptr(), B, T, D, Fn, _stream(xc.device)))
        return gx, gwr, gwi


class WirtingerSpectralFilter(nn.Module):
    """Low-pass complex filter on the first k = min(num_frequencies, T//2) bins of a (B,T,D) spectrum
    (wirtinger_ops.py:145-203); bins >= k are zeroed."""

    def __init__(self, num_channels: int, num_frequencies: int):
        super().__init__()
        self.num_channels = num_channels
        self.num_frequencies = num_frequencies
        self.weight = ComplexParameter(shape=(num_channels, num_frequencies), init_mode="ones")

    def forward(self, x_freq: torch.Tensor) -> torch.Tensor:
        B, T, D = x_freq.shape
        assert D == self.num_channels
        return _WirtingerFilterFn.apply(x_freq, self.weight.real, self.weight.imag)
