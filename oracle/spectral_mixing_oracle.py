"""CPU oracle for the spectral-mixing hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Two independent restatements of the reference algorithm:

* ``torch_port_*``  -- follows /root/reference/fft_tensor/spectral_layers.py:83-118 step by step with the
  same ``torch.fft`` calls (full complex fft along dim 1, low-pass complex filter on the first
  k = min(num_filters, T//2) bins, zero everything else, ifft, take .real, + bias) and lets autograd
  produce the backward, exactly as the reference does.  This is the "port" timed as the CPU baseline.
* ``closed_form_f64`` -- the algebraic closed form in numpy float64; shares no code
  with the port and is the high-precision yardstick for error measurements.

The arithmetic itself lives in third-party PyTorch (torch.fft -> ATen _fft_r2c/_fft_c2c -> MKL DFTI on
CPU); the reference pins only ``torch>=2.0.0`` (requirements.txt:4).  Pinned against the reference's own
outputs by tests/golden (see oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch


def live_bins(num_filters: int, T: int) -> int:
    """k of spectral_layers.py:94 / wirtinger_ops.py:187."""
    return min(int(num_filters), T // 2)


# ----------------------------------------------------------------------------------------------
# torch port (autograd backward), reference spectral_layers.py:83-118
# ----------------------------------------------------------------------------------------------
def torch_port_forward(x: torch.Tensor, w_re, w_im, bias) -> torch.Tensor:
    """y = Re(ifft(lowpass_filter(fft(x)))) + bias ; x: (B,T,D) real, w_*: (D,F) or None, bias: (D,) or None."""
    B, T, D = x.shape
    spec = torch.fft.fft(x, dim=1)                                   # :88
    if w_re is not None:
        k = live_bins(w_re.shape[1], T)                              # :94
        w = torch.complex(w_re, w_im)                                # :97
        kept = torch.zeros_like(spec)                                # :104
        kept[:, :k, :] = spec[:, :k, :] * w[:, :k].T.unsqueeze(0)    # :101-105
        spec = kept                                                  # :109
    y = torch.fft.ifft(spec, dim=1).real                             # :112
    if bias is not None:
        y = y + bias                                                 # :116
    return y


def torch_port_fwd_bwd(x, w_re, w_im, bias, g):
    """Returns (y, dL/dx, dL/dw_re, dL/dw_im, dL/dbias) for upstream gradient g, all via CPU autograd."""
    x = x.detach().clone().requires_grad_(True)
    params = [p.detach().clone().requires_grad_(True) for p in (w_re, w_im, bias)]
    y = torch_port_forward(x, *params)
    y.backward(g)
    return (y.detach(), x.grad, params[0].grad, params[1].grad, params[2].grad)


# ----------------------------------------------------------------------------------------------
# closed form, float64
# ----------------------------------------------------------------------------------------------
def closed_form_f64(x, w_re, w_im, bias, g=None):
    """numpy float64 closed form.  Inputs: array-likes.  Returns dict with y, X_low and (if g) gx, gw_re, gw_im, gb.

    y[b,t,d]   = (1/T) Re sum_{f<k} X[b,f,d] W[d,f] e^{+2 pi i f t/T} + bias[d]
    gx         = same operator with conj(W), no bias, applied to g
    dL/dW[d,f] = (1/T) sum_b G[b,f,d] conj(X[b,f,d])
    """
    x = np.asarray(x, dtype=np.float64)
    w_re = np.asarray(w_re, dtype=np.float64)
    w_im = np.asarray(w_im, dtype=np.float64)
    B, T, D = x.shape
    F = w_re.shape[1]
    k = live_bins(F, T)
    W = (w_re + 1j * w_im)[:, :k].T                       # (k, D)
    X = np.fft.fft(x, axis=1)[:, :k, :]                   # (B, k, D)

    def synth(A):
        full = np.zeros((B, T, D), dtype=np.complex128)
        full[:, :k, :] = A
        return np.fft.ifft(full, axis=1).real

    out = {"X_low": X, "k": k}
    y = synth(X * W[None])
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    out["y"] = y
    if g is not None:
        g = np.asarray(g, dtype=np.float64)
        G = np.fft.fft(g, axis=1)[:, :k, :]
        out["gx"] = synth(G * np.conj(W)[None])
        gW = (G * np.conj(X)).sum(axis=0) / T             # (k, D)
        gw_re = np.zeros((D, F)); gw_im = np.zeros((D, F))
        gw_re[:, :k] = gW.real.T
        gw_im[:, :k] = gW.imag.T
        out["gw_re"], out["gw_im"] = gw_re, gw_im
        out["gb"] = g.sum(axis=(0, 1))
    return out


# ----------------------------------------------------------------------------------------------
# Wirtinger filter multiply, reference wirtinger_ops.py:34-82 and :170-203
# ----------------------------------------------------------------------------------------------
def wirtinger_multiply_fwd_bwd(x_freq, w, g):
    """f = x_freq * w (w broadcast over dim 0); grad_x = g conj(w); grad_w = sum_b g conj(x) keepdim (:71-80)."""
    out = x_freq * w
    gx = g * torch.conj(w)
    gw = (g * torch.conj(x_freq)).sum(dim=0, keepdim=True)
    return out, gx, gw


def wirtinger_filter_forward(x_freq, w_re, w_im):
    """wirtinger_ops.py:170-203: low-pass filter of an already-transformed (B,T,D) complex tensor."""
    B, T, D = x_freq.shape
    k = live_bins(w_re.shape[1], T)
    w = torch.complex(w_re, w_im)[:, :k].T.unsqueeze(0)
    out = torch.zeros_like(x_freq)
    out[:, :k, :] = x_freq[:, :k, :] * w
    return out


def rel_l2(a, b) -> float:
    """||a-b|| / ||b|| in float64 (b is the yardstick)."""
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64)
    b = np.asarray(b, dtype=a.dtype)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return float(num / den) if den > 0 else float(num)
