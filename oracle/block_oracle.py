"""CPU restatement of the block-level callers of the hot path.  TEST INFRASTRUCTURE
(see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Each function follows the reference lines it cites with plain torch CPU ops (float32, like the reference); parity is
pinned by tests/golden/block_*.npz, generated from the UNMODIFIED reference by oracle/make_golden_blocks.py and checked
in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .spectral_mixing_oracle import torch_port_forward


def mlp_block_spectral_half(x, ln_w, ln_b, eps, w_re, w_im, bias):
    """x + spectral_mix(norm1(x)) -- /root/reference/fft_tensor/spectral_layers.py:185 (norm1: :161)."""
    xn = F.layer_norm(x, (x.shape[-1],), ln_w, ln_b, eps)
    return x + torch_port_forward(xn, w_re, w_im, bias)


def fixed_block_spectral_half(x, ln_w, ln_b, eps, kernel, gain, gate_freq_logits, gate_w, gate_b, cutoff=None, transition_bins=1):
    """residual + causal FFT convolution of ln(x) -- /root/reference/fft_lm/train_fixed_full.py:498-555, line by line
    (dropout inactive)."""
    residual = x
    x = F.layer_norm(x, (x.shape[-1],), ln_w, ln_b, eps)                       # :502-503
    B, T, C = x.shape
    K = kernel.shape[0]
    n_fft = 1
    while n_fft < (T + K - 1):                                                 # :507-510
        n_fft *= 2
    k = torch.zeros(n_fft, dtype=x.dtype)                                      # :513-515
    k[:K] = kernel
    k_freq = torch.fft.rfft(k)
    x_pad = F.pad(x, (0, 0, 0, n_fft - T))                                     # :518-519
    x_freq = torch.fft.rfft(x_pad, dim=1)
    y_freq = x_freq * k_freq.unsqueeze(0).unsqueeze(-1) * gain.unsqueeze(0).unsqueeze(0)   # :522
    Fbins = y_freq.size(1)
    g_freq = torch.sigmoid(gate_freq_logits[:Fbins]).to(dtype=y_freq.real.dtype)           # :528
    pooled = x.mean(dim=1)                                                     # :531
    g_ctx = torch.sigmoid(F.linear(pooled, gate_w, gate_b)).to(dtype=y_freq.real.dtype)    # :532
    y_freq = y_freq * g_freq.unsqueeze(0).unsqueeze(-1) * g_ctx.unsqueeze(1)   # :535
    if cutoff is not None:                                                     # :538-550
        cutoff_idx = min(int(cutoff), Fbins)
        if cutoff_idx < Fbins:
            trans = min(transition_bins, cutoff_idx)
            mask = torch.ones(Fbins, dtype=y_freq.real.dtype)
            start = cutoff_idx - trans
            if trans > 0:
                t = torch.linspace(0, 1, steps=trans, dtype=mask.dtype)
                mask[start:cutoff_idx] = 0.5 * (1.0 + torch.cos(torch.pi * t))
            mask[cutoff_idx:] = 0.0
            y_freq = y_freq * mask.unsqueeze(0).unsqueeze(-1)
    y_pad = torch.fft.irfft(y_freq, n=n_fft, dim=1)                            # :552
    return residual + y_pad[:, :T, :]                                          # :554-557


def overlap_save_conv(ctx_ln_new, ln_chunk, h_chunk, kernel, gain, gate_freq_logits, gate_w, gate_b, n_fft_full):
    """h_chunk + the chunk rows of the overlap-save convolution -- /root/reference/scripts/generate_chunked_overlap_save.py
    :113-170 (pooled context gate :113-115, segment :127-141, multiply :157-160, rows [K-1, K-1+B) :165-169)."""
    Bc = ln_chunk.size(1)
    K = kernel.shape[0]
    pooled = ctx_ln_new.sum(dim=1) / float(ctx_ln_new.size(1))
    g_ctx = torch.sigmoid(F.linear(pooled, gate_w, gate_b))
    Fbins = n_fft_full // 2 + 1
    g_freq = torch.sigmoid(gate_freq_logits[:Fbins])
    overlap = ctx_ln_new[:, -(K - 1 + Bc): -Bc, :] if K > 1 else ctx_ln_new[:, :0, :]
    x_seg = torch.cat([overlap, ln_chunk], dim=1)
    L = x_seg.size(1)
    x_pad = F.pad(x_seg, (0, 0, 0, n_fft_full - L)) if L < n_fft_full else x_seg[:, :n_fft_full, :]
    x_freq = torch.fft.rfft(x_pad.float(), dim=1)
    k = torch.zeros(n_fft_full)
    k[:K] = kernel
    y_freq = x_freq * torch.fft.rfft(k).view(1, -1, 1) * gain.view(1, 1, -1)
    y_freq = y_freq * g_freq.view(1, -1, 1) * g_ctx.view(1, 1, -1)
    y_pad = torch.fft.irfft(y_freq, n=n_fft_full, dim=1)
    return h_chunk + y_pad[:, K - 1: K - 1 + Bc, :]


def ema_scan(chunks, rho, theta, mode="aligned", init=None):
    """SpectralEMA.scan over .update -- /root/reference/fft_lm/spectral_ssm.py:78-125."""
    B, S, Fq = chunks.shape
    state = torch.zeros(B, Fq, dtype=torch.complex64) if init is None else init
    a = (rho * torch.exp(1j * theta)).to(torch.complex64)
    omr = 1.0 - rho
    for t in range(S):
        c = chunks[:, t, :]
        if mode == "polar":
            m_new = rho.unsqueeze(0) * torch.abs(state).float() + omr.unsqueeze(0) * torch.abs(c).float()
            state = m_new.to(torch.complex64) * torch.exp(1j * torch.angle(c).float()).to(torch.complex64)
        else:
            rot = torch.exp(1j * (torch.angle(c).float() - torch.angle(state).float())).to(torch.complex64)
            state = a.unsqueeze(0) * (state * rot) + omr.unsqueeze(0).to(torch.complex64) * c
    return state
