"""B200-native causal FFT-convolution core of fft_lm's spectral blocks, and its inference companions.

Mirrors (same class / parameter names, so the reference's checkpoints load unchanged):
  * ``FixedSpectralBlock``            /root/reference/fft_lm/train_fixed_full.py:425-560
  * ``overlap_save_block_update``     /root/reference/scripts/generate_chunked_overlap_save.py:78-177
  * ``EMAConfig`` / ``SpectralEMA``   /root/reference/fft_lm/spectral_ssm.py:30-125

The reference composes the block from a LayerNorm, ``F.pad`` to ``n_fft = next_pow2(T + K - 1)``, ``rfft``, four broadcast
multiplies, ``irfft``, a slice and a residual add.  Here the whole spectral half is the extended fused kernel of
csrc/sml_fast.cuh behind ``sml_forward_ext`` / ``sml_backward_ext`` (include/spectral_mix_b200.h):

  * the zero padding is TMA out-of-bounds fill and the ``[:T]`` slice is TMA store clipping (row windows) -- no padded copy;
  * LayerNorm is applied while the rows are loaded (row statistics from ``sml_ln_stats``); its affine part folds into the
    multiplier (gamma) and into a "spectral bias" (beta on a zero-padded window is ``beta * rfft(rect_T)``);
  * the multiplier is rank one in (frequency, channel) times a per-(batch, channel) gate: it enters the kernel as the
    layer's ``(D, F)`` filter (``2 H[f]`` on bins ``1 .. n/2-1``: the kernel keeps ``Re(ifft(.))`` semantics), a real weight
    for the bin ``n/2`` and the per-(b, c) factor ``chan_scale``;
  * the residual rows are added while the output rows are stored.

CUDA only; a shape the fused kernels do not take raises (there is no PyTorch fallback for the convolution itself).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native
from .spectral_layers import _IO_DTYPES, _f32c, _on_device, _ptr, _shape_info, _stream_handle


def conv_fft_len(T: int, kernel_len: int) -> int:
    """next power of two >= T + K - 1 (train_fixed_full.py:507-510)."""
    n = 1
    while n < T + kernel_len - 1:
        n *= 2
    return n


def _aligned(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


class _CausalSpectralConvFn(torch.autograd.Function):
    """y[:, :T] = residual + s[b,c] * irfft( (gamma x^ + beta rect)~ * gain[c] * H[f] )[:T]   (x^ = LayerNorm rows without affine)

    The multiplier is rank one in (channel, frequency): the kernel forms ``chan[c] * H[f]`` itself (``sml_ext`` rank-one mode), so no
    (C, F) array exists on either side.  Inputs (built by the caller with ordinary autograd ops from the block's parameters):
      x (B,T,C); h_re, h_im (n/2,): H with the factor 2 on bins >= 1 (the kernel keeps Re(ifft(.)) semantics); h_nyq (): Re H[n/2];
      chan (C,) = gamma*gain; beta_gain (C,) = beta*gain; q_re, q_im (n/2,), q_nyq (): rfft(rect_T)*H in the same convention (no
      gradient: the beta path's gradient with respect to H flows through ``u``); u (T,) = irfft(rfft(rect_T) H)[:T];
      gamma, beta (C,), gate_w (C,C), gate_b (C,): the context gate s = sigmoid(gate(mean_t LayerNorm(x))) (train_fixed_full.py:531-533);
      eps; n_fft; add_residual (False: return the convolution alone, for a dropout between it and the skip connection).
    """

    @staticmethod
    def forward(ctx, x, h_re, h_im, h_nyq, chan, beta_gain, q_re, q_im, q_nyq, u, gamma, beta, gate_w, gate_b, eps, n_fft, add_residual):
        B, T, C = x.shape
        io = _IO_DTYPES[x.dtype]
        Fn = n_fft // 2
        lib = _native.lib()
        dev = x.device
        xc = _aligned(x)
        hr, hi, hn = _f32c(h_re), _f32c(h_im), _f32c(h_nyq).reshape(1)
        qr, qi, qn = _f32c(q_re), _f32c(q_im), _f32c(q_nyq).reshape(1)
        ch, bg = _f32c(chan), _f32c(beta_gain)
        stats = torch.empty(B, n_fft, 2, dtype=torch.float32, device=dev)
        with _on_device(dev):
            _native.check(lib.sml_ln_stats(_ptr(xc), _ptr(stats), B, n_fft, T, 0, C, float(eps), io, _stream_handle(dev)))
        mean, rstd = stats[:, :T, 0], stats[:, :T, 1]
        # context gate: pooled = mean_t LayerNorm(x) = gamma * mean_t x^ + beta ; mean_t x^ from the row statistics and ONE
        # batched matrix-vector product over x (the normalised tensor is never materialised)
        mhat = (torch.bmm(rstd.to(xc.dtype).unsqueeze(1), xc).squeeze(1).float() - (mean * rstd).sum(1, keepdim=True)) / T
        pooled = gamma.float() * mhat + beta.float()
        s = torch.sigmoid(F.linear(pooled, gate_w.float(), gate_b.float())).contiguous()
        w_nyq = (ch * hn).contiguous()          # (C,) vectors for the bin n/2
        sb_nyq = (bg * qn).contiguous()
        y = torch.empty_like(xc)
        # the saved spectrum serves the filter gradient AND the gate gradient (d_core, which also feeds dL/dx through the pooled
        # context): needed whenever anything is differentiated
        need_spectrum = any(ctx.needs_input_grad)
        xlow = torch.empty(max(_shape_info(B, n_fft, C, Fn, io)[1] // 8, 1), dtype=torch.complex64, device=dev) if need_spectrum else None
        xnyq = torch.empty(B, C, dtype=torch.float32, device=dev)
        ext = _native.make_ext(row_stats=stats, residual=xc if add_residual else None, chan_scale=s, w_nyq=w_nyq, sb_nyq=sb_nyq, x_nyq=xnyq,
                               h_re=hr, h_im=hi, h_nyq=hn, chan=ch, bg=bg, q_re=qr, q_im=qi, q_nyq=qn, T_in=T, T_out=T)
        with _on_device(dev):
            _native.check(lib.sml_forward_ext(_ptr(xc), None, None, None, _ptr(y), _ptr(xlow), B, n_fft, C, Fn, io, ctypes.byref(ext),
                                              _stream_handle(dev)))
        ctx.save_for_backward(xc, stats, hr, hi, hn, ch, bg, qr, qi, qn, w_nyq, xlow, xnyq, s, pooled, mhat, gamma, gate_w)
        ctx.cfg = (B, T, C, Fn, io, n_fft, add_residual)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        xc, stats, hr, hi, hn, ch, bg, qr, qi, qn, w_nyq, xlow, xnyq, s, pooled, mhat, gamma, gate_w = ctx.saved_tensors
        B, T, C, Fn, io, n_fft, add_residual = ctx.cfg
        lib = _native.lib()
        dev = g.device
        gc = _aligned(g.to(xc.dtype))
        rows = int(lib.sml_ext_hpart_rows(B, n_fft, C, Fn, io))
        hpart = torch.empty(rows, Fn, dtype=torch.complex64, device=dev)      # channel-contracted filter gradient, one row per work item
        gnyq = torch.empty(B, C, dtype=torch.float32, device=dev)
        d_core = torch.empty(B, C, dtype=torch.float32, device=dev)
        d_q = torch.empty(B, C, dtype=torch.float32, device=dev)
        gh = torch.empty_like(gc)
        ext = _native.make_ext(chan_scale=s, w_nyq=w_nyq, x_nyq=xnyq, g_nyq=gnyq, T_in=T, T_out=T, d_core=d_core, d_q=d_q, q_re=qr, q_im=qi,
                               q_nyq=qn, h_re=hr, h_im=hi, h_nyq=hn, chan=ch, bg=bg, hpart=hpart)
        with _on_device(dev):
            _native.check(lib.sml_backward_ext(_ptr(gc), _ptr(xlow), None, None, _ptr(gh), None, None, None, None, 0, B, n_fft, C, Fn, io,
                                               ctypes.byref(ext), _stream_handle(dev)))
        dH = hpart.sum(0)                                       # dL/dH in the kernel's (gw_re + i gw_im) convention
        d_h_nyq = (gnyq * ch).sum()
        # gate: dL/ds[b,c] = sum_t g * (y - residual) / s, evaluated by the backward kernel in the spectral domain (it holds G, X_low and
        # H anyway): d_core = (1/T) sum_f Re(conj(G) X H), d_q = (1/T) sum_f Re(conj(G) Q) -- no pass over y, y is not kept
        ds = ch * d_core + bg * d_q
        d_chan = (s * d_core).sum(0)
        d_beta_gain = (s * d_q).sum(0)
        dz = ds * s * (1.0 - s)
        d_gate_w = dz.t() @ pooled
        d_gate_b = dz.sum(0)
        dpooled = dz @ gate_w.float()
        d_gamma = (dpooled * mhat).sum(0)
        d_beta = dpooled.sum(0)
        dmhat = dpooled * gamma.float()                       # d/d(mean_t x^): spreads over the T rows as dmhat / T
        # beta path, y_beta[b,t,c] = s[b,c] * beta_gain[c] * u[t]:  d_u[t] = sum_b (g[b] (s[b] * beta_gain))[t]  (one batched matrix-vector product)
        d_u = torch.bmm(gc, (s * bg).to(gc.dtype).unsqueeze(2)).squeeze(2).float().sum(0)
        # LayerNorm backward (+ the skip connection's gradient, + the pooled-mean term dmhat / T as chan_add) in one pass
        chan_add = (dmhat / T).contiguous()
        gx = torch.empty_like(gc)
        with _on_device(dev):
            _native.check(lib.sml_ln_backward(_ptr(gh), _ptr(xc), _ptr(stats), _ptr(gc) if add_residual else None, _ptr(chan_add),
                                              _ptr(gx), B, n_fft, T, 0, C, io, _stream_handle(dev)))
        need = ctx.needs_input_grad
        return (gx if need[0] else None,
                dH.real if need[1] else None, dH.imag if need[2] else None, d_h_nyq if need[3] else None,
                d_chan if need[4] else None, d_beta_gain if need[5] else None, None, None, None,
                d_u if need[9] else None,
                d_gamma if need[10] else None, d_beta if need[11] else None,
                d_gate_w if need[12] else None, d_gate_b if need[13] else None,
                None, None, None)


def _multiplier(kernel: torch.Tensor, gate_freq_logits: torch.Tensor, n_fft: int, kernel_len: int, cutoff: Optional[int],
                transition_bins: int) -> torch.Tensor:
    """H[f] = rfft(kernel padded to n_fft) * sigmoid(gate_freq)[f] * mask[f]   (train_fixed_full.py:512-516, :527-528, :538-550)."""
    k = torch.zeros(n_fft, device=kernel.device, dtype=torch.float32)
    k = torch.cat([kernel.float(), k[kernel_len:]]) if kernel_len <= n_fft else kernel.float()[:n_fft]
    H = torch.fft.rfft(k)
    Fb = H.shape[0]
    H = H * torch.sigmoid(gate_freq_logits[:Fb]).float()
    if cutoff is not None:
        cutoff_idx = min(int(cutoff), Fb)
        if cutoff_idx < Fb:
            trans = min(int(transition_bins), cutoff_idx)
            mask = torch.ones(Fb, device=H.device, dtype=torch.float32)
            start = cutoff_idx - trans
            if trans > 0:
                t = torch.linspace(0, 1, steps=trans, device=H.device, dtype=torch.float32)
                mask[start:cutoff_idx] = 0.5 * (1.0 + torch.cos(torch.pi * t))
            mask[cutoff_idx:] = 0.0
            H = H * mask
    return H


def _kernel_filter(H: torch.Tensor, chan: torch.Tensor):
    """(w_re, w_im, w_nyq) of the fused kernel for the irfft multiplier chan[c] * H[f]: the kernel takes Re(ifft(.)) of a
    one-sided spectrum (bins >= 1 at half weight, spectral_layers.py:112), so bins 1..n/2-1 enter doubled."""
    Fn = H.shape[0] - 1
    two = torch.cat([torch.ones(1, device=H.device), torch.full((Fn - 1,), 2.0, device=H.device)])   # (no indexed scalar stores: graph-capturable)
    Hs = H[:Fn] * two
    w_re = chan[:, None] * Hs.real[None, :]
    w_im = chan[:, None] * Hs.imag[None, :]
    return w_re, w_im, chan * H[Fn].real


def causal_spectral_conv_supported(x: torch.Tensor, kernel_len: int) -> bool:
    if not (x.is_cuda and x.dim() == 3 and x.dtype in _IO_DTYPES and x.numel() > 0):
        return False
    B, T, C = x.shape
    n = conv_fft_len(T, kernel_len)
    from .spectral_layers import _ext_supported
    return _ext_supported(B, n, C, n // 2, _IO_DTYPES[x.dtype], T_in=T, T_out=T, nyq=True)


class FixedSpectralBlock(nn.Module):
    """Drop-in for ``fft_lm.train_fixed_full.FixedSpectralBlock`` (train_fixed_full.py:425-560): pre-LayerNorm causal FFT
    convolution with frequency and context gates + residual, then the pointwise FFN (plain PyTorch, as in the reference)."""

    def __init__(self, d_model: int, seq_len: int, kernel_len: int, transition_bins: int, dropout: float = 0.1):
        super().__init__()
        self.ln = nn.LayerNorm(d_model)
        self.drop = nn.Dropout(dropout)
        self.seq_len = seq_len
        self.kernel_len = kernel_len
        self.transition_bins = int(max(1, transition_bins))
        self.kernel = nn.Parameter(torch.zeros(kernel_len))
        nn.init.normal_(self.kernel, mean=0.0, std=0.001)
        self.gain = nn.Parameter(torch.ones(d_model))
        self.max_freq_bins = conv_fft_len(seq_len, kernel_len) // 2 + 1
        self.gate_freq_logits = nn.Parameter(torch.ones(self.max_freq_bins) * 2.0)
        self.gate_ctx = nn.Linear(d_model, d_model)
        nn.init.zeros_(self.gate_ctx.weight)
        nn.init.constant_(self.gate_ctx.bias, 2.0)
        hidden = d_model * 2
        self.ffn_ln = nn.LayerNorm(d_model)
        self.ffn = nn.Sequential(nn.Linear(d_model, hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden, d_model))
        for m in self.ffn:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0.0, std=0.01)
                nn.init.zeros_(m.bias)

    def spectral_half(self, x: torch.Tensor, cutoff: Optional[int] = None) -> torch.Tensor:
        """``residual + drop(causal_conv(ln(x)))`` (train_fixed_full.py:498-555) through the fused kernel."""
        if not x.is_cuda:
            raise RuntimeError("FixedSpectralBlock (B200 build) needs a CUDA tensor; there is no CPU path")
        B, T, C = x.shape
        if not causal_spectral_conv_supported(x, self.kernel_len):
            raise RuntimeError("FixedSpectralBlock: shape not supported by the fused kernels: "
                               + _native.lib().sml_last_error().decode("utf-8", "replace"))
        n = conv_fft_len(T, self.kernel_len)
        H = _multiplier(self.kernel, self.gate_freq_logits, n, self.kernel_len, cutoff, self.transition_bins)
        gamma, beta = self.ln.weight.float(), self.ln.bias.float()
        gain = self.gain.float()
        Fn = n // 2
        two = torch.cat([torch.ones(1, device=x.device), torch.full((Fn - 1,), 2.0, device=x.device)])   # Re(ifft(.)) convention of the kernel
        H2 = H[:Fn] * two
        # beta path: a LayerNorm bias on the T real rows of a zero-padded window is beta * rect_T
        rect = torch.cat([torch.ones(T, device=x.device), torch.zeros(n - T, device=x.device)])
        Q = torch.fft.rfft(rect) * H
        u = torch.fft.irfft(Q, n=n)[:T]
        Qd = Q.detach()
        Q2 = Qd[:Fn] * two
        fuse_res = not (self.training and self.drop.p > 0.0)
        y = _CausalSpectralConvFn.apply(x, H2.real, H2.imag, H[Fn].real, gamma * gain, beta * gain, Q2.real, Q2.imag, Qd[Fn].real, u,
                                        gamma, beta, self.gate_ctx.weight, self.gate_ctx.bias, self.ln.eps, n, fuse_res)
        if not fuse_res:
            y = x + self.drop(y)
        return y

    def forward(self, x: torch.Tensor, cutoff: Optional[int] = None) -> torch.Tensor:
        x = self.spectral_half(x, cutoff)
        ff_in = self.ffn_ln(x)
        return x + self.drop(self.ffn(ff_in))

    def graphed(self, sample: torch.Tensor, cutoff: Optional[int] = None, half_only: bool = False, num_warmup_iters: int = 3):
        """A CUDA-graph-captured callable of this block (or of its spectral half) for inputs of ``sample``'s shape: forward and
        backward replay as graphs.  The block's spectral half is two fused kernels surrounded by a few dozen small parameter-side
        ops (effective filter, gates, reductions): at the reference's sizes their launch cost is most of the step, a graph
        removes it.  Usual CUDA-graph rules: fixed shape and cutoff, dropout inactive."""
        if not sample.is_cuda:
            raise RuntimeError("FixedSpectralBlock (B200 build) needs a CUDA tensor; there is no CPU path")
        if self.training and self.drop.p > 0.0:
            raise RuntimeError("graphed(): dropout draws new random numbers per call; use eval() or dropout=0")
        outer = self

        class _Half(nn.Module):      # only the parameters the spectral half uses (make_graphed_callables wants every one used)
            def __init__(self):
                super().__init__()
                self.ln, self.drop, self.gate_ctx = outer.ln, outer.drop, outer.gate_ctx
                self.kernel, self.gain, self.gate_freq_logits = outer.kernel, outer.gain, outer.gate_freq_logits
                self.kernel_len, self.transition_bins = outer.kernel_len, outer.transition_bins
                self.train(outer.training)

            def forward(self, x):
                return FixedSpectralBlock.spectral_half(self, x, cutoff)

        class _Whole(nn.Module):
            def __init__(self):
                super().__init__()
                self.blk = outer

            def forward(self, x):
                return self.blk(x, cutoff)

        mod = _Half() if half_only else _Whole()
        return torch.cuda.make_graphed_callables(mod, (sample.detach().clone().requires_grad_(True),), num_warmup_iters=num_warmup_iters)


@torch.no_grad()
def overlap_save_block_update(blk: FixedSpectralBlock, layer_state: dict, h_chunk: torch.Tensor, *, n_fft_full: int,
                              kernel_len: int, cache: Optional[dict] = None):
    """Chunked inference step of one block (generate_chunked_overlap_save.py:78-177): same state dictionary
    (``ctx_ln``: LayerNorm outputs of the last T rows, ``ctx_sum``), same outputs.  The convolution over the segment
    ``[last K-1 context rows | new chunk]`` is ONE fused kernel: zero padding to ``n_fft_full`` by TMA fill, the chunk rows
    ``[K-1, K-1+B)`` selected by the store window, the residual added on store."""
    Bc = h_chunk.size(1)
    C = h_chunk.size(2)
    T = layer_state["ctx_ln"].size(1)
    ln_chunk = blk.ln(h_chunk)
    ctx_ln = layer_state["ctx_ln"]
    ctx_ln_new = ln_chunk[:, -T:, :] if Bc >= T else torch.cat([ctx_ln[:, Bc:, :], ln_chunk], dim=1)
    ctx_sum_new = ctx_ln_new.sum(dim=1)
    pooled = (ctx_sum_new / float(ctx_ln_new.size(1))).float()
    g_ctx = torch.sigmoid(blk.gate_ctx(pooled.to(blk.gate_ctx.weight.dtype))).float().contiguous()
    if cache is not None and "w_re" in cache:
        w_re, w_im, w_nyq = cache["w_re"], cache["w_im"], cache["w_nyq"]
    else:
        H = _multiplier(blk.kernel, blk.gate_freq_logits, n_fft_full, kernel_len, None, blk.transition_bins)
        # the chunk rows are transform rows [K-1, K-1+B): a time shift of the output by o = K-1 is the phase ramp
        # exp(2 pi i f o / n) on the multiplier (the kernel then writes rows 0 .. B-1; TMA stores cannot start at a negative row)
        o = kernel_len - 1
        ang = 2.0 * math.pi * ((torch.arange(n_fft_full // 2 + 1, device=H.device) * o) % n_fft_full).double() / n_fft_full
        H = H * torch.complex(torch.cos(ang), torch.sin(ang)).to(torch.complex64)
        w_re, w_im, w_nyq = (t.contiguous() for t in _kernel_filter(H, blk.gain.float()))
        if cache is not None:
            cache.update(w_re=w_re, w_im=w_im, w_nyq=w_nyq)
    lib = _native.lib()
    io = _IO_DTYPES[h_chunk.dtype]
    R = _native.plan(1, n_fft_full, C, n_fft_full // 2, io)["R"]
    L = kernel_len - 1 + Bc
    if L > n_fft_full or Bc % max(R, 1) != 0:
        raise RuntimeError(f"overlap_save_block_update: chunk of {Bc} rows with K = {kernel_len} does not fit n_fft = {n_fft_full} (R = {R})")
    Lp = (L + R - 1) // R * R                     # the kernel's row windows are multiples of its R passes: pad with zero rows
    seg = torch.zeros(1, Lp, C, dtype=h_chunk.dtype, device=h_chunk.device)
    if kernel_len > 1:
        seg[:, : kernel_len - 1] = ctx_ln_new[:, -(kernel_len - 1 + Bc): -Bc, :]
    seg[:, kernel_len - 1: L] = ln_chunk
    res = _aligned(h_chunk)
    h_out = torch.empty_like(res)
    ext = _native.make_ext(residual=res, chan_scale=g_ctx, w_nyq=w_nyq, T_in=Lp, T_out=Bc)
    with _on_device(h_chunk.device):
        _native.check(lib.sml_forward_ext(_ptr(seg), _ptr(w_re), _ptr(w_im), None, _ptr(h_out), None, 1, n_fft_full, C,
                                          n_fft_full // 2, io, ctypes.byref(ext), _stream_handle(h_chunk.device)))
    ff_in = blk.ffn_ln(h_out)
    h_out = h_out + blk.ffn(ff_in)
    return h_out, {"ctx_ln": ctx_ln_new.contiguous(), "ctx_sum": ctx_sum_new.contiguous()}


@dataclass
class EMAConfig:
    n_freqs: int
    rho_init: float = 0.95
    theta_init: float = 0.0
    mode: str = "aligned"


class SpectralEMA(nn.Module):
    """Drop-in for ``fft_lm.spectral_ssm.SpectralEMA`` (spectral_ssm.py:38-125): the chunk scan runs as one kernel
    (``sml_spectral_ema_scan``, one thread per (batch element, frequency)) instead of a Python loop over the S chunks."""

    def __init__(self, cfg: EMAConfig):
        super().__init__()
        self.n_freqs = int(cfg.n_freqs)
        self.mode = str(cfg.mode)
        rho_init = min(max(float(cfg.rho_init), 1e-4), 1 - 1e-4)
        self.rho_logit = nn.Parameter(torch.full((self.n_freqs,), math.log(rho_init / (1 - rho_init)), dtype=torch.float32))
        self.theta_raw = nn.Parameter(torch.full((self.n_freqs,), float(cfg.theta_init), dtype=torch.float32))

    def decay_params(self, device=None, dtype=None):
        rho = torch.sigmoid(self.rho_logit)
        theta = math.pi * torch.tanh(self.theta_raw)
        if device is not None:
            rho, theta = rho.to(device=device), theta.to(device=device)
        if dtype is not None:
            rho, theta = rho.to(dtype=dtype), theta.to(dtype=dtype)
        return rho * torch.exp(1j * theta), rho, (1.0 - rho)

    @torch.no_grad()
    def init_state(self, batch: int, device, dtype) -> torch.Tensor:
        return torch.zeros((batch, self.n_freqs), device=device, dtype=torch.complex64)

    def _update_autograd(self, state: torch.Tensor, fft_chunk: torch.Tensor) -> torch.Tensor:
        """One EMA step as differentiable torch ops (spectral_ssm.py:78-105), for the training-time use of scan()."""
        a, rho, omr = self.decay_params(device=fft_chunk.device, dtype=torch.float32)
        a = a.to(torch.complex64)
        if self.mode == "polar":
            m_new = rho.unsqueeze(0) * torch.abs(state).float() + omr.unsqueeze(0) * torch.abs(fft_chunk).float()
            return m_new.to(torch.complex64) * torch.exp(1j * torch.angle(fft_chunk).float()).to(torch.complex64)
        rot = torch.exp(1j * (torch.angle(fft_chunk).float() - torch.angle(state).float())).to(torch.complex64)
        return a.unsqueeze(0) * (state * rot) + omr.unsqueeze(0).to(torch.complex64) * fft_chunk

    def scan(self, fft_chunks: torch.Tensor, init: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.mode not in ("aligned", "polar"):
            raise ValueError(f"Unknown SpectralEMA mode: {self.mode}")
        if torch.is_grad_enabled() and (fft_chunks.requires_grad or self.rho_logit.requires_grad or self.theta_raw.requires_grad
                                        or (init is not None and init.requires_grad)):
            # training: gradients flow into rho / theta and the chunks; the kernel is the inference path (no backward)
            state = torch.zeros((fft_chunks.shape[0], self.n_freqs), device=fft_chunks.device, dtype=torch.complex64) if init is None else init
            for t in range(fft_chunks.shape[1]):
                state = self._update_autograd(state, fft_chunks[:, t, :])
            return state
        with torch.no_grad():
            return self._scan_kernel(fft_chunks, init)

    def _scan_kernel(self, fft_chunks: torch.Tensor, init: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not fft_chunks.is_cuda:
            raise RuntimeError("SpectralEMA (B200 build) needs CUDA tensors; there is no CPU path")
        B, S, Fq = fft_chunks.shape
        assert Fq == self.n_freqs
        chunks = fft_chunks.to(torch.complex64).contiguous()
        st_in = None if init is None else init.to(torch.complex64).contiguous()
        _, rho, _ = self.decay_params(device=chunks.device, dtype=torch.float32)
        theta = (math.pi * torch.tanh(self.theta_raw)).to(device=chunks.device, dtype=torch.float32).contiguous()
        out = torch.empty(B, Fq, dtype=torch.complex64, device=chunks.device)
        with _on_device(chunks.device):
            _native.check(_native.lib().sml_spectral_ema_scan(_ptr(chunks), _ptr(st_in), _ptr(rho.contiguous()), _ptr(theta), _ptr(out),
                                                              B, S, Fq, 0 if self.mode == "aligned" else 1,
                                                              _stream_handle(chunks.device)))
        return out

    def update(self, state: torch.Tensor, fft_chunk: torch.Tensor) -> torch.Tensor:
        assert state.shape == fft_chunk.shape
        return self.scan(fft_chunk.unsqueeze(1), init=state)
