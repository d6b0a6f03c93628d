"""CPU: the host-side algebra of the fused blocks (no kernel call): how LayerNorm's affine part, the irfft multiplier, the bin
n/2 and the zero-padded window fold into the arrays the extended kernel takes (spectral_conv.py, spectral_layers.py).  The
kernel itself is modelled with torch.fft exactly as include/spectral_mix_b200.h documents sml_forward_ext; the result must
reproduce the fixtures generated from the UNMODIFIED reference (forward AND, through autograd of the model, the gradients)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import tensor_cuda_fft_b200.spectral_conv as sc
from oracle.spectral_mixing_oracle import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kernel_model(xhat, w_re, w_im, w_nyq, sb_re, sb_im, sb_nyq, scale, n, T_out, residual=None):
    """sml_forward_ext as documented: y = residual + Re(ifft(one-sided (X W + sb) * scale))[:T_out] + the bin n/2 at weight 1/n."""
    B, T, C = xhat.shape
    X = torch.fft.fft(F.pad(xhat, (0, 0, 0, n - T)), dim=1)                  # (B, n, C)
    k = n // 2
    W = torch.complex(w_re, w_im).t().unsqueeze(0)                           # (1, k, C)
    A = X[:, :k, :] * W
    if sb_re is not None:
        A = A + torch.complex(sb_re, sb_im).t().unsqueeze(0)
    A = A * scale.unsqueeze(1)
    Z = torch.zeros(B, n, C, dtype=torch.complex64)
    Z = torch.cat([A, torch.zeros(B, n - k, C, dtype=A.dtype)], dim=1)
    y = torch.fft.ifft(Z, dim=1).real
    tt = torch.arange(n, dtype=torch.float32)
    nyq = (X[:, k, :].real * w_nyq + (sb_nyq if sb_nyq is not None else 0.0)) * scale / n      # (B, C)
    y = y + nyq.unsqueeze(1) * torch.cos(torch.pi * tt).view(1, n, 1)
    y = y[:, :T_out, :]
    return y if residual is None else residual + y


@pytest.mark.parametrize("name", ["block_fixed_t64_k16_c32.npz", "block_fixed_t96_k24_c16.npz",
                                  "block_fixed_t512_k128_c32_cut.npz", "block_fixed_t1024_k128_c16.npz"])
def test_fixed_block_folding(name):
    d = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, name)).items()}
    x = d["x"].clone().requires_grad_(True)
    B, T, C = x.shape
    K = int(d["K"])
    blk = sc.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=int(d["trans"]), dropout=0.0)
    blk.load_state_dict({k[3:]: v for k, v in d.items() if k.startswith("sd.")}, strict=True)
    blk.eval()
    cutoff = None if int(d["cutoff"]) < 0 else int(d["cutoff"])
    n = sc.conv_fft_len(T, K)
    # the same steps as FixedSpectralBlock.spectral_half / _CausalSpectralConvFn.forward, differentiable on the CPU
    H = sc._multiplier(blk.kernel, blk.gate_freq_logits, n, K, cutoff, blk.transition_bins)
    gamma, beta, gain = blk.ln.weight, blk.ln.bias, blk.gain
    w_re, w_im, w_nyq = sc._kernel_filter(H, gamma * gain)
    rect = torch.zeros(n)
    rect[:T] = 1.0
    Q = torch.fft.rfft(rect) * H
    q_re, q_im, q_nyq = sc._kernel_filter(Q, torch.ones(1))
    bg = beta * gain
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + blk.ln.eps)
    xhat = (x - mean) * rstd
    pooled = gamma * xhat.mean(1) + beta
    s = torch.sigmoid(blk.gate_ctx(pooled))
    half = kernel_model(xhat, w_re, w_im, w_nyq, bg[:, None] * q_re, bg[:, None] * q_im, bg * q_nyq[0], s, n, T, residual=x)
    y = half + blk.ffn(blk.ffn_ln(half))
    assert rel_l2(y.detach().numpy(), d["y"].numpy()) <= 2e-6
    y.backward(d["g"])
    assert rel_l2(x.grad.numpy(), d["gx"].numpy()) <= 1e-5
    for k_, p in blk.named_parameters():
        want = d["grad." + k_].numpy()
        assert rel_l2(p.grad.numpy(), want) <= 2e-5, k_
