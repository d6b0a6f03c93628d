"""CPU: the block-level oracle restatements (oracle/block_oracle.py) against the fixtures generated from the UNMODIFIED
reference (oracle/make_golden_blocks.py): SpectralMLPBlock's spectral half, FixedSpectralBlock's causal FFT convolution,
the overlap-save chunk update and SpectralEMA.scan."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import block_oracle as bo
from oracle.spectral_mixing_oracle import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name)).items()}


def t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", ["block_mlp_t256_d64.npz", "block_mlp_t1024_d48.npz", "block_mlp_t100_d32.npz"])
def test_mlp_block_half(name):
    d = load(name)
    half = bo.mlp_block_spectral_half(t(d["x"]), t(d["sd.norm1.weight"]), t(d["sd.norm1.bias"]), 1e-5,
                                      t(d["sd.spectral_mix.weight_real"]), t(d["sd.spectral_mix.weight_imag"]),
                                      t(d["sd.spectral_mix.bias"]))
    assert rel_l2(half.numpy(), d["half"]) <= 1e-6


@pytest.mark.parametrize("name", ["block_fixed_t64_k16_c32.npz", "block_fixed_t96_k24_c16.npz",
                                  "block_fixed_t512_k128_c32_cut.npz", "block_fixed_t1024_k128_c16.npz"])
def test_fixed_block_half(name):
    d = load(name)
    x = t(d["x"])
    cutoff = None if int(d["cutoff"]) < 0 else int(d["cutoff"])
    half = bo.fixed_block_spectral_half(x, t(d["sd.ln.weight"]), t(d["sd.ln.bias"]), 1e-5, t(d["sd.kernel"]), t(d["sd.gain"]),
                                        t(d["sd.gate_freq_logits"]), t(d["sd.gate_ctx.weight"]), t(d["sd.gate_ctx.bias"]),
                                        cutoff=cutoff, transition_bins=int(d["trans"]))
    # the fixture holds the whole block: finish it with the reference's FFN (train_fixed_full.py:557-559)
    ff = torch.nn.functional.layer_norm(half, (x.shape[-1],), t(d["sd.ffn_ln.weight"]), t(d["sd.ffn_ln.bias"]), 1e-5)
    ff = torch.nn.functional.linear(ff, t(d["sd.ffn.0.weight"]), t(d["sd.ffn.0.bias"]))
    ff = torch.nn.functional.gelu(ff)
    ff = torch.nn.functional.linear(ff, t(d["sd.ffn.3.weight"]), t(d["sd.ffn.3.bias"]))
    assert rel_l2((half + ff).numpy(), d["y"]) <= 1e-6


def test_spectral_ema_scan():
    d = load("block_spectral_ema.npz")
    chunks, init = t(d["chunks"]), t(d["init"])
    for mode in ("aligned", "polar"):
        rho = torch.sigmoid(t(d[f"{mode}.rho_logit"]))
        theta = math.pi * torch.tanh(t(d[f"{mode}.theta_raw"]))
        assert rel_l2(bo.ema_scan(chunks, rho, theta, mode).numpy(), d[f"{mode}.scan"]) <= 1e-6
        assert rel_l2(bo.ema_scan(chunks, rho, theta, mode, init=init).numpy(), d[f"{mode}.scan_init"]) <= 1e-6
        assert rel_l2(bo.ema_scan(chunks[:, 5:6, :], rho, theta, mode, init=init).numpy(), d[f"{mode}.update"]) <= 1e-6


def test_spectral_ema_autograd_path_cpu():
    """SpectralEMA.scan with gradients enabled runs the differentiable torch loop (spectral_ssm.py:78-125 as ops): on the CPU it
    must reproduce the reference's states, and gradients must reach rho / theta (the kernel is the no_grad inference path)."""
    import tensor_cuda_fft_b200.spectral_conv as sc
    d = load("block_spectral_ema.npz")
    chunks, init = t(d["chunks"]), t(d["init"])
    for mode in ("aligned", "polar"):
        ema = sc.SpectralEMA(sc.EMAConfig(n_freqs=chunks.shape[2], mode=mode))
        with torch.no_grad():
            ema.rho_logit.copy_(t(d[f"{mode}.rho_logit"]))
            ema.theta_raw.copy_(t(d[f"{mode}.theta_raw"]))
        out = ema.scan(chunks, init=init)
        assert rel_l2(out.detach().numpy(), d[f"{mode}.scan_init"]) <= 1e-6
        out.abs().sum().backward()
        assert ema.rho_logit.grad is not None and float(ema.rho_logit.grad.abs().sum()) > 0
