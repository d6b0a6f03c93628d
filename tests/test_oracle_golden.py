"""Pins oracle/ against outputs of the UNMODIFIED reference (tests/golden, made by oracle/make_golden.py).

The reference's own tests hold no golden vectors for this path; the fixtures were produced by
importing /root/reference/fft_tensor/spectral_layers.py and wirtinger_ops.py in the build container.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import spectral_mixing_oracle as orc

LAYER_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "layer_*.npz")))
LAYER_FILES = [f for f in LAYER_FILES if "nonlearnable" not in f]
KEYS = [("y", "y"), ("gx", "gx"), ("gw_re", "gw_re"), ("gw_im", "gw_im"), ("gb", "gb")]


def test_fixtures_present():
    assert len(LAYER_FILES) >= 8


@pytest.mark.parametrize("path", LAYER_FILES, ids=[os.path.basename(p)[6:-4] for p in LAYER_FILES])
def test_torch_port_matches_reference(path):
    z = np.load(path)
    t = lambda k: torch.from_numpy(z[k])
    torch.set_num_threads(1)
    out = orc.torch_port_fwd_bwd(t("x"), t("w_re"), t("w_im"), t("bias"), t("g"))
    for (name, key), got in zip(KEYS, out):
        err = orc.rel_l2(got.numpy(), z[key])
        assert err <= 1e-6, (name, err)     # same torch.fft calls -> agreement to rounding


@pytest.mark.parametrize("path", LAYER_FILES, ids=[os.path.basename(p)[6:-4] for p in LAYER_FILES])
def test_closed_form_matches_reference(path):
    z = np.load(path)
    out = orc.closed_form_f64(z["x"], z["w_re"], z["w_im"], z["bias"], z["g"])
    for name, key in KEYS:
        err = orc.rel_l2(z[key], out[name])   # reference is fp32: error is the reference's own rounding
        assert err <= 2e-6, (name, err)
    k = out["k"]
    assert np.all(z["gw_re"][:, k:] == 0) and np.all(z["gw_im"][:, k:] == 0)   # dense zero columns >= k
    assert np.abs(z["gw_im"][:, 0]).max() <= 1e-5 * max(1.0, np.abs(z["gw_re"][:, 0]).max())


def test_known_answer_grad_norm(golden_dir):
    # spectral_layers.py:288-299: default init, loss = y.sum() -> ||x.grad|| = sqrt(B*T*D) = 256 at (2,128,256)
    z = np.load(os.path.join(golden_dir, "layer_default_init_ysum.npz"))
    assert abs(np.linalg.norm(z["gx"]) - 256.0) < 1e-2
    out = orc.closed_form_f64(z["x"], z["w_re"], z["w_im"], z["bias"], z["g"])
    assert abs(np.linalg.norm(out["gx"]) - 256.0) < 1e-6


def test_nonlearnable_is_identity(golden_dir):
    z = np.load(os.path.join(golden_dir, "layer_nonlearnable.npz"))
    assert int(z["n_params"]) == 0
    assert orc.rel_l2(z["y"], z["x"]) < 1e-6


def test_wirtinger_oracle_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "wirtinger.npz"))
    t = lambda k: torch.from_numpy(z[k])
    out, gx, gw = orc.wirtinger_multiply_fwd_bwd(t("mul_x"), t("mul_w"), t("mul_g"))
    assert orc.rel_l2(out.numpy(), z["mul_out"]) < 1e-6
    assert orc.rel_l2(gx.numpy(), z["mul_gx"]) < 1e-6
    assert orc.rel_l2(gw.numpy(), z["mul_gw"]) < 1e-6
    f = orc.wirtinger_filter_forward(t("filt_x"), t("filt_w_re"), t("filt_w_im"))
    assert orc.rel_l2(f.numpy(), z["filt_out"]) < 1e-6


def test_closed_form_half_weight_identity():
    # y = 0.5*irfft(A, n=T) + 0.5*Re(A0)/T + bias
    rng = np.random.default_rng(5)
    B, T, D, F = 2, 64, 6, 9
    x = rng.standard_normal((B, T, D)); wr = rng.standard_normal((D, F)); wi = rng.standard_normal((D, F))
    out = orc.closed_form_f64(x, wr, wi, None)
    A = np.zeros((B, T // 2 + 1, D), complex)
    A[:, :F] = np.fft.fft(x, axis=1)[:, :F] * (wr + 1j * wi).T[None]
    alt = 0.5 * np.fft.irfft(A, n=T, axis=1) + 0.5 * A[:, :1].real / T
    assert orc.rel_l2(alt, out["y"]) < 1e-13


@pytest.mark.parametrize("seed", range(12))
def test_two_restatements_agree_on_random_shapes(seed):
    # the torch port (same torch.fft calls as the reference) and the float64 closed form share no code: they must agree
    # on arbitrary shapes, including T = 1 (k = 0), T = 2 (DC only), odd T and num_filters > T // 2
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 4))
    T = int(rng.choice([1, 2, 3, 5, 16, 31, 64, 100, 257]))
    D = int(rng.integers(1, 9))
    F = int(rng.integers(1, 2 * D + 3))
    gen = torch.Generator().manual_seed(seed)
    x, g = torch.randn(B, T, D, generator=gen), torch.randn(B, T, D, generator=gen)
    w_re, w_im, bias = torch.randn(D, F, generator=gen), torch.randn(D, F, generator=gen), torch.randn(D, generator=gen)
    port = orc.torch_port_fwd_bwd(x, w_re, w_im, bias, g)
    ref = orc.closed_form_f64(x.numpy(), w_re.numpy(), w_im.numpy(), bias.numpy(), g.numpy())
    for (name, _), got in zip(KEYS, port):
        want = ref[name]
        scale = max(np.linalg.norm(want), 1e-6 * max(np.linalg.norm(ref["gw_re"]), 1.0))
        assert np.linalg.norm(got.numpy().astype(np.float64) - want) / scale <= 2e-5, (name, B, T, D, F)
