"""Generate tests/golden/*.npz by executing the UNMODIFIED reference in this container (CPU).

Run from the repo root:  python -m oracle.make_golden
Needs /root/reference (not present on the GPU box -- the fixtures are committed instead).
The reference module is loaded by file path so fft_tensor/__init__.py side effects are avoided
(SURVEY.md D9).  TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

REF_ROOT = os.environ.get("SML_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference(name: str):
    path = os.path.join(REF_ROOT, "fft_tensor", name + ".py")
    spec = importlib.util.spec_from_file_location("_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# (tag, B, T, D, num_filters or None, param init, upstream-grad kind)
LAYER_CASES = [
    ("t512_d64_f128", 1, 512, 64, 128, "randn", "randn"),        # BASELINE cfg-1 band (k=128 of T=512), narrow D
    ("pow2_t64_d32", 2, 64, 32, None, "randn", "randn"),
    ("nonpow2_t100_d32", 3, 100, 32, None, "randn", "randn"),    # reference accepts any T
    ("short_t16_d64", 2, 16, 64, None, "randn", "randn"),        # T//2 < num_filters -> zero grad columns
    ("odd_d_t48_d7", 2, 48, 7, 3, "randn", "randn"),             # odd embed, tiny filter bank
    ("nf_gt_half_t128_d16", 1, 128, 16, 100, "randn", "randn"),  # num_filters > T//2
    ("t1024_d48_f24", 1, 1024, 48, None, "randn", "randn"),      # streamed fast path (R>1)
    ("default_init_ysum", 2, 128, 256, None, "default", "ones"), # spectral_layers.py:288-299 known answer 256.0
]


def run_layer_case(ref_sl, tag, B, T, D, nf, init, gkind, seed):
    gen = torch.Generator().manual_seed(seed)
    layer = ref_sl.SpectralMixingLayer(D, num_filters=nf)
    if init == "randn":
        with torch.no_grad():
            layer.weight_real.copy_(torch.randn(layer.weight_real.shape, generator=gen))
            layer.weight_imag.copy_(torch.randn(layer.weight_imag.shape, generator=gen))
            layer.bias.copy_(torch.randn(layer.bias.shape, generator=gen))
    x = torch.randn(B, T, D, generator=gen).requires_grad_(True)
    g = torch.randn(B, T, D, generator=gen) if gkind == "randn" else torch.ones(B, T, D)
    y = layer(x)
    y.backward(g)
    return {
        "x": x.detach().numpy(), "g": g.numpy(),
        "w_re": layer.weight_real.detach().numpy(), "w_im": layer.weight_imag.detach().numpy(),
        "bias": layer.bias.detach().numpy(),
        "y": y.detach().numpy(), "gx": x.grad.numpy(),
        "gw_re": layer.weight_real.grad.numpy(), "gw_im": layer.weight_imag.grad.numpy(),
        "gb": layer.bias.grad.numpy(),
        "num_filters": np.int64(layer.num_filters),
    }


def run_wirtinger_cases(ref_w, seed):
    gen = torch.Generator().manual_seed(seed)
    out = {}
    # WirtingerGradient.apply on the broadcast shape used by WirtingerSpectralFilter (:192-194)
    B, k, D = 3, 10, 6
    xf = torch.complex(torch.randn(B, k, D, generator=gen), torch.randn(B, k, D, generator=gen)).requires_grad_(True)
    w = torch.complex(torch.randn(1, k, D, generator=gen), torch.randn(1, k, D, generator=gen)).requires_grad_(True)
    gup = torch.complex(torch.randn(B, k, D, generator=gen), torch.randn(B, k, D, generator=gen))
    f = ref_w.WirtingerGradient.apply(xf, w)
    f.backward(gup)
    out.update(mul_x=xf.detach().numpy(), mul_w=w.detach().numpy(), mul_g=gup.numpy(),
               mul_out=f.detach().numpy(), mul_gx=xf.grad.numpy(), mul_gw=w.grad.numpy())
    # WirtingerSpectralFilter module (:145-203)
    B, T, D, nf = 2, 24, 5, 9
    filt = ref_w.WirtingerSpectralFilter(D, nf)
    with torch.no_grad():
        filt.weight.real.copy_(torch.randn(D, nf, generator=gen))
        filt.weight.imag.copy_(torch.randn(D, nf, generator=gen))
    xf = torch.complex(torch.randn(B, T, D, generator=gen), torch.randn(B, T, D, generator=gen)).requires_grad_(True)
    gup = torch.complex(torch.randn(B, T, D, generator=gen), torch.randn(B, T, D, generator=gen))
    o = filt(xf)
    o.backward(gup)
    out.update(filt_x=xf.detach().numpy(), filt_g=gup.numpy(),
               filt_w_re=filt.weight.real.detach().numpy(), filt_w_im=filt.weight.imag.detach().numpy(),
               filt_out=o.detach().numpy(), filt_gx=xf.grad.numpy(),
               filt_gw_re=filt.weight.real.grad.numpy(), filt_gw_im=filt.weight.imag.grad.numpy())
    return out


def run_lm_case(seed):
    """SpectralLanguageModel (byte_spectral_model.py:105-161), the one reference LM that contains the layer: logits, next-byte
    cross-entropy and a few parameter gradients for a small dropout-free configuration."""
    import types
    ref_sl = load_reference("spectral_layers")
    pkg = types.ModuleType("fft_tensor")       # the model file does `from fft_tensor.spectral_layers import SpectralMLPBlock`
    pkg.__path__ = []
    sys.modules["fft_tensor"], sys.modules["fft_tensor.spectral_layers"] = pkg, ref_sl
    ref_bm = load_reference("byte_spectral_model")
    torch.manual_seed(seed)
    E, L, T, B = 32, 2, 64, 2
    model = ref_bm.SpectralLanguageModel(embed_dim=E, num_layers=L, max_seq_len=T, dropout=0.0)
    with torch.no_grad():     # move the filters and the band weights off their trivial initial values
        for blk in model.layers:
            blk.spectral_mix.weight_real.normal_()
            blk.spectral_mix.weight_imag.normal_()
            blk.spectral_mix.bias.normal_()
        model.byte_encoder.freq_bands.uniform_(0.5, 1.5)
    ids = torch.randint(0, 256, (B, T))
    emb = model.byte_encoder(ids)
    logits = model(ids)
    loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, 256), ids[:, 1:].reshape(-1))
    loss.backward()
    out = {"ids": ids.numpy(), "emb": emb.detach().numpy(), "logits": logits.detach().numpy(), "loss": np.float64(loss.item()),
           "cfg": np.array([E, L, T, B], dtype=np.int64)}
    for k, v in model.state_dict().items():
        out["sd." + k] = v.numpy()
    for name in ("layers.0.spectral_mix.weight_real", "layers.0.spectral_mix.weight_imag", "layers.0.spectral_mix.bias",
                 "layers.1.spectral_mix.weight_real", "byte_encoder.freq_bands", "output.weight", "layers.0.norm1.weight"):
        out["grad." + name] = dict(model.named_parameters())[name].grad.numpy()
    return out


def run_byte_encoder_edge_cases(seed):
    """ByteSpectralEmbedding (byte_spectral_model.py:20-102) on inputs whose spectrum has EXACTLY zero bins -- a constant
    byte, period-2 / period-4 sequences, a zero-padded prompt: the reference's per-position loop sees angle(0) = 0 there
    at every position, which a shift-theorem restatement has to reproduce."""
    import types
    ref_sl = load_reference("spectral_layers")
    pkg = types.ModuleType("fft_tensor")
    pkg.__path__ = []
    sys.modules["fft_tensor"], sys.modules["fft_tensor.spectral_layers"] = pkg, ref_sl
    ref_bm = load_reference("byte_spectral_model")
    torch.manual_seed(seed)
    E, T = 32, 64
    enc = ref_bm.ByteSpectralEmbedding(E, T)
    with torch.no_grad():
        enc.freq_bands.uniform_(0.5, 1.5)
    rows = [[65] * T, [97, 98] * (T // 2), [1, 2, 3, 4] * (T // 4), [0] * T,
            [ord(c) for c in "the quick brown fox"] + [0] * (T - 19), [200] * (T - 1) + [7]]
    ids = torch.tensor(rows, dtype=torch.long)
    with torch.no_grad():
        emb = enc(ids)
    out = {"ids": ids.numpy(), "emb": emb.numpy(), "cfg": np.array([E, T], dtype=np.int64)}
    for k, v in enc.state_dict().items():
        out["sd." + k] = v.numpy()
    return out


def run_hybrid_case(ref_sl, seed):
    """HybridSpectralAttention (spectral_layers.py:193-256), a caller of the layer: output and input gradient, dropout 0."""
    torch.manual_seed(seed)
    D, H, T, B = 32, 4, 64, 2
    mod = ref_sl.HybridSpectralAttention(D, num_heads=H, dropout=0.0)
    with torch.no_grad():
        mod.spectral.weight_real.normal_()
        mod.spectral.weight_imag.normal_()
        mod.spectral.bias.normal_()
    x = torch.randn(B, T, D).requires_grad_(True)
    g = torch.randn(B, T, D)
    y = mod(x)
    y.backward(g)
    out = {"x": x.detach().numpy(), "g": g.numpy(), "y": y.detach().numpy(), "gx": x.grad.numpy(),
           "cfg": np.array([D, H, T, B], dtype=np.int64),
           "grad.spectral.weight_real": mod.spectral.weight_real.grad.numpy(), "grad.qkv.weight": mod.qkv.weight.grad.numpy()}
    for k, v in mod.state_dict().items():
        out["sd." + k] = v.numpy()
    return out


def main():
    if not os.path.isdir(REF_ROOT):
        sys.exit(f"{REF_ROOT} not found: golden vectors can only be regenerated where the reference is mounted")
    torch.set_num_threads(1)
    os.makedirs(OUT_DIR, exist_ok=True)
    if "--only-byte-edge" in sys.argv:     # added in round 2: leaves the round-1 fixtures untouched
        np.savez_compressed(os.path.join(OUT_DIR, "byte_encoder_edge.npz"), **run_byte_encoder_edge_cases(seed=909))
        print("wrote byte_encoder_edge.npz")
        return
    ref_sl = load_reference("spectral_layers")
    ref_w = load_reference("wirtinger_ops")
    for i, case in enumerate(LAYER_CASES):
        data = run_layer_case(ref_sl, *case, seed=1000 + i)
        np.savez_compressed(os.path.join(OUT_DIR, f"layer_{case[0]}.npz"), **data)
        print("wrote", case[0], {k: v.shape for k, v in data.items() if hasattr(v, "shape") and v.ndim})
    # learnable=False (spectral_layers.py:62-66, :301-309): identity up to FFT round trip
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(2, 96, 12, generator=gen)
    lay = ref_sl.SpectralMixingLayer(12, learnable=False)
    np.savez_compressed(os.path.join(OUT_DIR, "layer_nonlearnable.npz"), x=x.numpy(), y=lay(x).numpy(),
                        n_params=np.int64(sum(p.numel() for p in lay.parameters())))
    np.savez_compressed(os.path.join(OUT_DIR, "wirtinger.npz"), **run_wirtinger_cases(ref_w, seed=4242))
    np.savez_compressed(os.path.join(OUT_DIR, "hybrid_attention.npz"), **run_hybrid_case(ref_sl, seed=2718))
    np.savez_compressed(os.path.join(OUT_DIR, "lm_small.npz"), **run_lm_case(seed=31337))
    np.savez_compressed(os.path.join(OUT_DIR, "byte_encoder_edge.npz"), **run_byte_encoder_edge_cases(seed=909))
    # parameter-count known answer, BENCHMARKS.md:86
    n = sum(p.numel() for p in ref_sl.SpectralMixingLayer(256).parameters())
    assert n == 65792, n
    print("golden vectors written to", OUT_DIR)


if __name__ == "__main__":
    main()
