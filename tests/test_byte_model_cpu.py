"""Host logic of the LM that contains the layer (no GPU): the vectorised byte encoder reproduces the reference's per-position
FFT loop (golden vectors from the unmodified reference, oracle/make_golden.py) and the reference state_dict loads strictly."""
import os

import numpy as np
import torch

from oracle import spectral_mixing_oracle as orc


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "lm_small.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, sd


def test_reference_state_dict_loads_strict(golden_dir):
    from tensor_cuda_fft_b200 import SpectralLanguageModel
    z, sd = _golden(golden_dir)
    E, L, T, _ = (int(v) for v in z["cfg"])
    model = SpectralLanguageModel(embed_dim=E, num_layers=L, max_seq_len=T, dropout=0.0)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)


def test_vectorised_byte_encoder_matches_reference_loop(golden_dir):
    from tensor_cuda_fft_b200 import SpectralLanguageModel
    z, sd = _golden(golden_dir)
    E, L, T, _ = (int(v) for v in z["cfg"])
    model = SpectralLanguageModel(embed_dim=E, num_layers=L, max_seq_len=T, dropout=0.0)
    model.load_state_dict(sd)
    emb = model.byte_encoder(torch.from_numpy(z["ids"]))      # pure torch: runs on the CPU
    assert orc.rel_l2(emb.detach().numpy(), z["emb"]) <= 2e-6


def test_byte_encoder_exactly_zero_bins(golden_dir):
    """Constant / periodic / zero-padded byte sequences have exactly-zero bins, where the reference's per-position loop sees
    angle(0) = 0 at every position (byte_spectral_model.py:63-94): the one-FFT encoder must not apply its phase ramp there."""
    from tensor_cuda_fft_b200.byte_spectral_model import ByteSpectralEmbedding
    z = np.load(os.path.join(golden_dir, "byte_encoder_edge.npz"))
    E, T = (int(v) for v in z["cfg"])
    enc = ByteSpectralEmbedding(E, T)
    enc.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}, strict=True)
    with torch.no_grad():
        emb = enc(torch.from_numpy(z["ids"]))
    for i in range(emb.shape[0]):
        err = orc.rel_l2(emb[i].numpy(), z["emb"][i])
        assert err <= 1e-5, f"row {i} ({z['ids'][i][:6]}...): rel-L2 {err:.3e}"
