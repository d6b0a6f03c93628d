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
