"""B200-native host of the layer inside a real model: drop-in for ``fft_tensor.byte_spectral_model``
(reference: /root/reference/fft_tensor/byte_spectral_model.py), the only LM in the reference that contains
``SpectralMixingLayer`` (through ``SpectralMLPBlock``, :130-133).

Same classes, constructor signatures and ``state_dict`` keys.  One deliberate difference: the reference's
``ByteSpectralEmbedding.forward`` loops over positions in Python and runs one roll + FFT per position (:63-94, the
documented 50-of-56 ms bottleneck).  By the shift theorem ``fft(roll(s, -p))[f] = fft(s)[f] * exp(+2 pi i f p / T)``, so the
magnitude is position independent and the phase advances linearly: the loop collapses to ONE FFT plus elementwise work.
The result is the same up to fp32 rounding (the reference's angle() wraps, sin/cos of it do not care).  Bins that are
EXACTLY zero (a constant byte, periodic or zero-padded sequences) are the one place where the shift theorem says nothing:
the reference sees angle(0) = 0 at every position there, so the phase ramp is masked on those bins
(tests/golden/byte_encoder_edge.npz pins it against the unmodified reference).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .spectral_layers import SpectralMLPBlock


class ByteSpectralEmbedding(nn.Module):
    """byte_spectral_model.py:20-102 -- bytes -> (|S_f| * band_f, sin phase_f(pos), cos phase_f(pos)) -> MLP."""

    def __init__(self, embed_dim: int = 256, max_seq_len: int = 512):
        super().__init__()
        self.embed_dim = embed_dim
        self.max_seq_len = max_seq_len
        self.freq_bands = nn.Parameter(torch.ones(embed_dim // 2))
        self.freq_proj = nn.Sequential(
            nn.Linear(embed_dim, embed_dim * 2),
            nn.LayerNorm(embed_dim * 2),
            nn.GELU(),
            nn.Linear(embed_dim * 2, embed_dim),
        )

    def forward(self, byte_ids: torch.Tensor) -> torch.Tensor:
        B, T = byte_ids.shape
        E = self.embed_dim
        signal = (byte_ids.float() / 127.5) - 1.0                          # :54
        k = min(E // 2, T // 2)                                           # :69
        spec = torch.fft.fft(signal, dim=1)[:, :k]                        # one FFT instead of T (shift theorem)
        mag = spec.abs() * self.freq_bands[:k]                            # :71, :75  (B, k), same for every position
        pos = torch.arange(T, device=byte_ids.device, dtype=torch.float32)
        f = torch.arange(k, device=byte_ids.device, dtype=torch.float32)
        ramp = (2.0 * math.pi / T) * torch.outer(pos, f)                  # (T, k): phase advance of bin f at position pos
        live = (spec != 0).unsqueeze(1)                                   # an exactly-zero bin stays zero under every roll: angle 0
        phase = torch.angle(spec).unsqueeze(1) + ramp.unsqueeze(0) * live  # (B, T, k)  == angle(fft(roll(s, -pos)))  :72
        feats = torch.cat([mag.unsqueeze(1).expand(B, T, k), torch.sin(phase), torch.cos(phase)], dim=-1)   # :80-84
        if feats.size(-1) < E:                                            # :87-91
            feats = F.pad(feats, (0, E - feats.size(-1)))
        else:
            feats = feats[..., :E]
        return self.freq_proj(feats)                                      # :99


class SpectralLanguageModel(nn.Module):
    """byte_spectral_model.py:105-208 -- byte encoder, N x SpectralMLPBlock (the fused layer inside), LayerNorm,
    Linear(embed_dim, 256)."""

    def __init__(self, embed_dim: int = 256, num_layers: int = 6, max_seq_len: int = 512, dropout: float = 0.1):
        super().__init__()
        self.embed_dim = embed_dim
        self.max_seq_len = max_seq_len
        self.byte_encoder = ByteSpectralEmbedding(embed_dim, max_seq_len)
        self.dropout = nn.Dropout(dropout)
        self.layers = nn.ModuleList([SpectralMLPBlock(embed_dim, dropout=dropout) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(embed_dim)
        self.output = nn.Linear(embed_dim, 256)

    def forward(self, byte_ids: torch.Tensor) -> torch.Tensor:
        x = self.dropout(self.byte_encoder(byte_ids))
        for layer in self.layers:
            x = layer(x)
        return self.output(self.norm(x))

    @torch.no_grad()
    def generate(self, prompt: str, max_new_bytes: int = 100, temperature: float = 1.0) -> str:
        """Sample bytes one at a time (byte_spectral_model.py:163-208)."""
        self.eval()
        device = next(self.parameters()).device
        generated = [ord(c) for c in prompt]
        byte_ids = torch.tensor([generated], dtype=torch.long, device=device)
        for _ in range(max_new_bytes):
            logits = self(byte_ids)
            probs = F.softmax(logits[0, -1, :] / temperature, dim=-1)
            nxt = int(torch.multinomial(probs, num_samples=1).item())
            generated.append(nxt)
            byte_ids = torch.tensor([generated[-self.max_seq_len:]], dtype=torch.long, device=device)
            if nxt == 0 or nxt > 127:
                break
        return "".join(chr(b) if b < 128 else "?" for b in generated)
