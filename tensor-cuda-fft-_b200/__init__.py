"""spectral-mix-b200: B200-native (sm_100a) SpectralMixingLayer forward/backward of FFT-Tensor.

Public surface mirrors the reference modules ``fft_tensor.spectral_layers`` and ``fft_tensor.wirtinger_ops``.
"""
from . import _native
from .spectral_layers import (HybridSpectralAttention, SpectralMLPBlock, SpectralMixingLayer, spectral_mix,
                              spectral_mix_fwd_bwd_host)
from .wirtinger_ops import (ComplexParameter, WirtingerGradient, WirtingerSpectralFilter)
from .byte_spectral_model import ByteSpectralEmbedding, SpectralLanguageModel
from .distributed import allreduce_filter_grads, attach_symmetric_grad_buffers, shard_batch
from .spectral_conv import EMAConfig, FixedSpectralBlock, SpectralEMA, overlap_save_block_update

__version__ = "0.1.0"
__all__ = [
    "SpectralMixingLayer", "SpectralMLPBlock", "HybridSpectralAttention", "spectral_mix", "spectral_mix_fwd_bwd_host",
    "WirtingerGradient", "ComplexParameter", "WirtingerSpectralFilter",
    "ByteSpectralEmbedding", "SpectralLanguageModel",
    "allreduce_filter_grads", "attach_symmetric_grad_buffers", "shard_batch",
    "FixedSpectralBlock", "overlap_save_block_update", "EMAConfig", "SpectralEMA",
]
