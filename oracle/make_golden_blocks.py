"""Generate tests/golden/block_*.npz by executing the UNMODIFIED reference in this container (CPU): the block-level callers
of the hot path (SURVEY.md section 8 f-1 SpectralMLPBlock, f-2 FixedSpectralBlock, f-4 overlap-save + SpectralEMA).

Run from the repo root:  python -m oracle.make_golden_blocks
Needs /root/reference (not present on the GPU box -- the fixtures are committed instead).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

REF_ROOT = os.environ.get("SML_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_by_path(modname: str, relpath: str):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def randomize(module, gen, scale=0.5):
    """Move every parameter away from its initial value (LayerNorm bias, gate weights, ... are zero at init)."""
    with torch.no_grad():
        for p in module.parameters():
            p.add_(scale * torch.randn(p.shape, generator=gen))


def sd_numpy(module, prefix="sd."):
    return {prefix + k: v.detach().numpy() for k, v in module.state_dict().items()}


def grads_numpy(module, prefix="grad."):
    return {prefix + k: p.grad.detach().numpy() for k, p in module.named_parameters() if p.grad is not None}


def run_mlp_block(ref_sl, B, T, D, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    blk = ref_sl.SpectralMLPBlock(D, mlp_ratio=2, dropout=0.0)
    randomize(blk, gen)
    blk.eval()
    x = torch.randn(B, T, D, generator=gen).requires_grad_(True)
    g = torch.randn(B, T, D, generator=gen)
    half = x + blk.spectral_mix(blk.norm1(x))          # spectral_layers.py:185
    y = blk(x)
    y.backward(g)
    out = {"x": x.detach().numpy(), "g": g.numpy(), "half": half.detach().numpy(), "y": y.detach().numpy(), "gx": x.grad.numpy()}
    out.update(sd_numpy(blk))
    out.update(grads_numpy(blk))
    return out


def run_fixed_block(tff, B, T, C, K, seed, cutoff=None, trans=4):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    blk = tff.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=trans, dropout=0.0)
    randomize(blk, gen, scale=0.3)
    with torch.no_grad():
        blk.kernel.copy_(0.2 * torch.randn(K, generator=gen))     # a kernel that matters (init std is 1e-3)
    blk.eval()
    x = torch.randn(B, T, C, generator=gen).requires_grad_(True)
    g = torch.randn(B, T, C, generator=gen)
    y = blk(x, cutoff=cutoff)
    y.backward(g)
    out = {"x": x.detach().numpy(), "g": g.numpy(), "y": y.detach().numpy(), "gx": x.grad.numpy(),
           "cutoff": np.int64(-1 if cutoff is None else cutoff), "trans": np.int64(trans), "K": np.int64(K)}
    out.update(sd_numpy(blk))
    out.update(grads_numpy(blk))
    return out


def run_overlap_save(tff, gen_mod, T, C, K, chunk, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    blk = tff.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=4, dropout=0.0)
    randomize(blk, gen, scale=0.3)
    with torch.no_grad():
        blk.kernel.copy_(0.2 * torch.randn(K, generator=gen))
    blk.eval()
    n_fft_full = tff.conv_freq_bins(T, K) * 2 - 2
    h_hist = torch.randn(1, T, C, generator=gen)
    with torch.no_grad():
        ln_in = blk.ln(h_hist)
    state = {"ctx_ln": ln_in.contiguous(), "ctx_sum": ln_in.sum(dim=1).contiguous()}
    out = {"h_hist": h_hist.numpy(), "K": np.int64(K), "n_fft_full": np.int64(n_fft_full), "chunk": np.int64(chunk)}
    for step in range(2):
        h_chunk = torch.randn(1, chunk, C, generator=gen)
        h_out, state = gen_mod.overlap_save_block_update(blk, state, h_chunk, n_fft_full=n_fft_full, kernel_len=K, cache=None)
        out[f"h_chunk{step}"] = h_chunk.numpy()
        out[f"h_out{step}"] = h_out.numpy()
        out[f"ctx_sum{step}"] = state["ctx_sum"].numpy()
    out.update(sd_numpy(blk))
    return out


def run_ema(ssm, seed):
    gen = torch.Generator().manual_seed(seed)
    out = {}
    B, S, Fq = 3, 17, 9
    chunks = torch.complex(torch.randn(B, S, Fq, generator=gen), torch.randn(B, S, Fq, generator=gen))
    chunks[0, 3, 2] = 0            # a zero bin: torch.angle(0) = 0
    chunks[1, 0, :] = 0            # zero first chunk: the state stays zero for one step
    init = torch.complex(torch.randn(B, Fq, generator=gen), torch.randn(B, Fq, generator=gen))
    out["chunks"] = chunks.numpy()
    out["init"] = init.numpy()
    for mode in ("aligned", "polar"):
        ema = ssm.SpectralEMA(ssm.EMAConfig(n_freqs=Fq, rho_init=0.9, theta_init=0.1, mode=mode))
        with torch.no_grad():
            ema.rho_logit.add_(0.5 * torch.randn(Fq, generator=gen))
            ema.theta_raw.add_(0.5 * torch.randn(Fq, generator=gen))
            out[f"{mode}.rho_logit"] = ema.rho_logit.numpy().copy()
            out[f"{mode}.theta_raw"] = ema.theta_raw.numpy().copy()
            out[f"{mode}.scan"] = ema.scan(chunks).numpy()
            out[f"{mode}.scan_init"] = ema.scan(chunks, init=init).numpy()
            out[f"{mode}.update"] = ema.update(init, chunks[:, 5, :]).numpy()
    return out


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    sys.path.insert(0, REF_ROOT)         # fft_lm is a plain package (no import side effects); scripts/ imports it by name
    ref_sl = load_by_path("_ref_spectral_layers", "fft_tensor/spectral_layers.py")
    from fft_lm import train_fixed_full as tff
    from fft_lm import spectral_ssm as ssm
    # generate_chunked_overlap_save.py imports fft_lm.chunk_head / ckpt_io at module level; only its pure function is used
    gen_mod = load_by_path("_ref_overlap_save", "scripts/generate_chunked_overlap_save.py")
    save = lambda name, d: (np.savez_compressed(os.path.join(OUT_DIR, name), **d), print("wrote", name, len(d), "arrays"))
    save("block_mlp_t256_d64.npz", run_mlp_block(ref_sl, 2, 256, 64, seed=11))              # fused kernels, M = 64 sub-transform
    save("block_mlp_t1024_d48.npz", run_mlp_block(ref_sl, 2, 1024, 48, seed=12))            # R > 1 (streamed passes)
    save("block_mlp_t100_d32.npz", run_mlp_block(ref_sl, 2, 100, 32, seed=13))              # generic path: unfused composition
    save("block_fixed_t64_k16_c32.npz", run_fixed_block(tff, 2, 64, 32, 16, seed=21))       # n_fft = 128
    save("block_fixed_t96_k24_c16.npz", run_fixed_block(tff, 3, 96, 16, 24, seed=22))       # n_fft = 128, T not a power of two
    save("block_fixed_t512_k128_c32_cut.npz", run_fixed_block(tff, 1, 512, 32, 128, seed=23, cutoff=200, trans=16))   # n_fft = 1024, cutoff mask
    save("block_fixed_t1024_k128_c16.npz", run_fixed_block(tff, 1, 1024, 16, 128, seed=24))  # the reference's default sizes: n_fft = 2048
    save("block_overlap_save.npz", run_overlap_save(tff, gen_mod, 256, 32, 32, 16, seed=31))
    save("block_spectral_ema.npz", run_ema(ssm, seed=41))


if __name__ == "__main__":
    main()
