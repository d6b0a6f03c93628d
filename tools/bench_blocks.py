"""Block-level timing on one B200: the spectral half of SpectralMLPBlock
(`x + spectral_mix(norm1(x))`, spectral_layers.py:185) fused (LayerNorm on load, residual on store) against the unfused
composition around the same fused layer, forward + backward, CUDA events, inputs larger than L2.  Also FixedSpectralBlock's
spectral half at the reference's default sizes against the reference composition run by PyTorch/cuFFT on the same GPU.
Prints one JSON line per row."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensor_cuda_fft_b200 as pkg                                    # noqa: E402
from tensor_cuda_fft_b200 import spectral_conv as sc                  # noqa: E402
from tensor_cuda_fft_b200 import spectral_layers as sl                # noqa: E402


def time_ms(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--peak", type=float, default=6554.2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    esz = 4 if dt == torch.float32 else 2
    torch.manual_seed(0)
    # ---- f-1: SpectralMLPBlock spectral half at BASELINE cfg-2
    B, T, D = 16, 8192, 768
    norm = torch.nn.LayerNorm(D).to(dev)
    layer = pkg.SpectralMixingLayer(D).to(dev)
    with torch.no_grad():
        layer.weight_real.normal_(); layer.weight_imag.normal_(); layer.bias.normal_()
    x = torch.randn(B, T, D, device=dev, dtype=dt)
    g = torch.randn(B, T, D, device=dev, dtype=dt)
    params = list(norm.parameters()) + list(layer.parameters())

    def step(fused):
        def run():
            for p in params:
                p.grad = None
            xr = x.detach().requires_grad_(True)
            y = sl.ln_spectral_mix_residual(xr, norm, layer) if fused else xr + layer(norm(xr))
            y.backward(g)
        return run

    def fwd(fused):
        def run():
            with torch.no_grad():
                return sl.ln_spectral_mix_residual(x, norm, layer) if fused else x + layer(norm(x))
        return run

    act = B * T * D * esz
    for name, fused in (("fused", True), ("unfused", False)):
        ms = time_ms(step(fused), args.steps, args.warmup)
        ms_f = time_ms(fwd(fused), args.steps, args.warmup)
        # floor of the block half: read x, write y (forward); read g, read x (LayerNorm backward), write gx (backward) = 5 passes
        print(json.dumps({"row": "f-1 SpectralMLPBlock spectral half", "variant": name, "shape": [B, T, D], "dtype": args.dtype,
                          "fwd_ms": round(ms_f, 4), "fwd_bwd_ms": round(ms, 4), "tokens_per_s": B * T / (ms * 1e-3),
                          "floor_passes": 5, "frac_of_hbm_roofline": 5 * act / (ms * 1e-3) / 1e9 / args.peak}))
    # ---- f-2: FixedSpectralBlock spectral half at the reference's default sizes (seq 1024, kernel 128, d_model 512)
    B, T, C, K = 64, 1024, 512, 128
    blk = sc.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=16, dropout=0.0).to(dev).eval()
    with torch.no_grad():
        blk.kernel.normal_(std=0.1)
    x2 = torch.randn(B, T, C, device=dev, dtype=dt)
    g2 = torch.randn(B, T, C, device=dev, dtype=dt)

    def ours():
        for p in blk.parameters():
            p.grad = None
        xr = x2.detach().requires_grad_(True)
        blk.spectral_half(xr).backward(g2)

    def ref_composition():            # the reference's ops (train_fixed_full.py:498-555) executed by PyTorch / cuFFT on this GPU
        for p in blk.parameters():
            p.grad = None
        xr = x2.detach().float().requires_grad_(True)
        xn = blk.ln(xr)
        n = sc.conv_fft_len(T, K)
        k = torch.zeros(n, device=dev)
        k[:K] = blk.kernel
        y_freq = torch.fft.rfft(torch.nn.functional.pad(xn, (0, 0, 0, n - T)), dim=1) * torch.fft.rfft(k).view(1, -1, 1) * blk.gain.view(1, 1, -1)
        g_freq = torch.sigmoid(blk.gate_freq_logits[: n // 2 + 1])
        g_ctx = torch.sigmoid(blk.gate_ctx(xn.mean(dim=1)))
        y_freq = y_freq * g_freq.view(1, -1, 1) * g_ctx.unsqueeze(1)
        y = xr + torch.fft.irfft(y_freq, n=n, dim=1)[:, :T, :]
        y.backward(g2.float())

    with torch.no_grad():
        fo = time_ms(lambda: blk.spectral_half(x2), args.steps, args.warmup)
    mo = time_ms(ours, args.steps, args.warmup)
    mr = time_ms(ref_composition, max(args.steps // 5, 3), 2)
    try:
        gh = blk.graphed(x2, half_only=True)

        def ours_graphed():
            for p in blk.parameters():
                p.grad = None
            xr = x2.detach().requires_grad_(True)
            gh(xr).backward(g2)

        mg = time_ms(ours_graphed, args.steps, args.warmup)
    except Exception as e:
        print("graphed f-2 unavailable:", str(e).splitlines()[0][:200], file=sys.stderr)
        mg = None
    act2 = B * T * C * esz
    print(json.dumps({"row": "f-2 FixedSpectralBlock spectral half", "shape": [B, T, C], "kernel_len": K, "n_fft": sc.conv_fft_len(T, K),
                      "dtype": args.dtype, "fwd_ms": round(fo, 4), "fwd_bwd_ms": round(mo, 4), "fwd_bwd_graphed_ms": None if mg is None else round(mg, 4),
                      "reference_composition_on_gpu_fwd_bwd_ms": round(mr, 4), "tokens_per_s": B * T / (mo * 1e-3),
                      "floor_passes": 5, "frac_of_hbm_roofline": 5 * act2 / (mo * 1e-3) / 1e9 / args.peak}))


if __name__ == "__main__":
    main()
