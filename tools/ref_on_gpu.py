#!/usr/bin/env python
"""Context number (not a bench arm): the reference ALGORITHM -- torch.fft.fft / complex filter / torch.fft.ifft(...).real + bias with
autograd, i.e. what fft_tensor/spectral_layers.py:88-116 executes -- run unchanged on the same B200 through cuFFT, next to this
library.  Prints one JSON line.  usage: python tools/ref_on_gpu.py [--batch 16 --seq 8192 --embed 768 --dtype f32]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def reference_forward(x, w_re, w_im, bias):          # spectral_layers.py:88-116, line for line
    T = x.shape[1]
    x_freq = torch.fft.fft(x, dim=1)
    k = min(w_re.shape[1], T // 2)
    w = torch.complex(w_re, w_im)
    filt = torch.zeros_like(x_freq)
    filt[:, :k, :] = x_freq[:, :k, :] * w[:, :k].T.unsqueeze(0)
    return torch.fft.ifft(filt, dim=1).real + bias


def timed(fn, steps, warmup):
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
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seq", type=int, default=8192)
    ap.add_argument("--embed", type=int, default=768)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    from tensor_cuda_fft_b200 import SpectralMixingLayer
    dev = torch.device("cuda:0")
    B, T, D = args.batch, args.seq, args.embed
    torch.manual_seed(0)
    layer = SpectralMixingLayer(D).to(dev)
    with torch.no_grad():
        layer.weight_real.normal_(); layer.weight_imag.normal_(); layer.bias.normal_()
    x = torch.randn(B, T, D, device=dev)
    g = torch.randn(B, T, D, device=dev)
    params = [p.detach().clone().requires_grad_(True) for p in (layer.weight_real, layer.weight_imag, layer.bias)]

    def ref_step():
        for p in params:
            p.grad = None
        xr = x.detach().requires_grad_(True)
        reference_forward(xr, *params).backward(g)
        return xr.grad

    def our_step():
        layer.zero_grad(set_to_none=True)
        xr = x.detach().requires_grad_(True)
        layer(xr).backward(g)
        return xr.grad

    gx_ref, gx_our = ref_step(), our_step()
    rel = ((gx_ref - gx_our).norm() / gx_ref.norm()).item()
    relw = ((params[0].grad - layer.weight_real.grad).norm() / params[0].grad.norm()).item()
    torch.cuda.reset_peak_memory_stats()
    ms_ref = timed(ref_step, args.steps, 3)
    mem_ref = torch.cuda.max_memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    ms_our = timed(our_step, args.steps, 3)
    mem_our = torch.cuda.max_memory_allocated()
    print(json.dumps({"shape": [B, T, D], "reference_algorithm_on_gpu": {"backend": "torch.fft (cuFFT) + ATen elementwise, autograd", "ms_per_step": ms_ref,
                      "tokens_per_s": B * T / ms_ref * 1e3, "peak_mem_mb": mem_ref / 2 ** 20},
                      "this_library": {"ms_per_step": ms_our, "tokens_per_s": B * T / ms_our * 1e3, "peak_mem_mb": mem_our / 2 ** 20},
                      "speedup": ms_ref / ms_our, "rel_l2_gx_vs_cufft_fp32": rel, "rel_l2_gw_re_vs_cufft_fp32": relw}))


if __name__ == "__main__":
    main()
