#!/usr/bin/env python
"""Randomised GPU stress: many random (B, T, D, F, dtype) problems through the module API, each checked against the reference
ALGORITHM (torch.fft composition of fft_tensor/spectral_layers.py:88-116 with autograd) run in fp32 on the same GPU.
usage: python tools/stress.py [n_cases] [seed]"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensor_cuda_fft_b200 import SpectralMixingLayer, _native


def reference(x, w_re, w_im, bias, g):
    x = x.detach().float().requires_grad_(True)
    ps = [p.detach().clone().requires_grad_(True) for p in (w_re, w_im, bias)]
    spec = torch.fft.fft(x, dim=1)
    k = min(ps[0].shape[1], x.shape[1] // 2)
    kept = torch.zeros_like(spec)
    kept[:, :k, :] = spec[:, :k, :] * torch.complex(ps[0], ps[1])[:, :k].T.unsqueeze(0)
    y = torch.fft.ifft(kept, dim=1).real + ps[2]
    y.backward(g.float())
    return y.detach(), x.grad, ps[0].grad, ps[1].grad, ps[2].grad


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = random.Random(seed)
    dev = torch.device("cuda:0")
    t_start = time.time()
    worst = {}
    paths = {}
    for case in range(n):
        fam = rng.choice(["pow2", "mult64", "mult256", "mult1024", "odd", "tiny"])
        if fam == "pow2":
            T = 2 ** rng.randint(6, 14)
        elif fam == "mult64":
            T = 64 * rng.randint(1, 40)
        elif fam == "mult256":
            T = 256 * rng.randint(1, 24)
        elif fam == "mult1024":
            T = 1024 * rng.randint(1, 12)
        elif fam == "odd":
            T = rng.randint(2, 700)
        else:
            T = rng.randint(1, 40)
        D = rng.choice([4, 8, 12, 16, 24, 32, 40, 64, 72, 96, 128, 200, 256, 384, 512, 768, 1024, 1536, 2048, 7, 30])
        F = rng.choice([None, None, None, max(1, D // 4), D, rng.randint(1, max(1, min(2 * D, 1100)))])
        Fn = F or max(D // 2, 1)
        # keep the problem small enough: <= ~64 M elements
        maxB = max(1, min(40, (1 << 26) // (T * D)))
        B = rng.randint(1, maxB)
        dtype = rng.choice([torch.float32, torch.float32, torch.bfloat16])
        io = _native.DTYPE_F32 if dtype == torch.float32 else _native.DTYPE_BF16
        plan = _native.plan(B, T, D, Fn, io)
        if plan["path"] == "generic" and T * plan["k"] * B * D > 6e10:      # bound the direct-DFT cost
            continue
        gen = torch.Generator(device=dev).manual_seed(seed * 100003 + case)
        layer = SpectralMixingLayer(D, num_filters=Fn).to(dev)
        with torch.no_grad():
            layer.weight_real.copy_(torch.randn(D, Fn, device=dev, generator=gen))
            layer.weight_imag.copy_(torch.randn(D, Fn, device=dev, generator=gen))
            layer.bias.copy_(torch.randn(D, device=dev, generator=gen))
        x = torch.randn(B, T, D, device=dev, generator=gen).to(dtype)
        g = torch.randn(B, T, D, device=dev, generator=gen).to(dtype)
        xr = x.clone().requires_grad_(True)
        y = layer(xr)
        y.backward(g)
        want = reference(x, layer.weight_real, layer.weight_imag, layer.bias, g)
        got = (y.detach(), xr.grad, layer.weight_real.grad, layer.weight_imag.grad, layer.bias.grad)
        tol = 2e-5 if dtype == torch.float32 else 1e-2
        key = f"{plan['path']} M={plan['M']}"
        paths[key] = paths.get(key, 0) + 1
        for name, a, b in zip(("y", "gx", "gw_re", "gw_im", "gb"), got, want):
            e = rel(a, b)
            if name == "gw_im":     # the DC column of weight_imag.grad is identically zero: with k = 1 only rounding noise is left
                e = ((a - b).norm() / max(b.norm().item(), 1e-2 * want[2].norm().item(), 1e-30)).item()
            worst[(str(dtype), name)] = max(worst.get((str(dtype), name), 0.0), e)
            if not (e <= tol):
                print(f"FAIL case {case}: B={B} T={T} D={D} F={Fn} {dtype} plan={plan} {name} rel={e:.3e}", flush=True)
                sys.exit(1)
        del layer, x, g, xr, y, want, got
    torch.cuda.synchronize()
    print(f"stress ok: {n} cases in {time.time() - t_start:.1f} s; plans {paths}")
    for k in sorted(worst):
        print("  worst", k, f"{worst[k]:.2e}")


if __name__ == "__main__":
    main()
