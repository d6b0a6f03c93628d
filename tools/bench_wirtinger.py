"""Stand-alone Wirtinger kernels on one B200: CUDA-event time and achieved HBM GB/s against the measured
peak.  Shapes: the (B, k, D) live slab and the full (B, T, D) complex spectrum of BASELINE cfg-2.  One JSON line per kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensor_cuda_fft_b200 import _native                      # noqa: E402

PEAK = 6554.2


def time_ms(fn, steps=30, warmup=5):
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
    lib = _native.lib()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    B, T, D, Fn = 16, 8192, 768, 384
    for name, shape in (("slab (16, 384*768)", (16, 384 * 768)), ("batch 64 (64, 512*1024)", (64, 512 * 1024))):
        Bm, N = shape
        x = torch.randn(Bm, N, dtype=torch.complex64, device=dev)
        g = torch.randn(Bm, N, dtype=torch.complex64, device=dev)
        w = torch.randn(N, dtype=torch.complex64, device=dev)
        out, gx, gw = torch.empty_like(x), torch.empty_like(x), torch.empty_like(w)
        ms = time_ms(lambda: _native.check(lib.sml_wirtinger_mul_forward(x.data_ptr(), w.data_ptr(), out.data_ptr(), Bm, N, st)))
        nbytes = 2 * x.numel() * 8 + N * 8
        rows.append({"kernel": "wirtinger_mul_forward", "shape": name, "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / PEAK})
        ms = time_ms(lambda: _native.check(lib.sml_wirtinger_mul_backward(g.data_ptr(), x.data_ptr(), w.data_ptr(), gx.data_ptr(), gw.data_ptr(), Bm, N, st)))
        nbytes = 3 * x.numel() * 8 + 2 * N * 8
        rows.append({"kernel": "wirtinger_mul_backward", "shape": name, "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / PEAK})
        del x, g, w, out, gx, gw
    xf = torch.randn(B, T, D, dtype=torch.complex64, device=dev)
    gf = torch.randn(B, T, D, dtype=torch.complex64, device=dev)
    w_re, w_im = torch.randn(D, Fn, device=dev), torch.randn(D, Fn, device=dev)
    out = torch.empty_like(xf)
    gwr, gwi = torch.empty(D, Fn, device=dev), torch.empty(D, Fn, device=dev)
    ms = time_ms(lambda: _native.check(lib.sml_wirtinger_filter_forward(xf.data_ptr(), w_re.data_ptr(), w_im.data_ptr(), out.data_ptr(), B, T, D, Fn, st)))
    k = min(Fn, T // 2)
    nbytes = B * k * D * 8 + B * T * D * 8           # read the live slab, write the whole (zero-filled) spectrum
    rows.append({"kernel": "wirtinger_filter_forward", "shape": "(16, 8192, 768) c64, k = 384", "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / PEAK})
    ms = time_ms(lambda: _native.check(lib.sml_wirtinger_filter_backward(gf.data_ptr(), xf.data_ptr(), w_re.data_ptr(), w_im.data_ptr(), out.data_ptr(),
                                                                           gwr.data_ptr(), gwi.data_ptr(), B, T, D, Fn, st)))
    nbytes = 2 * B * k * D * 8 + B * T * D * 8
    rows.append({"kernel": "wirtinger_filter_backward", "shape": "(16, 8192, 768) c64, k = 384", "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / PEAK})
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
