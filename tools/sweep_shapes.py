#!/usr/bin/env python
"""Shape sweep on one GPU through the raw C ABI: BASELINE.json configs[2] (cfg-3 geometry) and configs[4] (long-context sweep,
seq 1K-128K at embed 1024, ~2^19 tokens per launch).  Prints a markdown table: launch times, HBM fraction, tokens/s.
usage: python tools/sweep_shapes.py [f32|bf16]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensor_cuda_fft_b200 import _native

dt = sys.argv[1] if len(sys.argv) > 1 else "f32"
dtype, esz, io = (torch.float32, 4, _native.DTYPE_F32) if dt == "f32" else (torch.bfloat16, 2, _native.DTYPE_BF16)
lib = _native.lib()
dev = torch.device("cuda:0")
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
shapes = [(8, 512, 256), (16, 8192, 768), (64, 4096, 1024), (8, 4096, 1024)]
shapes += [(max(1, (1 << 19) // T), T, 1024) for T in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072)]
print(f"| B | T | D | k | plan | fwd ms | fwd HBM frac | bwd ms | bwd HBM frac | fwd+bwd Mtok/s | ({dt} I/O, peak {peak:.0f} GB/s)")
print("|---|---|---|---|---|---|---|---|---|---|")
for (B, T, D) in shapes:
    Fn = D // 2
    k = min(Fn, T // 2)
    plan = _native.plan(B, T, D, Fn, io)
    x = torch.randn(B, T, D, device=dev).to(dtype)
    g = torch.randn(B, T, D, device=dev).to(dtype)
    wr, wi, bs = torch.randn(D, Fn, device=dev), torch.randn(D, Fn, device=dev), torch.randn(D, device=dev)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    xlow = torch.empty(max(lib.sml_xlow_bytes(B, T, D, Fn), 8), dtype=torch.uint8, device=dev)
    gflat = torch.empty(2 * D * Fn + D, device=dev)
    gwr, gwi, gb = gflat[:D * Fn], gflat[D * Fn:2 * D * Fn], gflat[2 * D * Fn:]
    wsb = lib.sml_workspace_bytes(B, T, D, Fn, io)
    ws = torch.empty(max(wsb, 8), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    fwd = lambda: _native.check(lib.sml_forward(x.data_ptr(), wr.data_ptr(), wi.data_ptr(), bs.data_ptr(), y.data_ptr(), xlow.data_ptr(), B, T, D, Fn, io, st))
    bwd = lambda: _native.check(lib.sml_backward(g.data_ptr(), xlow.data_ptr(), wr.data_ptr(), wi.data_ptr(), gx.data_ptr(), gwr.data_ptr(), gwi.data_ptr(), gb.data_ptr(), ws.data_ptr(), wsb, B, T, D, Fn, io, st))

    def timeit(fn, n=30):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / n

    tf, tb = timeit(fwd), timeit(bwd)
    nb = 2 * B * T * D * esz + B * D * k * 8
    print(f"| {B} | {T} | {D} | {k} | {plan['path']} M={plan['M']} R={plan['R']} | {tf:.4f} | {nb / tf / 1e6 / peak:.3f} | {tb:.4f} | {nb / tb / 1e6 / peak:.3f} | {B * T / (tf + tb) / 1e3:.1f} |")
    del x, g, y, gx, xlow, ws
    torch.cuda.empty_cache()
