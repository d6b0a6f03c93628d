#!/usr/bin/env python
"""GPU diagnostic: run the warp-specialised kernel (SML_FAST_WS=1) against the lockstep kernel on multi-tile shapes.
On a trap, print the mbarrier-timeout record (SML_DEBUG=1).  usage: SML_DEBUG=1 python tools/ws_debug.py"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SML_DEBUG", "1")
import torch
from tensor_cuda_fft_b200 import _native

lib = _native.lib()
dev = torch.device("cuda:0")
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1, 1024, 768), (4, 1024, 768), (4, 2048, 768), (16, 8192, 768)]
for (B, T, D) in shapes:
    Fn = D // 2
    torch.manual_seed(0)
    x = torch.randn(B, T, D, device=dev)
    g = torch.randn(B, T, D, device=dev)
    wr, wi, bs = torch.randn(D, Fn, device=dev), torch.randn(D, Fn, device=dev), torch.randn(D, device=dev)
    res = {}
    for ws in ("0", "1"):
        os.environ["SML_FAST_WS"] = ws
        y, gx = torch.empty_like(x), torch.empty_like(x)
        xlow = torch.empty(lib.sml_xlow_bytes(B, T, D, Fn), dtype=torch.uint8, device=dev)
        gwr, gwi, gb = torch.empty(D, Fn, device=dev), torch.empty(D, Fn, device=dev), torch.empty(D, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        wsb = torch.empty(max(lib.sml_workspace_bytes(B, T, D, Fn, 0), 1), dtype=torch.uint8, device=dev)
        try:
            _native.check(lib.sml_forward(x.data_ptr(), wr.data_ptr(), wi.data_ptr(), bs.data_ptr(), y.data_ptr(), xlow.data_ptr(), B, T, D, Fn, 0, st))
            torch.cuda.synchronize()
            print(f"{(B, T, D)} ws={ws} fwd ok", flush=True)
            _native.check(lib.sml_backward(g.data_ptr(), xlow.data_ptr(), wr.data_ptr(), wi.data_ptr(), gx.data_ptr(), gwr.data_ptr(), gwi.data_ptr(), gb.data_ptr(), wsb.data_ptr(), wsb.numel(), B, T, D, Fn, 0, st))
            torch.cuda.synchronize()
            print(f"{(B, T, D)} ws={ws} bwd ok", flush=True)
        except Exception as e:
            print(f"{(B, T, D)} ws={ws} FAILED: {str(e).splitlines()[0]}", flush=True)
            lib.sml_debug_dump()
            sys.exit(1)
        res[ws] = (y, gx, gwr, gwi, gb)
    for name, a, b in zip(("y", "gx", "gw_re", "gw_im", "gb"), res["0"], res["1"]):
        err = ((a - b).norm() / b.norm()).item()
        print(f"   {name}: rel-L2(ws1 vs ws0) = {err:.3e}")
        if err > 1e-5 and name in ("y", "gx"):
            # localise: which (b, 16-channel tile) and which row residue r = t mod R differ
            R = max(T // 1024, 1)
            d = (a - b).float().view(B, T // R, R, D // 16, 16)
            per_tile = d.pow(2).sum(dim=(1, 2, 4)).sqrt()          # (B, ntd)
            bad = (per_tile > 1e-3 * b.float().norm() / (B * D / 16) ** 0.5).nonzero().tolist()
            print(f"      bad tiles (b, dt): {bad[:40]} ... total {len(bad)} of {B * D // 16}")
            for (bb, dt) in bad[:4]:
                e = d[bb, :, :, dt, :]                              # (T/R, R, 16)
                print(f"      tile b={bb} dt={dt} linear={bb * (D // 16) + dt}: err by r = {e.pow(2).sum(dim=(0, 2)).sqrt().tolist()},"
                      f" by channel = {[round(v, 3) for v in e.pow(2).sum(dim=(0, 1)).sqrt().tolist()]}")
                m = e.pow(2).sum(dim=(1, 2)).sqrt()                 # by m (row within pass)
                nz = (m > 1e-4).nonzero().flatten()
                print(f"         rows m with error: count {nz.numel()} first {nz[:8].tolist()} last {nz[-8:].tolist()}")
print("ws_debug done")
