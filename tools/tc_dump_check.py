#!/usr/bin/env python
"""Bring-up aid for csrc/sml_tc.cuh: runs ONE forward work item (B=1, D=32) on the GPU with SML_TC_DUMP set, and compares every
dumped intermediate (twiddled stage-1 output, band accumulator, band operand, twiddled stage-A output, y) with the numpy model
tests/tc_model.py (bf16 operand rounding on).  usage: python tools/tc_dump_check.py [T] [F]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Fn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
DUMP = f"/tmp/sml_tc_dump_{T}_{Fn}.bin"
os.environ["SML_TC"] = "1"
os.environ["SML_TC_DUMP"] = DUMP
os.environ["SML_DEBUG"] = "1"

import numpy as np
import torch
import tc_model
from tensor_cuda_fft_b200 import _native, spectral_mix

D, N2 = 32, T // 64
torch.manual_seed(T + Fn)
x = torch.randn(1, T, D).bfloat16()
wr, wi, bias = torch.randn(D, Fn), torch.randn(D, Fn), torch.randn(D)
dev = torch.device("cuda:0")
try:
    y = spectral_mix(x.to(dev), wr.to(dev), wi.to(dev), bias.to(dev))
    torch.cuda.synchronize()
except Exception as e:
    print("launch failed:", str(e).splitlines()[0])
    _native.lib().sml_debug_dump()
    sys.exit(2)
inter = {}
y_m, xlow_m, _ = tc_model.transform(x[0].float().double().numpy(), wr.double().numpy(), wi.double().numpy(), bias.double().numpy(),
                                    bf16_ops=True, intermediates=inter)
raw = np.fromfile(DUMP, dtype=np.float32).astype(np.float64)
nV = N2 * 2048
regions = {"V (stage 1 + twiddle)": (raw[:nV].reshape(N2, 32, 64), inter["V"]),
           "Z (band accumulator)": (raw[nV:nV + 36864].reshape(1152, 32)[:1088], inter["Z"]),
           "band operand": (raw[nV + 36864:nV + 2 * 36864].reshape(1152, 32)[:1088], inter["band"]),
           "AB (stage A + twiddle)": (raw[nV + 2 * 36864:].reshape(N2, 32, 64), inter["AB"])}


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


for name, (got, want) in regions.items():
    e = rel(got, want)
    print(f"{name:26s} rel-L2 {e:.3e}  |got| {np.linalg.norm(got):.3e} |want| {np.linalg.norm(want):.3e}  nan {int(np.isnan(got).sum())}")
    if e > 2e-2:
        err = np.abs(got - want)
        # where are the mismatches? summarise per leading index and per last index
        ax_last = err.reshape(-1, err.shape[-1]).max(axis=0)
        ax_first = err.reshape(err.shape[0], -1).max(axis=1)
        tol = 0.05 * np.abs(want).max()
        print("   bad last-axis indices :", np.nonzero(ax_last > tol)[0].tolist()[:64])
        print("   bad first-axis indices:", np.nonzero(ax_first > tol)[0].tolist()[:64])
        if err.ndim == 3:
            ax_mid = err.transpose(1, 0, 2).reshape(err.shape[1], -1).max(axis=1)
            print("   bad middle-axis indices:", np.nonzero(ax_mid > tol)[0].tolist()[:64])
        i = np.unravel_index(np.argmax(err), err.shape)
        print("   worst at", i, "got", got[i], "want", want[i])
yg = y[0].float().cpu().double().numpy()
print(f"{'y':26s} rel-L2 {rel(yg, y_m):.3e}  nan {int(np.isnan(yg).sum())}")
if rel(yg, y_m) > 2e-2:
    err = np.abs(yg - y_m)
    tol = 0.05 * np.abs(y_m).max()
    print("   bad t (first 64):", np.nonzero(err.max(axis=1) > tol)[0].tolist()[:64])
    print("   bad d:", np.nonzero(err.max(axis=0) > tol)[0].tolist())
