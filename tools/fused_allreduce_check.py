#!/usr/bin/env python
"""Parity of the fused reduce + multimem all-reduce (sml_backward_allreduce through SymmetricGradBucket) under torchrun:
every rank runs fwd+bwd on its own batch shard; the summed filter/bias gradients must equal (a) the NCCL all-reduce of the plain
backward and (b) the single-process gradient on the concatenated batch.  Several steps, so both buffer parities are exercised.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/fused_allreduce_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tensor_cuda_fft_b200 import SpectralMixingLayer, allreduce_filter_grads, attach_symmetric_grad_buffers
from tensor_cuda_fft_b200 import distributed as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


ok = True
for (B, T, Dm, dtype) in [(4, 2048, 96, torch.float32), (2, 8192, 768, torch.float32), (3, 1024, 64, torch.bfloat16)]:
    torch.manual_seed(11)
    ref = SpectralMixingLayer(Dm).to(dev)
    with torch.no_grad():
        ref.weight_real.normal_(); ref.weight_imag.normal_(); ref.bias.normal_()
    fused = SpectralMixingLayer(Dm).to(dev)
    fused.load_state_dict(ref.state_dict())
    bucket = attach_symmetric_grad_buffers([fused])
    for step in range(3):
        torch.manual_seed(100 * step + 7)
        xg = torch.randn(world * B, T, Dm, device=dev).to(dtype)       # the same global batch on every rank
        gg = torch.randn(world * B, T, Dm, device=dev).to(dtype)
        x, g = xg[rank * B:(rank + 1) * B], gg[rank * B:(rank + 1) * B]
        ref.zero_grad(set_to_none=True); fused.zero_grad(set_to_none=True)
        ref(x.clone().requires_grad_(True)).backward(g)
        flat = torch.cat([ref.weight_real.grad.reshape(-1), ref.weight_imag.grad.reshape(-1), ref.bias.grad])
        dist.all_reduce(flat)
        fused(x.clone().requires_grad_(True)).backward(g)
        allreduce_filter_grads([fused])
        got = torch.cat([fused.weight_real.grad.reshape(-1), fused.weight_imag.grad.reshape(-1), fused.bias.grad])
        torch.cuda.synchronize()
        e1 = rel(got, flat)
        # single-process gradient on the concatenated batch
        ref.zero_grad(set_to_none=True)
        ref(xg.clone().requires_grad_(True)).backward(gg)
        whole = torch.cat([ref.weight_real.grad.reshape(-1), ref.weight_imag.grad.reshape(-1), ref.bias.grad])
        e2 = rel(got, whole)
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        good = e1 <= tol and e2 <= tol
        ok = ok and good
        if rank == 0:
            print(f"shape {(B, T, Dm)} {str(dtype)[6:]} step {step}: path '{D.LAST_ALLREDUCE_PATH}', vs NCCL sum {e1:.2e}, vs single process {e2:.2e} -> {'ok' if good else 'MISMATCH'}", flush=True)
    D._BUCKETS.clear()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
