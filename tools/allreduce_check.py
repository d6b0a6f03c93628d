#!/usr/bin/env python
"""N-GPU check (torchrun): the symmetric-memory all-reduce of the filter gradients equals NCCL's, and how long each takes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tensor_cuda_fft_b200 import distributed as d

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


class Fake(torch.nn.Module):
    def __init__(self, D, F):
        super().__init__()
        self.weight_real = torch.nn.Parameter(torch.zeros(D, F, device=dev))
        self.weight_imag = torch.nn.Parameter(torch.zeros(D, F, device=dev))
        self.bias = torch.nn.Parameter(torch.zeros(D, device=dev))


for (D, F) in [(768, 384), (1024, 512), (32, 16)]:
    m = Fake(D, F)
    res = {}
    for mode in ("nccl", "symm"):
        os.environ["SML_ALLREDUCE"] = mode
        torch.manual_seed(100 + rank)
        flat = torch.randn(2 * D * F + D, device=dev)
        m.weight_real.grad, m.weight_imag.grad, m.bias.grad = flat[:D * F].view(D, F), flat[D * F:2 * D * F].view(D, F), flat[2 * D * F:]
        d.allreduce_filter_grads([m])
        torch.cuda.synchronize()
        res[mode] = (flat.clone(), d.LAST_ALLREDUCE_PATH)
        # timing
        for _ in range(5):
            d.allreduce_filter_grads([m])
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            d.allreduce_filter_grads([m])
        e1.record(); torch.cuda.synchronize()
        res[mode] += (e0.elapsed_time(e1) / 50 * 1e3,)
    err = ((res["symm"][0] - res["nccl"][0]).norm() / res["nccl"][0].norm()).item()
    if rank == 0:
        print(f"D={D} F={F} bytes={4 * (2 * D * F + D)}: nccl [{res['nccl'][1]}] {res['nccl'][2]:.1f} us | symm [{res['symm'][1]}] {res['symm'][2]:.1f} us | rel diff {err:.2e}", flush=True)
    assert err < 1e-6, err
if rank == 0 and not d._SymmetricAllReduce._cache:
    print("no symmetric path was created")
for v in d._SymmetricAllReduce._cache.values():
    if not v.ok and rank == 0:
        print("symmetric setup failed:", v.err)
dist.destroy_process_group()
