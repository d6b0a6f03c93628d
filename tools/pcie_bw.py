"""Pinned-memory H2D / D2H / simultaneous copy bandwidth of this box -- the ceiling of bench.py's `e2e` leg.
Single process: python tools/pcie_bw.py.  All GPUs at once (what bounds the 8-rank e2e number):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bw.py [--no-bind]
Each rank binds to its GPU's NUMA node first (tensor_cuda_fft_b200.distributed.bind_to_gpu_numa_node) unless --no-bind."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bind = {"how": "not bound"}
if "--no-bind" not in sys.argv:
    from tensor_cuda_fft_b200.distributed import bind_to_gpu_numa_node
    bind = bind_to_gpu_numa_node(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 403 * 2**20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(device_ids=[local])
def t(fn, reps=5):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    sync(); return (time.perf_counter() - t0) / reps
h2d = t(lambda: d_in.copy_(h_in, non_blocking=True))
d2h = t(lambda: h_out.copy_(d_out, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
bi = t(both)
print(f"rank {rank}/{world} [{bind.get('how')}; cpus {bind.get('cpus')}]: H2D {n/h2d/1e9:.1f} GB/s, D2H {n/d2h/1e9:.1f} GB/s, "
      f"simultaneous {n/bi/1e9:.1f} GB/s per direction (all {world} ranks copying at once)", flush=True)
if world > 1: dist.destroy_process_group()
