#!/usr/bin/env python
"""Training-step benchmark of the LM that hosts the fused layer (BASELINE.json configs[3], SURVEY.md D5 option (a)):
SpectralLanguageModel (reference byte_spectral_model.py:105-161), random init, synthetic byte tokens, seq 2048, AdamW,
bf16 autocast for the MLPs (LayerNorm keeps the spectral layer's input in fp32), DistributedDataParallel over NCCL.

  python tools/lm_train_step.py [--embed 768 --layers 6 --seq 2048 --batch 8 --steps 20 --warmup 5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/lm_train_step.py ...

Prints one JSON line (rank 0): tokens/s over all ranks (CUDA events, max over ranks) and the share of GPU kernel time spent in
this library's kernels during one profiled step."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--embed", type=int, default=768)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=8, help="per GPU")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-amp", action="store_true")
    args = ap.parse_args()
    from tensor_cuda_fft_b200 import SpectralLanguageModel, _native
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    model = SpectralLanguageModel(embed_dim=args.embed, num_layers=args.layers, max_seq_len=args.seq, dropout=0.0).to(dev)
    nparams = sum(p.numel() for p in model.parameters())
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(net.parameters(), lr=3e-4, fused=True)
    gen = torch.Generator(device=dev).manual_seed(rank)
    ids = torch.randint(0, 256, (args.batch, args.seq), device=dev, generator=gen)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.no_amp):
            logits = net(ids)
            loss = F.cross_entropy(logits[:, :-1].reshape(-1, 256).float(), ids[:, 1:].reshape(-1))
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    n0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _native.launch_count() - n0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    share = None
    try:    # one profiled step: share of GPU kernel time spent in this library's kernels
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        tot = ours = 0.0
        for ev in prof.key_averages():
            dt = float(getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0))
            tot += dt
            if "sml" in ev.key or "filtergrad" in ev.key:
                ours += dt
        share = ours / tot if tot > 0 else None
    except Exception as e:   # pragma: no cover
        share = None
        print("profiler unavailable:", e, file=sys.stderr)
    if rank == 0:
        print(json.dumps({
            "metric": "spectral_lm_train_tokens_per_sec", "value": world * args.batch * args.seq / (ms * 1e-3), "unit": "tokens/s",
            "n_gpus": world, "ms_per_step": ms, "loss": float(loss.item()), "steps": args.steps, "warmup": args.warmup,
            "config": {"model": "SpectralLanguageModel (byte_spectral_model.py:105)", "embed_dim": args.embed, "layers": args.layers,
                       "seq_len": args.seq, "batch_per_gpu": args.batch, "params": nparams, "amp": "bf16 autocast" if not args.no_amp else "fp32",
                       "optimizer": "AdamW(fused)", "data": "synthetic random bytes", "parallelism": f"DDP x{world}"},
            "library_kernel_launches_per_step": launches / args.steps, "library_share_of_gpu_time": share}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
