"""world_size-2 gloo test of the batch-sharding plumbing (no GPU): shard_batch partitions the batch and one flat
all-reduce of [weight_real.grad | weight_imag.grad | bias.grad] reproduces the single-process gradient sum."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, flat_layout=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensor_cuda_fft_b200.distributed import allreduce_filter_grads, shard_batch

    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.weight_real = torch.nn.Parameter(torch.zeros(4, 3))
            self.weight_imag = torch.nn.Parameter(torch.zeros(4, 3))
            self.bias = torch.nn.Parameter(torch.zeros(4))

    gen = torch.Generator().manual_seed(0)
    full = torch.randn(5, 7, 4, generator=gen)           # ragged: 5 samples over 2 ranks -> 3 + 2
    mine = shard_batch(full, rank, world)
    m = Fake()
    # a stand-in "local gradient" that is a sum over the local batch
    m.weight_real.grad = mine.sum(dim=(0, 1)).unsqueeze(1).repeat(1, 3)
    m.weight_imag.grad = 2 * m.weight_real.grad
    m.bias.grad = mine.sum(dim=(0, 1))
    if flat_layout:     # the layout the fused backward hands out: three views of one [gw_re | gw_im | gb] buffer
        flat = torch.cat([m.weight_real.grad.reshape(-1), m.weight_imag.grad.reshape(-1), m.bias.grad])
        m.weight_real.grad, m.weight_imag.grad, m.bias.grad = flat[:12].view(4, 3), flat[12:24].view(4, 3), flat[24:]
        ptr = flat.data_ptr()
    allreduce_filter_grads([m])
    if flat_layout:     # reduced in place: no concatenation, no copy back
        assert m.weight_real.grad.data_ptr() == ptr and m.bias.grad.data_ptr() == ptr + 24 * 4
    # plain lists: tensors in a multiprocessing queue travel as shared file descriptors, which die with the worker
    q.put((rank, mine.shape[0], m.weight_real.grad.tolist(), m.weight_imag.grad.tolist(), m.bias.grad.tolist()))
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("flat_layout", [False, True])
def test_sharded_filter_grad_allreduce(flat_layout):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, flat_layout)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gen = torch.Generator().manual_seed(0)
    full = torch.randn(5, 7, 4, generator=gen)
    want_b = full.sum(dim=(0, 1))
    assert [r[1] for r in res] == [3, 2]
    for _, _, gwr, gwi, gb in res:
        gwr, gwi, gb = torch.tensor(gwr), torch.tensor(gwi), torch.tensor(gb)
        assert torch.allclose(gb, want_b, atol=1e-5)
        assert torch.allclose(gwr, want_b.unsqueeze(1).repeat(1, 3), atol=1e-5)
        assert torch.allclose(gwi, 2 * want_b.unsqueeze(1).repeat(1, 3), atol=1e-5)


def test_symmetric_grad_bucket_layout_cpu():
    """SymmetricGradBucket hands every module a slice of ONE flat buffer laid out [gw_re | gw_im | gb] per module, back to back,
    in two alternating parities (the buffers are NVLink symmetric memory on a GPU job; here plain tensors through the test
    allocator)."""
    import torch
    from tensor_cuda_fft_b200 import SpectralMixingLayer
    from tensor_cuda_fft_b200.distributed import SymmetricGradBucket
    mods = [SpectralMixingLayer(8), SpectralMixingLayer(16, num_filters=4), SpectralMixingLayer(6, learnable=False)]
    bucket = SymmetricGradBucket(mods, group=None, allocator=lambda n, dev: torch.zeros(n, device=dev))
    assert bucket.numel == (2 * 8 * 4 + 8) + (2 * 16 * 4 + 16)
    assert mods[0]._grad_bucket == (bucket, 0) and mods[1]._grad_bucket == (bucket, 1) and not hasattr(mods[2], "_grad_bucket")
    cur, mc, nxt = bucket.slot(1)
    assert cur.numel() == 144 and nxt.numel() == 144 and mc == 0
    assert cur.data_ptr() == bucket.bufs[0].data_ptr() + 72 * 4 and nxt.data_ptr() == bucket.bufs[1].data_ptr() + 72 * 4
    views = [bucket.slot(0)[0][:32].view(8, 4), cur[128:]]
    assert bucket.covers(views) and not bucket.covers([torch.zeros(3)])
    bucket.parity ^= 1
    assert bucket.slot(1)[0].data_ptr() == bucket.bufs[1].data_ptr() + 72 * 4 and not bucket.covers(views)
