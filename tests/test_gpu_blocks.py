"""GPU parity of the block-level callers of the hot path through the extended fused
kernels (sml_forward_ext / sml_backward_ext, sml_ln_stats / sml_ln_backward, sml_spectral_ema_scan):
 (1) against fixtures generated from the UNMODIFIED reference (oracle/make_golden_blocks.py),
 (2) against the oracle restatement / the unfused composition on seeded inputs, incl. BASELINE's cfg-2 size.
Tolerance: rel-L2 <= 1e-5 fp32 (gradients of deep chains 2e-5), <= 1e-2 bf16 I/O."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import block_oracle as bo
from oracle.spectral_mixing_oracle import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def pkg():
    import tensor_cuda_fft_b200 as p
    from tensor_cuda_fft_b200 import _native
    _native.lib()
    return p


def load(name):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, name)).items()}


def state_dict_of(d):
    return {k[3:]: v for k, v in d.items() if k.startswith("sd.")}


def check_grads(module, d, tol):
    for k, p in module.named_parameters():
        want = d["grad." + k].numpy()
        got = p.grad.detach().float().cpu().numpy()
        assert rel_l2(got, want) <= tol, (k, rel_l2(got, want))


# ---------------------------------------------------------------- f-1: SpectralMLPBlock
@pytest.mark.parametrize("name,fused", [("block_mlp_t256_d64.npz", True), ("block_mlp_t1024_d48.npz", True),
                                        ("block_mlp_t100_d32.npz", False)])
def test_mlp_block_golden(pkg, dev, name, fused):
    from tensor_cuda_fft_b200 import spectral_layers as sl
    d = load(name)
    B, T, D = d["x"].shape
    blk = pkg.SpectralMLPBlock(D, mlp_ratio=2, dropout=0.0)
    blk.load_state_dict(state_dict_of(d), strict=True)
    blk = blk.to(dev).eval()
    x = d["x"].to(dev).requires_grad_(True)
    assert sl.fused_block_supported(x, blk.norm1, blk.spectral_mix) == fused
    calls = []
    orig = sl._LNSpectralResidualFn.apply
    sl._LNSpectralResidualFn.apply = staticmethod(lambda *a: (calls.append(1), orig(*a))[1])
    try:
        y = blk(x)
        y.backward(d["g"].to(dev))
    finally:
        sl._LNSpectralResidualFn.apply = orig
    torch.cuda.synchronize()
    assert (len(calls) == 1) == fused
    assert rel_l2(y.detach().cpu().numpy(), d["y"].numpy()) <= 1e-5
    assert rel_l2(x.grad.cpu().numpy(), d["gx"].numpy()) <= 1e-5
    check_grads(blk, d, 2e-5)
    # the spectral half alone
    with torch.no_grad():
        half = sl.ln_spectral_mix_residual(x.detach(), blk.norm1, blk.spectral_mix) if fused else x + blk.spectral_mix(blk.norm1(x))
    assert rel_l2(half.cpu().numpy(), d["half"].numpy()) <= 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 512, 40), (2, 2048, 96), (2, 192, 24)])
def test_ln_kernels_vs_torch(pkg, dev, shape, dtype):
    """sml_ln_stats / sml_ln_backward against torch.nn.functional.layer_norm and its autograd (no affine)."""
    from tensor_cuda_fft_b200 import _native
    from tensor_cuda_fft_b200.spectral_layers import _IO_DTYPES
    B, T, D = shape
    gen = torch.Generator().manual_seed(B * T + D)
    x = (torch.randn(B, T, D, generator=gen) * 2 + 0.5).to(dev, dtype)
    gh = torch.randn(B, T, D, generator=gen).to(dev, dtype)
    gres = torch.randn(B, T, D, generator=gen).to(dev, dtype)
    cadd = torch.randn(B, D, generator=gen).to(dev)
    lib = _native.lib()
    io = _IO_DTYPES[dtype]
    Tn = T + 64                                    # statistics laid out over a longer (zero-padded) transform, window at row 0
    stats = torch.full((B, Tn, 2), 7.0, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(lib.sml_ln_stats(x.data_ptr(), stats.data_ptr(), B, Tn, T, 0, D, 1e-5, io, st))
    xf = x.float().requires_grad_(True)
    mean = xf.mean(-1)
    rstd = torch.rsqrt(xf.var(-1, unbiased=False) + 1e-5)
    assert torch.all(stats[:, T:, :] == 0)
    assert rel_l2(stats[:, :T, 0].cpu().numpy(), mean.detach().cpu().numpy()) <= 1e-5
    assert rel_l2(stats[:, :T, 1].cpu().numpy(), rstd.detach().cpu().numpy()) <= 1e-5
    xhat = F.layer_norm(xf, (D,))
    (xhat * (gh.float() + cadd[:, None, :])).sum().backward()
    want = xf.grad + gres.float()
    gx = torch.empty_like(x)
    _native.check(lib.sml_ln_backward(gh.data_ptr(), x.data_ptr(), stats.data_ptr(), gres.data_ptr(), cadd.data_ptr(), gx.data_ptr(),
                                      B, Tn, T, 0, D, io, st))
    torch.cuda.synchronize()
    assert rel_l2(gx.float().cpu().numpy(), want.cpu().numpy()) <= (1e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype,B", [(torch.float32, 4), (torch.bfloat16, 4)])
def test_mlp_block_half_cfg2(pkg, dev, dtype, B):
    """BASELINE cfg-2 geometry (T = 8192, D = 768): the fused LayerNorm-on-load / residual-on-store kernel against the unfused
    composition around the same layer (torch LayerNorm + fused layer + add), forward and backward."""
    from tensor_cuda_fft_b200 import spectral_layers as sl
    T, D = 8192, 768
    gen = torch.Generator().manual_seed(5)
    norm = torch.nn.LayerNorm(D)
    layer = pkg.SpectralMixingLayer(D)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.3 * torch.randn(D, generator=gen))
        norm.bias.copy_(0.3 * torch.randn(D, generator=gen))
        layer.weight_real.copy_(torch.randn(layer.weight_real.shape, generator=gen))
        layer.weight_imag.copy_(torch.randn(layer.weight_imag.shape, generator=gen))
        layer.bias.copy_(torch.randn(D, generator=gen))
    norm, layer = norm.to(dev), layer.to(dev)
    x = (torch.randn(B, T, D, generator=gen) + 0.2).to(dev, dtype)
    g = torch.randn(B, T, D, generator=gen).to(dev, dtype)
    xa = x.clone().requires_grad_(True)
    assert sl.fused_block_supported(xa, norm, layer)
    ya = sl.ln_spectral_mix_residual(xa, norm, layer)
    ya.backward(g)
    ga = [xa.grad] + [p.grad.clone() for p in list(norm.parameters()) + list(layer.parameters())]
    for p in list(norm.parameters()) + list(layer.parameters()):
        p.grad = None
    xb = x.float().clone().requires_grad_(True)         # reference composition in fp32 on the same (rounded) inputs
    yb = xb + layer(norm(xb))
    yb.backward(g.float())
    gb = [xb.grad] + [p.grad.clone() for p in list(norm.parameters()) + list(layer.parameters())]
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_l2(ya.detach().float().cpu().numpy(), yb.detach().cpu().numpy()) <= tol
    for a, b in zip(ga, gb):
        assert rel_l2(a.float().cpu().numpy(), b.float().cpu().numpy()) <= (3e-5 if dtype == torch.float32 else 2e-2)


# ---------------------------------------------------------------- f-2: FixedSpectralBlock
@pytest.mark.parametrize("name", ["block_fixed_t64_k16_c32.npz", "block_fixed_t96_k24_c16.npz",
                                  "block_fixed_t512_k128_c32_cut.npz", "block_fixed_t1024_k128_c16.npz"])
def test_fixed_block_golden(pkg, dev, name):
    from tensor_cuda_fft_b200 import spectral_conv as sc
    d = load(name)
    B, T, C = d["x"].shape
    K = int(d["K"])
    blk = sc.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=int(d["trans"]), dropout=0.0)
    blk.load_state_dict(state_dict_of(d), strict=True)
    blk = blk.to(dev).eval()
    cutoff = None if int(d["cutoff"]) < 0 else int(d["cutoff"])
    x = d["x"].to(dev).requires_grad_(True)
    y = blk(x, cutoff=cutoff)
    y.backward(d["g"].to(dev))
    torch.cuda.synchronize()
    assert rel_l2(y.detach().cpu().numpy(), d["y"].numpy()) <= 1e-5
    assert rel_l2(x.grad.cpu().numpy(), d["gx"].numpy()) <= 2e-5
    check_grads(blk, d, 5e-5)


def test_fixed_block_dropout_path_and_bf16(pkg, dev):
    """dropout > 0 in training keeps the skip connection outside the kernel (same numbers with p -> identity in eval);
    bf16 I/O against the oracle restatement on the rounded input."""
    from tensor_cuda_fft_b200 import spectral_conv as sc
    torch.manual_seed(3)
    B, T, C, K = 2, 256, 64, 32
    blk = sc.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=4, dropout=0.0)
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.2 * torch.randn_like(p))
        blk.kernel.copy_(0.2 * torch.randn(K))
    sd = {k: v.clone() for k, v in blk.state_dict().items()}
    blk = blk.to(dev)
    x = torch.randn(B, T, C)
    want = bo.fixed_block_spectral_half(x, sd["ln.weight"], sd["ln.bias"], 1e-5, sd["kernel"], sd["gain"], sd["gate_freq_logits"],
                                        sd["gate_ctx.weight"], sd["gate_ctx.bias"])
    blk.train()                                        # p = 0: dropout is the identity but the unfused-residual branch runs
    blk.drop.p = 1e-12
    got_unfused = blk.spectral_half(x.to(dev))
    blk.eval()
    got_fused = blk.spectral_half(x.to(dev))
    assert rel_l2(got_fused.detach().cpu().numpy(), want.numpy()) <= 1e-5
    assert rel_l2(got_unfused.detach().cpu().numpy(), want.numpy()) <= 1e-5
    xb = x.to(torch.bfloat16)
    want_b = bo.fixed_block_spectral_half(xb.float(), sd["ln.weight"], sd["ln.bias"], 1e-5, sd["kernel"], sd["gain"],
                                          sd["gate_freq_logits"], sd["gate_ctx.weight"], sd["gate_ctx.bias"])
    got_b = blk.spectral_half(xb.to(dev))
    assert got_b.dtype == torch.bfloat16
    assert rel_l2(got_b.detach().float().cpu().numpy(), want_b.numpy()) <= 1e-2


# ---------------------------------------------------------------- f-4: inference path
def test_overlap_save_golden(pkg, dev):
    from tensor_cuda_fft_b200 import spectral_conv as sc
    d = load("block_overlap_save.npz")
    K, n_fft_full, chunk = int(d["K"]), int(d["n_fft_full"]), int(d["chunk"])
    T, C = d["h_hist"].shape[1:]
    blk = sc.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=4, dropout=0.0)
    blk.load_state_dict(state_dict_of(d), strict=True)
    blk = blk.to(dev).eval()
    with torch.no_grad():
        ln_in = blk.ln(d["h_hist"].to(dev))
    state = {"ctx_ln": ln_in.contiguous(), "ctx_sum": ln_in.sum(dim=1).contiguous()}
    cache = {}
    for step in range(2):
        h_out, state = sc.overlap_save_block_update(blk, state, d[f"h_chunk{step}"].to(dev), n_fft_full=n_fft_full, kernel_len=K,
                                                    cache=cache)
        torch.cuda.synchronize()
        assert rel_l2(h_out.cpu().numpy(), d[f"h_out{step}"].numpy()) <= 1e-5, step
        assert rel_l2(state["ctx_sum"].cpu().numpy(), d[f"ctx_sum{step}"].numpy()) <= 1e-5


def test_spectral_ema_golden(pkg, dev):
    from tensor_cuda_fft_b200 import spectral_conv as sc
    d = load("block_spectral_ema.npz")
    chunks, init = d["chunks"].to(dev), d["init"].to(dev)
    Fq = chunks.shape[2]
    for mode in ("aligned", "polar"):
        ema = sc.SpectralEMA(sc.EMAConfig(n_freqs=Fq, mode=mode)).to(dev)
        with torch.no_grad():
            ema.rho_logit.copy_(d[f"{mode}.rho_logit"])
            ema.theta_raw.copy_(d[f"{mode}.theta_raw"])
            got = [ema.scan(chunks), ema.scan(chunks, init=init), ema.update(init, chunks[:, 5, :])]
        torch.cuda.synchronize()
        for a, key in zip(got, ("scan", "scan_init", "update")):
            assert rel_l2(a.cpu().numpy(), d[f"{mode}.{key}"].numpy()) <= 1e-5, (mode, key)
    # a long scan at a realistic size against the oracle loop (S = 64 chunks of a 1024-token window, spectral_ssm.py:121)
    gen = torch.Generator().manual_seed(9)
    B, S, Fq = 8, 64, 513
    chunks = torch.complex(torch.randn(B, S, Fq, generator=gen), torch.randn(B, S, Fq, generator=gen))
    ema = sc.SpectralEMA(sc.EMAConfig(n_freqs=Fq)).to(dev)
    rho = torch.sigmoid(ema.rho_logit.detach().cpu())
    theta = math.pi * torch.tanh(ema.theta_raw.detach().cpu())
    with torch.no_grad():
        got = ema.scan(chunks.to(dev))
    assert rel_l2(got.cpu().numpy(), bo.ema_scan(chunks, rho, theta).numpy()) <= 1e-5


def test_ext_row_windows(pkg, dev):
    """sml_forward_ext row windows against the documented model: input window at an odd row offset, a shorter output,
    residual on the output rows, channel scale, the bin T/2; R = 2 passes so both parities of the shift are exercised.  An
    output window that does not start at row 0 is refused (TMA stores fault on negative coordinates)."""
    from tensor_cuda_fft_b200 import _native
    lib = _native.lib()
    B, T, D = 2, 512, 32                     # T = 2k -> M = 256, R = 2
    Fn = T // 2
    gen = torch.Generator().manual_seed(17)
    bad = _native.make_ext(T_in=64, T_out=64, out_row0=8)
    assert lib.sml_ext_supported(B, T, D, Fn, 0, ctypes.byref(bad)) != 0 and b"out_row0" in lib.sml_last_error()
    for (T_in, in0, T_out, out0) in [(128, 0, 64, 0), (64, 33, 128, 0), (512, 0, 512, 0), (2, 101, 2, 0), (48, 6, 16, 0)]:
        x = torch.randn(B, T_in, D, generator=gen)
        res = torch.randn(B, T_out, D, generator=gen)
        w_re, w_im = torch.randn(D, Fn, generator=gen), torch.randn(D, Fn, generator=gen)
        w_nyq = torch.randn(D, generator=gen)
        scale = torch.rand(B, D, generator=gen) + 0.5
        xp = torch.zeros(B, T, D)
        xp[:, in0:in0 + T_in] = x
        X = torch.fft.fft(xp, dim=1)
        A = X[:, :Fn, :] * torch.complex(w_re, w_im).t().unsqueeze(0) * scale.unsqueeze(1)
        Z = torch.cat([A, torch.zeros(B, T - Fn, D, dtype=A.dtype)], dim=1)
        y = torch.fft.ifft(Z, dim=1).real
        y = y + (X[:, Fn, :].real * w_nyq * scale / T).unsqueeze(1) * torch.cos(torch.pi * torch.arange(T)).view(1, T, 1)
        want = res + y[:, out0:out0 + T_out]
        xd, rd = x.to(dev), res.to(dev)
        wr, wi, wn, sc_ = w_re.to(dev), w_im.to(dev), w_nyq.to(dev), scale.to(dev).contiguous()
        xnyq = torch.empty(B, D, device=dev)
        out = torch.full((B, T_out, D), float("nan"), device=dev)
        ext = _native.make_ext(residual=rd, chan_scale=sc_, w_nyq=wn, x_nyq=xnyq, T_in=T_in, in_row0=in0, T_out=T_out, out_row0=out0)
        _native.check(lib.sml_forward_ext(xd.data_ptr(), wr.data_ptr(), wi.data_ptr(), None, out.data_ptr(), None, B, T, D, Fn, 0,
                                          ctypes.byref(ext), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert rel_l2(out.cpu().numpy(), want.numpy()) <= 1e-5, (T_in, in0, T_out, out0)
        assert rel_l2(xnyq.cpu().numpy(), X[:, Fn, :].real.numpy()) <= 1e-5


def test_blocks_random_shapes(pkg, dev):
    """Randomised shapes over every plan family the extended kernels take (sub-transform 64 / 256 / 1024, R = 1 .. 16, wide-band
    forms, fp32 and bf16): the fused SpectralMLPBlock half against the unfused composition, FixedSpectralBlock's half against
    the oracle restatement of the reference lines."""
    from tensor_cuda_fft_b200 import spectral_conv as sc
    from tensor_cuda_fft_b200 import spectral_layers as sl
    rng = np.random.default_rng(2024)
    n_fused = 0
    for case in range(40):
        T = int(rng.choice([64, 128, 192, 256, 512, 768, 1024, 2048, 3072, 4096]))
        D = int(rng.choice([8, 16, 24, 40, 64, 96, 136, 256]))
        B = int(rng.integers(1, 4))
        nf = int(rng.choice([max(1, D // 2), 5, 33, 100, 300]))
        dtype = torch.float32 if rng.random() < 0.7 else torch.bfloat16
        torch.manual_seed(case)
        norm = torch.nn.LayerNorm(D).to(dev)
        layer = pkg.SpectralMixingLayer(D, num_filters=nf).to(dev)
        with torch.no_grad():
            for p in list(norm.parameters()) + list(layer.parameters()):
                p.add_(0.5 * torch.randn_like(p))
        x = (torch.randn(B, T, D) * 1.5 + 0.3).to(dev, dtype)
        g = torch.randn(B, T, D).to(dev, dtype)
        xa = x.clone().requires_grad_(True)
        if not sl.fused_block_supported(xa, norm, layer):
            continue
        n_fused += 1
        ya = sl.ln_spectral_mix_residual(xa, norm, layer)
        ya.backward(g)
        ga = [xa.grad] + [p.grad.clone() for p in list(norm.parameters()) + list(layer.parameters())]
        for p in list(norm.parameters()) + list(layer.parameters()):
            p.grad = None
        xb = x.float().clone().requires_grad_(True)
        yb = xb + layer(norm(xb))
        yb.backward(g.float())
        gb = [xb.grad] + [p.grad.clone() for p in list(norm.parameters()) + list(layer.parameters())]
        tol = 2e-5 if dtype == torch.float32 else 2e-2
        tag = (case, B, T, D, nf, str(dtype))
        assert rel_l2(ya.detach().float().cpu().numpy(), yb.detach().cpu().numpy()) <= tol, tag
        for i, (a, b) in enumerate(zip(ga, gb)):
            if float(b.float().abs().max()) > 0:
                assert rel_l2(a.float().cpu().numpy(), b.float().cpu().numpy()) <= 5 * tol, tag + (i,)
    assert n_fused >= 20
    for case in range(12):
        T = int(rng.choice([32, 48, 64, 100, 128, 200, 256, 500, 1024]))
        K = int(rng.choice([4, 8, 16, 32, 64]))
        C = int(rng.choice([8, 16, 32, 48]))
        B = int(rng.integers(1, 4))
        x = torch.randn(B, T, C)
        if not sc.causal_spectral_conv_supported(x.to(dev), K):
            continue
        torch.manual_seed(100 + case)
        blk = sc.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=4, dropout=0.0)
        with torch.no_grad():
            for p in blk.parameters():
                p.add_(0.2 * torch.randn_like(p))
            blk.kernel.copy_(0.2 * torch.randn(K))
        sd = {k: v.clone() for k, v in blk.state_dict().items()}
        cutoff = None if case % 3 else int(rng.integers(2, sc.conv_fft_len(T, K) // 2))
        want = bo.fixed_block_spectral_half(x, sd["ln.weight"], sd["ln.bias"], 1e-5, sd["kernel"], sd["gain"], sd["gate_freq_logits"],
                                            sd["gate_ctx.weight"], sd["gate_ctx.bias"], cutoff=cutoff, transition_bins=4)
        got = blk.to(dev).eval().spectral_half(x.to(dev), cutoff)
        assert rel_l2(got.detach().cpu().numpy(), want.numpy()) <= 2e-5, (case, B, T, K, C, cutoff)
