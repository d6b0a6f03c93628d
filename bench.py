#!/usr/bin/env python
"""bench.py -- SpectralMixingLayer fwd+bwd tokens/s on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f32|bf16] [--impl ours|reference] [--config cfg1|cfg2|cfg3|cfg5]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              bench.py --gpus N --steps K --warmup W

A "step" is one forward + one backward of the layer over one batch of synthetic input.  Workload at N=1 is
BASELINE.json configs[1]: SpectralMixingLayer(embed_dim=768), x = (16, 8192, 768); weak scaling (each rank holds its
own batch of 16, i.e. the batch axis is sharded) with one NCCL all-reduce(sum) of the filter/bias gradients per step.

--config selects another BASELINE.json configuration (default cfg2 = configs[1], the one the metric is quoted on):
  cfg1  configs[0]  (8, 512, 256) fp32                       correctness-sized, launch bound
  cfg3  configs[2]  (64, 4096, 1024) fp32, STRONG scaling: the batch of 64 is split over the N ranks, filter-gradient all-reduce
  cfg5  configs[4]  long-context sweep T = 1K .. 128K at embed 1024, 2^20 tokens per step split over the N ranks by batch; one line
        with a `sweep` list (per T: tokens/s, HBM-roofline fraction, and at N=1 the CPU arm on a batch-1 sample)

One JSON line on stdout (rank 0):
  value        whole-job tokens/s (tokens = B*T summed over ranks), inputs resident in HBM, CUDA-event timed
  e2e          same metric through the C-ABI host-buffer call sml_fwd_bwd_host: HOST (pinned) x and g copied in and y, gx
               and the filter gradients copied back inside the timed region (chunked copy/compute pipeline)
  roofline     dominant kernel (fused backward): algorithmic bytes / measured launch time vs MEASURED_PEAKS.json
  cpu_baseline the unmodified reference module (baseline/_ref; oracle port if absent) timed on this box's host cores, bounded sample
  sub-records of the default line (N = 1): bf16 (bf16 I/O of the same config), blocks (SpectralMLPBlock's
               spectral half fused vs unfused, FixedSpectralBlock's half eager and graphed), graphed (launch-bound shapes:
               fwd+bwd replayed from CUDA graphs, and the whole step as ONE graph), reference_algorithm_on_gpu (context)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

print_line = print       # replaced in main(): the JSON line is held back until fd 1 is restored
METRIC = "spectral_mixing_fwd_bwd_tokens_per_sec"
UNIT = "tokens/s"
CFG = {"B": 16, "T": 8192, "D": 768}      # BASELINE.json configs[1]


def _fft_backend():
    return "MKL DFTI" if torch.backends.mkl.is_available() else "pocketfft"


def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown CPU"


def gpu_reference_algorithm(x, g, w_re, w_im, bias, steps=5):
    """Context only: the reference ALGORITHM (fft_tensor/spectral_layers.py:88-116: torch.fft.fft, complex filter into a zero
    tensor, torch.fft.ifft(...).real, + bias; autograd backward) executed by PyTorch/cuFFT on the same GPU and inputs."""
    params = [p.detach().clone().requires_grad_(True) for p in (w_re, w_im, bias)]

    def step():
        for p in params:
            p.grad = None
        xr = x.detach().float().requires_grad_(True)
        spec = torch.fft.fft(xr, dim=1)
        k = min(params[0].shape[1], xr.shape[1] // 2)
        kept = torch.zeros_like(spec)
        kept[:, :k, :] = spec[:, :k, :] * torch.complex(params[0], params[1])[:, :k].T.unsqueeze(0)
        (torch.fft.ifft(kept, dim=1).real + params[2]).backward(g.float())

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--dtype", choices=["f32", "bf16"], default="f32")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=CFG["B"], help="per-GPU batch")
    ap.add_argument("--seq", type=int, default=CFG["T"])
    ap.add_argument("--embed", type=int, default=CFG["D"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="CPU legs: batch of the timed sample (0 = reference arm: the full batch; cpu_baseline leg: 2)")
    ap.add_argument("--config", choices=["cfg1", "cfg2", "cfg3", "cfg5"], default="cfg2")
    ap.add_argument("--no-bf16", action="store_true", help="skip the bf16 I/O sub-record of the default line")
    ap.add_argument("--no-blocks", action="store_true", help="skip the block-level (f-1 / f-2) sub-record of the default line")
    args = ap.parse_args()
    args.scaling = "weak"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config == "cfg1":
        args.batch, args.seq, args.embed = 8, 512, 256
    elif args.config == "cfg3":      # strong scaling: global batch 64 split over the ranks
        args.batch, args.seq, args.embed, args.scaling = max(1, 64 // world), 4096, 1024, "strong"
    return args


def workload_config(args, world):
    """The `config` object, identical for both arms (the driver compares them)."""
    B, T, D = args.batch, args.seq, args.embed
    k = min(D // 2, T // 2)
    return {"workload": f"SpectralMixingLayer(embed_dim={D}) fwd+bwd, x=({B},{T},{D}) per GPU, {args.dtype} I/O, fp32 math, "
                        f"k={k} live bins, randn x/g/filter, dropout 0 (BASELINE.json {args.config})",
            "io_dtype": args.dtype, "global_batch": world * B, "seq_len": T, "embed_dim": D,
            "parallelism": f"batch-sharded x{world}"}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's torch port (== the reference's torch.fft algorithm) on the host cores
# ------------------------------------------------------------------------------------------------------------
def load_reference_module():
    """The UNMODIFIED fft_tensor/spectral_layers.py from baseline/_ref (placed there by __graft_entry__.build() where the
    reference is mounted; git-ignored, travels with gpurun), imported by file path (the package __init__ prints a banner
    and sets a global memory limit).  None if it is not there."""
    import importlib.util
    path = os.path.join(ROOT, "baseline", "_ref", "fft_tensor", "spectral_layers.py")
    if not os.path.exists(path):
        return None
    try:
        spec = importlib.util.spec_from_file_location("ref_spectral_layers", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


def cpu_reference_step_time(B, T, D, steps, warmup, budget_s=None):
    """Seconds per fwd+bwd step of the reference on the host cores.  Returns (dt, steps timed, threads, kind, what)."""
    torch.set_num_threads(os.cpu_count() or 1)
    gen = torch.Generator().manual_seed(0)
    Fn = D // 2
    w_re, w_im, bias = torch.randn(D, Fn, generator=gen), torch.randn(D, Fn, generator=gen), torch.randn(D, generator=gen)
    x, g = torch.randn(B, T, D, generator=gen), torch.randn(B, T, D, generator=gen)
    ref = load_reference_module()
    if ref is not None:
        layer = ref.SpectralMixingLayer(D)
        with torch.no_grad():
            layer.weight_real.copy_(w_re); layer.weight_imag.copy_(w_im); layer.bias.copy_(bias)
        kind, what = "reference", "unmodified fft_tensor/spectral_layers.py (baseline/_ref)"

        def step():
            layer.zero_grad(set_to_none=True)
            xr = x.detach().requires_grad_(True)
            layer(xr).backward(g)
    else:
        from oracle import spectral_mixing_oracle as orc   # allowed here: cpu_baseline / --impl reference legs only
        kind, what = "port", "oracle torch port of spectral_layers.py:83-118"

        def step():
            orc.torch_port_fwd_bwd(x, w_re, w_im, bias, g)
    for _ in range(warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    return sum(times) / len(times), len(times), torch.get_num_threads(), kind, what


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # the SAME configuration as the GPU arm: the full per-GPU batch, --steps and --warmup as given (a step of cfg-2 costs about
    # 0.8 s on a 16-core host).  --cpu-sample-batch < batch keeps a smaller sample for smoke tests; a 240 s budget bounds a
    # run that was started with a huge --steps (the line reports the steps actually timed).
    B = args.batch if args.cpu_sample_batch <= 0 else min(args.cpu_sample_batch, args.batch)
    T, D = args.seq, args.embed
    dt, n, threads, kind, what = cpu_reference_step_time(B, T, D, args.steps, args.warmup, budget_s=240.0)
    val = B * T / dt
    cfg = workload_config(args, world)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{what} (torch.fft, {_fft_backend()}, {_cpu_model()}), x=({B},{T},{D}) fp32 fwd+bwd"
                                   + ("" if B == args.batch else f" (batch {B} of {args.batch}; columns are independent)")
                                   + f", mean of {n} steps"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print_line(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields through NVML)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  A separate `nvidia-smi -lms` process does the sampling (a
    Python thread in this process is starved of the GIL by the launch loop); samples are kept if their timestamp falls
    inside [mark_begin, mark_end]."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index, interval_ms=10):
        import subprocess
        import tempfile
        self.proc, self.t0, self.t1 = None, None, None
        self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", str(interval_ms)], stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception as e:   # pragma: no cover
            self.err = str(e)

    def start(self):      # kept for symmetry: the process is already running (its start-up takes ~0.1 s)
        pass

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi_unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.out.flush()
        rows = []
        for ln in open(self.out.name):
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
            except Exception:
                continue
        try:
            os.unlink(self.out.name)
        except OSError:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.005 <= r[0] <= self.t1 + 0.005] or rows[-3:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        reasons = sorted({n for r in inside for n, v in zip(self.NAMES, r[4]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": inside[0][2], "reasons": reasons,
                "samples": len(inside), "power_w_max": max(r[3] for r in inside),
                "sampler": "nvidia-smi -lms 10, samples inside the timed region" + ("; " + self.note if getattr(self, "note", None) else "")}


def _collective_path():
    from tensor_cuda_fft_b200 import distributed as d
    return d.LAST_ALLREDUCE_PATH


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "of measured: MEASURED_PEAKS.json hbm_gbs (STREAM-style copy)"
        except Exception:
            pass
    return 6650.0, "of fallback: 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class Problem:
    """One (B, T, D, dtype) problem resident on the GPU: the module, its inputs, and timers for the step and the two C-ABI calls."""

    def __init__(self, B, T, D, dtype_name, dev, rank, world):
        from tensor_cuda_fft_b200 import SpectralMixingLayer, _native
        self.B, self.T, self.D, self.dtype_name, self.dev, self.world = B, T, D, dtype_name, dev, world
        self.Fn = D // 2
        self.k = min(self.Fn, T // 2)
        self.dtype = torch.float32 if dtype_name == "f32" else torch.bfloat16
        self.esz = 4 if dtype_name == "f32" else 2
        self.io = _native.DTYPE_F32 if dtype_name == "f32" else _native.DTYPE_BF16
        self.native = _native
        self.lib = _native.lib()
        self.plan = _native.plan(B, T, D, self.Fn, self.io)
        torch.manual_seed(0 + rank)
        self.layer = SpectralMixingLayer(D).to(dev)
        with torch.no_grad():
            self.layer.weight_real.normal_()
            self.layer.weight_imag.normal_()
            self.layer.bias.normal_()
        self.x = torch.randn(B, T, D, device=dev).to(self.dtype)
        self.g = torch.randn(B, T, D, device=dev).to(self.dtype)
        mode = os.environ.get("SML_ALLREDUCE", "fused")
        if world > 1 and mode != "nccl":
            # filter/bias gradients live in an NVLink symmetric-memory bucket: "fused" (default) = the batch-reduction kernel pushes
            # its sums with multimem.red, the all-reduce is that store plus one barrier; "symm" = local sums + one in-place multimem
            # all-reduce; any failure to set it up leaves the NCCL all-reduce in place
            try:
                from tensor_cuda_fft_b200 import attach_symmetric_grad_buffers
                attach_symmetric_grad_buffers([self.layer], fused=None if mode == "fused" else False)
            except Exception as e:
                sys.stderr.write(f"symmetric-memory gradient bucket unavailable ({str(e).splitlines()[0][:160]}): using NCCL\n")

    def step(self):
        """forward + backward through the module API (autograd included); at N > 1 one flat all-reduce of the filter gradients."""
        from tensor_cuda_fft_b200 import allreduce_filter_grads
        self.layer.zero_grad(set_to_none=True)
        xr = self.x.detach().requires_grad_(True)
        y = self.layer(xr)
        y.backward(self.g)
        if self.world > 1:
            allreduce_filter_grads([self.layer])
        gx = xr.grad
        xr.grad = None
        return y, gx

    def abi_calls(self):
        """(fwd, bwd) closures over the raw C ABI on the current stream (per-kernel timing for the roofline)."""
        lib, native = self.lib, self.native
        B, T, D, Fn, io = self.B, self.T, self.D, self.Fn, self.io
        stream = torch.cuda.current_stream().cuda_stream
        wr, wi, bs = self.layer.weight_real.detach(), self.layer.weight_imag.detach(), self.layer.bias.detach()
        y, gx = torch.empty_like(self.x), torch.empty_like(self.x)
        xlow = torch.empty(max(lib.sml_xlow_bytes(B, T, D, Fn), 8), dtype=torch.uint8, device=self.dev)
        gflat = torch.empty(2 * D * Fn + D, device=self.dev)     # [gw_re | gw_im | gb], the host module's layout
        gwr, gwi, gb = gflat[:D * Fn], gflat[D * Fn:2 * D * Fn], gflat[2 * D * Fn:]
        ws_bytes = lib.sml_workspace_bytes(B, T, D, Fn, io)
        ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=self.dev)
        x, g = self.x, self.g
        self._keep = (y, gx, xlow, gflat, ws)

        def fwd():
            native.check(lib.sml_forward(x.data_ptr(), wr.data_ptr(), wi.data_ptr(), bs.data_ptr(), y.data_ptr(),
                                         xlow.data_ptr(), B, T, D, Fn, io, stream))

        def bwd():
            native.check(lib.sml_backward(g.data_ptr(), xlow.data_ptr(), wr.data_ptr(), wi.data_ptr(), gx.data_ptr(),
                                          gwr.data_ptr(), gwi.data_ptr(), gb.data_ptr(), ws.data_ptr(), ws_bytes,
                                          B, T, D, Fn, io, stream))
        return fwd, bwd

    # byte counts: the ALGORITHMIC floor does not count what the implementation adds (X_low, partial sums)
    def bytes_step(self):
        return 4 * self.B * self.T * self.D * self.esz

    def bytes_fwd(self):
        return 2 * self.B * self.T * self.D * self.esz + (2 * self.D * self.Fn + self.D) * 4

    def bytes_bwd(self):
        return 2 * self.B * self.T * self.D * self.esz + 2 * (2 * self.D * self.Fn + self.D) * 4

    def bytes_impl_extra(self):
        xl = self.B * self.D * self.k * 8
        return {"fwd_xlow_write": xl, "bwd_xlow_read": xl, "bwd_partial_terms_write_read": 2 * xl}


def time_steps(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def time_kernel(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return sum(ts) / len(ts), min(ts)


def run_ours(args):
    import torch.distributed as dist
    from tensor_cuda_fft_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a GPU (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    peak, peak_src = load_peaks()

    def roof(nbytes, ms):
        ach = nbytes / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None}

    if args.config == "cfg5":
        run_sweep(args, dev, rank, world, barrier, max_over_ranks, roof)
        if world > 1:
            dist.destroy_process_group()
        return

    B, T, D = args.batch, args.seq, args.embed
    prob = Problem(B, T, D, args.dtype, dev, rank, world)
    sampler = ClockSampler(local)      # separate process; started before the warm-up so it is sampling by the time we measure
    for _ in range(max(args.warmup, 3)):
        prob.step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events, max over ranks ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _native.launch_count()
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        prob.step()
    e1.record()
    barrier()
    sampler.mark_end()
    launches = _native.launch_count() - n0
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    if ms_step * args.steps < 100.0:
        # the timed region is shorter than a few sampling periods: keep the same load running (untimed, the same number of
        # extra steps on every rank) until the sampler has seen ~0.2 s of it, and say so in the clocks record
        for _ in range(min(5000, int(200.0 / ms_step) + 1)):
            prob.step()
        barrier()
        sampler.t1 = time.time()
        sampler.note = "timed region < 0.1 s: window extended over an untimed continuation of the same steps"
    clocks = sampler.stop()
    value = world * B * T / (ms_step * 1e-3)

    # ---- per-kernel timing through the raw C ABI (events on the launching stream), for the roofline ----
    fwd, bwd = prob.abi_calls()
    nk = min(max(args.steps, 10), 100)
    fwd_ms, fwd_min = time_kernel(fwd, nk)
    bwd_ms, bwd_min = time_kernel(bwd, nk)
    traffic = {}
    try:    # DRAM bytes per launch QUOTED from the committed ncu capture of exactly this configuration (not measured in this run)
        if (B, T, D, args.dtype) == (16, 8192, 768, "f32"):
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["cfg2_f32"]
    except Exception:
        traffic = {}
    fast = prob.plan["path"] == "fast"
    roofline = roof(prob.bytes_bwd(), bwd_ms)
    roofline.update({"kernel": "sml_backward = fused analysis + Wirtinger filter-gradient terms + synthesis kernel, + filtergrad_reduce_kernel" if fast else "generic kernels",
                     "launch_ms": bwd_ms, "launch_ms_min": bwd_min, "algorithmic_bytes": prob.bytes_bwd(), "peak_source": peak_src,
                     "algorithmic_bytes_note": "SURVEY.md 8(d): read g + write gx + filter read + gradients written; X_low and the per-batch partial terms are implementation traffic (implementation_extra_bytes), not counted",
                     "implementation_extra_bytes": prob.bytes_impl_extra(),
                     "traffic": traffic.get("sml_backward"),
                     "traffic_source": "QUOTED from profiles/traffic.json (ncu --set full capture of this configuration, round 2), not measured in this run" if traffic else None})
    roofline_fwd = roof(prob.bytes_fwd(), fwd_ms)
    roofline_fwd.update({"kernel": "sml_forward (one fused kernel)", "launch_ms": fwd_ms, "launch_ms_min": fwd_min,
                         "algorithmic_bytes": prob.bytes_fwd(), "traffic": traffic.get("sml_forward")})
    roofline_step = roof(prob.bytes_step(), ms_step)
    roofline_step.update({"note": "4-pass floor bytes / whole fwd+bwd step time (includes launch gaps and, at N>1, the all-reduce)"})

    # ---- bf16 I/O of the same configuration (BASELINE.json configs[1] names both): a sub-record of the default line ----
    bf16 = None
    if args.dtype == "f32" and args.config == "cfg2" and not args.no_bf16:
        try:
            pb = Problem(B, T, D, "bf16", dev, rank, world)
            ms_b = max_over_ranks(time_steps(pb.step, min(args.steps, 100), 5, barrier))
            fb, bb = pb.abi_calls()
            fb_ms, _ = time_kernel(fb, 30)
            bb_ms, _ = time_kernel(bb, 30)
            bf16 = {"value": world * B * T / (ms_b * 1e-3), "unit": UNIT, "ms_per_step": ms_b, "io_dtype": "bf16",
                    "kernels": "tensor-core (tcgen05) kernels, csrc/sml_tc.cuh" if os.environ.get("SML_TC", "") == "1" or _tc_default() else "CUDA-core butterfly kernels, csrc/sml_fast.cuh",
                    "roofline_step": roof(pb.bytes_step(), ms_b), "roofline_fwd": dict(roof(pb.bytes_fwd(), fb_ms), launch_ms=fb_ms),
                    "roofline_bwd": dict(roof(pb.bytes_bwd(), bb_ms), launch_ms=bb_ms)}
            del pb, fb, bb
            torch.cuda.empty_cache()
        except Exception as e:   # the sub-record must never cost the headline line
            bf16 = {"unavailable": str(e).splitlines()[0][:200]}

    # ---- launch-bound shapes (configs[0]): the same step replayed from CUDA graphs (SpectralMixingLayer.graphed) ----
    graphed = None
    if world == 1 and 2 * B * T * D * prob.esz < 64e6:
        try:
            glayer = prob.layer.graphed(prob.x)

            def gstep():
                prob.layer.zero_grad(set_to_none=True)
                xr = prob.x.detach().requires_grad_(True)
                glayer(xr).backward(prob.g)

            ms_g = time_steps(gstep, max(args.steps, 50), 5, barrier)
            graphed = {"ms_per_step": ms_g, "value": B * T / (ms_g * 1e-3), "unit": UNIT,
                       "what": "the same fwd+bwd step with forward and backward replayed from CUDA graphs (SpectralMixingLayer.graphed)"}
            # the whole step (forward + backward, all gradients) as ONE graph on static buffers: one launch per step
            replay, bufs = prob.layer.graphed_step(prob.x, prob.g)
            ms_w = time_steps(replay, max(args.steps, 50), 5, barrier)
            graphed["whole_step_graph"] = {"ms_per_step": ms_w, "value": B * T / (ms_w * 1e-3), "unit": UNIT,
                                           "what": "SpectralMixingLayer.graphed_step: forward + backward captured as one CUDA graph on static buffers"}
            del replay, bufs
        except Exception as e:
            graphed = {"unavailable": str(e).splitlines()[0][:200]}

    # ---- block rows: the layer inside its callers, fused prologue / epilogue vs the unfused composition ----
    blocks = None
    if world == 1 and args.config == "cfg2" and args.dtype == "f32" and not args.no_blocks:
        try:
            blocks = run_blocks(prob, B, T, D, roof, barrier)
        except Exception as e:
            blocks = {"unavailable": str(e).splitlines()[0][:200]}

    # ---- end to end: the reference-facing C-ABI call with HOST buffers (sml_fwd_bwd_host): pinned host x, g in;
    #      y, gx and the filter/bias gradients back in host memory; all copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(prob, args, world, dev, barrier, max_over_ranks)

    # ---- CPU baseline (rank 0, N=1 only): a bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bs = min(args.cpu_sample_batch if args.cpu_sample_batch > 0 else 2, B)
        dt, n, threads, kind, what = cpu_reference_step_time(Bs, T, D, steps=40, warmup=1, budget_s=12.0)
        cpu = {"value": Bs * T / dt, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{what} (torch.fft on host, {_fft_backend()}, {_cpu_model()}), "
                         f"x=({Bs},{T},{D}) fp32 fwd+bwd (batch {Bs} of {B}; columns independent), mean of {n} steps",
               "ms_per_step": dt * 1e3}

    # ---- context: the reference algorithm through PyTorch/cuFFT on this same GPU (rank 0, N=1 only) ----
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            wr, wi, bs = prob.layer.weight_real.detach(), prob.layer.weight_imag.detach(), prob.layer.bias.detach()
            ms_ref = gpu_reference_algorithm(prob.x, prob.g, wr, wi, bs)
            gpu_ref = {"value": B * T / (ms_ref * 1e-3), "unit": UNIT, "ms_per_step": ms_ref,
                       "what": "reference algorithm (torch.fft/cuFFT + ATen elementwise + autograd, fp32) on the same B200 and shape; context, not an arm"}
        except Exception as e:   # e.g. out of memory at very large shapes: context only
            gpu_ref = {"unavailable": str(e).splitlines()[0][:200]}

    if rank == 0:
        esz = prob.esz
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",      # arithmetic type of the path (bf16 is an I/O type only)
            "config": workload_config(args, world),
            "impl_detail": {"plan": prob.plan, "l2": f"inputs larger than L2 ({2 * B * T * D * esz / 1e6:.0f} MB read per step vs 126 MB), no flush"
                                                    if 2 * B * T * D * esz > 126e6 else "inputs smaller than L2, no flush: launch-bound configuration",
                            "collective": "none" if world == 1 else f"1 all-reduce(sum) of [gw_re|gw_im|gb] per step ({_collective_path()})"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline, "roofline_fwd": roofline_fwd, "roofline_step": roofline_step,
            "bf16": bf16, "graphed": graphed, "blocks": blocks,
            "e2e": e2e, "cpu_baseline": cpu, "reference_algorithm_on_gpu": gpu_ref,
        }
        print_line(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_blocks(prob, B, T, D, roof, barrier):
    """Block-level rows of the default line: SpectralMLPBlock's spectral half `x + spectral_mix(norm1(x))` (spectral_layers.py:185)
    fused (LayerNorm on load, residual on store: sml_ln_stats + sml_forward_ext, sml_backward + sml_ln_backward) against the unfused
    composition around the same fused layer; FixedSpectralBlock's spectral half (train_fixed_full.py:498-555) at the reference's
    default sizes.  Floor = 5 passes over the activation (x, y | g, x, gx)."""
    import tensor_cuda_fft_b200 as pkg
    from tensor_cuda_fft_b200 import spectral_conv as sc
    from tensor_cuda_fft_b200 import spectral_layers as sl
    dev = prob.x.device
    norm = torch.nn.LayerNorm(D).to(dev)
    layer = prob.layer
    params = list(norm.parameters()) + list(layer.parameters())

    def step(fused):
        def run():
            for p in params:
                p.grad = None
            xr = prob.x.detach().requires_grad_(True)
            y = sl.ln_spectral_mix_residual(xr, norm, layer) if fused else xr + layer(norm(xr))
            y.backward(prob.g)
        return run

    out = {}
    act = B * T * D * prob.esz
    for name, fused in (("fused", True), ("unfused", False)):
        ms = time_steps(step(fused), 30, 5, barrier)
        out["mlp_block_half_" + name] = {"ms_per_step": ms, "value": B * T / (ms * 1e-3), "unit": UNIT, "roofline_5pass": roof(5 * act, ms)}
    for p in params:
        p.grad = None
    Bc, Tc, C, K = 64, 1024, 512, 128
    blk = sc.FixedSpectralBlock(C, seq_len=Tc, kernel_len=K, transition_bins=16, dropout=0.0).to(dev).eval()
    with torch.no_grad():
        blk.kernel.normal_(std=0.1)
    x2 = torch.randn(Bc, Tc, C, device=dev)
    g2 = torch.randn(Bc, Tc, C, device=dev)

    def conv_step():
        for p in blk.parameters():
            p.grad = None
        xr = x2.detach().requires_grad_(True)
        blk.spectral_half(xr).backward(g2)

    ms = time_steps(conv_step, 30, 5, barrier)
    out["fixed_block_half"] = {"shape": [Bc, Tc, C], "kernel_len": K, "n_fft": sc.conv_fft_len(Tc, K), "ms_per_step": ms,
                               "value": Bc * Tc / (ms * 1e-3), "unit": UNIT, "roofline_5pass": roof(5 * Bc * Tc * C * 4, ms)}
    try:      # the same step replayed from CUDA graphs: the half is two fused kernels plus a few dozen small parameter-side ops
        gh = blk.graphed(x2, half_only=True)

        def conv_step_graphed():
            for p in blk.parameters():
                p.grad = None
            xr = x2.detach().requires_grad_(True)
            gh(xr).backward(g2)

        ms_g = time_steps(conv_step_graphed, 30, 5, barrier)
        out["fixed_block_half"]["graphed"] = {"ms_per_step": ms_g, "value": Bc * Tc / (ms_g * 1e-3), "roofline_5pass": roof(5 * Bc * Tc * C * 4, ms_g)}
    except Exception as e:
        out["fixed_block_half"]["graphed"] = {"unavailable": str(e).splitlines()[0][:200]}
    out["what"] = ("mlp_block_half: x + spectral_mix(LayerNorm(x)) fwd+bwd at the headline shape, fused prologue/epilogue vs torch LayerNorm + fused "
                   "layer + add; fixed_block_half: pre-LN causal FFT convolution + gates + residual of fft_lm's FixedSpectralBlock, fwd+bwd")
    return out


def _tc_default():
    try:
        from tensor_cuda_fft_b200 import _native
        return bool(getattr(_native, "TC_DEFAULT", False))
    except Exception:
        return False


def run_e2e(prob, args, world, dev, barrier, max_over_ranks):
    import torch.distributed as dist
    from tensor_cuda_fft_b200 import spectral_mix_fwd_bwd_host
    B, T, D, dtype, esz = prob.B, prob.T, prob.D, prob.dtype, prob.esz
    param_bytes = (2 * D * prob.Fn + D) * 4
    # the host side of this leg is PCIe/host-memory bound: sit on the GPU's NUMA node BEFORE the pinned buffers are touched
    from tensor_cuda_fft_b200.distributed import bind_to_gpu_numa_node
    old_affinity = os.sched_getaffinity(0)      # restored at the end of the leg: the CPU baseline wants every core again
    numa = bind_to_gpu_numa_node(dev.index) if os.environ.get("SML_NUMA_BIND", "1") != "0" else {"how": "disabled (SML_NUMA_BIND=0)"}
    xh = torch.empty(B, T, D, dtype=dtype).pin_memory()
    gh = torch.empty(B, T, D, dtype=dtype).pin_memory()
    xh.copy_(prob.x.detach())
    gh.copy_(prob.g.detach())
    yh = torch.empty(B, T, D, dtype=dtype).pin_memory()
    gxh = torch.empty(B, T, D, dtype=dtype).pin_memory()
    wr_h, wi_h, bs_h = prob.layer.weight_real.detach().cpu(), prob.layer.weight_imag.detach().cpu(), prob.layer.bias.detach().cpu()
    torch.cuda.synchronize()

    def e2e_step():
        _, _, gwr_h, gwi_h, gb_h = spectral_mix_fwd_bwd_host(xh, gh, wr_h, wi_h, bs_h, out=(yh, gxh), device=dev)
        if world > 1:   # filter-gradient sum across ranks (tiny: staged through the device for NCCL)
            flat = torch.cat([gwr_h.reshape(-1), gwi_h.reshape(-1), gb_h]).to(dev)
            dist.all_reduce(flat)
            flat.cpu()

    n_e2e = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()      # the call is synchronous (host buffers complete on return): wall clock is exact
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_e2e
    try:
        os.sched_setaffinity(0, old_affinity)
    except Exception:
        pass
    return {"value": world * B * T / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * B * T * D * esz + param_bytes,
            "d2h_bytes_per_step": 2 * B * T * D * esz + param_bytes, "ms_per_step": ms, "steps": n_e2e,
            "api": "sml_fwd_bwd_host (C ABI, host pointers): pinned host x,g,params in; y, gx, filter/bias grads out; "
                   "batch-chunked H2D / kernels / D2H pipeline on three streams",
            "host_affinity": numa}


def run_sweep(args, dev, rank, world, barrier, max_over_ranks, roof):
    """BASELINE.json configs[4]: long-context sweep T = 1K .. 128K, embed 1024, 2^20 tokens per step split over the ranks by
    batch, fp32 (or --dtype bf16); at N = 1 the CPU arm runs next to it on a batch-1 sample of every length."""
    D = 1024
    rows, tok_total, ms_total = [], 0, 0.0
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    first = True
    for T in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072):
        Bg = max(world, (1 << 20) // T)
        B = max(1, Bg // world)
        prob = Problem(B, T, D, args.dtype, dev, rank, world)
        if first:
            sampler.mark_begin()
            first = False
        ms = max_over_ranks(time_steps(prob.step, max(5, min(args.steps, 30)), max(args.warmup, 3), barrier))
        row = {"seq_len": T, "batch_per_gpu": B, "ms_per_step": ms, "value": world * B * T / (ms * 1e-3), "unit": UNIT,
               "roofline_step_frac": roof(prob.bytes_step(), ms)["frac"], "plan": prob.plan}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            dt, n, threads, kind, _ = cpu_reference_step_time(1, T, D, steps=3, warmup=1, budget_s=4.0)
            row["cpu_reference"] = {"value": T / dt, "unit": UNIT, "cores": threads, "kind": kind, "sample": f"x=(1,{T},{D}) fp32, {n} steps"}
        rows.append(row)
        tok_total += world * B * T
        ms_total += ms
        del prob
        torch.cuda.empty_cache()
    sampler.mark_end()
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": tok_total / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": max(5, min(args.steps, 30)),
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / len(rows), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"long-context sweep T=1K..128K, embed 1024, 2^20 tokens per step, {args.dtype} I/O (BASELINE.json cfg5)",
                           "io_dtype": args.dtype, "parallelism": f"batch-sharded x{world}"},
                "value_note": "tokens of all eight lengths / sum of their step times", "clocks": clocks, "sweep": rows}
        print_line(json.dumps(line))


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: everything else a library prints there (e.g. NCCL's version banner) is sent
    # to stderr by pointing fd 1 at fd 2 for the duration of the run.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    global print_line
    print_line = lines.append
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
