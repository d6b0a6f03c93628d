"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/spectral_mix_b200.h declares, and its host-side planning logic is right.  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as ge
    from tensor_cuda_fft_b200 import _native
    ge.build()      # incremental (make): a no-op when the library is newer than its sources, a rebuild when it is stale or missing
    return _native


def test_header_symbols_exported(native):
    hdr = open(os.path.join(ROOT, "include", "spectral_mix_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sml_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(native.EXPORTED_SYMBOLS)
    raw = ctypes.CDLL(native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert native.lib().sml_abi_version() == 3


def test_plan_selection(native):
    p = native.plan(16, 8192, 768, 384)
    assert p == {"path": "fast", "M": 1024, "R": 8, "k": 384}
    assert native.plan(8, 512, 256, 128) == {"path": "fast", "M": 256, "R": 2, "k": 128}
    assert native.plan(64, 4096, 1024, 512) == {"path": "fast", "M": 1024, "R": 4, "k": 512}
    assert native.plan(1, 131072, 1024, 512)["R"] == 128
    assert native.plan(3, 100, 32, 16)["path"] == "generic"      # T is not a multiple of any sub-transform length
    assert native.plan(2, 3072, 768, 384) == {"path": "fast", "M": 1024, "R": 3, "k": 384}    # T = R * M, R = 3
    assert native.plan(2, 1536, 96, 48) == {"path": "fast", "M": 256, "R": 6, "k": 48}
    assert native.plan(2, 1536, 768, 384) == {"path": "fast", "M": 256, "R": 6, "k": 384}     # three band bins per sub-bin
    assert native.plan(2, 1280, 1024, 512) == {"path": "fast", "M": 256, "R": 5, "k": 512}    # four
    assert native.plan(2, 1536, 1536, 768)["path"] == "generic"  # k = 768 needs M = 1024, which does not divide 1536
    assert native.plan(2, 48, 7, 3)["path"] == "generic"         # odd D
    assert native.plan(2, 16, 64, 32)["k"] == 8                  # k = min(F, T//2)
    assert native.plan(2, 16, 64, 32)["path"] == "generic"       # sub-transform would not fit in T
    assert native.plan(2, 64, 32, 16) == {"path": "fast", "M": 64, "R": 1, "k": 16}
    # wide band (k <= M < 2k): full half-spectrum shapes run with M = T/2, R = 2
    assert native.plan(4, 512, 768, 384) == {"path": "fast", "M": 256, "R": 2, "k": 256}
    assert native.plan(2, 128, 256, 128) == {"path": "fast", "M": 64, "R": 2, "k": 64}
    assert native.plan(2, 512, 384, 192) == {"path": "fast", "M": 256, "R": 2, "k": 192}
    assert native.plan(2, 2048, 2048, 1024) == {"path": "fast", "M": 1024, "R": 2, "k": 1024}   # embed 2048
    assert native.plan(2, 8192, 1536, 768) == {"path": "fast", "M": 1024, "R": 8, "k": 768}
    assert native.plan(2, 4096, 4096, 2048)["path"] == "generic"   # k = 2048 > 1024: no kernel variant
    assert native.plan(16, 8192, 768, 384, native.DTYPE_BF16)["path"] == "fast"


def test_buffer_sizes(native):
    lib = native.lib()
    assert lib.sml_xlow_bytes(16, 8192, 768, 384) == 16 * 768 * 384 * 8
    # fast path: per-batch-element filter-gradient terms (B,D,k) c64 + bias-gradient terms (B,D) f32
    assert lib.sml_workspace_bytes(16, 8192, 768, 384, 0) == 16 * 768 * 384 * 8 + 16 * 768 * 4
    assert lib.sml_workspace_bytes(3, 100, 32, 16, 0) == 3 * 32 * 16 * 8
    assert lib.sml_xlow_bytes(2, 1, 4, 2) == 0


def test_bad_arguments_report_errors(native):
    with pytest.raises(RuntimeError, match="invalid shape"):
        native.plan(0, 8, 8, 4)
    with pytest.raises(RuntimeError, match="io_dtype"):
        native.plan(1, 8, 8, 4, 7)


def test_module_surface_matches_reference():
    from tensor_cuda_fft_b200 import SpectralMixingLayer, SpectralMLPBlock, HybridSpectralAttention
    from tensor_cuda_fft_b200 import ComplexParameter, WirtingerSpectralFilter
    layer = SpectralMixingLayer(256)
    assert sum(p.numel() for p in layer.parameters()) == 65792          # BENCHMARKS.md:86
    assert list(layer.state_dict().keys()) == ["weight_real", "weight_imag", "bias"]
    assert layer.weight_real.shape == (256, 128) and layer.num_filters == 128
    assert torch.all(layer.weight_real == 1) and torch.all(layer.weight_imag == 0) and torch.all(layer.bias == 0)
    assert layer._verify_gradients is True and isinstance(layer.dropout, torch.nn.Dropout)
    frozen = SpectralMixingLayer(64, learnable=False)
    assert len(list(frozen.parameters())) == 0 and frozen.weight_real is None and frozen.bias is None
    assert SpectralMixingLayer(64, num_filters=5, dropout=0.25).dropout.p == 0.25
    blk = SpectralMLPBlock(64)
    assert blk.spectral_mix.dropout.p == 0.1 and blk.mlp[0].out_features == 256
    assert HybridSpectralAttention(64).spectral.embed_dim == 64
    assert ComplexParameter((4, 3), "ones")().dtype == torch.complex64
    assert WirtingerSpectralFilter(8, 3).weight.real.shape == (8, 3)
    with pytest.raises(ValueError):
        ComplexParameter((2, 2), "nope")


def test_no_cpu_fallback():
    from tensor_cuda_fft_b200 import SpectralMixingLayer, WirtingerGradient
    layer = SpectralMixingLayer(8)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.randn(1, 4, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        SpectralMixingLayer(8, learnable=False)(torch.randn(1, 4, 8))
    with pytest.raises(AssertionError, match="Expected embed_dim=8, got 6"):
        layer(torch.randn(1, 4, 6))
    with pytest.raises(RuntimeError, match="CUDA"):
        WirtingerGradient.apply(torch.randn(2, 3, dtype=torch.complex64), torch.randn(1, 3, dtype=torch.complex64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tensor-cuda-fft-_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_ext_plan_checks_without_gpu():
    """sml_ext_supported is pure host logic (plan + window checks): callable on the CPU box."""
    import ctypes
    from tensor_cuda_fft_b200 import _native as native
    lib = native.lib()
    ok = lambda B, T, D, F, **kw: lib.sml_ext_supported(B, T, D, F, 0, ctypes.byref(native.make_ext(**kw)))
    assert ok(16, 8192, 768, 384) == 0                                   # SpectralMLPBlock at the headline shape
    assert ok(2, 100, 32, 16) != 0 and b"fused kernels" in lib.sml_last_error()     # generic-path shape: callers compose the unfused ops
    assert ok(64, 2048, 512, 1024, T_in=1024, T_out=1024) == 0            # FixedSpectralBlock defaults: n_fft 2048 over 1024 rows
    ext = native.make_ext(T_in=1024, T_out=1024)
    ext.w_nyq = 1
    assert lib.sml_ext_supported(64, 2048, 512, 1024, 0, ctypes.byref(ext)) == 0    # full half-spectrum: the bin T/2 is carried
    assert lib.sml_ext_supported(16, 8192, 768, 384, 0, ctypes.byref(ext)) != 0     # band-limited plan: no bin T/2
    assert ok(2, 2048, 64, 1024, T_in=1023, T_out=1024) != 0 and b"multiples of R" in lib.sml_last_error()
    assert ok(2, 2048, 64, 1024, T_in=1024, T_out=512, out_row0=4) != 0 and b"out_row0" in lib.sml_last_error()
    assert ok(2, 2048, 64, 1024, T_in=2000, in_row0=100) != 0              # window does not fit the transform


def test_sml_ext_struct_matches_header():
    """The ctypes mirror of sml_ext (_native.SmlExt) has the header's fields in the header's order with matching kinds."""
    import ctypes
    import re
    from tensor_cuda_fft_b200 import _native as native
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "spectral_mix_b200.h")).read()
    body = hdr[hdr.index("typedef struct sml_ext {"): hdr.index("} sml_ext;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(const\s+)?(void|float|int)\s*(\*?)\s*(.+)$", decl)
        assert m, decl
        for name in m.group(4).split(","):
            fields.append((name.strip().lstrip("*").strip(), "ptr" if m.group(3) == "*" or name.strip().startswith("*") else m.group(2)))
    mirror = [(n, "ptr" if t is ctypes.c_void_p else ("int" if t is ctypes.c_int else "float")) for n, t in native.SmlExt._fields_]
    assert fields == mirror
