"""ctypes binding of libspectral_mix_b200.so (C ABI in include/spectral_mix_b200.h).

The library is the product: if it is missing or fails to load, every op raises -- there is no PyTorch or
CPU fallback (the reference's "try-import, else warn and fall back" convention, fft_tensor/tensor.py:13-18,
is deliberately NOT reproduced).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SML_LIB_PATH") or os.path.join(_PKG_DIR, "libspectral_mix_b200.so")   # override: A/B of builds
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

DTYPE_F32 = 0
DTYPE_BF16 = 1
PATH_FAST = 1
PATH_GENERIC = 2

_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-j8", "-C", CSRC_DIR], capture_output=True, text=True)
    if res.returncode != 0:
        raise NativeLibraryError("building libspectral_mix_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


class SmlExt(ctypes.Structure):
    """``sml_ext`` of include/spectral_mix_b200.h: optional block prologue / epilogue of the extended entry points."""
    _fields_ = [("row_stats", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("chan_scale", ctypes.c_void_p),
                ("w_nyq", ctypes.c_void_p), ("sb_re", ctypes.c_void_p), ("sb_im", ctypes.c_void_p), ("sb_nyq", ctypes.c_void_p),
                ("x_nyq", ctypes.c_void_p), ("g_nyq", ctypes.c_void_p),
                ("d_core", ctypes.c_void_p), ("d_q", ctypes.c_void_p), ("q_re", ctypes.c_void_p), ("q_im", ctypes.c_void_p),
                ("q_nyq", ctypes.c_void_p),
                ("h_re", ctypes.c_void_p), ("h_im", ctypes.c_void_p), ("h_nyq", ctypes.c_void_p), ("chan", ctypes.c_void_p),
                ("bg", ctypes.c_void_p), ("hpart", ctypes.c_void_p),
                ("T_in", ctypes.c_int), ("in_row0", ctypes.c_int), ("T_out", ctypes.c_int), ("out_row0", ctypes.c_int)]


def make_ext(row_stats=None, residual=None, chan_scale=None, w_nyq=None, sb_re=None, sb_im=None, sb_nyq=None, x_nyq=None,
             g_nyq=None, T_in=0, in_row0=0, T_out=0, out_row0=0, d_core=None, d_q=None, q_re=None, q_im=None, q_nyq=None, h_re=None, h_im=None, h_nyq=None, chan=None, bg=None, hpart=None) -> SmlExt:
    """Build an ``sml_ext`` from torch tensors (or None).  The caller keeps the tensors alive for the duration of the call."""
    p = lambda t: None if t is None else t.data_ptr()
    return SmlExt(p(row_stats), p(residual), p(chan_scale), p(w_nyq), p(sb_re), p(sb_im), p(sb_nyq), p(x_nyq), p(g_nyq),
                  p(d_core), p(d_q), p(q_re), p(q_im), p(q_nyq), p(h_re), p(h_im), p(h_nyq), p(chan), p(bg), p(hpart), int(T_in), int(in_row0), int(T_out), int(out_row0))


def _declare(lib):
    c_int, c_void_p, c_size_t, c_ll = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_longlong
    ip = ctypes.POINTER(ctypes.c_int)
    lib.sml_abi_version.restype = c_int
    lib.sml_abi_version.argtypes = []
    lib.sml_last_error.restype = ctypes.c_char_p
    lib.sml_last_error.argtypes = []
    lib.sml_launch_count.restype = ctypes.c_ulonglong
    lib.sml_launch_count.argtypes = []
    lib.sml_debug_dump.restype = c_int
    lib.sml_debug_dump.argtypes = []
    lib.sml_plan.restype = c_int
    lib.sml_plan.argtypes = [c_int, c_int, c_int, c_int, c_int, ip, ip, ip, ip]
    lib.sml_xlow_bytes.restype = c_size_t
    lib.sml_xlow_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.sml_workspace_bytes.restype = c_size_t
    lib.sml_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.sml_forward.restype = c_int
    lib.sml_forward.argtypes = [c_void_p] * 6 + [c_int] * 5 + [c_void_p]
    lib.sml_backward.restype = c_int
    lib.sml_backward.argtypes = [c_void_p] * 9 + [c_size_t] + [c_int] * 5 + [c_void_p]
    lib.sml_backward_allreduce.restype = c_int
    lib.sml_backward_allreduce.argtypes = [c_void_p] * 9 + [c_size_t] + [c_int] * 5 + [c_void_p] * 3
    lib.sml_fwd_bwd_host.restype = c_int
    lib.sml_fwd_bwd_host.argtypes = [c_void_p] * 10 + [c_int] * 6
    lib.sml_host_release.restype = c_int
    lib.sml_host_release.argtypes = []
    lib.sml_release.restype = c_int
    lib.sml_release.argtypes = []
    lib.sml_wirtinger_mul_forward.restype = c_int
    lib.sml_wirtinger_mul_forward.argtypes = [c_void_p] * 3 + [c_ll, c_ll, c_void_p]
    lib.sml_wirtinger_mul_backward.restype = c_int
    lib.sml_wirtinger_mul_backward.argtypes = [c_void_p] * 5 + [c_ll, c_ll, c_void_p]
    lib.sml_wirtinger_filter_forward.restype = c_int
    lib.sml_wirtinger_filter_forward.argtypes = [c_void_p] * 4 + [c_int] * 4 + [c_void_p]
    lib.sml_wirtinger_filter_backward.restype = c_int
    lib.sml_wirtinger_filter_backward.argtypes = [c_void_p] * 7 + [c_int] * 4 + [c_void_p]
    ext_p = ctypes.POINTER(SmlExt)
    lib.sml_ext_hpart_rows.restype = c_int
    lib.sml_ext_hpart_rows.argtypes = [c_int] * 5
    lib.sml_ext_supported.restype = c_int
    lib.sml_ext_supported.argtypes = [c_int] * 5 + [ext_p]
    lib.sml_forward_ext.restype = c_int
    lib.sml_forward_ext.argtypes = [c_void_p] * 6 + [c_int] * 5 + [ext_p, c_void_p]
    lib.sml_backward_ext.restype = c_int
    lib.sml_backward_ext.argtypes = [c_void_p] * 9 + [c_size_t] + [c_int] * 5 + [ext_p, c_void_p]
    lib.sml_ln_stats.restype = c_int
    lib.sml_ln_stats.argtypes = [c_void_p] * 2 + [c_int] * 5 + [ctypes.c_float, c_int, c_void_p]
    lib.sml_ln_backward.restype = c_int
    lib.sml_ln_backward.argtypes = [c_void_p] * 6 + [c_int] * 6 + [c_void_p]
    lib.sml_spectral_ema_scan.restype = c_int
    lib.sml_spectral_ema_scan.argtypes = [c_void_p] * 5 + [c_int] * 4 + [c_void_p]


EXPORTED_SYMBOLS = (
    "sml_abi_version", "sml_last_error", "sml_plan", "sml_xlow_bytes", "sml_workspace_bytes",
    "sml_forward", "sml_backward", "sml_fwd_bwd_host", "sml_host_release",
    "sml_wirtinger_mul_forward", "sml_wirtinger_mul_backward",
    "sml_wirtinger_filter_forward", "sml_wirtinger_filter_backward",
    "sml_launch_count", "sml_debug_dump", "sml_release", "sml_backward_allreduce",
    "sml_ext_supported", "sml_ext_hpart_rows", "sml_forward_ext", "sml_backward_ext", "sml_ln_stats", "sml_ln_backward", "sml_spectral_ema_scan",
)


def lib():
    """Load (once) and return the native library; raises NativeLibraryError if it is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NativeLibraryError(
                        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"or `make -C {CSRC_DIR}`. There is no fallback path.")
                try:
                    handle = ctypes.CDLL(LIB_PATH)
                except OSError as e:   # pragma: no cover
                    raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
                _declare(handle)
                _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("spectral_mix_b200: " + lib().sml_last_error().decode("utf-8", "replace"))


def plan(B: int, T: int, D: int, F: int, io_dtype: int = DTYPE_F32) -> dict:
    path, M, R, k = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib().sml_plan(B, T, D, F, io_dtype, ctypes.byref(path), ctypes.byref(M), ctypes.byref(R), ctypes.byref(k)))
    return {"path": "fast" if path.value == PATH_FAST else "generic", "M": M.value, "R": R.value, "k": k.value}


def launch_count() -> int:
    return int(lib().sml_launch_count())
