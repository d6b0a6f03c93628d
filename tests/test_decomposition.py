"""Host-logic check: the streamed two-stage band-limited decomposition the CUDA fast path implements
(pair packing, register band accumulation, partner bins, filter, transpose synthesis, Wirtinger grads)
reproduces the closed form in float64.  See tests/decomposition_model.py."""
import numpy as np
import pytest

from decomposition_model import closed_form, fast_path

CASES = [(2, 256, 4, 20, 64, 8, 8), (1, 64, 4, 32, 64, 8, 8), (2, 128, 6, 3, 64, 8, 8),
         (1, 512, 2, 100, 256, 16, 16), (1, 256, 2, 128, 256, 16, 16), (1, 2048, 2, 384, 1024, 32, 32)]


@pytest.mark.parametrize("B,T,D,F,M,N1,N2", CASES)
def test_fast_path_model(B, T, D, F, M, N1, N2):
    rng = np.random.default_rng(B * 1000 + T + D)
    x = rng.standard_normal((B, T, D)); g = rng.standard_normal((B, T, D))
    wr = rng.standard_normal((D, F)); wi = rng.standard_normal((D, F)); bias = rng.standard_normal(D)
    ref = closed_form(x, wr, wi, bias, g)
    out = fast_path(x, wr, wi, bias, g, M, N1, N2)
    for n, a, b in zip(["y", "gx", "gwr", "gwi", "gb", "X"], ref, out):
        if n == "X":
            a = np.transpose(a, (0, 2, 1))
        err = np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30)
        assert err < 1e-11, (n, err)
