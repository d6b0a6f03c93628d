"""CPU proof of the tensor-core kernel's decomposition (tests/tc_model.py mirrors csrc/sml_tc.cuh index for index):
in float64 it must reproduce the oracle's closed form to rounding; with bf16 operand rounding it must stay inside the
bf16 gate of north_star (1e-2) with margin."""
import numpy as np
import pytest

from oracle import spectral_mixing_oracle as orc
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tc_model


def _case(T, DC, F, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((1, T, DC))
    g = rng.standard_normal((1, T, DC))
    wr, wi, bias = rng.standard_normal((DC, F)), rng.standard_normal((DC, F)), rng.standard_normal(DC)
    return x, g, wr, wi, bias


@pytest.mark.parametrize("T,DC,F", [(512, 4, 16), (1024, 3, 384), (2048, 2, 300), (1024, 2, 512), (512, 2, 256)])
def test_tc_decomposition_float64(T, DC, F):
    x, g, wr, wi, bias = _case(T, DC, F, seed=T + F)
    ref = orc.closed_form_f64(x, wr, wi, bias, g)
    y_ref, gx_ref, gwr_ref, gwi_ref = ref["y"], ref["gx"], ref["gw_re"], ref["gw_im"]
    y, xlow, _ = tc_model.transform(x[0], wr, wi, bias)
    assert orc.rel_l2(y, y_ref[0]) < 1e-12
    k = min(F, T // 2)
    assert orc.rel_l2(xlow.T, np.fft.fft(x[0], axis=0)[:k]) < 1e-12
    gx, _, gterms = tc_model.transform(g[0], wr, wi, None, backward=True, xlow=xlow)
    assert orc.rel_l2(gx, gx_ref[0]) < 1e-12
    assert orc.rel_l2(gterms.real, gwr_ref[:, :k]) < 1e-12      # batch of one: the per-batch term IS the gradient
    assert orc.rel_l2(gterms.imag, gwi_ref[:, :k]) < 1e-12


def test_tc_bf16_operand_rounding_inside_gate():
    T, DC, F = 2048, 4, 384
    x, g, wr, wi, bias = _case(T, DC, F, seed=7)
    xb = tc_model.bf16_round(x)
    y_ref = orc.closed_form_f64(xb, wr, wi, bias)["y"]
    y, xlow, _ = tc_model.transform(xb[0], wr, wi, bias, bf16_ops=True)
    err = orc.rel_l2(tc_model.bf16_round(y), y_ref[0])
    assert err < 6e-3, err
