"""Numpy model (float64, optional bf16 operand rounding) of the tensor-core kernel csrc/sml_tc.cuh.

Mirrors the kernel's index conventions one to one -- the 64 stage-1 slots, the 34 band rows per channel, the 16 two-sided
band slots, the constant tables of csrc/sml_inst_tc.cu (B1, B2, inter-stage twiddles) and the mid-phase row rules -- so the
decomposition is proved on the CPU (tests/test_tc_model.py compares it with the oracle) before any GPU time is spent.
T = 64 * N2, t = N2*m1 + n, f = f1 + 64*f2.
"""
import numpy as np

RPD = 34


def bf16_round(a):
    """Round float64/float32 array to bf16 (nearest even), returned as float64."""
    f = np.asarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def table_b1():
    """B1[slot, m1]: slot 0 = Re S_0, slot 1 = Re S_32, slot 2 f1 + c = Re / Im of S_f1."""
    b1 = np.zeros((64, 64))
    m1 = np.arange(64)
    for s in range(64):
        f1 = 0 if s == 0 else 32 if s == 1 else s // 2
        imag = s >= 2 and (s & 1)
        ang = 2 * np.pi * ((f1 * m1) % 64) / 64
        b1[s] = -np.sin(ang) if imag else np.cos(ang)
    return b1


def table_b2(N2):
    """B2[(q, c'), (n, c)], f2 = q - 8."""
    b2 = np.zeros((32, 2 * N2))
    n = np.arange(N2)
    for q in range(16):
        ang = 2 * np.pi * ((n * (q - 8)) % N2) / N2
        cs, sn = np.cos(ang), np.sin(ang)
        b2[2 * q, 0::2] = cs
        b2[2 * q, 1::2] = sn
        b2[2 * q + 1, 0::2] = -sn
        b2[2 * q + 1, 1::2] = cs
    return b2


def table_tw(T, N2):
    """tw[n, j] = W_T^{n j} (j >= 1), W_T^{32 n} (j = 0)."""
    n = np.arange(N2)[:, None]
    j = np.arange(32)[None, :]
    e = (n * np.where(j == 0, 32, j)) % T
    return np.exp(-2j * np.pi * e / T)


def row_class(p):
    """band row p of a channel -> class f1 (None for the pad row)."""
    return 0 if p == 0 else 32 if p == 1 else p - 1 if p <= 32 else None


def mid_row(Z, p, k, invT, w, bias, backward, xlow_row=None):
    """One band row: Z = 16 complex accumulator slots.  Returns (band slots (16 complex), dict f -> X_f, dict f -> grad term)."""
    out = np.zeros(16, complex)
    spec, gterm = {}, {}

    def bin_(f, X, scale):
        spec[f] = X
        if backward:
            if xlow_row is not None:
                gterm[f] = X * np.conj(xlow_row[f]) * invT
            return X * np.conj(w[f]) * scale
        return X * w[f] * scale

    if 2 <= p <= 32:
        f1 = p - 1
        for f2 in range(8):
            f = f1 + 64 * f2
            if f < k:
                out[8 + f2] = bin_(f, Z[8 + f2], invT)
        for g2 in range(1, 9):
            f = 64 * g2 - f1
            if f < k:
                out[8 - g2] = np.conj(bin_(f, np.conj(Z[8 - g2]), invT))
    elif p == 0:
        a = bin_(0, Z[8], invT)
        out[8] = a.real + (0.0 if backward or bias is None else bias)
        for f2 in range(1, 8):
            f = 64 * f2
            if f < k:
                a = bin_(f, Z[8 + f2], 0.5 * invT)
                out[8 + f2] = a
                out[8 - f2] = np.conj(a)
    elif p == 1:
        for f2 in range(8):
            f = 32 + 64 * f2
            if f < k:
                a = bin_(f, Z[8 + f2], 0.5 * invT)
                out[8 + f2] = a
                out[7 - f2] = np.conj(a)
    return out, spec, gterm


def transform(x, w_re, w_im, bias, backward=False, xlow=None, bf16_ops=False, intermediates=None):
    """x: (T, DC) real (one work item: one batch element, DC channels).  w_*: (DC, F).  Returns (y, X_low (DC, k), gterms (DC, k))."""
    T, DC = x.shape
    N2 = T // 64
    F = w_re.shape[1]
    k = min(F, T // 2)
    rnd = bf16_round if bf16_ops else (lambda a: a)
    B1, B2, TW = rnd(table_b1()), rnd(table_b2(N2)), table_tw(T, N2)
    xs = rnd(x).reshape(64, N2, DC)                       # [m1, n, d]
    # stage 1: S[(n,d), slot]
    S = np.einsum("mnd,sm->nds", xs, B1)
    # twiddle + transpose: V[(d,p), (n,c)]
    V = np.zeros((DC, RPD, N2), complex)
    V[:, 0, :] = S[:, :, 0].T                             # class 0: Re S_0, not twiddled
    V[:, 1, :] = (S[:, :, 1] * TW[:, 0][:, None]).T       # class 32: Re S_32 * W_T^{32 n}
    for f1 in range(1, 32):
        V[:, f1 + 1, :] = ((S[:, :, 2 * f1] + 1j * S[:, :, 2 * f1 + 1]) * TW[:, f1][:, None]).T
    A2 = np.zeros((DC, RPD, 2 * N2))
    A2[:, :, 0::2], A2[:, :, 1::2] = rnd(V.real), rnd(V.imag)
    # stage 2: Z[(d,p), (q,c')]
    Zr = A2 @ B2.T
    Z = Zr[:, :, 0::2] + 1j * Zr[:, :, 1::2]
    # mid phase
    W = w_re + 1j * w_im
    band = np.zeros((DC, RPD, 16), complex)
    X_low = np.zeros((DC, k), complex)
    G = np.zeros((DC, k), complex)
    for d in range(DC):
        for p in range(33):
            out, spec, gterm = mid_row(Z[d, p], p, k, 1.0 / T, W[d], None if bias is None else bias[d], backward,
                                       None if xlow is None else xlow[d])
            band[d, p] = out
            for f, v in spec.items():
                X_low[d, f] = v
            for f, v in gterm.items():
                G[d, f] = v
    Ab = np.zeros((DC, RPD, 32))
    Ab[:, :, 0::2], Ab[:, :, 1::2] = rnd(band.real), rnd(band.imag)
    # stage A: Y[(d,p), (n,c)] = band @ B2 (the MN-major view of the same table)
    Yr = Ab @ B2
    Y = Yr[:, :, 0::2] + 1j * Yr[:, :, 1::2]
    # twiddle back + transpose: stage-B operand AB[(n,d), slot]
    AB = np.zeros((N2, DC, 64))
    AB[:, :, 0] = Y[:, 0, :].real.T                                   # class 0 is not twiddled
    AB[:, :, 1] = (Y[:, 1, :] * np.conj(TW[:, 0])[None, :]).real.T    # class 32
    for f1 in range(1, 32):
        v = Y[:, f1 + 1, :] * np.conj(TW[:, f1])[None, :]
        AB[:, :, 2 * f1], AB[:, :, 2 * f1 + 1] = v.real.T, v.imag.T
    AB = rnd(AB)
    # stage B: y[(n,d), m1]
    yv = np.einsum("nds,sm->mnd", AB, B1)
    if intermediates is not None:     # layouts of the kernel's SML_TC_DUMP regions (tools/tc_dump_check.py)
        Vd = np.zeros((N2, DC, 64))
        Vd[:, :, 0], Vd[:, :, 1] = S[:, :, 0], S[:, :, 1]
        for f1 in range(1, 32):
            Vd[:, :, 2 * f1], Vd[:, :, 2 * f1 + 1] = V[:, f1 + 1, :].real.T, V[:, f1 + 1, :].imag.T
        intermediates.update(V=Vd, Z=Zr.reshape(DC * RPD, 32), band=Ab.reshape(DC * RPD, 32), AB=AB, S=S)
    return yv.reshape(T, DC), X_low, G
