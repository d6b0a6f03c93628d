"""Design prototype (numpy, float64): validates the streamed two-stage band-limited DFT decomposition
used by the CUDA fast path, including pair packing, the spectral mid phase and the backward formulas.
Mirrors thread/register index conventions of csrc/sml_fast.cuh."""
import numpy as np

def closed_form(x, wr, wi, bias, g):
    B, T, D = x.shape
    F = wr.shape[1]
    k = min(F, T // 2)
    W = (wr + 1j * wi)[:, :k].T  # (k, D)
    X = np.fft.fft(x, axis=1)[:, :k, :]
    A = X * W[None]
    t = np.arange(T)
    E = np.exp(2j * np.pi * np.outer(t, np.arange(k)) / T)  # (T,k)
    y = np.real(np.einsum('tf,bfd->btd', E, A)) / T + bias
    G = np.fft.fft(g, axis=1)[:, :k, :]
    gx = np.real(np.einsum('tf,bfd->btd', E, G * np.conj(W)[None])) / T
    gW = (G * np.conj(X)).sum(0) / T  # (k, D)
    gwr = np.zeros_like(wr); gwi = np.zeros_like(wi)
    gwr[:, :k] = gW.real.T; gwi[:, :k] = gW.imag.T
    gb = g.sum((0, 1))
    return y, gx, gwr, gwi, gb, X

def analysis_band(z, T, M, N1, N2):
    """z: (T,) complex. returns acc[f1, f2] = band value at signed freq fs=f1+N1*f2s."""
    R = T // M
    acc = np.zeros((N1, N2), complex)
    WT = lambda n: np.exp(-2j * np.pi * (n % T) / T)
    for r in range(R):
        sub = z[r::R]                       # sub[m] = z[R*m + r]
        a = sub.reshape(N1, N2)             # a[m1, m2] = sub[N2*m1+m2]
        u = np.fft.fft(a, axis=0)           # u[f1, m2]
        f1 = np.arange(N1)[:, None]; m2 = np.arange(N2)[None, :]
        u = u * WT((R * m2 + r) * f1)       # combined twiddle W_T^{(R m2 + r) f1}
        Y = np.fft.fft(u, axis=1)           # Y[f1, f2]
        f2 = np.arange(N2)
        f2s = np.where(f2 < N2 // 2, f2, f2 - N2)
        c = WT(r * N1 * f2s)                # uniform twiddle
        acc += Y * c[None, :]
    return acc

def synthesis_band(C, T, M, N1, N2):
    """C[f1,f2] band spectrum (signed). returns z (T,) = sum_fs C[fs] e^{+2 pi i fs t/T}."""
    R = T // M
    z = np.zeros(T, complex)
    WTc = lambda n: np.exp(+2j * np.pi * (n % T) / T)
    f2 = np.arange(N2)
    f2s = np.where(f2 < N2 // 2, f2, f2 - N2)
    for r in range(R):
        v = C * WTc(r * N1 * f2s)[None, :]
        v = np.fft.ifft(v, axis=1) * N2      # over f2 -> m2 : v[f1, m2]
        f1 = np.arange(N1)[:, None]; m2 = np.arange(N2)[None, :]
        v = v * WTc((R * m2 + r) * f1)
        s = np.fft.ifft(v, axis=0) * N1      # over f1 -> m1 : s[m1, m2]
        z[r::R] = s.reshape(M)
    return z

def partner(Zacc, N1, N2):
    """P[f1,f2] = Zacc at frequency -(fs)."""
    M = N1 * N2
    f = (np.arange(N1)[:, None] + N1 * np.arange(N2)[None, :])
    pf = (M - f) % M
    return Zacc[pf % N1, pf // N1]

def fast_path(x, wr, wi, bias, g, M, N1, N2):
    B, T, D = x.shape
    F = wr.shape[1]; k = min(F, T // 2)
    f = (np.arange(N1)[:, None] + N1 * np.arange(N2)[None, :])
    fs = np.where(f < M // 2, f, f - M)
    af = np.abs(fs)
    live = af < k
    if k == M // 2:
        pass
    y = np.zeros_like(x); gx = np.zeros_like(x)
    gwr = np.zeros_like(wr); gwi = np.zeros_like(wi); gb = np.zeros(D)
    Xs = np.zeros((B, D, k), complex)
    def spectral(Z, Wd0, Wd1, conjw):
        P = partner(Z, N1, N2)
        # value at +|f| and -|f|
        Zp = np.where(fs >= 0, Z, P); Zm = np.where(fs >= 0, P, Z)
        X0 = 0.5 * (Zp + np.conj(Zm)); X1 = -0.5j * (Zp - np.conj(Zm))
        w0 = np.where(live, Wd0[np.minimum(af, k - 1)], 0); w1 = np.where(live, Wd1[np.minimum(af, k - 1)], 0)
        if conjw: w0 = np.conj(w0); w1 = np.conj(w1)
        A0 = X0 * w0; A1 = X1 * w1
        C = np.where(fs > 0, 0.5 * (A0 + 1j * A1), np.where(fs < 0, 0.5 * (np.conj(A0) + 1j * np.conj(A1)), A0.real + 1j * A1.real))
        C = np.where(live, C, 0) / T
        return X0, X1, C
    for b in range(B):
        for d in range(0, D, 2):
            W0 = (wr[d] + 1j * wi[d]); W1 = (wr[d + 1] + 1j * wi[d + 1])
            Z = analysis_band(x[b, :, d] + 1j * x[b, :, d + 1], T, M, N1, N2)
            X0, X1, C = spectral(Z, W0, W1, False)
            pos = (fs >= 0) & live
            Xs[b, d, fs[pos]] = X0[pos]; Xs[b, d + 1, fs[pos]] = X1[pos]
            zz = synthesis_band(C, T, M, N1, N2)
            y[b, :, d] = zz.real + bias[d]; y[b, :, d + 1] = zz.imag + bias[d + 1]
            Zg = analysis_band(g[b, :, d] + 1j * g[b, :, d + 1], T, M, N1, N2)
            G0, G1, Cg = spectral(Zg, W0, W1, True)
            zz = synthesis_band(Cg, T, M, N1, N2)
            gx[b, :, d] = zz.real; gx[b, :, d + 1] = zz.imag
            gW0 = G0 * np.conj(X0) / T; gW1 = G1 * np.conj(X1) / T
            np.add.at(gwr[d], fs[pos], gW0[pos].real); np.add.at(gwi[d], fs[pos], gW0[pos].imag)
            np.add.at(gwr[d + 1], fs[pos], gW1[pos].real); np.add.at(gwi[d + 1], fs[pos], gW1[pos].imag)
            gb[d] += G0[0, 0].real; gb[d + 1] += G1[0, 0].real
    return y, gx, gwr, gwi, gb, Xs

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for (B, T, D, F, M, N1, N2) in [(2, 256, 4, 20, 64, 8, 8), (1, 64, 4, 32, 64, 8, 8), (2, 128, 6, 3, 64, 8, 8), (1, 512, 2, 100, 256, 16, 16), (1, 256, 2, 128, 256, 16, 16)]:
        x = rng.standard_normal((B, T, D)); g = rng.standard_normal((B, T, D))
        wr = rng.standard_normal((D, F)); wi = rng.standard_normal((D, F)); bias = rng.standard_normal(D)
        ref = closed_form(x, wr, wi, bias, g)
        out = fast_path(x, wr, wi, bias, g, M, N1, N2)
        k = min(F, T // 2)
        names = ["y", "gx", "gwr", "gwi", "gb", "X"]
        for n, a, b_ in zip(names, ref, out):
            if n == "X": a = np.transpose(a, (0, 2, 1))
            err = np.linalg.norm(a - b_) / max(np.linalg.norm(a), 1e-30)
            print((B, T, D, F, M), n, f"{err:.2e}")
            assert err < 1e-12, n
    print("OK")
