// sml_dft.cuh -- fully unrolled in-register DFTs of length 2..64 (natural order in, natural order out).
//
// Building block of the fused SpectralMixingLayer kernels: every thread owns NR complex points of one
// (batch, channel-pair) column and runs a radix-2 decimation-in-time recursion on them entirely in
// registers.  All twiddles are compile-time constants; non-trivial butterflies use the 6-FMA
// "tangent" form  e +- c*(o*(1 + i*t)),  t = tan(theta), instead of the 8-op textbook form.
// Unused outputs are removed by the compiler (dead-code elimination), which is how the kernels get
// output-pruned transforms for free.
//
// Replaces (together with sml_fast.cuh) the library FFT calls of the reference:
//   torch.fft.fft  -- /root/reference/fft_tensor/spectral_layers.py:88
//   torch.fft.ifft -- /root/reference/fft_tensor/spectral_layers.py:112
// The header is host-compilable so the butterflies are unit-tested on the CPU (tests/test_dft_host.py).
#pragma once

#if defined(__CUDACC__)
#define SML_HD __host__ __device__ __forceinline__
#else
#define SML_HD inline
#endif

namespace sml {

struct cf {
    float re, im;
};

SML_HD cf cmul(cf a, cf b) { return cf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
// a * conj(b)
SML_HD cf cmulc(cf a, cf b) { return cf{a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im}; }
SML_HD cf cadd(cf a, cf b) { return cf{a.re + b.re, a.im + b.im}; }
SML_HD cf csub(cf a, cf b) { return cf{a.re - b.re, a.im - b.im}; }

#if defined(__CUDA_ARCH__)
#define SML_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define SML_FMA(a, b, c) ((a) * (b) + (c))
#endif

namespace detail {
// cos(2*pi*j/64), j = 0..16 (quarter wave), rounded from float64
constexpr double kCos64[17] = {1.0,
                               0.9951847266721969,
                               0.9807852804032304,
                               0.9569403357322088,
                               0.9238795325112867,
                               0.881921264348355,
                               0.8314696123025452,
                               0.773010453362737,
                               0.7071067811865476,
                               0.6343932841636455,
                               0.5555702330196023,
                               0.4713967368259978,
                               0.38268343236508984,
                               0.29028467725446233,
                               0.19509032201612833,
                               0.09801714032956077,
                               0.0};
constexpr double cos64(int j) {
    j = ((j % 64) + 64) % 64;
    return j <= 16 ? kCos64[j] : j <= 32 ? -kCos64[32 - j] : j <= 48 ? -kCos64[j - 32] : kCos64[64 - j];
}
constexpr double sin64(int j) { return cos64(j - 16); }

// w = W_N^{K} for DIR=-1 (forward, e^{-2 pi i K/N}) or its conjugate for DIR=+1.   w = c + i*s
template <int N, int K, int DIR>
struct Tw {
    static_assert(64 % N == 0, "N must divide 64");
    static constexpr int J = K * (64 / N);
    static constexpr float c = (float)cos64(J);
    static constexpr float s = (float)(DIR * sin64(J));
    static constexpr double cd = cos64(J);
    static constexpr float t = (float)(cd != 0.0 ? DIR * sin64(J) / cd : 0.0);   // tan form, unused when c == 0
};

// lo = e + w*o ; hi = e - w*o
template <int N, int K, int DIR>
SML_HD void butterfly(const cf e, const cf o, cf& lo, cf& hi) {
    if constexpr (K == 0) {
        lo = cadd(e, o);
        hi = csub(e, o);
    } else if constexpr (4 * K == N) {
        // w = DIR * i : w*o = DIR * (-o.im, o.re)
        const cf t = (DIR < 0) ? cf{o.im, -o.re} : cf{-o.im, o.re};
        lo = cadd(e, t);
        hi = csub(e, t);
    } else {
        constexpr float c = Tw<N, K, DIR>::c;
        constexpr float t = Tw<N, K, DIR>::t;
        const float ur = SML_FMA(-t, o.im, o.re);
        const float ui = SML_FMA(t, o.re, o.im);
        lo.re = SML_FMA(c, ur, e.re);
        lo.im = SML_FMA(c, ui, e.im);
        hi.re = SML_FMA(-c, ur, e.re);
        hi.im = SML_FMA(-c, ui, e.im);
    }
}

template <int N, int DIR, int K>
struct CombineLoop {
    SML_HD static void run(cf (&v)[N], const cf (&e)[N / 2], const cf (&o)[N / 2]) {
        butterfly<N, K, DIR>(e[K], o[K], v[K], v[K + N / 2]);
        if constexpr (K + 1 < N / 2) CombineLoop<N, DIR, K + 1>::run(v, e, o);
    }
};
}   // namespace detail

// In-place DFT of v[0..N): v[k] <- sum_n v[n] * exp(DIR * 2 pi i n k / N).  No normalisation.
template <int N, int DIR>
struct Dft {
    SML_HD static void run(cf (&v)[N]) {
        cf e[N / 2], o[N / 2];
#pragma unroll
        for (int n = 0; n < N / 2; ++n) {
            e[n] = v[2 * n];
            o[n] = v[2 * n + 1];
        }
        Dft<N / 2, DIR>::run(e);
        Dft<N / 2, DIR>::run(o);
        detail::CombineLoop<N, DIR, 0>::run(v, e, o);
    }
};

template <int DIR>
struct Dft<1, DIR> {
    SML_HD static void run(cf (&)[1]) {}
};

}   // namespace sml
