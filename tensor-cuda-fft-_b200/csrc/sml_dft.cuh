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

// ------------------------------------------------------------------------------------------------
// complex primitives.  Device code maps every one of them onto Blackwell's packed fp32x2 pipe
// (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2): a complex (re, im) is one 64-bit register
// pair, and ptxas folds the half-swap (im, re), the (-,+)/(+,-) sign patterns and 32-bit scalar
// broadcasts into operand modifiers (.LO_HI, .NP/.PN, .F32), so a complex multiply is 2 issue slots and a
// twiddled butterfly 3 instead of 4 and 6.  Host code (unit tests) uses the scalar definitions.
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define SML_X2 1
#else
#define SML_X2 0
#endif

#if SML_X2
__device__ __forceinline__ unsigned long long x2_pack(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ cf x2_unpack(unsigned long long v) {
    cf r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.re), "=f"(r.im) : "l"(v));
    return r;
}
// elementwise (a.re*b.re + c.re, a.im*b.im + c.im)
__device__ __forceinline__ cf x2_fma(cf a, cf b, cf c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x2_pack(a.re, a.im)), "l"(x2_pack(b.re, b.im)), "l"(x2_pack(c.re, c.im)));
    return x2_unpack(d);
}
__device__ __forceinline__ cf x2_mul(cf a, cf b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x2_pack(a.re, a.im)), "l"(x2_pack(b.re, b.im)));
    return x2_unpack(d);
}
__device__ __forceinline__ cf x2_add(cf a, cf b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x2_pack(a.re, a.im)), "l"(x2_pack(b.re, b.im)));
    return x2_unpack(d);
}
__device__ __forceinline__ cf x2_sub(cf a, cf b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x2_pack(a.re, a.im)), "l"(x2_pack(b.re, b.im)));
    return x2_unpack(d);
}
#endif

SML_HD cf cadd(cf a, cf b) {
#if SML_X2
    return x2_add(a, b);
#else
    return cf{a.re + b.re, a.im + b.im};
#endif
}
SML_HD cf csub(cf a, cf b) {
#if SML_X2
    return x2_sub(a, b);
#else
    return cf{a.re - b.re, a.im - b.im};
#endif
}
// e + s*u  (real scalar s)
SML_HD cf caxpy(float s, cf u, cf e) {
#if SML_X2
    return x2_fma(u, cf{s, s}, e);
#else
    return cf{s * u.re + e.re, s * u.im + e.im};
#endif
}
// o * (1 + i*t) = (o.re - t*o.im, o.im + t*o.re)
SML_HD cf crot(cf o, float t) {
#if SML_X2
    return x2_fma(cf{o.im, o.re}, cf{-t, t}, o);
#else
    return cf{o.re - t * o.im, o.im + t * o.re};
#endif
}
// a + i*b  and  a - i*b
SML_HD cf cadd_i(cf a, cf b) {
#if SML_X2
    return x2_add(a, cf{-b.im, b.re});
#else
    return cf{a.re - b.im, a.im + b.re};
#endif
}
SML_HD cf csub_i(cf a, cf b) {
#if SML_X2
    return x2_add(a, cf{b.im, -b.re});
#else
    return cf{a.re + b.im, a.im - b.re};
#endif
}
// a * b
SML_HD cf cmul(cf a, cf b) {
#if SML_X2
    return x2_fma(cf{a.im, a.re}, cf{-b.im, b.im}, x2_mul(a, cf{b.re, b.re}));
#else
    return cf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
#endif
}
// a * conj(b)
SML_HD cf cmulc(cf a, cf b) {
#if SML_X2
    return x2_fma(cf{a.im, a.re}, cf{b.im, -b.im}, x2_mul(a, cf{b.re, b.re}));
#else
    return cf{a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im};
#endif
}
// acc + a * b
SML_HD cf cmac(cf acc, cf a, cf b) {
#if SML_X2
    return x2_fma(cf{a.im, a.re}, cf{-b.im, b.im}, x2_fma(a, cf{b.re, b.re}, acc));
#else
    return cf{acc.re + a.re * b.re - a.im * b.im, acc.im + a.re * b.im + a.im * b.re};
#endif
}
// s * a  (real scalar)
SML_HD cf cscale(float s, cf a) {
#if SML_X2
    return x2_mul(a, cf{s, s});
#else
    return cf{s * a.re, s * a.im};
#endif
}

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
        // w = DIR * i
        if constexpr (DIR < 0) {
            lo = csub_i(e, o);
            hi = cadd_i(e, o);
        } else {
            lo = cadd_i(e, o);
            hi = csub_i(e, o);
        }
    } else {
        // w = c (1 + i t):  lo/hi = e +- c * (o (1 + i t))      -- 3 packed FMAs
        constexpr float c = Tw<N, K, DIR>::c;
        constexpr float t = Tw<N, K, DIR>::t;
        const cf u = crot(o, t);
        lo = caxpy(c, u, e);
        hi = caxpy(-c, u, e);
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
