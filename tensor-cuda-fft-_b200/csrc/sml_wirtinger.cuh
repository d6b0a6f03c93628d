// sml_wirtinger.cuh -- the complex filter multiply of the hot path as stand-alone kernels, for callers that
// already hold a spectrum (reference: /root/reference/fft_tensor/wirtinger_ops.py).
//   WirtingerGradient.forward  :34-50   out = x * w
//   WirtingerGradient.backward :53-82   gx = g conj(w) ; gw = sum_b g conj(x)
//   WirtingerSpectralFilter    :170-203 low-pass scatter of the first k = min(F, T//2) bins
// All of them are single-pass HBM-bound streams with 8- or 16-byte vector accesses.
#pragma once

#include <cuda_runtime.h>

#include "sml_dft.cuh"

namespace sml {

// x, out: (B, N) complex64; w: (N,)
__global__ void wirtinger_mul_fwd_kernel(const float2* __restrict__ x, const float2* __restrict__ w,
                                         float2* __restrict__ out, long long B, long long N) {
    const long long total = B * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const float2 a = x[i];
        const float2 c = __ldg(w + (i % N));
        out[i] = make_float2(a.x * c.x - a.y * c.y, a.x * c.y + a.y * c.x);
    }
}

// one thread per n: loops over the batch -> gx written once, gw reduced deterministically in registers
__global__ void wirtinger_mul_bwd_kernel(const float2* __restrict__ g, const float2* __restrict__ x,
                                         const float2* __restrict__ w, float2* __restrict__ gx,
                                         float2* __restrict__ gw, long long B, long long N) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float2 c = __ldg(w + n);
    float sr = 0.f, si = 0.f;
    for (long long b = 0; b < B; ++b) {
        const float2 gg = g[b * N + n];
        const float2 xx = x[b * N + n];
        gx[b * N + n] = make_float2(gg.x * c.x + gg.y * c.y, gg.y * c.x - gg.x * c.y);   // g conj(w)
        sr += gg.x * xx.x + gg.y * xx.y;                                                  // g conj(x)
        si += gg.y * xx.x - gg.x * xx.y;
    }
    gw[n] = make_float2(sr, si);
}

// out[b,f,d] = f < k ? x[b,f,d] * W[d,f] : 0        grid-stride over (B*T*D)
__global__ void wirtinger_filter_fwd_kernel(const float2* __restrict__ x, const float* __restrict__ w_re,
                                            const float* __restrict__ w_im, float2* __restrict__ out, int B, int T,
                                            int D, int F, int k) {
    const long long total = (long long)B * T * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const int f = (int)((i / D) % T);
        float2 o = make_float2(0.f, 0.f);
        if (f < k) {
            const float2 a = x[i];
            const float wr = __ldg(w_re + (size_t)d * F + f), wi = __ldg(w_im + (size_t)d * F + f);
            o = make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
        }
        out[i] = o;
    }
}

// grid (ceil(D/32), ceil(T/8)), block (32, 8); thread = (f, d), loops over b.  Also zero-fills gw columns >= k
// through the separate fill below.
__global__ void wirtinger_filter_bwd_kernel(const float2* __restrict__ g, const float2* __restrict__ x,
                                            const float* __restrict__ w_re, const float* __restrict__ w_im,
                                            float2* __restrict__ gx, float* __restrict__ gw_re,
                                            float* __restrict__ gw_im, int B, int T, int D, int F, int k) {
    const int d = blockIdx.x * 32 + threadIdx.x;
    const int f = blockIdx.y * 8 + threadIdx.y;
    if (d >= D || f >= T) return;
    if (f >= k) {
        for (int b = 0; b < B; ++b) gx[((size_t)b * T + f) * D + d] = make_float2(0.f, 0.f);
        return;
    }
    const float wr = __ldg(w_re + (size_t)d * F + f), wi = __ldg(w_im + (size_t)d * F + f);
    float sr = 0.f, si = 0.f;
    for (int b = 0; b < B; ++b) {
        const size_t i = ((size_t)b * T + f) * D + d;
        const float2 gg = g[i];
        const float2 xx = x[i];
        gx[i] = make_float2(gg.x * wr + gg.y * wi, gg.y * wr - gg.x * wi);
        sr += gg.x * xx.x + gg.y * xx.y;
        si += gg.y * xx.x - gg.x * xx.y;
    }
    gw_re[(size_t)d * F + f] = sr;
    gw_im[(size_t)d * F + f] = si;
}

}   // namespace sml
