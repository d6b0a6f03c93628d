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

// x, out: (B, N) complex64; w: (N,).  grid (ceil(N/2 / 256), batch slices): a thread owns two neighbouring bins (16-byte accesses,
// the filter pair stays in registers) and walks its slice of the batch -- no integer division per element.
__global__ void __launch_bounds__(256) wirtinger_mul_fwd_kernel(const float2* __restrict__ x, const float2* __restrict__ w,
                                                                float2* __restrict__ out, long long B, long long N) {
    const long long n = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (n >= N) return;
    const bool pair = (n + 1 < N) && (N % 2 == 0);     // rows stay 16-byte aligned only for even N
    const float2 c0 = __ldg(w + n), c1 = (n + 1 < N) ? __ldg(w + n + 1) : make_float2(0.f, 0.f);
    for (long long b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t i = (size_t)b * N + n;
        if (pair) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(x + i));
            __stcs(reinterpret_cast<float4*>(out + i),
                   make_float4(a.x * c0.x - a.y * c0.y, a.x * c0.y + a.y * c0.x, a.z * c1.x - a.w * c1.y, a.z * c1.y + a.w * c1.x));
        } else {
            const float2 a = x[i];
            out[i] = make_float2(a.x * c0.x - a.y * c0.y, a.x * c0.y + a.y * c0.x);
            if (n + 1 < N) {
                const float2 a1 = x[i + 1];
                out[i + 1] = make_float2(a1.x * c1.x - a1.y * c1.y, a1.x * c1.y + a1.y * c1.x);
            }
        }
    }
}

// gx = g conj(w) ; gw = sum_b g conj(x)  (wirtinger_ops.py:71, :77-80).  One thread per bin, the batch walked four elements at a
// time (eight independent 8-byte loads in flight per thread); gw is reduced in registers in a fixed order: deterministic.
__global__ void __launch_bounds__(128) wirtinger_mul_bwd_kernel(const float2* __restrict__ g, const float2* __restrict__ x,
                                                                const float2* __restrict__ w, float2* __restrict__ gx,
                                                                float2* __restrict__ gw, long long B, long long N) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float2 c = __ldg(w + n);
    float sr = 0.f, si = 0.f;
    long long b = 0;
    for (; b + 4 <= B; b += 4) {
        float2 gg[4], xx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            gg[u] = __ldcs(g + (size_t)(b + u) * N + n);
            xx[u] = __ldcs(x + (size_t)(b + u) * N + n);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            __stcs(gx + (size_t)(b + u) * N + n, make_float2(gg[u].x * c.x + gg[u].y * c.y, gg[u].y * c.x - gg[u].x * c.y));   // g conj(w)
            sr += gg[u].x * xx[u].x + gg[u].y * xx[u].y;                                                                        // g conj(x)
            si += gg[u].y * xx[u].x - gg[u].x * xx[u].y;
        }
    }
    for (; b < B; ++b) {
        const float2 gg = g[(size_t)b * N + n];
        const float2 xx = x[(size_t)b * N + n];
        gx[(size_t)b * N + n] = make_float2(gg.x * c.x + gg.y * c.y, gg.y * c.x - gg.x * c.y);
        sr += gg.x * xx.x + gg.y * xx.y;
        si += gg.y * xx.x - gg.x * xx.y;
    }
    gw[n] = make_float2(sr, si);
}

// out[b,f,d] = f < k ? x[b,f,d] * W[d,f] : 0.  grid (ceil(D/32), ceil(T/8), batch slices), block (32, 8): thread = (d, f) walks
// its slice of the batch with the filter value in registers (no integer division, rows of 32 channels = 256 contiguous bytes).
__global__ void __launch_bounds__(256) wirtinger_filter_fwd_kernel(const float2* __restrict__ x, const float* __restrict__ w_re,
                                                                   const float* __restrict__ w_im, float2* __restrict__ out, int B, int T,
                                                                   int D, int F, int k) {
    const int d = blockIdx.x * 32 + threadIdx.x;
    const int f = blockIdx.y * 8 + threadIdx.y;
    if (d >= D || f >= T) return;
    const bool live = f < k;
    const float wr = live ? __ldg(w_re + (size_t)d * F + f) : 0.f, wi = live ? __ldg(w_im + (size_t)d * F + f) : 0.f;
    for (int b = blockIdx.z; b < B; b += gridDim.z) {
        const size_t i = ((size_t)b * T + f) * D + d;
        float2 o = make_float2(0.f, 0.f);
        if (live) {
            const float2 a = __ldcs(x + i);
            o = make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
        }
        __stcs(out + i, o);
    }
}

// grid (ceil(D/32), ceil(max(T, F)/8)), block (32, 8); thread = (d, f) loops over b (four batch elements in flight).  Rows
// f >= k: gx is zero-filled; gradient columns f in [k, F) are zero-filled here too (no memset launches).
__global__ void __launch_bounds__(256) wirtinger_filter_bwd_kernel(const float2* __restrict__ g, const float2* __restrict__ x,
                                                                   const float* __restrict__ w_re, const float* __restrict__ w_im,
                                                                   float2* __restrict__ gx, float* __restrict__ gw_re,
                                                                   float* __restrict__ gw_im, int B, int T, int D, int F, int k) {
    const int d = blockIdx.x * 32 + threadIdx.x;
    const int f = blockIdx.y * 8 + threadIdx.y;
    if (d >= D) return;
    if (f >= k) {
        if (f < F) {
            gw_re[(size_t)d * F + f] = 0.f;
            gw_im[(size_t)d * F + f] = 0.f;
        }
        if (f < T)
            for (int b = 0; b < B; ++b) __stcs(gx + ((size_t)b * T + f) * D + d, make_float2(0.f, 0.f));
        return;
    }
    const float wr = __ldg(w_re + (size_t)d * F + f), wi = __ldg(w_im + (size_t)d * F + f);
    float sr = 0.f, si = 0.f;
    const size_t bs = (size_t)T * D, i0 = (size_t)f * D + d;
    int b = 0;
    for (; b + 4 <= B; b += 4) {
        float2 gg[4], xx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            gg[u] = __ldcs(g + i0 + (size_t)(b + u) * bs);
            xx[u] = __ldcs(x + i0 + (size_t)(b + u) * bs);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            __stcs(gx + i0 + (size_t)(b + u) * bs, make_float2(gg[u].x * wr + gg[u].y * wi, gg[u].y * wr - gg[u].x * wi));
            sr += gg[u].x * xx[u].x + gg[u].y * xx[u].y;
            si += gg[u].y * xx[u].x - gg[u].x * xx[u].y;
        }
    }
    for (; b < B; ++b) {
        const size_t i = i0 + (size_t)b * bs;
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
