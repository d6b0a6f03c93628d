// sml_generic.cuh -- direct band-limited DFT kernels: correct for ANY (T, D, F), used when the fused fast
// path does not apply (T not a multiple of the sub-transform length, odd D, very wide filter banks).  Still CUDA-only: there is no
// CPU fallback anywhere in the library.
//
//   analysis  : X[b,d,f] = sum_t x[b,t,d] W_T^{f t}                      f < k   (spectral_layers.py:88, :101)
//   synthesis : y[b,t,d] = (1/T) sum_{f<k} Re(X[b,d,f] W~[d,f] e^{+2 pi i f t/T}) + bias[d]   (:105-116)
//               with W~ = W (forward) or conj(W) (backward, wirtinger_ops.py:71)
//   filtergrad: gW[d,f]  = (1/T) sum_b G[b,d,f] conj(X[b,d,f]); gb[d] = sum_b Re G[b,d,0]     (wirtinger_ops.py:77-80)
// Twiddles come from an exact-phase table W_T^n (n = f*t mod T tracked in integers), so accuracy does not
// degrade with T.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "sml_dft.cuh"

namespace sml {

template <typename IO>
__device__ __forceinline__ float io_load(const IO* p);
template <>
__device__ __forceinline__ float io_load<float>(const float* p) {
    return __ldg(p);
}
template <>
__device__ __forceinline__ float io_load<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}
template <typename IO>
__device__ __forceinline__ void io_store(IO* p, float v);
template <>
__device__ __forceinline__ void io_store<float>(float* p, float v) {
    *p = v;
}
template <>
__device__ __forceinline__ void io_store<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
}

// grid (ceil(D/32), ceil(k/8), B), block (32, 8): thread = one (d, f) bin, loops over t.
template <typename IO>
__global__ void generic_analysis_kernel(const IO* __restrict__ x, cf* __restrict__ X, const cf* __restrict__ gtab,
                                        int T, int D, int k) {
    const int d = blockIdx.x * 32 + threadIdx.x;
    const int f = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (d >= D || f >= k) return;
    const IO* xp = x + (size_t)b * T * D + d;
    float ar = 0.f, ai = 0.f;
    int n = 0;   // (f * t) mod T
    for (int t = 0; t < T; ++t) {
        const float v = io_load<IO>(xp + (size_t)t * D);
        const float2 w = __ldg(reinterpret_cast<const float2*>(gtab) + n);
        ar = SML_FMA(v, w.x, ar);
        ai = SML_FMA(v, w.y, ai);
        n += f;
        if (n >= T) n -= T;
    }
    reinterpret_cast<float2*>(X)[((size_t)b * D + d) * k + f] = make_float2(ar, ai);
}

// grid (ceil(D/32), ceil(T/8), B), block (32, 8): thread = one output element (t, d), loops over f.
template <typename IO, bool CONJW>
__global__ void generic_synthesis_kernel(const cf* __restrict__ X, const float* __restrict__ w_re,
                                         const float* __restrict__ w_im, const float* __restrict__ bias,
                                         IO* __restrict__ out, const cf* __restrict__ gtab, int T, int D, int F, int k,
                                         float invT) {
    const int d = blockIdx.x * 32 + threadIdx.x;
    const int t = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (d >= D || t >= T) return;
    const float2* Xp = reinterpret_cast<const float2*>(X) + ((size_t)b * D + d) * k;
    float acc = 0.f;
    int n = 0;   // (f * t) mod T
    for (int f = 0; f < k; ++f) {
        const float2 xv = Xp[f];
        const float wr = __ldg(w_re + (size_t)d * F + f);
        const float wi = CONJW ? -__ldg(w_im + (size_t)d * F + f) : __ldg(w_im + (size_t)d * F + f);
        const float a_re = xv.x * wr - xv.y * wi;
        const float a_im = xv.x * wi + xv.y * wr;
        const float2 w = __ldg(reinterpret_cast<const float2*>(gtab) + n);   // (cos, -sin)
        // Re(A * e^{+i theta}) = a_re cos - a_im sin = a_re*w.x + a_im*w.y
        acc = SML_FMA(a_re, w.x, acc);
        acc = SML_FMA(a_im, w.y, acc);
        n += t;
        if (n >= T) n -= T;
    }
    float y = acc * invT;
    if (bias != nullptr) y += __ldg(bias + d);
    io_store<IO>(out + ((size_t)b * T + t) * D + d, y);
}

// one thread per (d, f < F): deterministic batch reduction; also zero-fills the columns f >= k.
__global__ void generic_filtergrad_kernel(const cf* __restrict__ G, const cf* __restrict__ X,
                                          float* __restrict__ gw_re, float* __restrict__ gw_im,
                                          float* __restrict__ gb, int B, int D, int F, int k, float invT) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)D * F) return;
    const int d = (int)(idx / F), f = (int)(idx % F);
    float sr = 0.f, si = 0.f, sb = 0.f;
    if (f < k) {
        for (int b = 0; b < B; ++b) {
            const float2 g = reinterpret_cast<const float2*>(G)[((size_t)b * D + d) * k + f];
            const float2 x = reinterpret_cast<const float2*>(X)[((size_t)b * D + d) * k + f];
            sr += g.x * x.x + g.y * x.y;
            si += g.y * x.x - g.x * x.y;
            sb += g.x;
        }
    }
    gw_re[idx] = sr * invT;
    gw_im[idx] = si * invT;
    if (f == 0) gb[d] = sb;
}

// bias gradient when k == 0 (T == 1 or an empty filter bank): gb[d] = sum_{b,t} g
template <typename IO>
__global__ void generic_biasgrad_kernel(const IO* __restrict__ g, float* __restrict__ gb, long long rows, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float s = 0.f;
    for (long long r = 0; r < rows; ++r) s += io_load<IO>(g + r * D + d);
    gb[d] = s;
}

// W_T^n = exp(-2 pi i n / T) computed in float64, stored as float2
__global__ void twiddle_table_kernel(cf* __restrict__ tab, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= T) return;
    double s, c;
    sincospi(2.0 * (double)n / (double)T, &s, &c);
    tab[n] = cf{(float)c, (float)(-s)};
}

}   // namespace sml
