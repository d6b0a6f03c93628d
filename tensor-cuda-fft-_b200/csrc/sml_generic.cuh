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

// Both transforms are dense "DFT as GEMM" products tiled like an fp32 GEMM: a 256-thread CTA owns a 64 x 64 output tile,
// every thread a 4 x 4 register block; the reduction axis is streamed through shared memory in chunks of GK together with
// the matching twiddles W_T^{(f t) mod T} (exact integer phase, gathered from the table once per chunk and tile).
constexpr int GK = 16;   // reduction chunk

__device__ __forceinline__ float2 generic_twiddle(const cf* __restrict__ gtab, int f, int t, int T) {
    const unsigned long long n = ((unsigned long long)(unsigned)f * (unsigned long long)(unsigned)t) % (unsigned)T;
    return __ldg(reinterpret_cast<const float2*>(gtab) + n);   // (cos, -sin)
}

// analysis: X[b,d,f] = sum_t x[b,t,d] W_T^{f t}.  grid (ceil(D/64), ceil(k/64), B), block 256: thread = 4 f x 4 d.
template <typename IO>
__global__ void __launch_bounds__(256) generic_analysis_kernel(const IO* __restrict__ x, cf* __restrict__ X,
                                                               const cf* __restrict__ gtab, int T, int D, int k) {
    __shared__ __align__(16) float xs[GK][64];     // [t][d]
    __shared__ __align__(16) float2 ws[GK][64];    // [t][f]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int d0 = blockIdx.x * 64, f0 = blockIdx.y * 64, b = blockIdx.z;
    const IO* xb = x + (size_t)b * T * D;
    float ar[4][4], ai[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ar[i][j] = ai[i][j] = 0.f;
    for (int t0 = 0; t0 < T; t0 += GK) {
#pragma unroll
        for (int e = tid; e < GK * 64; e += 256) {
            const int tt = e >> 6, c = e & 63;
            const int t = t0 + tt;
            xs[tt][c] = (t < T && d0 + c < D) ? io_load<IO>(xb + (size_t)t * D + d0 + c) : 0.f;
            ws[tt][c] = (t < T && f0 + c < k) ? generic_twiddle(gtab, f0 + c, t, T) : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < GK; ++tt) {
            const float4 xv = *reinterpret_cast<const float4*>(&xs[tt][4 * tx]);
            const float4 w01 = *reinterpret_cast<const float4*>(&ws[tt][4 * ty]);
            const float4 w23 = *reinterpret_cast<const float4*>(&ws[tt][4 * ty + 2]);
            const float xd[4] = {xv.x, xv.y, xv.z, xv.w};
            const float wr[4] = {w01.x, w01.z, w23.x, w23.z}, wi[4] = {w01.y, w01.w, w23.y, w23.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ar[i][j] = SML_FMA(xd[j], wr[i], ar[i][j]);
                    ai[i][j] = SML_FMA(xd[j], wi[i], ai[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int d = d0 + 4 * tx + j;
        if (d >= D) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = f0 + 4 * ty + i;
            if (f < k) reinterpret_cast<float2*>(X)[((size_t)b * D + d) * k + f] = make_float2(ar[i][j], ai[i][j]);
        }
    }
}

// synthesis: y[b,t,d] = (1/T) sum_{f<k} Re(X[b,d,f] W~[d,f] e^{+2 pi i f t/T}) + bias[d].
// grid (ceil(D/64), ceil(T/64), B), block 256: thread = 4 t x 4 d; the filter is applied while the spectrum chunk is staged.
template <typename IO, bool CONJW>
__global__ void __launch_bounds__(256) generic_synthesis_kernel(const cf* __restrict__ X, const float* __restrict__ w_re,
                                                                const float* __restrict__ w_im, const float* __restrict__ bias,
                                                                IO* __restrict__ out, const cf* __restrict__ gtab, int T, int D,
                                                                int F, int k, float invT) {
    __shared__ __align__(16) float2 as[GK][64];    // [f][d]  A = X * W~
    __shared__ __align__(16) float2 ws[GK][64];    // [f][t]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int d0 = blockIdx.x * 64, t0 = blockIdx.y * 64, b = blockIdx.z;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int f0 = 0; f0 < k; f0 += GK) {
#pragma unroll
        for (int e = tid; e < GK * 64; e += 256) {
            {   // spectrum chunk: consecutive threads take consecutive f of one channel (X is (B, D, k))
                const int c = e / GK, ff = e % GK;
                const int d = d0 + c, f = f0 + ff;
                float2 a = make_float2(0.f, 0.f);
                if (d < D && f < k) {
                    const float2 xv = reinterpret_cast<const float2*>(X)[((size_t)b * D + d) * k + f];
                    const float wr = __ldg(w_re + (size_t)d * F + f);
                    const float wi = CONJW ? -__ldg(w_im + (size_t)d * F + f) : __ldg(w_im + (size_t)d * F + f);
                    a = make_float2(xv.x * wr - xv.y * wi, xv.x * wi + xv.y * wr);
                }
                as[ff][c] = a;
            }
            {
                const int ff = e >> 6, c = e & 63;
                const int f = f0 + ff, t = t0 + c;
                ws[ff][c] = (f < k && t < T) ? generic_twiddle(gtab, f, t, T) : make_float2(0.f, 0.f);
            }
        }
        __syncthreads();
#pragma unroll
        for (int ff = 0; ff < GK; ++ff) {
            const float4 a01 = *reinterpret_cast<const float4*>(&as[ff][4 * tx]);
            const float4 a23 = *reinterpret_cast<const float4*>(&as[ff][4 * tx + 2]);
            const float4 w01 = *reinterpret_cast<const float4*>(&ws[ff][4 * ty]);
            const float4 w23 = *reinterpret_cast<const float4*>(&ws[ff][4 * ty + 2]);
            const float are[4] = {a01.x, a01.z, a23.x, a23.z}, aim[4] = {a01.y, a01.w, a23.y, a23.w};
            const float wr[4] = {w01.x, w01.z, w23.x, w23.z}, wi[4] = {w01.y, w01.w, w23.y, w23.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // Re(A * e^{+i theta}) = a_re cos - a_im sin = a_re*w.x + a_im*w.y   (table holds (cos, -sin))
                    acc[i][j] = SML_FMA(are[j], wr[i], acc[i][j]);
                    acc[i][j] = SML_FMA(aim[j], wi[i], acc[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + 4 * ty + i;
        if (t >= T) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = d0 + 4 * tx + j;
            if (d >= D) continue;
            float y = acc[i][j] * invT;
            if (bias != nullptr) y += __ldg(bias + d);
            io_store<IO>(out + ((size_t)b * T + t) * D + d, y);
        }
    }
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
