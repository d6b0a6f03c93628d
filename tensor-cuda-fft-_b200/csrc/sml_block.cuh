// sml_block.cuh -- the row-wise halves of the blocks that host the transform:
//   * LayerNorm row statistics for the normalise-on-load prologue of the fused kernel (sml_fast.cuh, EXT), reference:
//     SpectralMLPBlock.forward `x + spectral_mix(norm1(x))`, /root/reference/fft_tensor/spectral_layers.py:161, :185, and
//     FixedSpectralBlock.forward `x = self.ln(x)`, /root/reference/fft_lm/train_fixed_full.py:502-503;
//   * the LayerNorm backward fused with the residual branch's gradient add;
//   * SpectralEMA.scan, /root/reference/fft_lm/spectral_ssm.py:78-125: the chunk-by-chunk phase-aligned / polar EMA of the
//     inference path, one thread per (batch element, frequency) walking the S chunks (the recurrence is not associative:
//     the rotation depends on the running state's phase, so it is parallel over (b, f) only).
// All kernels are HBM-bound streams: one warp per row, 16-byte accesses where the row allows it.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sml_generic.cuh"   // io_load / io_store

namespace sml {

// 16 bytes of a row as floats: 4 (fp32) or 8 (bf16) elements
template <typename IO>
struct Vec16;
template <>
struct Vec16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// stats[b, t] = {mean, rstd} of input row t - row0 of batch element b (over D), {0, 0} for transform rows outside the input.
// One warp per transform row; two sweeps over the row (the second one hits L1): mean, then the centred second moment
// (the same biased variance torch.nn.LayerNorm uses, computed without cancellation).
template <typename IO, bool VEC>
__global__ void __launch_bounds__(256) ln_stats_kernel(const IO* __restrict__ x, float2* __restrict__ stats, long long nrows, int T,
                                                        int T_in, int row0, int D, float eps) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const long long b = row / T;
    const int t = (int)(row - b * T) - row0;
    if (t < 0 || t >= T_in) {
        if (lane == 0) stats[row] = make_float2(0.f, 0.f);
        return;
    }
    const IO* xr = x + ((size_t)b * T_in + t) * D;
    float s = 0.f;
    if constexpr (VEC) {
        constexpr int N = Vec16<IO>::N;
        for (int d = lane * N; d < D; d += 32 * N) {
            float v[N];
            Vec16<IO>::load(xr + d, v);
#pragma unroll
            for (int i = 0; i < N; ++i) s += v[i];
        }
    } else {
        for (int d = lane; d < D; d += 32) s += io_load<IO>(xr + d);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
    if constexpr (VEC) {
        constexpr int N = Vec16<IO>::N;
        for (int d = lane * N; d < D; d += 32 * N) {
            float v[N];
            Vec16<IO>::load(xr + d, v);
#pragma unroll
            for (int i = 0; i < N; ++i) q = fmaf(v[i] - mean, v[i] - mean, q);
        }
    } else {
        for (int d = lane; d < D; d += 32) {
            const float c = io_load<IO>(xr + d) - mean;
            q = fmaf(c, c, q);
        }
    }
    const float var = warp_sum(q) / (float)D;
    if (lane == 0) stats[row] = make_float2(mean, rsqrtf(var + eps));
}

// LayerNorm backward (no affine: gamma / beta were folded into the filter on the host) fused with the residual branch:
//   gx[b,i,:] = rstd * (gh' - mean_d(gh') - x^ * mean_d(gh' * x^)) + g_res[b,i,:]     x^ = (x - mean) * rstd,  gh' = gh + cadd[b,:]
// gh = dL/dx^ from sml_backward_ext; g_res (nullable) = the gradient that reached the block output (the skip connection);
// cadd (nullable, (B, D) fp32) = a per-(batch element, channel) constant on dL/dx^ (the gradient of a mean over the rows,
// e.g. the pooled context of FixedSpectralBlock's gate).
template <typename IO, bool VEC>
__global__ void __launch_bounds__(256) ln_backward_kernel(const IO* __restrict__ gh, const IO* __restrict__ x, const float2* __restrict__ stats,
                                                           const IO* __restrict__ gres, const float* __restrict__ cadd, IO* __restrict__ gx,
                                                           long long nrows, int T, int T_in, int row0, int D) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;   // rows of the INPUT tensor: (b, i)
    const int lane = threadIdx.x & 31;
    const long long b = row / T_in;
    const int i = (int)(row - b * T_in);
    const float2 st = __ldg(stats + (size_t)b * T + i + row0);
    const float mean = st.x, rstd = st.y;
    const IO* ghr = gh + (size_t)row * D;
    const IO* xr = x + (size_t)row * D;
    const float* car = cadd != nullptr ? cadd + (size_t)b * D : nullptr;
    float s1 = 0.f, s2 = 0.f;
    if constexpr (VEC) {
        constexpr int N = Vec16<IO>::N;
        for (int d = lane * N; d < D; d += 32 * N) {
            float g[N], v[N];
            Vec16<IO>::load(ghr + d, g);
            Vec16<IO>::load(xr + d, v);
            if (car != nullptr) {
#pragma unroll
                for (int j = 0; j < N; ++j) g[j] += __ldg(car + d + j);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) { s1 += g[j]; s2 = fmaf(g[j], (v[j] - mean) * rstd, s2); }
        }
    } else {
        for (int d = lane; d < D; d += 32) {
            const float g = io_load<IO>(ghr + d) + (car != nullptr ? __ldg(car + d) : 0.f);
            s1 += g;
            s2 = fmaf(g, (io_load<IO>(xr + d) - mean) * rstd, s2);
        }
    }
    const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
    IO* outr = gx + (size_t)row * D;
    const IO* rr = gres != nullptr ? gres + (size_t)row * D : nullptr;
    if constexpr (VEC) {
        constexpr int N = Vec16<IO>::N;
        for (int d = lane * N; d < D; d += 32 * N) {
            float g[N], v[N], r[N], o[N];
            Vec16<IO>::load(ghr + d, g);
            Vec16<IO>::load(xr + d, v);
            if (car != nullptr) {
#pragma unroll
                for (int j = 0; j < N; ++j) g[j] += __ldg(car + d + j);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) r[j] = 0.f;
            if (rr != nullptr) Vec16<IO>::load(rr + d, r);
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = fmaf(rstd, g[j] - m1 - (v[j] - mean) * rstd * m2, r[j]);
            Vec16<IO>::store(outr + d, o);
        }
    } else {
        for (int d = lane; d < D; d += 32) {
            const float g = io_load<IO>(ghr + d) + (car != nullptr ? __ldg(car + d) : 0.f);
            const float xh = (io_load<IO>(xr + d) - mean) * rstd;
            const float r = rr != nullptr ? io_load<IO>(rr + d) : 0.f;
            io_store<IO>(outr + d, fmaf(rstd, g - m1 - xh * m2, r));
        }
    }
}

// SpectralEMA.scan (spectral_ssm.py:107-125 over .update, :78-105): state <- update(state, chunk_s) for s = 0..S-1.
//   aligned: state <- a * state * exp(i (angle(chunk) - angle(state))) + (1 - rho) * chunk ,  a = rho * exp(i theta)
//   polar  : state <- (rho |state| + (1 - rho) |chunk|) * exp(i angle(chunk))
// exp(i (angle(c) - angle(s))) = (c / |c|) * conj(s / |s|) with torch.angle(0) = 0 (unit phase for a zero operand): no
// transcendental is needed.  chunks: (B, S, F) complex64; state_in (nullable = zeros) / state_out: (B, F) complex64;
// rho, theta: (F,).  Threads run along F (coalesced), one per (b, f).
__device__ __forceinline__ float2 unit_phase(float2 z) {
    const float m = hypotf(z.x, z.y);
    // torch.angle(0 + 0j) = atan2(0, 0) = 0 (for +0 real part); a zero with a negative-zero real part has angle pi
    if (m == 0.f) return make_float2(signbit(z.x) ? -1.f : 1.f, 0.f);
    return make_float2(z.x / m, z.y / m);
}
__global__ void __launch_bounds__(256) spectral_ema_scan_kernel(const float2* __restrict__ chunks, const float2* __restrict__ state_in,
                                                                 const float* __restrict__ rho, const float* __restrict__ theta,
                                                                 float2* __restrict__ state_out, int B, int S, int F, int polar) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F) return;
    const int b = (int)(idx / F), f = (int)(idx - (long long)b * F);
    const float r = __ldg(rho + f);
    float sn, cs;
    sincosf(__ldg(theta + f), &sn, &cs);
    const float2 a = make_float2(r * cs, r * sn);
    const float omr = 1.f - r;
    float2 st = state_in != nullptr ? __ldg(state_in + idx) : make_float2(0.f, 0.f);
    const float2* cp = chunks + (size_t)b * S * F + f;
    for (int s = 0; s < S; ++s) {
        const float2 c = __ldg(cp + (size_t)s * F);
        const float2 uc = unit_phase(c);
        if (polar) {
            const float m = r * hypotf(st.x, st.y) + omr * hypotf(c.x, c.y);
            st = make_float2(m * uc.x, m * uc.y);
        } else {
            const float2 us = unit_phase(st);
            // rot = uc * conj(us); aligned = st * rot; st = a * aligned + (1 - rho) * c
            const float2 rot = make_float2(uc.x * us.x + uc.y * us.y, uc.y * us.x - uc.x * us.y);
            const float2 al = make_float2(st.x * rot.x - st.y * rot.y, st.x * rot.y + st.y * rot.x);
            st = make_float2(a.x * al.x - a.y * al.y + omr * c.x, a.x * al.y + a.y * al.x + omr * c.y);
        }
    }
    state_out[idx] = st;
}

}   // namespace sml
