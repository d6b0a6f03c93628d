// probe: SASS of a DFT32 built from f32x2 primitives
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../tensor-cuda-fft-_b200/csrc/sml_dft.cuh"
using namespace sml;
__global__ void probe(const float2* in, float2* out) {
    cf v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { float2 a = in[threadIdx.x + 32 * i]; v[i] = cf{a.x, a.y}; }
    Dft<32, -1>::run(v);
#pragma unroll
    for (int i = 0; i < 32; ++i) out[threadIdx.x + 32 * i] = make_float2(v[i].re, v[i].im);
}
__global__ void probe_cmul(const float2* in, float2* out) {
    cf v[8], w[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 a = in[threadIdx.x + 32 * i]; v[i] = cf{a.x, a.y}; a = in[threadIdx.x + 32 * (i+8)]; w[i] = cf{a.x, a.y}; a = in[threadIdx.x + 32 * (i+16)]; acc[i] = cf{a.x, a.y};}
#pragma unroll
    for (int i = 0; i < 8; ++i) { cf p = cmul(v[i], w[i]); cf q = cmulc(v[i], w[i]); acc[i] = cmac(acc[i], p, w[(i+1)&7]); acc[i] = cadd(acc[i], q); }
#pragma unroll
    for (int i = 0; i < 8; ++i) out[threadIdx.x + 32 * i] = make_float2(acc[i].re, acc[i].im);
}
