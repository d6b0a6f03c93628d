#include <stdint.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pack(float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__global__ void swz(const float2* in, float2* out, float s, float t) {
    float2 a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64];
    uint64_t A = pack(a.x, a.y), B = pack(b.x, b.y), C = pack(c.x, c.y);
    uint64_t S = pack(s, s);              // broadcast scalar
    uint64_t Bs = pack(b.y, b.x);         // swapped
    uint64_t NT = pack(-t, t);            // sign pattern
    uint64_t r1 = fma2(A, S, C);          // scalar broadcast multiply
    uint64_t r2 = fma2(Bs, NT, B);        // (b.re - t b.im, b.im + t b.re)
    uint64_t r3 = fma2(r2, S, r1);
    float x, y; unpack(r3, x, y);
    out[threadIdx.x] = make_float2(x, y);
}
