// Microbenchmark: issue rate of packed fp32x2 (fma/add/mul .f32x2, sm_100) vs scalar FFMA/FADD on one B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096
#define NACC 8

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int MODE>
__global__ void k(float* out, float s) {
    // MODE 0: scalar FFMA x 2*NACC, 1: fma2 x NACC, 2: scalar FADD x 2*NACC, 3: add2 x NACC, 4: mixed scalar fma+add, 5: mixed fma2+add2
    float a[2 * NACC];
    uint64_t p[NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float2 v = make_float2(a[2 * i], a[2 * i + 1]);
        p[i] = *reinterpret_cast<uint64_t*>(&v);
    }
    float2 sv = make_float2(s, s * 0.5f);
    uint64_t s2 = *reinterpret_cast<uint64_t*>(&sv);
    float2 cv = make_float2(0.25f, 0.125f);
    uint64_t c2 = *reinterpret_cast<uint64_t*>(&cv);
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 2 * NACC; ++i) a[i] = fmaf(a[i], s, 0.25f);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) p[i] = fma2(p[i], s2, c2);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 2 * NACC; ++i) a[i] = a[i] + s;
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) p[i] = add2(p[i], s2);
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 2 * NACC; i += 2) { a[i] = fmaf(a[i], s, a[i + 1]); a[i + 1] = a[i + 1] + a[i]; }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < NACC; i += 2) { p[i] = fma2(p[i], s2, p[i + 1]); p[i + 1] = add2(p[i + 1], p[i]); }
        } else if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) p[i] = mul2(p[i], s2);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) r += a[i];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { float2 v = *reinterpret_cast<float2*>(&p[i]); r += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int nthreads, float flop_per_iter_thread) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * (2048 / nthreads > 0 ? 1 : 1);
    float* out; cudaMalloc(&out, sizeof(float) * blocks * nthreads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, nthreads>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, nthreads>>>(out, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * nthreads * ITERS * flop_per_iter_thread;   // fp32 lane-ops (fma counts 1)
    printf("%-28s threads/SM=%4d  %.3f ms  %.1f Glane-op/s  = %.1f lane-ops/clk/SM @1.965GHz  err=%s\n", name, nthreads, ms,
           ops / ms / 1e6, ops / (ms * 1e-3) / sms / 1.965e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int nt : {128, 256, 512, 1024}) {
        run<0>("scalar FFMA", nt, 2 * NACC);
        run<1>("fma.rn.f32x2", nt, 2 * NACC);
        run<2>("scalar FADD", nt, 2 * NACC);
        run<3>("add.rn.f32x2", nt, 2 * NACC);
        run<6>("mul.rn.f32x2", nt, 2 * NACC);
        run<4>("scalar FFMA+FADD dep", nt, 2 * NACC);
        run<5>("fma2+add2 dep", nt, 2 * NACC);
    }
    return 0;
}
