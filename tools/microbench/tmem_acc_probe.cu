// tmem_acc_probe.cu -- what does it cost to keep the kernel's band accumulators in TMEM instead of registers?
// (a P = 8 CTA of 256 threads needs <= 128 registers for two CTAs per SM; the 24 complex
//  accumulators of a thread are 48 of them.)  Each thread owns 48 32-bit TMEM words (its lane, 48 columns); per "pass" it loads
// them (3 x tcgen05.ld.32x32b.x16), adds 24 complex products, stores them back (3 x tcgen05.st.32x32b.x16).  Compared with the
// same arithmetic on a register-resident array, alone and next to a block of independent FFMA work (the transforms).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_acc_probe tmem_acc_probe.cu ; run: ./tmem_acc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD16(r, addr)                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"  \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                 \
                 : "r"(addr))
#define ST16(r, addr)                                                                                                     \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"  \
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),     \
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(addr)           \
                 : "memory")

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: accumulators in registers; 1: in TMEM.  WORK: independent FFMA instructions per pass next to the accumulation.
template <int MODE, int WORK>
__global__ void __launch_bounds__(256, 2) probe(float* out, long long* cycles, int passes) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (MODE == 1) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(&tmem_base)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // lane quadrant of this warp, and its 48-column slice (two warps share a quadrant)
    const uint32_t taddr = (MODE == 1 ? tmem_base : 0u) + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 48);
    uint32_t acc[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) acc[i] = __float_as_uint(0.f);
    if (MODE == 1) {
        ST16((acc + 0), taddr); ST16((acc + 16), taddr + 16); ST16((acc + 32), taddr + 32);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    float vr = 1e-3f * threadIdx.x, vi = 2e-3f;
    const long long t0 = clock64();
    for (int p = 0; p < passes; ++p) {
        if (MODE == 1) {
            LD16((acc + 0), taddr); LD16((acc + 16), taddr + 16); LD16((acc + 32), taddr + 32);
        }
        // the "transform" of the pass: independent FFMA chains (8-way ILP)
#pragma unroll
        for (int k = 0; k < WORK / 8; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = fmaf(w[i], 1.0000001f, 1e-7f);
        }
        if (MODE == 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float cr = w[0] * 1e-3f, ci = w[1] * 1e-3f;
#pragma unroll
        for (int j = 0; j < 24; ++j) {   // acc[j] += (vr + i vi) * (cr + i ci)
            float ar = __uint_as_float(acc[2 * j]), ai = __uint_as_float(acc[2 * j + 1]);
            ar = fmaf(vr, cr, fmaf(-vi, ci, ar));
            ai = fmaf(vr, ci, fmaf(vi, cr, ai));
            acc[2 * j] = __float_as_uint(ar);
            acc[2 * j + 1] = __float_as_uint(ai);
            vr += 1e-6f;
        }
        if (MODE == 1) {
            ST16((acc + 0), taddr); ST16((acc + 16), taddr + 16); ST16((acc + 32), taddr + 32);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 48; ++i) s += __uint_as_float(acc[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (MODE == 1) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base));
    }
}

template <int MODE, int WORK>
void run(const char* name, float* out, long long* cyc, int passes) {
    const int grid = 148 * 2;
    probe<MODE, WORK><<<grid, 256>>>(out, cyc, passes);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE, WORK><<<grid, 256>>>(out, cyc, passes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    printf("%-44s : %8.1f cycles per pass (CTA 0 clock64), kernel %.3f ms  (%s)\n", name, mean / passes, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 296 * 256 * sizeof(float));
    cudaMalloc(&cyc, 296 * sizeof(long long));
    const int passes = 2000;
    run<0, 0>("registers, accumulate only", out, cyc, passes);
    run<1, 0>("TMEM (ld 48 + st 48 per pass), accumulate only", out, cyc, passes);
    run<0, 1024>("registers + 1024 independent FFMA per pass", out, cyc, passes);
    run<1, 1024>("TMEM + 1024 independent FFMA per pass", out, cyc, passes);
    run<0, 4096>("registers + 4096 independent FFMA per pass", out, cyc, passes);
    run<1, 4096>("TMEM + 4096 independent FFMA per pass", out, cyc, passes);
    return 0;
}
