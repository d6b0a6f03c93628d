// Microbenchmark: which data path streams the fast kernel's access pattern fastest on B200?
// Pattern (sml_fast_kernel, P = 4 pairs): a CTA of 128 threads copies tiles of [1024 rows x ROWB bytes]; consecutive rows of a tile
// are R*D*esz bytes apart in a (B, T, D) tensor (t = R*m + r).  Per pass: land the tile in shared memory, every thread pulls its
// 32 elements (rows 32*m1 + m2, channel pair p) into registers, then the rows go back out to a second tensor.
//   load  path 0 = TMA (cp.async.bulk.tensor.4d + mbarrier)       1 = LSU (cp.async 16 B per thread + mbarrier arrive)
//   store path 0 = shared staging + TMA bulk store                 1 = LSU (st.global straight from registers)
// No compute.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldst_stream ldst_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
    return ok;
}

// ROWB: bytes per row (32 = fp32 P=4, 16 = bf16 P=4).  element per thread per row: ROWB/4 bytes
template <int ROWB, int LD, int ST>
__global__ void __launch_bounds__(128, 3)
    stream_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, const char* __restrict__ x,
                  char* __restrict__ y, int ntd, int ntiles, int R, long long row_stride, long long batch_stride, long long pass_stride) {
    constexpr int M = 1024, BOX = 256, EB = ROWB / 4;   // EB bytes per thread per row (8 or 4)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* X = smem;                    // landing
    unsigned char* S = smem + M * ROWB;         // staging (TMA store path only)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * M * ROWB);
    const int tid = threadIdx.x, tp = tid % 4, tm2 = tid / 4;
    uint64_t pol = 0;
    if (LD == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"((LD == 0 || LD == 4) ? 1 : 128));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const int my = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my * R;
    auto coords = [&](int L, int& b, int& dt, int& r) {
        int it = L / R;
        r = L % R;
        int tile = blockIdx.x + it * gridDim.x;
        b = tile / ntd;
        dt = tile % ntd;
    };
    auto issue = [&](int L) {
        if (L >= total) return;
        int b, dt, r;
        coords(L, b, dt, r);
        if (LD == 0 || LD == 4) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(M * ROWB) : "memory");
                for (int bx = 0; bx < M / BOX; ++bx) {
                    if (LD == 4)
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
                                     ::"r"(s32(X + bx * BOX * ROWB)), "l"((uint64_t)&tin), "r"(s32(bar)), "r"(dt * 8), "r"(r), "r"(bx * BOX), "r"(b), "l"(pol) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                     ::"r"(s32(X + bx * BOX * ROWB)), "l"((uint64_t)&tin), "r"(s32(bar)), "r"(dt * 8), "r"(r), "r"(bx * BOX), "r"(b) : "memory");
                }
            }
        } else {
            // 16-byte chunks: chunk c -> row c / (ROWB/16), part c % (ROWB/16); 128 threads, consecutive threads take consecutive chunks
            const char* src = x + (long long)b * batch_stride + (long long)r * pass_stride + (long long)dt * ROWB;
            constexpr int CPR = ROWB / 16, NCH = M * CPR;
#pragma unroll 4
            for (int c = tid; c < NCH; c += 128) {
                const int row = c / CPR, part = c % CPR;
                if (LD == 2)
                    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(s32(X + row * ROWB + part * 16)), "l"(src + (long long)row * row_stride + part * 16) : "memory");
                else if (LD == 3)
                    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(s32(X + row * ROWB + part * 16)), "l"(src + (long long)row * row_stride + part * 16) : "memory");
                else
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(X + row * ROWB + part * 16)), "l"(src + (long long)row * row_stride + part * 16) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(bar)) : "memory");
        }
    };
    issue(0);
    float sink = 0.f;
    for (int L = 0; L < total; ++L) {
        while (!try_wait(bar, L & 1)) {}
        uint32_t v[32][EB / 4];
#pragma unroll
        for (int m1 = 0; m1 < 32; ++m1) {
            const unsigned char* p = X + (32 * m1 + tm2) * ROWB + tp * EB;
            if (EB == 8) { uint2 q = *reinterpret_cast<const uint2*>(p); v[m1][0] = q.x; v[m1][EB / 4 - 1] = q.y; }
            else v[m1][0] = *reinterpret_cast<const uint32_t*>(p);
        }
        __syncthreads();   // X drained by everyone
        if (ST == 0 && tid == 0 && L > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        issue(L + 1);
        int b, dt, r;
        coords(L, b, dt, r);
        if (ST == 0) {
            __syncthreads();   // staging free (thread 0 waited for the previous store)
#pragma unroll
            for (int m1 = 0; m1 < 32; ++m1) {
                unsigned char* p = S + (32 * m1 + tm2) * ROWB + tp * EB;
                if (EB == 8) *reinterpret_cast<uint2*>(p) = make_uint2(v[m1][0], v[m1][EB / 4 - 1]);
                else *reinterpret_cast<uint32_t*>(p) = v[m1][0];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                for (int bx = 0; bx < M / BOX; ++bx) {
                    if (LD == 4)
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
                                     ::"l"((uint64_t)&tout), "r"(s32(S + bx * BOX * ROWB)), "r"(dt * 8), "r"(r), "r"(bx * BOX), "r"(b), "l"(pol) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"((uint64_t)&tout), "r"(s32(S + bx * BOX * ROWB)), "r"(dt * 8), "r"(r), "r"(bx * BOX), "r"(b) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            char* dst = y + (long long)b * batch_stride + (long long)r * pass_stride + (long long)dt * ROWB + tp * EB;
#pragma unroll
            for (int m1 = 0; m1 < 32; ++m1) {
                char* p = dst + (long long)(32 * m1 + tm2) * row_stride;
                if (EB == 8) *reinterpret_cast<uint2*>(p) = make_uint2(v[m1][0], v[m1][EB / 4 - 1]);
                else *reinterpret_cast<uint32_t*>(p) = v[m1][0];
            }
        }
    }
    if (ST == 0 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (sink == 123.f) y[0] = 1;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// element type: 4-byte "float" units for ROWB=32 (8 per row), 2-byte units for ROWB=16 (8 per row): always 8 elements per row
CUtensorMap make_map(EncFn enc, void* base, int B, int T, int D, int esz) {
    CUtensorMap m;
    int M = 1024, R = T / M;
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)R, (cuuint64_t)M, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)D * esz, (cuuint64_t)R * D * esz, (cuuint64_t)T * D * esz};
    cuuint32_t box[4] = {8, 1, 256, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

template <int ROWB, int LD, int ST>
void run(EncFn enc, char* x, char* y, int B, int T, int D) {
    const int esz = ROWB / 8;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    CUtensorMap tin = make_map(enc, x, B, T, D, esz), tout = make_map(enc, y, B, T, D, esz);
    int ntd = D / 8, ntiles = B * ntd, R = T / 1024;
    size_t smem = (size_t)2 * 1024 * ROWB + 64;
    auto k = stream_kernel<ROWB, LD, ST>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = sms * 3; if (grid > ntiles) grid = ntiles;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    cudaMemset(y, 0, (size_t)B * T * D * esz);
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, 128, smem>>>(tin, tout, x, y, ntd, ntiles, R, (long long)R * D * esz, (long long)T * D * esz, (long long)D * esz);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // verify the copy
    size_t n = (size_t)B * T * D * esz;
    unsigned char* hx = (unsigned char*)malloc(n); unsigned char* hy = (unsigned char*)malloc(n);
    cudaMemcpy(hx, x, n, cudaMemcpyDeviceToHost); cudaMemcpy(hy, y, n, cudaMemcpyDeviceToHost);
    size_t bad = 0; for (size_t i = 0; i < n; ++i) bad += hx[i] != hy[i];
    free(hx); free(hy);
    double bytes = (double)n * 2;
    printf("rowB=%2d load=%s store=%s : %.3f ms  %.0f GB/s (read+write)  mismatches=%zu  (%s)\n", ROWB, LD == 0 ? "TMA" : LD == 4 ? "TMA(evict_first ld+st)" : LD == 1 ? "LSU(cp.async16)" : LD == 2 ? "LSU(cp.async16.L2::128B)" : "LSU(cp.async16.L2::256B)", ST ? "LSU(st.global)" : "TMA",
           best, bytes / best / 1e6, bad, cudaGetErrorString(cudaGetLastError()));
}

__global__ void fill(uint32_t* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)(i * 2654435761u);
}

int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncFn enc = (EncFn)p;
    const int B = 16, T = 8192, D = 768;
    char *x, *y; cudaMalloc(&x, (size_t)B * T * D * 4); cudaMalloc(&y, (size_t)B * T * D * 4);
    fill<<<1024, 256>>>((uint32_t*)x, (size_t)B * T * D);
    run<32, 0, 0>(enc, x, y, B, T, D);
    run<32, 4, 0>(enc, x, y, B, T, D);
    run<32, 1, 0>(enc, x, y, B, T, D);
    run<32, 0, 1>(enc, x, y, B, T, D);
    run<32, 1, 1>(enc, x, y, B, T, D);
    run<32, 2, 0>(enc, x, y, B, T, D);
    run<32, 3, 0>(enc, x, y, B, T, D);
    run<32, 2, 1>(enc, x, y, B, T, D);
    run<16, 0, 0>(enc, x, y, B, T, D);
    run<16, 4, 0>(enc, x, y, B, T, D);
    run<16, 0, 1>(enc, x, y, B, T, D);
    run<16, 2, 0>(enc, x, y, B, T, D);
    return 0;
}
