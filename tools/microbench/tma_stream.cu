// Microbenchmark: achievable HBM streaming rate of the fast kernel's TMA access pattern on B200.
// Each CTA streams tiles of [M rows x W bytes] (rows R*D*4 bytes apart, as in sml_fast_kernel) from a (B,T,D) fp32
// tensor with cp.async.bulk.tensor.4d, optionally storing each tile back to a second tensor.  No compute.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu
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

template <int PAIRS, int NSTAGE, bool STORE>
__global__ void stream_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int ntd, int ntiles, int R,
                              float* sink) {
    constexpr int M = 1024, W = PAIRS * 8, BOX = 256;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * M * W);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + s)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const int my = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my * R;
    auto issue = [&](int L) {
        if (L >= total) return;
        int it = L / R, r = L % R, tile = blockIdx.x + it * gridDim.x, b = tile / ntd, dt = tile % ntd, s = L % NSTAGE;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + s)), "r"(M * W) : "memory");
        for (int bx = 0; bx < M / BOX; ++bx)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(smem + s * M * W + bx * BOX * W)), "l"((uint64_t)&tin), "r"(s32(bar + s)), "r"(dt * PAIRS * 2), "r"(r), "r"(bx * BOX), "r"(b) : "memory");
    };
    if (tid == 0) for (int s = 0; s < NSTAGE; ++s) issue(s);
    float acc = 0.f;
    for (int L = 0; L < total; ++L) {
        const int s = L % NSTAGE;
        while (!try_wait(bar + s, (L / NSTAGE) & 1)) {}
        acc += reinterpret_cast<float*>(smem + s * M * W)[tid];
        __syncthreads();
        if (tid == 0) {
            if (STORE) {
                int it = L / R, r = L % R, tile = blockIdx.x + it * gridDim.x, b = tile / ntd, dt = tile % ntd;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                for (int bx = 0; bx < M / BOX; ++bx)
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"((uint64_t)&tout), "r"(s32(smem + s * M * W + bx * BOX * W)), "r"(dt * PAIRS * 2), "r"(r), "r"(bx * BOX), "r"(b) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            issue(L + NSTAGE);
        }
    }
    if (tid == 0 && STORE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CUtensorMap make_map(EncFn enc, void* base, int B, int T, int D, int pairs, CUtensorMapL2promotion prom) {
    CUtensorMap m;
    int M = 1024, R = T / M;
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)R, (cuuint64_t)M, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)D * 4, (cuuint64_t)R * D * 4, (cuuint64_t)T * D * 4};
    cuuint32_t box[4] = {(cuuint32_t)(2 * pairs), 1, 256, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

template <int PAIRS, int NSTAGE, bool STORE>
void run(EncFn enc, float* x, float* y, int B, int T, int D, int ctas_per_sm, CUtensorMapL2promotion prom, const char* pname) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    CUtensorMap tin = make_map(enc, x, B, T, D, PAIRS, prom), tout = make_map(enc, y, B, T, D, PAIRS, prom);
    int ntd = D / (2 * PAIRS), ntiles = B * ntd, R = T / 1024;
    size_t smem = (size_t)NSTAGE * 1024 * PAIRS * 8 + 64;
    auto k = stream_kernel<PAIRS, NSTAGE, STORE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = sms * ctas_per_sm; if (grid > ntiles) grid = ntiles;
    float* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, 128, smem>>>(tin, tout, ntd, ntiles, R, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double bytes = (double)B * T * D * 4 * (STORE ? 2 : 1);
    printf("rowB=%3d stages=%d ctas/SM=%d store=%d prom=%s : %.3f ms  %.0f GB/s  (%s)\n", PAIRS * 8, NSTAGE, ctas_per_sm, (int)STORE, pname, best, bytes / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncFn enc = (EncFn)p;
    const int B = 16, T = 8192, D = 768;
    float *x, *y; cudaMalloc(&x, (size_t)B * T * D * 4); cudaMalloc(&y, (size_t)B * T * D * 4);
    cudaMemset(x, 0, (size_t)B * T * D * 4);
    auto P128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; auto P256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B; auto PN = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    // loads only
    run<4, 1, false>(enc, x, y, B, T, D, 3, P128, "128B");
    run<4, 2, false>(enc, x, y, B, T, D, 3, P128, "128B");
    run<4, 2, false>(enc, x, y, B, T, D, 3, PN, "none");
    run<4, 2, false>(enc, x, y, B, T, D, 3, P256, "256B");
    run<4, 2, false>(enc, x, y, B, T, D, 6, P128, "128B");
    run<8, 2, false>(enc, x, y, B, T, D, 1, P128, "128B");
    run<8, 2, false>(enc, x, y, B, T, D, 3, P128, "128B");
    run<16, 1, false>(enc, x, y, B, T, D, 1, P128, "128B");
    run<16, 1, false>(enc, x, y, B, T, D, 3, P128, "128B");
    // load + store (copy)
    run<4, 1, true>(enc, x, y, B, T, D, 3, P128, "128B");
    run<4, 2, true>(enc, x, y, B, T, D, 3, P128, "128B");
    run<8, 1, true>(enc, x, y, B, T, D, 3, P128, "128B");
    run<8, 2, true>(enc, x, y, B, T, D, 3, P128, "128B");
    run<16, 1, true>(enc, x, y, B, T, D, 3, P128, "128B");
    run<16, 1, true>(enc, x, y, B, T, D, 1, P128, "128B");
    return 0;
}
