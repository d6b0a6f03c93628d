// umma_probe.cu -- checks, on a B200, every tcgen05 operand-layout / descriptor convention the tensor-core spectral kernel
// (csrc/sml_tc.cuh) relies on, against a CPU reference.  One test per process invocation (a bad descriptor can fault):
//     umma_probe <test> [variant]
//   1  stage-1 form : A = MN-major SWIZZLE_128B (a TMA-landed [k][128 B] tile, two 64-row MN groups), B = K-major SWIZZLE_128B
//   2  stage-2 form : A = K-major no-swizzle "planes" (row*16 B + kchunk*plane), B = K-major SWIZZLE_128B, k-step inside the atom
//   3  stage-A form : A = K-major no-swizzle planes, B = MN-major SWIZZLE_128B view of a K-major table, N offset inside the 128-B row
//   4  stage-B form : A = K-major SWIZZLE_128B rows of 128 B, B = MN-major SWIZZLE_128B view of a [64][128 B] table
//   5  TMA          : 4-D box {32 d, 2 n, 64 m1} of a bf16 (T, D) tensor with CU_TENSOR_MAP_SWIZZLE_128B -> shared image dump
//   6  accumulate   : fp32 accumulation rounding of 3xTF32-like chains (kind::tf32, 64 accumulating MMAs)
// The host builds the exact shared-memory image (swizzle applied) and the descriptors; the kernel only copies the image,
// issues the MMAs, and reads the accumulator back with tcgen05.ld.32x32b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct ProbeParams {
    int nk;               // MMAs (k-steps)
    int N;                // accumulator columns
    uint32_t idesc;       // instruction descriptor
    int tf32;             // kind::tf32 instead of kind::f16
    uint32_t a_bytes, b_bytes;   // image sizes
    uint64_t adesc[64], bdesc[64];   // descriptors with start address relative to the image base (added in the kernel)
};

__global__ void __launch_bounds__(128) probe_kernel(const unsigned char* a_img, const unsigned char* b_img, float* d_out, ProbeParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sa = smem;
    unsigned char* sb = smem + ((p.a_bytes + 1023u) & ~1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < p.a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (uint32_t i = tid; i < p.b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint64_t abase = (uint64_t)((s32(sa) >> 4) & 0x3FFF), bbase = (uint64_t)((s32(sb) >> 4) & 0x3FFF);
        for (int s = 0; s < p.nk; ++s) {
            const uint64_t ad = p.adesc[s] + abase, bd = p.bdesc[s] + bbase;
            const uint32_t acc = s > 0 ? 1u : 0u;
            if (p.tf32)
                asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n}\n"
                             ::"r"(tm), "l"(ad), "l"(bd), "r"(p.idesc), "r"(acc) : "memory");
            else
                asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n}\n"
                             ::"r"(tm), "l"(ad), "l"(bd), "r"(p.idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    // wait for the MMAs (bounded)
    {
        uint32_t ok = 0;
        for (uint32_t spin = 0; !ok && spin < (1u << 22); ++spin)
            asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.u32 %0,1,0,q;\n}\n" : "=r"(ok) : "r"(s32(&bar)) : "memory");
        if (!ok) { if (tid == 0) printf("probe: MMA completion timed out\n"); __trap(); }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < p.N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) d_out[(size_t)tid * p.N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}

// ---------------------------------------------------------------------------------------------------------------------
static uint16_t f2bf(float f) {   // round to nearest even
    uint32_t u; memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

static uint64_t make_desc(uint32_t start_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((start_bytes >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)(layout_type & 7) << 61;       // 0 none, 2 SW128, 4 SW64, 6 SW32
    return d;
}
static uint32_t make_idesc(int M, int N, int a_mn, int b_mn, int fmt /*1 bf16, 2 tf32*/) {
    uint32_t d = 0;
    d |= 1u << 4;                 // D = f32
    d |= (uint32_t)fmt << 7;      // A format
    d |= (uint32_t)fmt << 10;     // B format
    d |= (uint32_t)a_mn << 15;
    d |= (uint32_t)b_mn << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
static uint32_t sw128(uint32_t byte_off) {   // Swizzle<3,4,3>: 16-byte chunk index (bits 4-6) ^= bits 7-9
    return byte_off ^ (((byte_off >> 7) & 7u) << 4);
}

static int run_mma_test(const char* name, int M, int N, int K, const std::vector<float>& A, const std::vector<float>& B,
                        const std::vector<unsigned char>& aimg, const std::vector<unsigned char>& bimg, ProbeParams& p, double tol) {
    unsigned char *da, *db; float* dd;
    CK(cudaMalloc(&da, aimg.size())); CK(cudaMalloc(&db, bimg.size())); CK(cudaMalloc(&dd, sizeof(float) * 128 * N));
    CK(cudaMemcpy(da, aimg.data(), aimg.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dd, 0xFF, sizeof(float) * 128 * N));
    p.a_bytes = (uint32_t)aimg.size(); p.b_bytes = (uint32_t)bimg.size(); p.N = N;
    const size_t smem = ((aimg.size() + 1023) & ~(size_t)1023) + bimg.size() + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(da, db, dd, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: KERNEL FAILED: %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * N);
    CK(cudaMemcpy(D.data(), dd, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; int bad = 0, firstbad = -1;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
            const double err = fabs((double)D[(size_t)m * N + n] - ref);
            if (!(err <= tol)) { if (firstbad < 0) firstbad = m * N + n; ++bad; }
            if (err > maxerr || err != err) maxerr = err;
            if (fabs(ref) > maxref) maxref = fabs(ref);
        }
    printf("%s: max|err| %.3e (max|ref| %.2f), %d / %d elements outside %.1e%s\n", name, maxerr, maxref, bad, M * N, tol,
           bad ? "  -> MISMATCH" : "  -> OK");
    if (bad) {
        const int m = firstbad / N, n = firstbad % N;
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
        printf("   first mismatch at (m=%d, n=%d): got %.5f want %.5f;  row0: got %.4f %.4f %.4f %.4f\n", m, n, D[firstbad], ref, D[0], D[1], D[2], D[3]);
    }
    return bad != 0;
}

static void fill(std::vector<float>& v, unsigned seed, bool bf) {
    uint32_t s = seed * 2654435761u + 12345u;
    for (auto& x : v) {
        s = s * 1664525u + 1013904223u;
        float f = ((int)((s >> 8) & 0xFFFF) - 32768) / 32768.0f;
        x = bf ? bf2f(f2bf(f)) : f;
    }
}
static void put_bf(std::vector<unsigned char>& img, uint32_t off, float v) {
    if (off + 2 > img.size()) { printf("image overflow at %u\n", off); exit(3); }
    uint16_t h = f2bf(v); memcpy(&img[off], &h, 2);
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void tma_dump_kernel(const __grid_constant__ CUtensorMap tm, unsigned char* out, int n0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(8192) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(s32(smem)), "l"((uint64_t)&tm), "r"(s32(&bar)), "r"(32), "r"(n0), "r"(0) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok && spin < (1u << 22); ++spin)
        asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.u32 %0,1,0,q;\n}\n" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    if (!ok) __trap();
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char** argv) {
    const int test = argc > 1 ? atoi(argv[1]) : 1;
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    ProbeParams p{};
    if (test == 1) {
        // D[128 x 64] = A[128 x 64] * B[64 x 64]^T ; A MN-major SW128: two MN groups of 64 rows, each [64 k][128 B]
        const int M = 128, N = 64, K = 64;
        std::vector<float> A(M * K), B(N * K);
        fill(A, 1, true); fill(B, 2, true);
        std::vector<unsigned char> aimg(2 * 8192, 0), bimg(64 * 128, 0);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) {
                const uint32_t g = m / 64, mm = m % 64;
                put_bf(aimg, g * 8192 + sw128(k * 128 + mm * 2), A[m * K + k]);
            }
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) put_bf(bimg, sw128(n * 128 + k * 2), B[n * K + k]);
        p.nk = 4; p.idesc = make_idesc(M, N, 1, 0, 1);
        for (int s = 0; s < 4; ++s) {
            p.adesc[s] = variant == 0 ? make_desc(s * 2048, 8192, 1024, 2) : make_desc(s * 2048, 1024, 8192, 2);
            p.bdesc[s] = make_desc(s * 32, 16, 1024, 2);
        }
        return run_mma_test(variant == 0 ? "T1 A=MN/SW128 (LBO=group, SBO=kgroup), B=K/SW128" : "T1' (LBO/SBO swapped)", M, N, K, A, B, aimg, bimg, p, 2e-3);
    }
    if (test == 2) {
        // D[128 x 32] += over 2 k-steps of K=16; A planes: tile 1 of a 256-row operand (row*16 + kchunk*PL); B atoms of 64 k
        const int M = 128, N = 32, K = 32, ROWS = 256, PL = ROWS * 16;
        std::vector<float> A(M * K), B(N * K);
        fill(A, 3, true); fill(B, 4, true);
        // two chunk buffers (one per k-step), each 2 planes
        std::vector<unsigned char> aimg(2 * 2 * PL, 0), bimg(4096, 0);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) {
                const int s = k / 16, kk = k % 16;
                put_bf(aimg, s * 2 * PL + (128 + m) * 16 + (kk / 8) * PL + (kk % 8) * 2, A[m * K + k]);
            }
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) put_bf(bimg, sw128(n * 128 + (32 + k) * 2), B[n * K + k]);   // k-steps 2,3 of the atom
        p.nk = 2; p.idesc = make_idesc(M, N, 0, 0, 1);
        for (int s = 0; s < 2; ++s) {
            p.adesc[s] = variant == 0 ? make_desc(s * 2 * PL + 128 * 16, PL, 128, 0) : make_desc(s * 2 * PL + 128 * 16, 128, PL, 0);
            p.bdesc[s] = make_desc((2 + s) * 32, 16, 1024, 2);
        }
        return run_mma_test(variant == 0 ? "T2 A=K/none planes (LBO=plane, SBO=128), B=K/SW128 k-offset" : "T2' (A LBO/SBO swapped)", M, N, K, A, B, aimg, bimg, p, 2e-3);
    }
    if (test == 3) {
        // D[128 x 16] = A[128 x 32] * Bv^T, Bv[n'][k'] = table[k'][16 j + n'] : MN-major SW128 view of a K-major SW128 table
        const int M = 128, N = 16, K = 32, PL = 128 * 16;
        const int j = variant & 3;           // N offset inside the 128-byte row: j * 32 bytes
        const int lbo_mode = variant >> 2;   // 0: LBO = 4096, 1: LBO = 0
        std::vector<float> A(M * K), tab(32 * 64), B(N * K);
        fill(A, 5, true); fill(tab, 6, true);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) B[n * K + k] = tab[k * 64 + 16 * j + n];
        std::vector<unsigned char> aimg(4 * PL, 0), bimg(4096, 0);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) put_bf(aimg, m * 16 + (k / 8) * PL + (k % 8) * 2, A[m * K + k]);
        for (int r = 0; r < 32; ++r)
            for (int c = 0; c < 64; ++c) put_bf(bimg, sw128(r * 128 + c * 2), tab[r * 64 + c]);
        p.nk = 2; p.idesc = make_idesc(M, N, 0, 1, 1);
        for (int s = 0; s < 2; ++s) {
            p.adesc[s] = make_desc(s * 2 * PL, PL, 128, 0);
            p.bdesc[s] = make_desc(s * 2048 + j * 32, lbo_mode ? 0 : 4096, 1024, 2);
        }
        char nm[128]; snprintf(nm, sizeof nm, "T3 B=MN/SW128 view, N offset %d B inside the row, LBO=%d", j * 32, lbo_mode ? 0 : 4096);
        return run_mma_test(nm, M, N, K, A, B, aimg, bimg, p, 2e-3);
    }
    if (test == 4) {
        // D[128 x 64] = A[128 x 64] * Bv^T, A K-major SW128 rows of 128 B; Bv[n'][k'] = table[k'][n'] (table = [64][128 B] K-major SW128)
        const int M = 128, N = 64, K = 64;
        std::vector<float> A(M * K), tab(64 * 64), B(N * K);
        fill(A, 7, true); fill(tab, 8, true);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) B[n * K + k] = tab[k * 64 + n];
        std::vector<unsigned char> aimg(128 * 128, 0), bimg(64 * 128, 0);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) put_bf(aimg, sw128(m * 128 + k * 2), A[m * K + k]);
        for (int r = 0; r < 64; ++r)
            for (int c = 0; c < 64; ++c) put_bf(bimg, sw128(r * 128 + c * 2), tab[r * 64 + c]);
        p.nk = 4; p.idesc = make_idesc(M, N, 0, 1, 1);
        for (int s = 0; s < 4; ++s) {
            p.adesc[s] = make_desc(s * 32, 16, 1024, 2);
            p.bdesc[s] = make_desc(s * 2048, variant ? 0 : 8192, 1024, 2);
        }
        return run_mma_test("T4 A=K/SW128 rows, B=MN/SW128 view of a [64][128B] table", M, N, K, A, B, aimg, bimg, p, 2e-3);
    }
    if (test == 5) {
        // TMA: bf16 tensor (T = 64*16 rows, D = 64), view {d, n, m1} with t = 16*m1 + n; box {32, 2, 64}, SWIZZLE_128B
        const int N2 = 16, T = 64 * N2, D = 64;
        std::vector<uint16_t> x((size_t)T * D);
        for (int t = 0; t < T; ++t) for (int d = 0; d < D; ++d) x[(size_t)t * D + d] = (uint16_t)(t * 64 + d);   // unique tags
        uint16_t* dx; unsigned char* dout;
        CK(cudaMalloc(&dx, x.size() * 2)); CK(cudaMalloc(&dout, 8192));
        CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
        EncFn enc = (EncFn)fp;
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N2, 64};
        cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)N2 * D * 2};
        cuuint32_t box[3] = {32, 2, 64}, es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("T5: encode failed %d\n", (int)r); return 1; }
        const int n0 = 6;
        CK(cudaFuncSetAttribute(tma_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 1024));
        tma_dump_kernel<<<1, 128, 8192 + 1024>>>(tm, dout, n0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T5: KERNEL FAILED: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<unsigned char> img(8192);
        CK(cudaMemcpy(img.data(), dout, 8192, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m1 = 0; m1 < 64; ++m1) for (int nl = 0; nl < 2; ++nl) for (int d = 0; d < 32; ++d) {
            const uint32_t off = sw128(m1 * 128 + nl * 64 + d * 2);
            uint16_t got; memcpy(&got, &img[off], 2);
            const int t = N2 * m1 + n0 + nl;
            const uint16_t want = (uint16_t)(t * 64 + 32 + d);
            if (got != want) { if (bad < 4) printf("T5: (m1=%d n=%d d=%d) at %u: got %u want %u\n", m1, nl, d, off, got, want); ++bad; }
        }
        printf("T5 TMA box {32 d, 2 n, 64 m1} SWIZZLE_128B -> [m1][n][d] image with chunk ^= (row & 7): %d mismatches%s\n", bad, bad ? "  -> MISMATCH" : "  -> OK");
        return bad != 0;
    }
    if (test == 6) {
        // fp32 accumulation in the tensor core: D[128 x 16] over 64 accumulating tf32 MMAs (K = 8 each); values exactly representable in
        // tf32, so the only error is the accumulator's rounding.  Compared with a double-precision sum.
        // (data for K = 256 is held in shared memory and swept twice: 64 MMAs)
        const int M = 128, N = 16, K = 256;
        std::vector<float> A(M * K), B(N * K);
        fill(A, 9, false); fill(B, 10, false);
        for (auto& v : A) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); }
        for (auto& v : B) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); }
        // A: K-major no-swizzle planes of 4 tf32 (16 B): row*16 + kchunk*PL ; B the same
        const int PLA = 128 * 16, PLB = 16 * 16;
        std::vector<unsigned char> aimg((size_t)(K / 4) * PLA, 0), bimg((size_t)(K / 4) * PLB, 0);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) memcpy(&aimg[(size_t)(k / 4) * PLA + m * 16 + (k % 4) * 4], &A[m * K + k], 4);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) memcpy(&bimg[(size_t)(k / 4) * PLB + n * 16 + (k % 4) * 4], &B[n * K + k], 4);
        p.nk = 64; p.tf32 = 1; p.idesc = make_idesc(M, N, 0, 0, 2);
        for (int s = 0; s < 64; ++s) {
            p.adesc[s] = make_desc((s % 32) * 2 * PLA, PLA, 128, 0);
            p.bdesc[s] = make_desc((s % 32) * 2 * PLB, PLB, 128, 0);
        }
        // report relative rms error and the mean signed relative error (a systematic shrink shows round-toward-zero accumulation)
        unsigned char *da, *db; float* dd;
        CK(cudaMalloc(&da, aimg.size())); CK(cudaMalloc(&db, bimg.size())); CK(cudaMalloc(&dd, sizeof(float) * 128 * N));
        CK(cudaMemcpy(da, aimg.data(), aimg.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
        p.a_bytes = (uint32_t)aimg.size(); p.b_bytes = (uint32_t)bimg.size(); p.N = N;
        const size_t smem = ((aimg.size() + 1023) & ~(size_t)1023) + bimg.size() + 1024;
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe_kernel<<<1, 128, smem>>>(da, db, dd, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T6: KERNEL FAILED: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> D(128 * N);
        CK(cudaMemcpy(D.data(), dd, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
        double se = 0, sr = 0, shrink = 0; int cnt = 0;
        double se32 = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; float f32 = 0.f;
            for (int rep = 0; rep < 2; ++rep)
                for (int k = 0; k < K; ++k) { ref += (double)A[m * K + k] * (double)B[n * K + k]; f32 = fmaf(A[m * K + k], B[n * K + k], f32); }
            const double err = (double)D[m * N + n] - ref;
            se += err * err; sr += ref * ref; se32 += ((double)f32 - ref) * ((double)f32 - ref);
            if (fabs(ref) > 1.0) { shrink += err / ref; ++cnt; }
        }
        printf("T6 tf32 x 64 accumulating MMAs (K=512): rel-L2 error %.3e (sequential fp32 FMA chain: %.3e); mean signed relative error %.3e over %d elements\n",
               sqrt(se / sr), sqrt(se32 / sr), cnt ? shrink / cnt : 0.0, cnt);
        return 0;
    }
    printf("unknown test %d\n", test);
    return 1;
}
