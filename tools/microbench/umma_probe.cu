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

__global__ void tma_dump_kernel(const __grid_constant__ CUtensorMap tm, unsigned char* out, int c0, int c1, int c2, int bytes) {
    extern __shared__ __align__(1024) unsigned char smem[];   // 32 KB data area (pre-filled with 0xFF) + barrier; no static shared memory
    uint64_t& bar = *reinterpret_cast<uint64_t*>(smem + 32768);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xFFFFFFFFu;
    if (threadIdx.x == 0 && (s32(smem) & 1023u)) printf("T5: dynamic shared base 0x%x is not 1024-byte aligned\n", s32(smem));
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(s32(smem)), "l"((uint64_t)&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok && spin < (1u << 22); ++spin)
        asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.u32 %0,1,0,q;\n}\n" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    if (!ok) { if (threadIdx.x == 0) printf("T5: TMA completion timed out\n"); }
    __syncthreads();
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) out[i] = smem[i];
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
    if (test == 7) {
        // D[128 x 64] = A[128 x 64] * B[64 x 64]^T ; A MN-major SWIZZLE_64B: four MN groups of 32 rows (one per n), each [64 k][64 B]
        const int M = 128, N = 64, K = 64;
        std::vector<float> A(M * K), B(N * K);
        fill(A, 11, true); fill(B, 12, true);
        std::vector<unsigned char> aimg(4 * 4096, 0), bimg(64 * 128, 0);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) {
                const uint32_t g = m / 32, mm = m % 32;
                const uint32_t lin = k * 64 + mm * 2;
                put_bf(aimg, g * 4096 + (lin ^ (((lin >> 7) & 3u) << 4)), A[m * K + k]);
            }
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) put_bf(bimg, sw128(n * 128 + k * 2), B[n * K + k]);
        p.nk = 4; p.idesc = make_idesc(M, N, 1, 0, 1);
        for (int s = 0; s < 4; ++s) {
            p.adesc[s] = variant == 0 ? make_desc(s * 1024, 4096, 512, 4) : make_desc(s * 1024, 512, 4096, 4);
            p.bdesc[s] = make_desc(s * 32, 16, 1024, 2);
        }
        return run_mma_test(variant == 0 ? "T7 A=MN/SW64 (LBO=group 4096, SBO=kgroup 512), B=K/SW128" : "T7' (LBO/SBO swapped)", M, N, K, A, B, aimg, bimg, p, 2e-3);
    }
    if (test == 5) {
        // TMA landing layout: bf16 tensor (T = 64*16 rows, D = 64), t = 16*m1 + n.  The whole 32 KB buffer is dumped and every
        // element is searched for, so the landing rule is measured instead of assumed.
        //   variant 0: dims {d, n, m1}, box {32, 2, 64}, SWIZZLE_128B (inner 64 B < span)     1: same, SWIZZLE_NONE
        //   variant 2: dims {d, m1, n}, box {32, 64, 4}, SWIZZLE_64B  (inner 64 B = span)     3: dims {d, n, m1}, box {64, 1, 64}, SWIZZLE_128B (inner 128 B)
        const int N2 = 16, T = 64 * N2, D = 64;
        std::vector<uint16_t> x((size_t)T * D);
        for (int t = 0; t < T; ++t) for (int d = 0; d < D; ++d) x[(size_t)t * D + d] = (uint16_t)(t * 64 + d);   // unique tags (< 0xFFFF)
        uint16_t* dx; unsigned char* dout;
        CK(cudaMalloc(&dx, x.size() * 2)); CK(cudaMalloc(&dout, 32768));
        CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
        EncFn enc = (EncFn)fp;
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N2, 64};
        cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)N2 * D * 2};
        cuuint32_t box[3] = {32, 2, 64}, es[3] = {1, 1, 1};
        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
        int c1 = 6, c2 = 0, d0 = 32;
        if (variant == 1) swz = CU_TENSOR_MAP_SWIZZLE_NONE;
        if (variant == 2) {
            dims[1] = 64; dims[2] = N2; strides[0] = (cuuint64_t)N2 * D * 2; strides[1] = (cuuint64_t)D * 2;
            box[1] = 64; box[2] = 4; swz = CU_TENSOR_MAP_SWIZZLE_64B; c1 = 0; c2 = 4;
        }
        if (variant == 3) { box[0] = 64; box[1] = 1; d0 = 0; }
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("T5: encode failed %d\n", (int)r); return 1; }
        const int bytes = (int)(box[0] * box[1] * box[2] * 2);
        CK(cudaFuncSetAttribute(tma_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
        tma_dump_kernel<<<1, 128, 32768 + 1024>>>(tm, dout, d0, c1, c2, bytes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T5 variant %d: KERNEL FAILED: %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<unsigned char> img(32768);
        CK(cudaMemcpy(img.data(), dout, 32768, cudaMemcpyDeviceToHost));
        // where did each tag land?
        std::vector<int> where(65536, -1);
        int landed = 0, maxoff = 0;
        for (int off = 0; off < 32768; off += 2) {
            uint16_t v; memcpy(&v, &img[off], 2);
            if (v != 0xFFFF) { where[v] = off; ++landed; if (off > maxoff) maxoff = off; }
        }
        printf("T5 variant %d: %d elements landed (expected %d), highest byte offset %d\n", variant, landed, bytes / 2, maxoff);
        // hypotheses for the byte offset of (i2, i1, d) = box coordinates (outer, middle, inner)
        const int nb1 = box[1], nb2 = box[2], nb0 = box[0];
        int bad_lin = 0, bad_sw128 = 0, bad_sw64 = 0, bad_pad128 = 0, shown = 0;
        for (int i2 = 0; i2 < nb2; ++i2) for (int i1 = 0; i1 < nb1; ++i1) for (int d = 0; d < nb0; ++d) {
            int t, dd;
            if (variant == 2) { t = N2 * i1 + (c2 + i2); dd = 32 + d; }          // {d, m1, n}: i1 = m1, i2 = n
            else { t = N2 * i2 + (c1 + i1); dd = (variant == 3 ? 0 : 32) + d; }   // {d, n, m1}: i1 = n, i2 = m1
            const int tag = t * 64 + dd;
            const int got = where[tag];
            const uint32_t lin = (uint32_t)((i2 * nb1 + i1) * nb0 + d) * 2;
            const uint32_t sw64 = lin ^ (((lin >> 7) & 3u) << 4);
            const uint32_t padlin = (uint32_t)(i2 * nb1 + i1) * 128 + d * 2;
            if (got != (int)lin) ++bad_lin;
            if (got != (int)sw128(lin)) ++bad_sw128;
            if (got != (int)sw64) ++bad_sw64;
            if (got != (int)sw128(padlin)) ++bad_pad128;
            if (shown < 12 && (d % 8) == 0 && i2 < 2) { printf("   (i2=%d i1=%d d=%d) lin %u -> landed at %d\n", i2, i1, d, lin, got); ++shown; }
        }
        printf("   mismatches vs: linear %d | Swizzle<3,4,3>(linear) %d | Swizzle<2,4,3>(linear) %d | Swizzle<3,4,3>(rows padded to 128 B) %d\n",
               bad_lin, bad_sw128, bad_sw64, bad_pad128);
        return 0;
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
