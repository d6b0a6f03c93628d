// sml_tc.cuh -- tensor-core ("DFT as GEMM", tcgen05 + TMEM) SpectralMixingLayer forward / backward for bf16 I/O on sm_100a.
//
// What one launch computes is identical to sml_fast_kernel (reference: /root/reference/fft_tensor/spectral_layers.py:88-116
// and its autograd graph == /root/reference/fft_tensor/wirtinger_ops.py:53-82); what differs is where the arithmetic runs.
// With bf16 I/O the CUDA-core butterflies of sml_fast.cuh are compute-bound (same time as fp32 at half the bytes), so all
// four DFT stages of the band-limited transform run on the 5th-generation tensor cores here and the CUDA cores only move
// data between them (twiddle, transpose, filter).  T = 64 * N2, t = N2*m1 + n, f = f1 + 64*f2:
//
//   analysis   stage 1  S[(n,d), slot]   = sum_m1 x[m1,(n,d)] * B1[slot, m1]          DFT-64 over m1 of REAL columns; the 64 slots are
//                                                                                     Re S_f1 (f1 = 0..32) and Im S_f1 (f1 = 1..31)
//              twiddle  V[n,f1]          = S[n,f1] * W_T^{n f1}                        CUDA cores, TMEM -> registers -> shared (transposed)
//              stage 2  Z[(d,f1), q]    += sum_n V[(d,f1), n] * W_N2^{n f2}            f2 = q - 8 two-sided, accumulated in TMEM over all n
//   mid phase  each thread owns one row (channel d, class f1): X_f = Z (f2 >= 0) or conj Z (f2 < 0: the bin 64|f2| - f1), filter,
//              X_low / Wirtinger gradient terms, Hermitian-extended band with the 1/T, 1/2 and bias rules folded in
//   synthesis  stage A  Y[(d,f1), n]     = sum_q band[(d,f1), q] * W_N2^{-n f2}
//              twiddle  V'[n,f1]         = Y * W_T^{-n f1}                              CUDA cores, transposed back
//              stage B  y[(n,d), m1]     = sum_slot V'[(n,d), slot] * B1[slot, m1]      the transpose of stage 1
//
// Operands: x lands by TMA (one box {32 d, 64 m1, 4 n} per tile, SWIZZLE_64B) and IS the MN-major A operand of stage 1; B1 (64 x 64) and B2 (32 x 2 N2) are
// bf16 tables resident in shared memory, used K-major by the analysis and through their MN-major (transposed) view by the
// synthesis; everything accumulates in fp32 in TMEM (512 columns: two stage-1 / stage-B tiles + the 9 x 32-column band
// accumulator, which synthesis re-uses for the stage-A output).  The descriptor conventions are the ones verified on
// hardware by tools/microbench/umma_probe.cu (tests 2-4, 7; test 5 measured the TMA landing rule -- SWIZZLE_128B with a
// 64-byte inner box pads every row to 128 bytes, which is why the x tile uses SWIZZLE_64B).
//
// One CTA per SM, 20 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator, warps 4-19 =
// four compute warpgroups (each warp reads its own 32-lane TMEM quadrant).  Work item = (batch element, 32 channels):
// 64-byte global rows.  Precision: bf16 operands, fp32 accumulation: ~3e-3 relative L2 (gate 1e-2 for bf16 I/O).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sml_fast.cuh"

namespace sml {

struct TcParams {
    const float* w_re;     // (D,F)
    const float* w_im;     // (D,F)
    const float* bias;     // (D,) or null (FWD)
    cf* xlow;              // (B,D,k) complex64: written by FWD (nullable) / read by BWD (nullable)
    float* gw_re;          // BWD: non-null = filter gradients wanted
    cf* gpart;             // (B,D,k) per-batch-element filter-gradient terms (BWD)
    float* gbpart;         // (B,D)   per-batch-element bias-gradient terms (BWD)
    const unsigned char* b1_img;   // 8 KB   bf16 [64 slots][64 m1], SWIZZLE_128B image
    const unsigned char* b2_img;   // NA * 4 KB  bf16 [32 rows][2 N2], atoms of 64 columns, SWIZZLE_128B image
    const float2* tw;      // [N2][32]: j >= 1: W_T^{n j};  j = 0: W_T^{32 n}
    int B, T, D, F, k;
    int N2;                // T / 64
    int ntd;               // channel tiles = D / 32
    int nitems;            // B * ntd
    int consumer_fence;    // see the kernel
    int nslot;             // x-tile landing slots (4..8): as many as shared memory allows, the analysis phase is TMA-latency bound
    void* out;             // y (FWD) / gx (BWD): (B,T,D) bf16, written with plain global stores
    float invT;
    unsigned int* dbg;
    float* dump;           // bring-up aid (SML_TC_DUMP): intermediates of work item 0, see tools/tc_dump_check.py; null in production
};

namespace tc {

constexpr int THREADS = 640;
constexpr int RPD = 34;                  // band rows per channel: class 0, class 32, classes 1..31, one pad row
constexpr int ROWS = 32 * RPD;           // 1088 live rows
constexpr int NTILE = 9;                 // 128-row tiles of the band operand
constexpr uint32_t PLANE = 1152u * 16u;  // one 8-element k-chunk of all 9 tiles: row * 16 B

// shared-memory map (bytes; the sizes depend on N2 and on the number of x-tile slots, so the offsets are computed at run time).
// Synthesis re-uses the analysis buffers: band operand <-> A2 chunks, stage-B operand (2 x 32 KB) <-> x tiles.
//   B1 8 KB | B2 NA*4 KB | x tiles nslot*16 KB | A2 chunks 2 x 36 KB | twiddle rows nslot*1 KB | synthesis twiddle rows 2 x 2 KB | barriers
constexpr int MAX_SLOT = 8;
constexpr uint32_t OFF_B1 = 0;
constexpr uint32_t OFF_B2 = 8192;
struct SmemMap {
    uint32_t x, a2, tw, tws, bar, total;
};
__host__ __device__ inline SmemMap smem_map(int N2, int nslot) {
    SmemMap m;
    m.x = OFF_B2 + (uint32_t)((N2 + 31) / 32) * 4096u;
    m.a2 = m.x + (uint32_t)nslot * 16384u;
    m.tw = m.a2 + 4u * PLANE;
    m.tws = m.tw + (uint32_t)nslot * 1024u;
    m.bar = m.tws + 4096u;
    m.total = m.bar + 1024u;
    return m;
}

struct Bars {
    uint64_t x_full[MAX_SLOT], x_free[MAX_SLOT], d1_full[2], d1_free[2], a2_full[2], a2_free[2], d2_full, aband_full;
    uint64_t tws_full[2], tws_free[2], da_full[2], da_free[2], ab_full[2], ab_free[2], db_full[2], db_free[2];
    uint64_t xregion_free;   // once per work item: the last stage-B MMAs have read the x-tile region (it doubles as their operand)
    uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 1024, "barrier block");

// TMEM columns
constexpr uint32_t COL_D1 = 0;      // 2 x 64 (stage 1) | 2 x 64 (stage B)
constexpr uint32_t COL_D2 = 128;    // 9 x 32 (stage 2) | 2 x 9 x 16 (stage A)

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
constexpr uint32_t LAYOUT_NONE = 0, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);   // D = f32, A = B = bf16
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void stg_bf16(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// mbarrier wait for the tensor-core kernel: the try_wait carries a suspend-time hint, so a waiting warp sleeps in hardware
// instead of burning issue slots in a poll loop (the compute warps share their schedulers with the warps that are working);
// bounded like mbar_wait (10 s of %globaltimer), a lost arrival traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_tc(uint64_t* bar, uint32_t parity, unsigned int* dbg, uint32_t tag, uint32_t aux) {
    uint64_t t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
        if ((spins & 0xFFu) == 0u) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) mbar_timeout(dbg, tag, parity, aux);
        }
    }
}
// warp-level arrive: every lane has finished its part (and fenced it) before lane 0 signals
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// ------------------------------------------------------------------------------------------------
// mid phase for one band row (channel d, class f1 = 0..32): in = the 16 two-sided accumulator slots Z[q], f2 = q - 8;
// out = the 16 band slots of the synthesis operand as packed bf16 (re, im).
//   slot q >= 8 holds the bin f = f1 + 64 (q - 8):                X_f = Z[q]
//   slot q <  8 holds the NEGATIVE frequency f1 + 64 (q - 8):     X_f = conj Z[q] for the bin f = 64 (8 - q) - f1
// Classes 0 and 32 are self-conjugate (every bin shows up on both sides, half weight each); the DC bin carries the bias.
// All filter / X_low loads of eight slots are issued before the first use, so their latencies overlap.
// ------------------------------------------------------------------------------------------------
template <bool BWD>
__device__ __forceinline__ void mid_row(const uint32_t (&z)[32], uint32_t (&outw)[16], const TcParams& prm, int b, int d, int f1) {
    const int k = prm.k;
    const float invT = prm.invT;
    const bool selfconj = f1 == 0 || f1 == 32;
    const float scale = selfconj ? 0.5f * invT : invT;
    const bool grads = BWD && prm.gw_re != nullptr;
    const float* const wre = prm.w_re + (size_t)d * prm.F;
    const float* const wim = prm.w_im + (size_t)d * prm.F;
    const size_t xrow = ((size_t)b * prm.D + d) * (size_t)k;
    float2* const xl = reinterpret_cast<float2*>(prm.xlow) + xrow;
    float2* const gp = reinterpret_cast<float2*>(prm.gpart) + xrow;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float wr[8], wi[8];
        float2 xs[8];
        int fq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int q = 8 * half + j;
            const int f = q >= 8 ? f1 + 64 * (q - 8) : 64 * (8 - q) - f1;
            fq[j] = f < k ? f : -1;
            wr[j] = 0.f; wi[j] = 0.f; xs[j] = make_float2(0.f, 0.f);
            if (fq[j] >= 0) {
                wr[j] = __ldg(wre + f);
                wi[j] = __ldg(wim + f);
                if (grads) xs[j] = __ldg(xl + f);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int q = 8 * half + j;
            const bool neg = q < 8;
            const float zr = __uint_as_float(z[2 * q]), zi = __uint_as_float(z[2 * q + 1]);
            const cf X = cf{zr, neg ? -zi : zi};
            uint32_t o = 0u;
            if (fq[j] >= 0) {
                const bool owner = !(selfconj && neg);   // a self-conjugate class sees each bin twice: the positive side writes it
                cf a;
                if constexpr (!BWD) {
                    if (owner && prm.xlow != nullptr) xl[fq[j]] = make_float2(X.re, X.im);
                    a = cmul(X, cf{wr[j], wi[j]});
                } else {
                    if (grads && owner) {
                        const cf gt = cmulc(X, cf{xs[j].x, xs[j].y});   // G conj(X_low), wirtinger_ops.py:77
                        gp[fq[j]] = make_float2(gt.re * invT, gt.im * invT);
                        if (fq[j] == 0) prm.gbpart[(size_t)b * prm.D + d] = X.re;
                    }
                    a = cmulc(X, cf{wr[j], wi[j]});
                }
                float ore = a.re * scale, oim = (neg ? -a.im : a.im) * scale;
                if (f1 == 0 && q == 8) {          // DC: full weight, real, + bias (a constant in time)
                    ore = a.re * invT;
                    oim = 0.f;
                    if constexpr (!BWD) {
                        if (prm.bias != nullptr) ore += __ldg(prm.bias + d);
                    }
                }
                o = pack_bf16(ore, oim);
            }
            outw[q] = o;
        }
    }
}

}   // namespace tc

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <bool BWD, int MODE>   // MODE 0 = production, 1 = intermediate dump (bring-up), 2 = phase timing (SML_DEBUG)
__global__ void __launch_bounds__(tc::THREADS, 1)
    sml_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const TcParams prm) {
    using namespace tc;
    constexpr bool DUMP = MODE == 1;
    constexpr bool TIMING = MODE == 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int N2 = prm.N2;
    const int nslot = prm.nslot;
    const SmemMap sm = smem_map(N2, nslot);
    Bars* const bars = reinterpret_cast<Bars*>(smem + sm.bar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Where the generic-proxy writes of the compute warps (operand tiles in shared memory) are fenced against the tensor core's
    // async-proxy reads: by every writer before it arrives (0), or once by the MMA thread after it has acquired the barrier (1).
    const bool consumer_fence = prm.consumer_fence != 0;
    const int NT1 = N2 >> 2;   // stage-1 / stage-B tiles (4 n each) per work item: even
    const int NCH = N2 >> 3;   // stage-2 / stage-A chunks (8 n each)
    unsigned int* const dbg = prm.dbg;

    // ---- one-time setup: barriers, TMEM, resident tables ----
    if (tid == 0) {
        for (int i = 0; i < MAX_SLOT; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_free[i], 9); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->d1_full[i], 1);  mbar_init(&bars->d1_free[i], 8);
            mbar_init(&bars->a2_full[i], 16); mbar_init(&bars->a2_free[i], 1);
            mbar_init(&bars->tws_full[i], 1); mbar_init(&bars->tws_free[i], 16);
            mbar_init(&bars->da_full[i], 1);  mbar_init(&bars->da_free[i], 16);
            mbar_init(&bars->ab_full[i], 16); mbar_init(&bars->ab_free[i], 1);
            mbar_init(&bars->db_full[i], 1);  mbar_init(&bars->db_free[i], 8);
        }
        mbar_init(&bars->d2_full, 1);
        mbar_init(&bars->aband_full, 16);
        mbar_init(&bars->xregion_free, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // B1 / B2 images: plain copies (the images are already swizzled)
        const uint4* s1 = reinterpret_cast<const uint4*>(prm.b1_img);
        uint4* d1 = reinterpret_cast<uint4*>(smem + OFF_B1);
        for (int i = tid; i < 8192 / 16; i += THREADS) d1[i] = __ldg(s1 + i);
        const int nb2 = ((N2 + 31) / 32) * 4096 / 16;
        const uint4* s2 = reinterpret_cast<const uint4*>(prm.b2_img);
        uint4* d2 = reinterpret_cast<uint4*>(smem + OFF_B2);
        for (int i = tid; i < nb2; i += THREADS) d2[i] = __ldg(s2 + i);
        // rows >= ROWS of the band operand are read by the last tile's MMAs: keep them finite
        uint4* za = reinterpret_cast<uint4*>(smem + sm.a2);
        for (int i = tid; i < (int)(4 * PLANE / 16); i += THREADS) za[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // the setup above (barriers, TMEM, constant tables) overlapped the tail of the previous kernel in the stream (PDL)
    griddep_wait();
    griddep_launch_dependents();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t sbase = smem_u32(smem);

    const int my_items = (prm.nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    constexpr uint32_t IDESC_S1 = instr_desc(128, 64, 1, 0);   // A = x tile (MN-major), B = B1 (K-major)
    constexpr uint32_t IDESC_S2 = instr_desc(128, 32, 0, 0);   // A = A2 chunk (K-major), B = B2 (K-major)
    constexpr uint32_t IDESC_SA = instr_desc(128, 16, 0, 1);   // A = band (K-major), B = B2 viewed MN-major
    constexpr uint32_t IDESC_SB = instr_desc(128, 64, 0, 1);   // A = stage-B operand (K-major), B = B1 viewed MN-major

    // compute-warp coordinates
    const int cw = warp - 4;             // 0..15 for compute warps
    const int wg = cw >> 2;              // warpgroup 0..3
    const int quad = warp & 3;           // TMEM lane quadrant of this warp
    const uint32_t tq = tmem + ((uint32_t)(quad * 32) << 16);
    const int tp = wg >> 1, h = wg & 1;  // tile parity and column half for the per-tile epilogues

    // x-tile slot rings: every role walks the same sequence of (slot, phase); the epilogue pairs take every second tile
    int xs = warp >= 4 ? tp : 0;         // slot of this role's next tile
    uint32_t xph = 0;                    // its phase bit
    auto xadvance = [&](int step) {
        xs += step;
        if (xs >= nslot) { xs -= nslot; xph ^= 1u; }
    };
    int prefetched = 0;                  // producer: tiles of the CURRENT item already issued at the end of the previous one

    auto issue_x = [&](int b_, int d0_, int i) {   // producer thread: tile i of work item (b_, d0_) into the next slot
        mbar_wait_tc(&bars->x_free[xs], xph ^ 1u, dbg, 1u, (uint32_t)i);
        mbar_expect_tx(&bars->x_full[xs], 16384u + 1024u);
        tma_load_4d(smem + sm.x + (uint32_t)xs * 16384u, &tmap_in, &bars->x_full[xs], d0_, 0, 4 * i, b_);   // box {32 d, 64 m1, 4 n}: [n][m1][d], SWIZZLE_64B
        bulk_g2s(smem + sm.tw + (uint32_t)xs * 1024u, prm.tw + (size_t)(4 * i) * 32, 1024u, &bars->x_full[xs]);
        xadvance(1);
    };

    // SML_DEBUG=1: CTA 0 records %globaltimer at its phase boundaries (words 1024.. of the host-mapped debug record)
    auto mark = [&](int it, int phase) {
        if (TIMING && dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0 && it < 8) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            reinterpret_cast<volatile uint64_t*>(dbg + 1024)[it * 4 + phase] = now;
        }
    };
    // finer accounting (same switch): time warp 4 spends waiting / working inside the epilogues, accumulated per work item
    const bool timing = TIMING && dbg != nullptr && blockIdx.x == 0 && warp == 4;
    uint64_t tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
    auto tick = [&](int slot) {   // adds the time since the previous tick to tacc[slot]
        if (timing) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            tacc[slot] += now - tlast;
            tlast = now;
        }
    };
    for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int b = item / prm.ntd, dt = item - b * prm.ntd;
        const int d0 = dt * 32;
        mark(it, 0);
        if (timing) {
#pragma unroll
            for (int q = 0; q < 8; ++q) tacc[q] = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tlast));
        }
        const uint32_t gi0 = (uint32_t)it * (uint32_t)NT1;   // global tile / chunk use counters (barrier phases)
        const uint32_t gc0 = (uint32_t)it * (uint32_t)NCH;

        if (warp == 0) {
            // =========================== TMA producer ===========================
            if (lane == 0) {
                for (int i = prefetched; i < NT1; ++i) issue_x(b, d0, i);
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t u = gc0 + c, s = u & 1u;
                    mbar_wait_tc(&bars->tws_free[s], ((u >> 1) & 1u) ^ 1u, dbg, 2u, u);
                    mbar_expect_tx(&bars->tws_full[s], 2048u);
                    bulk_g2s(smem + sm.tws + s * 2048u, prm.tw + (size_t)(8 * c) * 32, 2048u, &bars->tws_full[s]);
                }
                // the x-tile region doubles as the stage-B operand: once the last stage-B MMAs have read it, the first tiles of the
                // NEXT work item can land while this item's last outputs are still being written
                prefetched = 0;
                if (it + 1 < my_items) {
                    // (a barrier of its own, one phase per work item: this thread is many phases behind ab_free, and a parity
                    //  wait cannot tell phase j from phase j - 2)
                    mbar_wait_tc(&bars->xregion_free, (uint32_t)it & 1u, dbg, 19u, (uint32_t)it);
                    const int item2 = item + (int)gridDim.x;
                    const int b2 = item2 / prm.ntd, d02 = (item2 - b2 * prm.ntd) * 32;
                    const int npre = nslot < NT1 ? nslot : NT1;
                    for (int i = 0; i < npre; ++i) issue_x(b2, d02, i);
                    prefetched = npre;
                }
            }
        } else if (warp == 1) {
            // =========================== MMA issuer (one thread) ===========================
            if (lane == 0) {
                auto stage2 = [&](int c) {
                    const uint32_t uc = gc0 + c, s2 = uc & 1u;
                    mbar_wait_tc(&bars->a2_full[s2], (uc >> 1) & 1u, dbg, 3u, uc);
                    if (consumer_fence) fence_proxy_async();
                    tc_fence_after();
                    const uint64_t bd = smem_desc(sbase + OFF_B2 + (uint32_t)(c >> 2) * 4096u + (uint32_t)(c & 3) * 32u, 16, 1024, LAYOUT_SW128);
#pragma unroll 1
                    for (int t = 0; t < NTILE; ++t) {
                        const uint64_t ad = smem_desc(sbase + sm.a2 + s2 * 2u * PLANE + (uint32_t)t * 2048u, PLANE, 128, LAYOUT_NONE);
                        mma_bf16(tmem + COL_D2 + 32u * t, ad, bd, IDESC_S2, c > 0 ? 1u : 0u);
                    }
                    mma_commit(&bars->a2_free[s2]);
                };
                // ---- analysis ----
                int c2 = 0;   // next stage-2 chunk: issued as soon as both of its tiles have been transposed (polled before every tile)
                for (int i = 0; i < NT1; ++i) {
                    const uint32_t u = gi0 + i, p = u & 1u;
                    if (c2 < NCH && 2 * c2 + 1 < i) {
                        const uint32_t uc = gc0 + c2;
                        if (mbar_try_wait(&bars->a2_full[uc & 1u], (uc >> 1) & 1u)) { stage2(c2); ++c2; }
                    }
                    mbar_wait_tc(&bars->x_full[xs], xph, dbg, 4u, u);
                    mbar_wait_tc(&bars->d1_free[p], ((u >> 1) & 1u) ^ 1u, dbg, 5u, u);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        // MN-major SWIZZLE_64B: 32 channels (64 B) contiguous, 8 k-rows (m1) at 64 B, one MN group per n
                        const uint64_t ad = smem_desc(sbase + sm.x + (uint32_t)xs * 16384u + ks * 1024u, 4096, 512, LAYOUT_SW64);
                        const uint64_t bd = smem_desc(sbase + OFF_B1 + ks * 32u, 16, 1024, LAYOUT_SW128);
                        mma_bf16(tmem + COL_D1 + 64u * p, ad, bd, IDESC_S1, ks > 0 ? 1u : 0u);
                    }
                    mma_commit(&bars->x_free[xs]);
                    mma_commit(&bars->d1_full[p]);
                    xadvance(1);
                    // never fall more than one chunk behind: the epilogue of chunk c2 + 2 needs the operand buffer of chunk c2
                    if ((i & 1) && i >= 3 && c2 <= ((i - 3) >> 1)) { stage2(c2); ++c2; }
                }
                for (; c2 < NCH; ++c2) stage2(c2);
                mma_commit(&bars->d2_full);
                // ---- synthesis ----
                auto stageA = [&](int c) {
                    const uint32_t uc = gc0 + c, s = uc & 1u;
                    mbar_wait_tc(&bars->da_free[s], ((uc >> 1) & 1u) ^ 1u, dbg, 6u, uc);
                    tc_fence_after();
#pragma unroll 1
                    for (int t = 0; t < NTILE; ++t) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint64_t ad = smem_desc(sbase + sm.a2 + (uint32_t)t * 2048u + ks * 2u * PLANE, PLANE, 128, LAYOUT_NONE);
                            const uint64_t bd = smem_desc(sbase + OFF_B2 + (uint32_t)(c >> 2) * 4096u + ks * 2048u + (uint32_t)(c & 3) * 32u, 0, 1024, LAYOUT_SW128);
                            mma_bf16(tmem + COL_D2 + s * 144u + 16u * t, ad, bd, IDESC_SA, ks > 0 ? 1u : 0u);
                        }
                    }
                    mma_commit(&bars->da_full[s]);
                };
                auto stageB = [&](int i) {
                    const uint32_t u = gi0 + i, p = u & 1u;
                    const int c = i >> 1;
                    const uint32_t uc = gc0 + c, s = uc & 1u;
                    if ((i & 1) == 0) {
                        mbar_wait_tc(&bars->ab_full[s], (uc >> 1) & 1u, dbg, 7u, uc);
                        if (consumer_fence) fence_proxy_async();
                    }
                    mbar_wait_tc(&bars->db_free[p], ((u >> 1) & 1u) ^ 1u, dbg, 8u, u);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = smem_desc(sbase + sm.x + s * 32768u + (uint32_t)(i & 1) * 16384u + ks * 32u, 16, 1024, LAYOUT_SW128);
                        const uint64_t bd = smem_desc(sbase + OFF_B1 + ks * 2048u, 0, 1024, LAYOUT_SW128);
                        mma_bf16(tmem + COL_D1 + 64u * p, ad, bd, IDESC_SB, ks > 0 ? 1u : 0u);
                    }
                    mma_commit(&bars->db_full[p]);
                };
                mbar_wait_tc(&bars->aband_full, (uint32_t)it & 1u, dbg, 9u, (uint32_t)it);
                if (consumer_fence) fence_proxy_async();
                tc_fence_after();
                stageA(0);
                if (NCH > 1) stageA(1);
                for (int c = 0; c < NCH; ++c) {
                    stageB(2 * c);
                    stageB(2 * c + 1);
                    mma_commit(&bars->ab_free[(gc0 + c) & 1u]);
                    if (c + 2 < NCH) stageA(c + 2);
                }
                mma_commit(&bars->xregion_free);
                // every MMA of this work item has completed when the last commit has arrived
                const uint32_t ul = gc0 + NCH - 1;
                mbar_wait_tc(&bars->ab_free[ul & 1u], (ul >> 1) & 1u, dbg, 10u, ul);
            }
        } else if (warp >= 4) {
            // =========================== compute warps ===========================
            // the mid phase reads this item's filter rows (BWD: and its X_low rows) once, cold: pull them into L2 now, one 128-byte
            // line per thread and step, while the analysis streams
            {
                const int ct = cw * 32 + lane;                      // 0..511
                const int lines = (prm.k * 4 + 127) / 128;          // lines per filter row
                for (int idx = ct; idx < 32 * lines; idx += 512) {
                    const int dl = idx / lines, ln = idx - dl * lines;
                    const size_t off = (size_t)(d0 + dl) * prm.F + (size_t)ln * 32;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.w_re + off));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.w_im + off));
                    if (BWD && prm.gw_re != nullptr) {
                        const float2* xl = reinterpret_cast<const float2*>(prm.xlow) + ((size_t)b * prm.D + d0 + dl) * (size_t)prm.k + (size_t)ln * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xl));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xl + 16));
                    }
                }
            }
            // ---- analysis epilogue: stage-1 accumulator -> twiddle -> A2 chunk (transposed) ----
            for (int i = tp; i < NT1; i += 2) {
                const uint32_t u = gi0 + i, p = u & 1u;
                mbar_wait_tc(&bars->x_full[xs], xph, dbg, 11u, u);
                float2 twv[16];
                {
                    const float4* tws = reinterpret_cast<const float4*>(smem + sm.tw + (uint32_t)xs * 1024u + quad * 256 + h * 128);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 q4 = tws[j];
                        twv[2 * j] = make_float2(q4.x, q4.y);
                        twv[2 * j + 1] = make_float2(q4.z, q4.w);
                    }
                }
                tick(0);   // [0] E1: x tile + twiddles landed
                mbar_wait_tc(&bars->d1_full[p], (u >> 1) & 1u, dbg, 12u, u);
                tick(1);   // [1] E1: stage-1 MMA done
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tq + COL_D1 + 64u * p + 32u * h, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&bars->d1_free[p]); mbar_arrive(&bars->x_free[xs]); }
                xadvance(2);
                const int c = i >> 1;
                const uint32_t uc = gc0 + c, s2 = uc & 1u;
                mbar_wait_tc(&bars->a2_free[s2], ((uc >> 1) & 1u) ^ 1u, dbg, 13u, uc);
                // row of (channel lane, class): (lane*34 + p)*16 bytes; k slot of n = 4i + quad inside the chunk: plane i&1, word quad
                unsigned char* const a2 = smem + sm.a2 + s2 * 2u * PLANE + (uint32_t)(i & 1) * PLANE + (uint32_t)lane * (RPD * 16) + quad * 4;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float re = __uint_as_float(v[2 * j]), im = __uint_as_float(v[2 * j + 1]);
                    if (h == 0 && j == 0) {
                        // slots 0 / 1 = Re S_0 / Re S_32 (both real): class 0 is not twiddled, class 32 gets W_T^{32 n}
                        *reinterpret_cast<uint32_t*>(a2 + 0 * 16) = pack_bf16(re, 0.f);
                        *reinterpret_cast<uint32_t*>(a2 + 1 * 16) = pack_bf16(im * twv[0].x, im * twv[0].y);
                        if constexpr (DUMP) {
                            if (item == 0) {
                                float* dp = prm.dump + ((size_t)(4 * i + quad) * 32 + lane) * 64;
                                dp[0] = re; dp[1] = im;
                            }
                        }
                    } else {
                        const cf w = cmul(cf{re, im}, cf{twv[j].x, twv[j].y});
                        const int prow = 16 * h + j + 1;   // class f1 = 16h + j sits in row f1 + 1
                        *reinterpret_cast<uint32_t*>(a2 + prow * 16) = pack_bf16(w.re, w.im);
                        if constexpr (DUMP) {
                            if (item == 0) {
                                float* dp = prm.dump + ((size_t)(4 * i + quad) * 32 + lane) * 64 + 32 * h + 2 * j;
                                dp[0] = w.re; dp[1] = w.im;
                            }
                        }
                    }
                }
                if (!consumer_fence) fence_proxy_async();
                warp_arrive(&bars->a2_full[s2], lane);
                tick(2);   // [2] E1: TMEM load, twiddle, stores, fence
            }

            tick(7);
            mark(it, 1);
            // ---- mid phase: band accumulator -> filter -> band operand of the synthesis ----
            mbar_wait_tc(&bars->d2_full, (uint32_t)it & 1u, dbg, 14u, (uint32_t)it);
            tc_fence_after();
            for (int t = (wg + it) & 3; t < NTILE; t += 4) {
                if (128 * t + 32 * quad >= ROWS) continue;   // rows >= ROWS stay zero (initial fill; nothing ever writes them)
                const int r = 128 * t + 32 * quad + lane;
                uint32_t z[32];
                tmem_ld32(tq + COL_D2 + 32u * t, z);
                tmem_ld_wait();
                uint32_t outw[16];
                const int dl = r / RPD, p = r - dl * RPD;
                if (r < ROWS && p <= 32) {
                    mid_row<BWD>(z, outw, prm, b, d0 + dl, p == 0 ? 0 : p == 1 ? 32 : p - 1);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) outw[q] = 0u;
                }
                if constexpr (DUMP) {
                    if (item == 0) {
                        float* dz = prm.dump + (size_t)N2 * 2048 + (size_t)r * 32;
                        float* dbn = prm.dump + (size_t)N2 * 2048 + 36864 + (size_t)r * 32;
#pragma unroll
                        for (int q = 0; q < 32; ++q) dz[q] = __uint_as_float(z[q]);
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&outw[q]);
                            dbn[2 * q] = __low2float(hh); dbn[2 * q + 1] = __high2float(hh);
                        }
                    }
                }
                unsigned char* const ab = smem + sm.a2 + (uint32_t)r * 16u;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
                    *reinterpret_cast<uint4*>(ab + pl * PLANE) = make_uint4(outw[4 * pl], outw[4 * pl + 1], outw[4 * pl + 2], outw[4 * pl + 3]);
            }
            tc_fence_before();
            if (!consumer_fence) fence_proxy_async();
            warp_arrive(&bars->aband_full, lane);
            mark(it, 2);
            tick(7);   // [7] everything between the per-tile epilogues (mid phase, waits for d2 / band)

            // ---- synthesis epilogues, software-pipelined by one chunk: EA(c) then EB(tiles of chunk c-1) ----
            auto epilogueA = [&](int c) {
                const uint32_t uc = gc0 + c, s = uc & 1u;
                mbar_wait_tc(&bars->tws_full[s], (uc >> 1) & 1u, dbg, 15u, uc);
                mbar_wait_tc(&bars->da_full[s], (uc >> 1) & 1u, dbg, 16u, uc);
                tc_fence_after();
                mbar_wait_tc(&bars->ab_free[s], ((uc >> 1) & 1u) ^ 1u, dbg, 17u, uc);
                tick(3);   // [3] EA: waits (twiddles, stage-A MMA, operand buffer free)
                const float2* const twc = reinterpret_cast<const float2*>(smem + sm.tws + s * 2048u);
                // nine tiles over four warpgroups: the group that takes three rotates with the chunk, so that over the double-buffered
                // chunks the load evens out (every warp waits on the same barriers: a fixed assignment makes one group the critical path)
                for (int t = (wg + c) & 3; t < NTILE; t += 4) {
                    if (128 * t + 32 * quad >= ROWS) continue;   // the last tile is half empty
                    const int r = 128 * t + 32 * quad + lane;
                    const int dl = r / RPD, p = r - dl * RPD;
                    uint32_t v[16];
                    tmem_ld16(tq + COL_D2 + s * 144u + 16u * t, v);
                    tmem_ld_wait();
                    const bool live = r < ROWS && p <= 32 && p != 1;
                    const int word = p >= 2 ? p - 1 : 0;                 // twiddle column AND 4-byte k word of the stage-B operand row
                    // row R = (nn & 3) * 32 + dl of tile nn >> 2: byte nn * 4096 + dl * 128; the swizzle term only depends on dl
                    unsigned char* const dst = smem + sm.x + s * 32768u + (uint32_t)dl * 128u + (uint32_t)(((word >> 2) ^ (dl & 7)) << 4) +
                                               (uint32_t)(word & 3) * 4u;
                    const float2* const twp = twc + word;
#pragma unroll
                    for (int nn = 0; nn < 8; ++nn) {
                        const float yr = __uint_as_float(v[2 * nn]), yi = __uint_as_float(v[2 * nn + 1]);
                        const float2 w = twp[nn * 32];
                        // Y * conj(W_T^{n f1})
                        float vr = yr * w.x + yi * w.y;
                        float vi = yi * w.x - yr * w.y;
                        if (p == 0) vr = yr;                             // class 0 is not twiddled
                        const float up = __shfl_down_sync(0xffffffffu, vr, 1);   // class 32 (row p = 1) hands Re V_32 to its class-0 neighbour
                        if (p == 0) vi = up;
                        if (live) *reinterpret_cast<uint32_t*>(dst + nn * 4096) = pack_bf16(vr, vi);
                        if constexpr (DUMP) {
                            if (item == 0 && live) {
                                float* dp = prm.dump + (size_t)N2 * 2048 + 2 * 36864 + ((size_t)(8 * c + nn) * 32 + dl) * 64 + 2 * word;
                                dp[0] = vr; dp[1] = vi;
                            }
                        }
                    }
                }
                tc_fence_before();
                if (!consumer_fence) fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&bars->ab_full[s]); mbar_arrive(&bars->da_free[s]); mbar_arrive(&bars->tws_free[s]); }
                tick(4);   // [4] EA: work
            };
            // stage-B accumulator -> bf16 -> global: thread (n = 4i + quad, d = lane) owns the 32 rows t = N2*m1 + n, m1 = 32h .. 32h+31;
            // a warp store covers 32 channels = 64 contiguous bytes
            const size_t row_stride = (size_t)N2 * prm.D;
            __nv_bfloat16* const obase = reinterpret_cast<__nv_bfloat16*>(prm.out) + ((size_t)b * prm.T + quad) * prm.D + d0 + lane +
                                         (size_t)(32 * h) * row_stride;
            auto epilogueB = [&](int i) {
                const uint32_t u = gi0 + i, p = u & 1u;
                mbar_wait_tc(&bars->db_full[p], (u >> 1) & 1u, dbg, 18u, u);
                tick(5);   // [5] EB: wait for the stage-B MMA
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tq + COL_D1 + 64u * p + 32u * h, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->db_free[p]);
                __nv_bfloat16* o = obase + (size_t)(4 * i) * prm.D;
#pragma unroll
                for (int j = 0; j < 32; ++j) stg_bf16(o + (size_t)j * row_stride, __uint_as_float(v[j]));
                tick(6);   // [6] EB: work
            };
            epilogueA(0);
            for (int c = 1; c < NCH; ++c) {
                epilogueA(c);
                epilogueB(2 * (c - 1) + tp);
            }
            epilogueB(2 * (NCH - 1) + tp);
            mark(it, 3);
            if (timing && lane == 0 && it < 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) reinterpret_cast<volatile uint64_t*>(dbg + 1024 + 64)[it * 8 + q] = tacc[q];
            }
        }
        // ---- end of the work item: every role has drained; the aliased buffers and TMEM columns change hands ----
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}   // namespace sml
