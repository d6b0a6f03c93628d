// sml_fast.cuh -- fused SpectralMixingLayer forward / backward kernels for sm_100a (T = R * M, M in {64, 256, 1024}).
//
// What one launch computes (reference: /root/reference/fft_tensor/spectral_layers.py:88-116 and the
// autograd graph it implies, == /root/reference/fft_tensor/wirtinger_ops.py:53-82):
//   FWD : y  = Re(ifft(lowpass_k(fft(x) * W))) + bias               (+ saves X_low = fft(x)[:, :k, :])
//   BWD : gx = Re(ifft(lowpass_k(fft(g) * conj(W)))) ,  gW += (1/T) fft(g)[:k] * conj(X_low) ,  gb += sum g
//
// Algorithm.  Only |f| < k <= M/2 bins are live, so with T = R*M the length-T
// transform is streamed as R passes of a length-M = NR*NR four-step FFT over the decimated rows
// t = R*m + r.  Two real channels (d, d+1) ride in one complex sequence z = x_d + i x_{d+1}; the
// two-sided band of z (2k-1 bins) is accumulated IN REGISTERS across passes:
//     Z[fs] = sum_r W_T^{r fs} FFT_M(z[r::R])[fs mod M]
// The spectral "mid phase" un-mixes the two channels (Hermitian split through warp shuffles), applies the
// complex filter, and re-packs a band spectrum C whose inverse transform is y_d + i y_{d+1}; synthesis is
// the exact transpose of analysis and writes each output row once.  HBM traffic = read x + write y
// (+ 2k/T of that for X_low).
//
// Thread mappings (NT = NR*P threads, P channel pairs per CTA):
//   time side  : tid -> (p = tid % P, m2 = tid / P)   rows m = NR*m1 + m2 ; coalesced 8*P-byte row segments
//   freq side  : tid -> (f1 = tid % NR, p = tid / NR) bins f = f1 + NR*f2 ; partner bin -f is in the same warp
// Global <-> shared traffic is all TMA (cp.async.bulk.tensor.4d loads with mbarrier completion, bulk-group
// stores).  Each CTA owns two shared buffers: X = TMA landing zone (analysis) / output staging (synthesis),
// Y = padded exchange buffer.  The load of pass r+1 is issued as soon as every thread has pulled pass r out of X,
// the store of pass r-1 drains while pass r is computed.  Inter-stage twiddles W_T^{(R m2 + r) f1} are powers of
// one per-thread base and are generated in registers (no table traffic through shared memory).
// The kernel is persistent (static round-robin over (batch, channel-tile) work items); with P = 4 pairs a CTA
// is 4 warps and three CTAs share an SM and drift out of phase, which is what keeps the FMA, shared-memory
// and TMA pipes busy at the same time.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sml_dft.cuh"

namespace sml {

struct FastParams {
    void* out;             // y (FWD) or gx (BWD): (B,T,D) IO
    const float* w_re;     // (D,F)
    const float* w_im;     // (D,F)
    const float* bias;     // (D,) or null            (FWD)
    cf* xlow;              // (B,D,k) complex64: written by FWD (nullable) / read by BWD (nullable)
    float* gw_re;          // (D,F): written by the reduction kernel that follows BWD; here only "filter grads wanted" (nullable)
    cf* gpart;             // (B,D,k) per-batch-element filter-gradient terms (1/T) G conj(X_low), written by BWD
    float* gbpart;         // (B,D)   per-batch-element bias-gradient terms Re G_0 = sum_t g
    const cf* gtab;        // W_T^n = exp(-2 pi i n / T), n < T
    int B, T, D, F, k, R;  // R = T / M passes
    int ntd;               // channel tiles = ceil(D / 2P)
    int ntiles;            // B * ntd
    float invT;
    unsigned int* dbg;     // host-mapped debug record (null unless SML_DEBUG is set): filled by a timed-out mbarrier wait
    // ---- pass splitting: `split` CTAs share one work item, each streams R / split consecutive passes; the partial bands meet in
    //      `xch` (one [2 KJ][threads] complex block per unit, L2-resident) and are summed by every CTA of the group in a fixed order
    int l2_hint;           // 1: TMA loads / stores carry an L2 evict_first policy (tuning knob SML_L2_HINT)
    int split;             // 1 = off
    cf* xch;               // (ntiles * split) partial bands
    unsigned int* xflag;   // (ntiles) arrival counters, zero before the launch
    // ---- extended kernels (template flag EXT) only: the hosting block's prologue / epilogue fused around the transform ----
    const float2* stats;   // (B, T) {mean, rstd} per TRANSFORM row ({0,0} on zero-padding rows): LayerNorm on load, or null
    const float* scale;    // (B, D) per (batch element, channel) factor on the filtered spectrum (context gate), or null
    const float* sb_re;    // (D, F) "spectral bias": added to the filtered spectrum X W before the channel factor (FWD only) --
    const float* sb_im;    //        the image of a LayerNorm beta on a zero-padded window (beta * rfft(rect) * W0), or null
    const float* sb_nyq;   // (D,) its bin T/2, or null
    const float* wnyq;     // (D,) real weight of the bin T/2 (full half-spectrum filters, irfft semantics), or null
    float* xnyq;           // (B, D) spectrum at the bin T/2 (real): written by FWD, read by BWD
    float* gnyqpart;       // (B, D) per-batch-element gradient terms of wnyq, written by BWD
    // BWD: gradient of the per-(b, c) factor from the spectra the mid phase holds anyway (no pass over y):
    //   d_core[b,c] = (1/T) sum_f Re(conj(G_f) X_f W_f) (+ bin T/2),  d_q[b,c] = (1/T) sum_f Re(conj(G_f) Q_f)  -- dL/dscale = d_core + bg[c] d_q
    float* d_core;         // (B, D) or null
    float* d_q;            // (B, D) or null (needs q_re / q_im)
    const float* q_re;     // (F,) rank-one spectral bias: sb[c,f] = bg[c] * Q[f]
    const float* q_im;
    const float* q_nyq;    // (1,) its bin T/2 (device scalar), or null
    // rank-one filter mode (h_re != null): w[d,f] = chan[d] * H[f] (and sb[d,f] = bg[d] * Q[f], FWD) are formed in the mid phase
    // from two (F,) arrays and two (D,) vectors instead of being read as (D, F) arrays; in BWD the filter gradient is contracted
    // over the channels in the kernel: hpart[item, f] = sum_{c in item} chan[c] (1/T) scale G conj(X)  (k complex per work item,
    // summed over the items by the host) -- no (B, D, k) spill -- and d_core is evaluated with H instead of W (dL/dchan = sum_b
    // scale * d_core, dL/dscale = chan * d_core + bg * d_q).  The FastCfg exchange buffer carries the sum over the P channel pairs.
    const float* h_re;
    const float* h_im;
    const float* h_nyq;    // (1,) H at the bin T/2 (real)
    const float* chan;     // (D,)
    const float* bg;       // (D,) or null
    cf* hpart;             // (ntiles, k), BWD
    int res;               // 1: tmap_res describes a residual tensor (output geometry) that is added to the output rows
    int in_q, in_r;        // input row i is transform row in_row0 + i, in_row0 = R * in_q + in_r (rows outside the input read as
                           // zero); output row i is transform row i (rows past the output tensor are not written)
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// cold path of mbar_wait: leave a record in host-mapped memory (SML_DEBUG=1), then trap
static __device__ __noinline__ void mbar_timeout(unsigned int* dbg, uint32_t tag, uint32_t parity, uint32_t aux) {
    if (dbg != nullptr) {   // layout: [0] total, [1..31] count per tag, records of 8 words from word 32: 4 per tag
        atomicAdd(dbg, 1u);
        const unsigned int n = atomicAdd(dbg + 1 + (tag & 31u) % 31u, 1u);
        if (n < 4u) {
            unsigned int* r = dbg + 32 + 8 * (4 * ((tag & 31u) % 31u) + n);
            r[0] = tag; r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = parity; r[4] = aux;
        }
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg = nullptr, uint32_t tag = 0,
                                          uint32_t aux = 0) {
    // bounded wait: a lost completion becomes a trap (reported as a launch failure) instead of a hung GPU.  The bound is
    // wall-clock time (10 s on %globaltimer, sampled every 2^16 polls), not a poll count, so a healthy kernel that is
    // time-sliced (MPS, preemption, a debugger) is not killed by it.
    uint64_t t0 = 0;
    for (uint32_t spins = 1; !mbar_try_wait(bar, parity); ++spins) {
        if ((spins & 0xFFFFu) == 0u) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) mbar_timeout(dbg, tag, parity, aux);
        }
    }
}
// 4-D tiled TMA load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}

// 4-D tiled TMA store shared -> global (bulk async group; SASS: UTMASTG).  Out-of-range elements are clipped.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): activations are streamed once, evict_first keeps them from displacing
// the filter rows / saved spectra that ARE re-used
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_4d_hint(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d_hint(const CUtensorMap* tmap, const void* src, int c0, int c1, int c2, int c3, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
        ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores issued by this thread have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// programmatic dependent launch: wait for the predecessor grid in the stream (all of its memory operations are visible
// afterwards); let the successor grid start its own prologue
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// element I/O: one channel pair (d, d+1) <-> complex
// ------------------------------------------------------------------------------------------------
template <typename IO>
struct PairIO;
template <>
struct PairIO<float> {
    static __device__ __forceinline__ cf load_s(const float* p) {   // shared, 8-byte aligned
        const float2 v = *reinterpret_cast<const float2*>(p);
        return cf{v.x, v.y};
    }
    static __device__ __forceinline__ void store_s(float* p, cf v) {
        *reinterpret_cast<float2*>(p) = make_float2(v.re, v.im);
    }
};
template <>
struct PairIO<__nv_bfloat16> {
    static __device__ __forceinline__ cf load_s(const __nv_bfloat16* p) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
        return cf{v.x, v.y};
    }
    static __device__ __forceinline__ void store_s(__nv_bfloat16* p, cf v) {
        *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.re, v.im);
    }
};

template <int NR, int P, typename IO, int XB = 1>
struct FastCfg {
    static constexpr int NT = NR * P;
    static constexpr int M = NR * NR;
    static constexpr int XS = NR + 2;    // exchange row stride (complex): rows stay 16-byte aligned (128-bit row access)
    static constexpr int CJN = 4 * NR;   // pass twiddles per slot: one per band column (up to 4 NR columns, see KJ)
    static constexpr int BOXROWS = M < 256 ? M : 256;
    static constexpr int NBOX = M / BOXROWS;
    static constexpr uint32_t LOAD_BYTES = (uint32_t)M * 2u * P * sizeof(IO);            // X: one TMA stage, dense [M][2P]
    static constexpr uint32_t XBUF_BYTES = (LOAD_BYTES + 127u) & ~127u;
    static constexpr uint32_t YBUF_BYTES = ((uint32_t)NT * XS * sizeof(cf) + 127u) & ~127u;   // Y: exchange [NT][XS]
    static constexpr size_t SMEM_BYTES = (size_t)XB * XBUF_BYTES + YBUF_BYTES + 2u * CJN * sizeof(cf) + 4 * sizeof(uint64_t);   // 2 mbarriers + 2 drain counters
};

// v[i] *= (or *= conj of) base0 * step^i for i in [0, NR): twiddle powers generated in registers, 8 at a time.
// q holds base0*step^{8h+lo}; step8 = step^8.  CONJ multiplies by the conjugate instead.
template <int NR, bool CONJ, bool UNIT_BASE>
__device__ __forceinline__ void apply_power_twiddles(cf (&v)[NR], const cf base0, const cf step) {
    static_assert(NR % 8 == 0, "NR must be a multiple of 8");
    const cf s2 = cmul(step, step);
    const cf s4 = cmul(s2, s2);
    const cf s8 = cmul(s4, s4);
    cf q[8];
    q[0] = base0;
    q[1] = UNIT_BASE ? step : cmul(base0, step);
    q[2] = UNIT_BASE ? s2 : cmul(base0, s2);
    q[3] = cmul(q[1], s2);
    q[4] = UNIT_BASE ? s4 : cmul(base0, s4);
    q[5] = cmul(q[1], s4);
    q[6] = cmul(q[2], s4);
    q[7] = cmul(q[3], s4);
#pragma unroll
    for (int h = 0; h < NR / 8; ++h) {
#pragma unroll
        for (int lo = 0; lo < 8; ++lo) {
            if (UNIT_BASE && h == 0 && lo == 0) continue;   // multiply by 1
            v[8 * h + lo] = CONJ ? cmulc(v[8 * h + lo], q[lo]) : cmul(v[8 * h + lo], q[lo]);
        }
        if (h + 1 < NR / 8) {
#pragma unroll
            for (int lo = 0; lo < 8; ++lo) q[lo] = (UNIT_BASE && h == 0 && lo == 0) ? s8 : cmul(q[lo], s8);
        }
    }
}

// BWD: pull this tile's X_low rows (read once per tile in the mid phase, cold in HBM) into L2 while the analysis
// passes run.  One lane per 128-byte line.
template <int NR, int KJ>
__device__ __forceinline__ void prefetch_xlow_l2(const FastParams& prm, int b, int d0, int ff1) {
    if (prm.gw_re == nullptr || prm.xlow == nullptr || d0 >= prm.D || (ff1 & 15) != 0) return;
    const float2* row0 = reinterpret_cast<const float2*>(prm.xlow) + ((size_t)b * prm.D + d0) * prm.k;
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
        const int af = ff1 + NR * j;
        if (af < prm.k) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row0 + af));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row0 + prm.k + af));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mid phase (freq-side threads): acc holds the two-sided band Z of z = x_d + i x_{d+1}; on return it holds the band
// C whose inverse transform is y_d + i y_{d+1}.  Hermitian split through warp shuffles (the partner bin -f lives in
// lane NR - f1 of the same warp), complex filter (reference spectral_layers.py:101-105), and for BWD the Wirtinger
// filter gradient terms G conj(X) (reference wirtinger_ops.py:77-80), one per batch element (summed by
// filtergrad_reduce_kernel).
// ------------------------------------------------------------------------------------------------
// side = false (pass splitting: every CTA of a group runs the mid phase on the same summed band): skip the global side outputs
// (X_low, gradient terms, the bin T/2), which the first CTA of the group writes.
template <int NR, int KJ, bool BWD, bool EXT = false, int P = 1>
__device__ __forceinline__ void spectral_mid_phase(cf (&acc)[2 * KJ], const FastParams& prm, int b, int d0, int ff1, int lane, bool side = true,
                                                   cf* ybuf = nullptr, int tile = 0, int tid = 0) {
    constexpr int NJ = 2 * KJ;
    const int D = prm.D;
    const bool pvalid = d0 < D;
    const bool rank1 = EXT && prm.h_re != nullptr;
    const bool hgrads = EXT && BWD && rank1 && prm.hpart != nullptr && side;   // channel-contracted filter gradient through ybuf
    const bool grads = BWD && prm.gw_re != nullptr && side && !hgrads;
    float ch0 = 0.f, ch1 = 0.f, bg0 = 0.f, bg1 = 0.f;
    if constexpr (EXT) {
        if (rank1 && pvalid) {
            ch0 = __ldg(prm.chan + d0); ch1 = __ldg(prm.chan + d0 + 1);
            if (prm.bg != nullptr) { bg0 = __ldg(prm.bg + d0); bg1 = __ldg(prm.bg + d0 + 1); }
        }
        if (hgrads) __syncthreads();   // every warp has finished reading the exchange buffer of the last analysis pass
    }
    const bool want_ds = EXT && BWD && prm.d_core != nullptr && side;   // needs X_low like the filter gradient
    float ec0 = 0.f, ec1 = 0.f, eq0 = 0.f, eq1 = 0.f;
    // EXT: per-(batch element, channel) factor on the filtered spectrum, and the raw analysis value of the bin -T/2
    float sc0 = 1.f, sc1 = 1.f;
    cf znyq = cf{0.f, 0.f};
    if constexpr (EXT) {
        if (prm.scale != nullptr && pvalid) {
            sc0 = __ldg(prm.scale + (size_t)b * D + d0);
            sc1 = __ldg(prm.scale + (size_t)b * D + d0 + 1);
        }
        // the bin -T/2 = -NR*KJ sits at index KJ of the f1 = 0 lanes (the host only offers wnyq on plans with NR*KJ = T/2)
        if (prm.wnyq != nullptr) znyq = acc[KJ];
    }
    // the bin -fs of (ff1, j) lives in lane NR - ff1 of the same warp at index NJ - 1 - j (ff1 = 0: own index NJ - j)
    const int src_lane = (lane & ~(NR - 1)) | ((NR - ff1) & (NR - 1));
    const size_t wrow0 = (size_t)d0 * prm.F, wrow1 = wrow0 + prm.F;
    const size_t xrow0 = ((size_t)b * D + d0) * prm.k, xrow1 = xrow0 + prm.k;
    // Per chunk of JB bins: 1. issue every global load of the chunk up front (filter rows; BWD: saved X_low rows) so that
    // their latencies overlap -- the transform registers are dead here, so the values fit; 2. each thread filters its
    // non-negative bins and serves the mirror bins of its partner lane.
    constexpr int JB = KJ <= 12 ? KJ : 8;   // (chunks of 4 for the extended backward at KJ = 32 remove its ~400 bytes of spills but cost more load rounds: 473 vs 408 us)
    static_assert(KJ % JB == 0, "chunking of the mid phase");
    cf self_mirror = acc[0];   // ff1 == 0: the mirror of bin NR*j is the own bin at index NJ - j (DC mirrors itself)
#pragma unroll
    for (int j0 = 0; j0 < KJ; j0 += JB) {
        float wv[JB][4];   // (D, F) filter of the two channels -- or, in rank-one mode, H[f] in [0..1] and Q[f] in [2..3]
        float2 xv[JB][2];
        // all global loads of the chunk are issued before anything depends on them; the mode branch sits OUTSIDE the unrolled
        // loops (a branch per bin serialises the loads: measured 2x on the whole kernel)
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) wv[jj][0] = wv[jj][1] = wv[jj][2] = wv[jj][3] = 0.f;
        bool done = false;
        if constexpr (EXT) {
            if (rank1) {
#pragma unroll
                for (int jj = 0; jj < JB; ++jj) {
                    const int af = ff1 + NR * (j0 + jj);
                    if (pvalid && af < prm.k) {
                        wv[jj][0] = __ldg(prm.h_re + af);
                        wv[jj][1] = __ldg(prm.h_im + af);
                    }
                }
                if (prm.q_re != nullptr) {
#pragma unroll
                    for (int jj = 0; jj < JB; ++jj) {
                        const int af = ff1 + NR * (j0 + jj);
                        if (pvalid && af < prm.k) {
                            wv[jj][2] = __ldg(prm.q_re + af);
                            wv[jj][3] = __ldg(prm.q_im + af);
                        }
                    }
                }
                done = true;
            }
        }
        if (!done) {
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
                const int af = ff1 + NR * (j0 + jj);
                if (pvalid && af < prm.k) {
                    wv[jj][0] = __ldg(prm.w_re + wrow0 + af);
                    wv[jj][1] = __ldg(prm.w_im + wrow0 + af);
                    wv[jj][2] = __ldg(prm.w_re + wrow1 + af);
                    wv[jj][3] = __ldg(prm.w_im + wrow1 + af);
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
            const int af = ff1 + NR * (j0 + jj);
            xv[jj][0] = xv[jj][1] = make_float2(0.f, 0.f);
            if (pvalid && af < prm.k && (grads || want_ds || hgrads)) {
                xv[jj][0] = __ldg(reinterpret_cast<const float2*>(prm.xlow) + xrow0 + af);
                xv[jj][1] = __ldg(reinterpret_cast<const float2*>(prm.xlow) + xrow1 + af);
            }
        }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
            const int j = j0 + jj;
            const int af = ff1 + NR * j;   // fs >= 0
            const bool live = pvalid && af < prm.k;
            float mre = __shfl_sync(0xffffffffu, acc[NJ - 1 - j].re, src_lane);
            float mim = __shfl_sync(0xffffffffu, acc[NJ - 1 - j].im, src_lane);
            if (ff1 == 0) {
                mre = self_mirror.re;
                mim = self_mirror.im;
            }
            self_mirror = acc[NJ - 1 - j];   // = acc[NJ - (j + 1)], read before this iteration overwrites it
            const cf zp = acc[j], zm = cf{mre, mim};
            // Hermitian split: spectra of the two real channels at +af
            cf s0 = cf{0.5f * (zp.re + zm.re), 0.5f * (zp.im - zm.im)};
            cf s1 = cf{0.5f * (zp.im + zm.im), 0.5f * (zm.re - zp.re)};
            const cf hq = cf{wv[jj][0], wv[jj][1]};                                  // rank-one mode: H[f]
            const cf qq = rank1 ? cf{wv[jj][2], wv[jj][3]} : cf{0.f, 0.f};           // rank-one mode: Q[f]
            const cf w0 = rank1 ? cf{ch0 * hq.re, ch0 * hq.im} : hq;
            const cf w1 = rank1 ? cf{ch1 * hq.re, ch1 * hq.im} : cf{wv[jj][2], wv[jj][3]};
            const float gdc0 = s0.re, gdc1 = s1.re;   // BWD: sum_t g of the two channels when af == 0 (before any scaling)
            if constexpr (EXT && BWD) {
                if (want_ds && live) {   // dL/dscale terms of this bin, from the unscaled G
                    const cf wd0 = rank1 ? hq : w0, wd1 = rank1 ? hq : w1;
                    const cf A0 = cmul(cf{xv[jj][0].x, xv[jj][0].y}, wd0), A1 = cmul(cf{xv[jj][1].x, xv[jj][1].y}, wd1);
                    ec0 = fmaf(s0.re, A0.re, fmaf(s0.im, A0.im, ec0));
                    ec1 = fmaf(s1.re, A1.re, fmaf(s1.im, A1.im, ec1));
                    eq0 = fmaf(s0.re, qq.re, fmaf(s0.im, qq.im, eq0));   // (Q = 0 outside the rank-one mode)
                    eq1 = fmaf(s1.re, qq.re, fmaf(s1.im, qq.im, eq1));
                }
                // y = scale * ifft(W X): the gradient entering the filter is scale * G
                s0 = cf{s0.re * sc0, s0.im * sc0};
                s1 = cf{s1.re * sc1, s1.im * sc1};
            }
            cf a0, a1;
            if constexpr (!BWD) {
                if (live && prm.xlow != nullptr && side) {
                    reinterpret_cast<float2*>(prm.xlow)[xrow0 + af] = make_float2(s0.re, s0.im);
                    reinterpret_cast<float2*>(prm.xlow)[xrow1 + af] = make_float2(s1.re, s1.im);
                }
                a0 = cmul(s0, w0);
                a1 = cmul(s1, w1);
                if constexpr (EXT) {
                    if (prm.sb_re != nullptr && live) {
                        a0 = cf{a0.re + __ldg(prm.sb_re + wrow0 + af), a0.im + __ldg(prm.sb_im + wrow0 + af)};
                        a1 = cf{a1.re + __ldg(prm.sb_re + wrow1 + af), a1.im + __ldg(prm.sb_im + wrow1 + af)};
                    } else if (rank1) {   // bg = 0 / qv = 0 where they are not given
                        a0 = cf{fmaf(bg0, qq.re, a0.re), fmaf(bg0, qq.im, a0.im)};
                        a1 = cf{fmaf(bg1, qq.re, a1.re), fmaf(bg1, qq.im, a1.im)};
                    }
                    a0 = cf{a0.re * sc0, a0.im * sc0};
                    a1 = cf{a1.re * sc1, a1.im * sc1};
                }
            } else {
                if constexpr (EXT) {
                    if (hgrads && af < prm.k) {   // sum over this thread's two channels; the P pairs meet in the exchange buffer
                        const cf g0 = cmulc(s0, cf{xv[jj][0].x, xv[jj][0].y}), g1 = cmulc(s1, cf{xv[jj][1].x, xv[jj][1].y});   // zero when !pvalid
                        ybuf[(size_t)(tid / NR) * (KJ * NR) + af] = cf{(ch0 * g0.re + ch1 * g1.re) * prm.invT, (ch0 * g0.im + ch1 * g1.im) * prm.invT};
                    }
                }
                if (live && grads) {
                    const cf g0 = cmulc(s0, cf{xv[jj][0].x, xv[jj][0].y});   // G conj(X)
                    const cf g1 = cmulc(s1, cf{xv[jj][1].x, xv[jj][1].y});
                    // per-batch-element terms; summed over the batch by filtergrad_reduce_kernel (deterministic, and no
                    // fp32 atomics, which serialise in the LSU at ~1.3 cycles per lane)
                    reinterpret_cast<float2*>(prm.gpart)[xrow0 + af] = make_float2(g0.re * prm.invT, g0.im * prm.invT);
                    reinterpret_cast<float2*>(prm.gpart)[xrow1 + af] = make_float2(g1.re * prm.invT, g1.im * prm.invT);
                    if (af == 0) {
                        prm.gbpart[(size_t)b * D + d0] = gdc0;
                        prm.gbpart[(size_t)b * D + d0 + 1] = gdc1;
                    }
                }
                a0 = cmulc(s0, w0);   // G conj(W)
                a1 = cmulc(s1, w1);
            }
            const float h = 0.5f * prm.invT;
            cf c = cf{h * (a0.re - a1.im), h * (a0.im + a1.re)};
            if (af == 0) {   // DC bin; the bias (reference :116) rides on it: a constant in time is a DC term of y_d + i y_{d+1}
                c = cf{a0.re * prm.invT, a1.re * prm.invT};
                if constexpr (!BWD) {
                    if (prm.bias != nullptr && pvalid) c = cf{c.re + __ldg(prm.bias + d0), c.im + __ldg(prm.bias + d0 + 1)};
                }
            }
            acc[j] = c;
            // hand the value for the mirror bin -fs to its owner (the partner does the same for this thread)
            const cf cn = cf{h * (a0.re + a1.im), h * (a1.re - a0.im)};
            acc[NJ - 1 - j].re = __shfl_sync(0xffffffffu, cn.re, src_lane);
            acc[NJ - 1 - j].im = __shfl_sync(0xffffffffu, cn.im, src_lane);
        }
    }
    // ff1 == 0 lanes received their own values one index too low (bin -NR*j belongs at NJ - j); index KJ is the bin
    // -NR*KJ, outside the band
    if (ff1 == 0) {
#pragma unroll
        for (int idx = NJ - 1; idx > KJ; --idx) acc[idx] = acc[idx - 1];
        acc[KJ] = cf{0.f, 0.f};
    }
    if constexpr (EXT && BWD) {
        if (hgrads) {   // (uniform over the CTA) sum the P partial spectra and write this work item's k bins
            __syncthreads();
            for (int idx = tid; idx < prm.k; idx += NR * P) {
                cf sum = ybuf[idx];
#pragma unroll
                for (int p = 1; p < P; ++p) sum = cadd(sum, ybuf[(size_t)p * (KJ * NR) + idx]);
                reinterpret_cast<float2*>(prm.hpart)[(size_t)tile * prm.k + idx] = make_float2(sum.re, sum.im);
            }
            // (the first synthesis pass writes the exchange buffer only after its barrier (A'))
        }
    }
    if constexpr (EXT && BWD) {
        if (prm.d_core != nullptr) {   // (uniform branch: every lane takes part in the shuffles)
#pragma unroll
            for (int o = NR / 2; o > 0; o >>= 1) {   // sum over the NR lanes (f1) that share this channel pair
                ec0 += __shfl_xor_sync(0xffffffffu, ec0, o);
                ec1 += __shfl_xor_sync(0xffffffffu, ec1, o);
                eq0 += __shfl_xor_sync(0xffffffffu, eq0, o);
                eq1 += __shfl_xor_sync(0xffffffffu, eq1, o);
            }
            if (want_ds && ff1 == 0 && pvalid) {
                const size_t o = (size_t)b * D + d0;
                if (prm.wnyq != nullptr && prm.xnyq != nullptr) {   // bin T/2: G and X real there
                    const float wd0 = rank1 ? (prm.h_nyq != nullptr ? __ldg(prm.h_nyq) : 0.f) : __ldg(prm.wnyq + d0);
                    const float wd1 = rank1 ? wd0 : __ldg(prm.wnyq + d0 + 1);
                    ec0 = fmaf(znyq.re, prm.xnyq[o] * wd0, ec0);
                    ec1 = fmaf(znyq.im, prm.xnyq[o + 1] * wd1, ec1);
                    const float qn = prm.q_nyq != nullptr ? __ldg(prm.q_nyq) : 0.f;
                    eq0 = fmaf(znyq.re, qn, eq0);
                    eq1 = fmaf(znyq.im, qn, eq1);
                }
                prm.d_core[o] = ec0 * prm.invT;
                prm.d_core[o + 1] = ec1 * prm.invT;
                if (prm.d_q != nullptr) { prm.d_q[o] = eq0 * prm.invT; prm.d_q[o + 1] = eq1 * prm.invT; }
            }
        }
    }
    if constexpr (EXT) {
        // bin T/2 (== -T/2): Z = X_d + i X_{d+1} with both spectra real there; irfft semantics: weight 1/T, real filter weight
        if (prm.wnyq != nullptr && ff1 == 0) {
            cf c = cf{0.f, 0.f};
            if (pvalid) {
                const float wn0 = __ldg(prm.wnyq + d0), wn1 = __ldg(prm.wnyq + d0 + 1);
                const size_t o = (size_t)b * D + d0;
                if constexpr (!BWD) {
                    if (prm.xnyq != nullptr && side) { prm.xnyq[o] = znyq.re; prm.xnyq[o + 1] = znyq.im; }
                    float b0 = 0.f, b1 = 0.f;
                    if (prm.sb_nyq != nullptr) { b0 = __ldg(prm.sb_nyq + d0); b1 = __ldg(prm.sb_nyq + d0 + 1); }
                    c = cf{(znyq.re * wn0 + b0) * sc0 * prm.invT, (znyq.im * wn1 + b1) * sc1 * prm.invT};
                } else {
                    const float g0 = znyq.re * sc0, g1 = znyq.im * sc1;
                    if (prm.gnyqpart != nullptr && prm.xnyq != nullptr && side) {
                        prm.gnyqpart[o] = g0 * prm.xnyq[o] * prm.invT;
                        prm.gnyqpart[o + 1] = g1 * prm.xnyq[o + 1] * prm.invT;
                    }
                    c = cf{g0 * wn0 * prm.invT, g1 * wn1 * prm.invT};
                }
            }
            acc[KJ] = c;
        }
    }
}

// band column j -> sub-bin column f2 (the signed column index f2s = j or j - NJ, wrapped into [0, NR))
template <int NR, int KJ>
__host__ __device__ constexpr int band_f2(int j) {
    const int f2s = j < KJ ? j : j - 2 * KJ;
    return ((f2s % NR) + NR) % NR;
}
// true if no earlier band column lands on the same sub-bin column (synthesis: assign instead of accumulate)
template <int NR, int KJ>
__host__ __device__ constexpr bool band_first(int j) {
    for (int jj = 0; jj < j; ++jj)
        if (band_f2<NR, KJ>(jj) == band_f2<NR, KJ>(j)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// XB = number of TMA landing tiles.  XB = 1: the load of pass r+1 is issued when pass r has been drained and has one pass
// of compute to land.  XB = 2 (where shared memory allows three CTAs per SM with it, i.e. bf16 I/O at M = 1024 and the
// small sub-transforms): loads run two passes ahead; during synthesis one tile is the store staging buffer while the other
// already receives pass 0 of the next work item.
// (A per-thread cp.async.cg load path was measured too: 0.157 vs 0.183 ms in a copy-only stream of 32-byte rows,
//  tools/microbench/ldst_stream.cu, but 0.227 vs 0.198 ms inside this kernel, where the copies compete with the exchange
//  traffic for the LSU / shared-memory pipe -- TMA stays the load path.)
// EXT = true: the same kernel with the hosting block's prologue / epilogue fused in (all optional at run time, FastParams):
//   * LayerNorm on load: x^ = (x - mean[t]) * rstd[t] from a per-row statistics array (the affine part folds into the filter
//     and the bias on the host: gamma scales the filter rows, beta is a DC term);
//   * residual add on store: the staging tile first receives the residual rows of the pass by TMA (prm.res, tmap_res: the
//     tile is idle between two stores), the second-stage outputs are added in place;
//   * input / output row windows (in_row0 and the row counts of the tensor maps): rows outside the input are zero-filled
//     by TMA, rows past the output are clipped by TMA -- zero-padded ("linear") convolution and overlap-save without a
//     padded copy (an output window that starts at row o > 0 is a phase ramp exp(2 pi i f o / T) on the filter: host side);
//   * a per-(batch element, channel) factor on the filtered spectrum and a real weight for the bin T/2 (irfft semantics of a
//     full half-spectrum multiplier) -- the causal FFT-convolution core of fft_lm's FixedSpectralBlock.
// With a residual the load stream of a work item is R analysis loads followed by R residual loads, all through the same XB
// landing tiles and mbarriers (load n -> tile n % XB).
// SPLIT = true: pass splitting (FastParams::split CTAs share one work item); compiled only for the largest sub-transform, where
// long sequences leave the grid under-filled -- as a run-time option it cost the default forward kernel 2.7 % (10 more registers).
template <int NR, int KJ, int P, int MINB, typename IO, bool BWD, int XB, bool EXT = false, bool SPLIT = false>
__global__ void __launch_bounds__(NR* P, MINB)
    sml_fast_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                    const __grid_constant__ CUtensorMap tmap_res, const FastParams prm) {
    using C = FastCfg<NR, P, IO, XB>;
    static_assert(XB == 1 || XB == 2, "one or two landing tiles");
    constexpr int NT = C::NT, XS = C::XS;
    constexpr int NJ = 2 * KJ;   // live f2 columns held per thread: [0,KJ) and [NR-KJ, NR)
    // KJ <= NR/2: the band (2k-1 bins) fits into one period of the sub-transform, every sub-bin f2 carries at most one band
    // column.  NR/2 < KJ <= NR ("wide band", k <= M): a sub-bin f1 + NR f2 carries two band bins, fs = f1 + NR f2 >= 0 and
    // fs - M < 0, i.e. columns j = f2 and j = f2 + NJ - NR accumulate the same DFT output with different pass twiddles --
    // this is what lets T = 2k (full half-spectrum, e.g. T = 512 with embed >= 512) run with M = T/2, R = 2.
    // In general a sub-bin carries ceil(NJ / NR) band bins ("aliases", fs = f1 + NR f2 + q M): up to four periods are
    // instantiated for the two small sub-transforms, which keeps T = R * 256 fused for bands up to k = 512.
    static_assert(NJ <= 4 * NR, "band wider than four periods of the sub-transform");
    static_assert(NT % 32 == 0 && 32 % NR == 0, "freq-side partner bin must live in the same warp");
    constexpr int CJN = C::CJN;

    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* const xbuf0 = smem;                                                   // XB landing / staging tiles
    cf* const ybuf = reinterpret_cast<cf*>(smem + XB * C::XBUF_BYTES);                   // exchange [NT][XS]
    cf* const cj = reinterpret_cast<cf*>(smem + XB * C::XBUF_BYTES + C::YBUF_BYTES);     // [2][CJN]  W_T^{NR r f2s(j)}
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(cj + 2 * CJN);                    // [2] one per landing tile
    unsigned int* const xdone = reinterpret_cast<unsigned int*>(mbar + 2);               // [2] warps that have drained X[slot]
    // load n lands in tile n % XB and completes phase n / XB of that tile's mbarrier
    auto xslot = [&](int n) -> int { return XB == 2 ? (n & 1) : 0; };
    auto xbuf = [&](int n) -> unsigned char* { return xbuf0 + (size_t)xslot(n) * C::XBUF_BYTES; };
    auto xparity = [&](int n) -> uint32_t { return XB == 2 ? (((uint32_t)n >> 1) & 1u) : ((uint32_t)n & 1u); };

    const int tid = threadIdx.x;
    const int tp = tid % P, tm2 = tid / P;      // time-side mapping
    const int ff1 = tid % NR, fp2 = tid / NR;   // freq-side mapping
    const int R = prm.R, T = prm.T, D = prm.D;
    const float2* const gtab = reinterpret_cast<const float2*>(prm.gtab);

    // work units: (work item, group member j); member j streams the passes [j Rs, (j + 1) Rs) of the item (split = 1: Rs = R)
    const int S = SPLIT ? prm.split : 1, Rs = SPLIT ? R / S : R;
    const int nunits = prm.ntiles * S;
    const int my_ntiles = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // units of this CTA
    const bool res = EXT && prm.res != 0;        // residual rows ride through the landing tiles during synthesis
    const int lpi = res ? 2 * Rs : Rs;           // loads per unit
    const int total_loads = my_ntiles * lpi;

    // rows of pass r in a tensor whose row i is transform row row0 + i (row0 = R*q0 + r0): t - row0 = R*(m - q0 - borrow) + rs
    auto shifted = [&](int r, int q0, int r0, int& rs, int& mshift) {
        rs = r - r0;
        mshift = q0;
        if (rs < 0) { rs += R; mshift += 1; }
    };
    auto issue_load = [&](int L) {   // thread 0 only.  load L = (work item L / lpi, pass) -> X; passes >= R fetch the residual rows
        if (L >= total_loads) return;
        const int it = L / lpi;
        int r = L - it * lpi;
        const bool resload = r >= Rs;
        if (resload) r -= Rs;
        const int unit = (int)blockIdx.x + it * (int)gridDim.x;
        const int tile = unit / S;
        r += (unit - tile * S) * Rs;
        const int b = tile / prm.ntd, dt = tile - b * prm.ntd;
        int rs = r, mshift = 0;
        if constexpr (EXT) { if (!resload) shifted(r, prm.in_q, prm.in_r, rs, mshift); }   // residual rows: output geometry, no shift
        const CUtensorMap* const tm = (EXT && resload) ? &tmap_res : &tmap_in;
        fence_proxy_async();
#ifdef SML_DIAG_NO_LOAD   // diagnostic build (tools/diag_no_io.sh): the SM side alone, tiles "land" at once with whatever is in shared memory
        (void)tm; (void)rs; (void)mshift; (void)b; (void)dt;
        mbar_arrive(mbar + xslot(L));
#else
        mbar_expect_tx(mbar + xslot(L), C::LOAD_BYTES);
        if (prm.l2_hint) {
            const uint64_t pol = l2_policy_evict_first();
#pragma unroll
            for (int bx = 0; bx < C::NBOX; ++bx)
                tma_load_4d_hint(xbuf(L) + (size_t)bx * C::BOXROWS * 2 * P * sizeof(IO), tm, mbar + xslot(L), dt * 2 * P, rs, bx * C::BOXROWS - mshift, b, pol);
        } else {
#pragma unroll
            for (int bx = 0; bx < C::NBOX; ++bx)
                tma_load_4d(xbuf(L) + (size_t)bx * C::BOXROWS * 2 * P * sizeof(IO), tm, mbar + xslot(L), dt * 2 * P, rs, bx * C::BOXROWS - mshift, b);
        }
#endif
    };
    // (stores are never shifted: TMA tensor stores fault on negative coordinates -- measured -- so an output window that
    //  does not start at transform row 0 is expressed as a phase ramp on the filter by the host instead)
    auto issue_store = [&](const unsigned char* stage, int b, int dt, int r) {   // thread 0 only: staging tile -> rows r + R*m
#ifdef SML_DIAG_NO_STORE
        if (r >= 0) return;
#endif
        if (prm.l2_hint) {
            const uint64_t pol = l2_policy_evict_first();
#pragma unroll
            for (int bx = 0; bx < C::NBOX; ++bx)
                tma_store_4d_hint(&tmap_out, stage + (size_t)bx * C::BOXROWS * 2 * P * sizeof(IO), dt * 2 * P, r, bx * C::BOXROWS, b, pol);
        } else {
#pragma unroll
            for (int bx = 0; bx < C::NBOX; ++bx)
                tma_store_4d(&tmap_out, stage + (size_t)bx * C::BOXROWS * 2 * P * sizeof(IO), dt * 2 * P, r, bx * C::BOXROWS, b);
        }
        tma_store_commit();
    };
    // uniform twiddle of pass r for band column j = tid (threads < NJ): W_T^{NR r f2s}, f2s = j (j < KJ) or j - NJ
    auto cj_load = [&](int r) -> float2 {
        const int f2s = tid < KJ ? tid : tid - NJ;
        // |NR r f2s| < 2 NR^2 R = 2T (|f2s| <= 2 NR): reduce mod T (not a power of two in general) without a division
        int idx = NR * r * f2s;
        if (idx < 0) idx += T;
        if (idx < 0) idx += T;
        if (idx >= T) idx -= T;
        return __ldg(gtab + idx);
    };

    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar + 1, 1);
        fence_mbar_init();
        xdone[0] = 0u;
        xdone[1] = 0u;
    }
    __syncthreads();
    // everything above ran concurrently with the tail of the previous kernel in the stream (PDL); from here on this
    // kernel reads what that kernel may have written (x / g, X_low) and overwrites what it may still be reading
    griddep_wait();
    griddep_launch_dependents();
    if (tid == 0) {
        issue_load(0);
        if constexpr (XB == 2) issue_load(1);
    }

    int L = 0;      // loads consumed so far
    int slot = 0;   // cj slot of the current pass

    for (int it = 0; it < my_ntiles; ++it) {
        const int unit = (int)blockIdx.x + it * (int)gridDim.x;
        const int tile = unit / S;
        const int member = unit - tile * S;
        const int r_begin = member * Rs, r_end = r_begin + Rs;
        const int b = tile / prm.ntd, dt = tile - b * prm.ntd;

        cf acc[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[j] = cf{0.f, 0.f};
        if constexpr (BWD) { if (member == 0) prefetch_xlow_l2<NR, KJ>(prm, b, dt * 2 * P + 2 * fp2, ff1); }

        // ===================== analysis: R streamed passes, band accumulated in registers =====================
        for (int r = r_begin; r < r_end; ++r) {
            // twiddle seeds for this pass (consumed after the first DFT / after barrier (A))
            const float2 wb = __ldg(gtab + (R * tm2 + r));   // W_T^{R m2 + r}
            float2 cjv = make_float2(1.f, 0.f), cjn = make_float2(1.f, 0.f);
            if (tid < NJ) {
                cjv = cj_load(r);
                if (r + 1 == r_end) cjn = cj_load(r_begin);   // first synthesis pass
            }
            cf v[NR];
            bool ln_done = false;
            if constexpr (EXT) {
                if (prm.stats != nullptr) {
                    // LayerNorm on load: the row statistics of this pass are fetched before the tile wait (L2 hits: the 96
                    // channel tiles of a batch element share them), x^ = (x - mean) * rstd as one FMA per element
                    float2 st[NR];
                    const float2* sp = prm.stats + (size_t)b * T + (size_t)R * tm2 + r;
#pragma unroll
                    for (int m1 = 0; m1 < NR; ++m1) st[m1] = __ldg(sp + (size_t)m1 * NR * R);
                    mbar_wait(mbar + xslot(L), xparity(L), prm.dbg, 1u, (uint32_t)L);
                    const IO* src = reinterpret_cast<const IO*>(xbuf(L)) + tm2 * 2 * P + 2 * tp;
#pragma unroll
                    for (int m1 = 0; m1 < NR; ++m1) {
                        const cf x = PairIO<IO>::load_s(src + m1 * NR * 2 * P);
                        const float a = st[m1].y, c = -st[m1].x * st[m1].y;
                        v[m1] = cf{fmaf(x.re, a, c), fmaf(x.im, a, c)};
                    }
                    ln_done = true;
                }
            }
            if (!ln_done) {
                mbar_wait(mbar + xslot(L), xparity(L), prm.dbg, 1u, (uint32_t)L);
                const IO* src = reinterpret_cast<const IO*>(xbuf(L)) + tm2 * 2 * P + 2 * tp;
#pragma unroll
                for (int m1 = 0; m1 < NR; ++m1) v[m1] = PairIO<IO>::load_s(src + m1 * NR * 2 * P);
            }
            // the last warp to drain the tile re-arms it right away with the next load that lands there (XB ahead).  The tile
            // of a work item's last pass becomes the store staging buffer instead and is re-armed after the last store
            // (with a residual and two tiles it takes a residual load right away: staging alternates between the tiles).
            if (r + 1 < r_end || (res && XB == 2)) {
                __syncwarp();
                if ((tid & 31) == 0) {
                    __threadfence_block();
                    if ((atomicAdd(xdone + xslot(L), 1u) % (NT / 32)) == NT / 32 - 1) issue_load(L + XB);
                }
            }
            Dft<NR, -1>::run(v);   // over m1 -> f1
            apply_power_twiddles<NR, false, true>(v, cf{1.f, 0.f}, cf{wb.x, wb.y});   // v[f1] *= W_T^{(R m2 + r) f1}
            __syncthreads();   // (A) last pass's exchange reads are done
            {
                float4* xrow = reinterpret_cast<float4*>(ybuf + tid * XS);
#pragma unroll
                for (int h = 0; h < NR / 2; ++h) xrow[h] = make_float4(v[2 * h].re, v[2 * h].im, v[2 * h + 1].re, v[2 * h + 1].im);
            }
            if (tid < NJ) {
                cj[slot * CJN + tid] = cf{cjv.x, cjv.y};
                if (r + 1 == r_end) cj[(slot ^ 1) * CJN + tid] = cf{cjn.x, cjn.y};
            }
            __syncthreads();   // (B)
            {
                const cf* xb = ybuf + fp2 * XS + ff1;
#pragma unroll
                for (int m2 = 0; m2 < NR; ++m2) v[m2] = xb[m2 * P * XS];
            }
            Dft<NR, -1>::run(v);   // over m2 -> f2   (outputs outside the band are dead code)
            {
                const cf* cjs = cj + slot * CJN;
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[j] = cmac(acc[j], v[band_f2<NR, KJ>(j)], cjs[j]);
            }
            slot ^= 1;
            ++L;
        }

        // ===================== mid phase: un-mix the channel pair, filter, re-pack =====================
        if (SPLIT && S > 1) {
            // pass splitting: publish this unit's partial band, wait for the group, sum the S partials in member order (every
            // CTA of the group gets the same bits); all CTAs of the grid are co-resident (grid <= resident slots), so the wait
            // cannot starve.  Bounded like the mbarrier waits: a lost arrival traps instead of hanging the GPU.
            cf* const mine = prm.xch + (size_t)unit * NJ * NT;
#pragma unroll
            for (int j = 0; j < NJ; ++j) __stcg(reinterpret_cast<float2*>(mine + (size_t)j * NT + tid), make_float2(acc[j].re, acc[j].im));
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                atomicAdd(prm.xflag + tile, 1u);
                uint64_t t0 = 0;
                for (uint32_t spins = 1;; ++spins) {
                    unsigned int seen;
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(prm.xflag + tile) : "memory");
                    if (seen >= (unsigned int)S) break;
                    __nanosleep(64);
                    if ((spins & 0x3FFFu) == 0u) {
                        uint64_t now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 10000000000ull) mbar_timeout(prm.dbg, 3u, seen, (uint32_t)tile);
                    }
                }
            }
            __syncthreads();
            const cf* const grp = prm.xch + (size_t)tile * S * NJ * NT;
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[j] = cf{0.f, 0.f};
            for (int m = 0; m < S; ++m) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float2 v2 = __ldcg(reinterpret_cast<const float2*>(grp + ((size_t)m * NJ + j) * NT + tid));
                    acc[j] = cf{acc[j].re + v2.x, acc[j].im + v2.y};
                }
            }
        }
        spectral_mid_phase<NR, KJ, BWD, EXT, P>(acc, prm, b, dt * 2 * P + 2 * fp2, ff1, tid & 31, member == 0, ybuf, tile, tid);

        // ===================== synthesis: transpose of analysis; rows leave through a TMA store from X =====================
        // staging tile: the tile of this work item's last load, drained by every warp (barrier (A')); with a residual the
        // tile that received (or is about to receive) the residual rows of the pass: load L
        for (int r = r_begin; r < r_end; ++r) {
            unsigned char* const stage = res ? xbuf(L) : xbuf(L - 1);
            // twiddle seeds: v[m2] *= conj(W_T^{r f1} * (W_T^{R f1})^{m2})
            const float2 sr = __ldg(gtab + r * ff1);
            const float2 beta = __ldg(gtab + R * ff1);
            float2 cjn = make_float2(1.f, 0.f);
            if (tid < NJ && r + 1 < r_end) cjn = cj_load(r + 1);
            cf v[NR];
            {
                const cf* cjs = cj + slot * CJN;
#pragma unroll
                for (int f2 = 0; f2 < NR; ++f2) v[f2] = cf{0.f, 0.f};
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int f2 = band_f2<NR, KJ>(j);
                    const cf term = cmulc(acc[j], cjs[j]);
                    // wide band: later columns land on a sub-bin that already holds an alias
                    v[f2] = band_first<NR, KJ>(j) ? term : cadd(v[f2], term);
                }
            }
            Dft<NR, +1>::run(v);   // over f2 -> m2
            apply_power_twiddles<NR, true, false>(v, cf{sr.x, sr.y}, cf{beta.x, beta.y});
            __syncthreads();   // (A') previous exchange readers are done
            {
                cf* xb = ybuf + fp2 * XS + ff1;
#pragma unroll
                for (int m2 = 0; m2 < NR; ++m2) xb[m2 * P * XS] = v[m2];
            }
            if (tid < NJ && r + 1 < r_end) cj[(slot ^ 1) * CJN + tid] = cf{cjn.x, cjn.y};
            if (tid == 0) {
                if (r > r_begin) tma_store_wait_read();   // X may be overwritten after (B')
                if (res) {
                    // one tile: fetch this pass's residual rows now (the tile is free: drained / its last store has been read);
                    // two tiles: the tile of pass r - 1 is free, it takes the residual rows of pass r + 1
                    if (XB == 1) issue_load(L);
                    else if (r > r_begin) issue_load(L + 1);
                }
            }
            __syncthreads();   // (B')
            {
                const float4* xrow = reinterpret_cast<const float4*>(ybuf + tid * XS);
#pragma unroll
                for (int h = 0; h < NR / 2; ++h) {
                    const float4 q = xrow[h];
                    v[2 * h] = cf{q.x, q.y};
                    v[2 * h + 1] = cf{q.z, q.w};
                }
            }
            Dft<NR, +1>::run(v);   // over f1 -> m1
            if (res) {
                mbar_wait(mbar + xslot(L), xparity(L), prm.dbg, 2u, (uint32_t)L);   // residual rows of this pass have landed
                IO* dst = reinterpret_cast<IO*>(stage) + tm2 * 2 * P + 2 * tp;
#pragma unroll
                for (int m1 = 0; m1 < NR; ++m1) PairIO<IO>::store_s(dst + m1 * NR * 2 * P, cadd(v[m1], PairIO<IO>::load_s(dst + m1 * NR * 2 * P)));
                ++L;
            } else {
                IO* dst = reinterpret_cast<IO*>(stage) + tm2 * 2 * P + 2 * tp;
#pragma unroll
                for (int m1 = 0; m1 < NR; ++m1) PairIO<IO>::store_s(dst + m1 * NR * 2 * P, v[m1]);
            }
            fence_proxy_async();
            __syncthreads();   // (C') staging tile complete: the store drains while the next pass computes
            if (tid == 0) issue_store(stage, b, dt, r);
            slot ^= 1;
        }
        if (tid == 0) {
            tma_store_wait_read();
            issue_load(L + XB - 1);   // the staging tile's next load: pass XB - 1 of the next work item
        }
    }
    if (tid == 0) tma_store_wait_all();
}

// batch reduction of the filter / bias gradient terms written by the BWD kernel (wirtinger_ops.py:77-80: sum over dim 0).
// One thread per (d, pair of bins f, f+1 < F) -- pairs that are not both live or not 16-byte aligned take the scalar tail; also zero-fills the columns f >= k, so no memset is needed.
// Deterministic (fixed summation order over b).
// Fused collective (mc != null): gw_re / gw_im / gb are the three blocks of ONE flat buffer that lives in NVLink symmetric
// memory, `mc` is the MULTICAST alias of that buffer: instead of storing its sums locally the kernel pushes them with
// multimem.red.add -- the NVSwitch adds the value into every rank's copy (in-switch reduction, NVLS) -- so after a
// cross-rank barrier every rank holds the global sum: the batch reduction's own store IS the all-reduce.  The buffers
// alternate between two parities; `znext` (the local copy of the other parity) is cleared here for the next step.
__device__ __forceinline__ void multimem_red_add(float* mc_addr, float v) {
    asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}
static __global__ void __launch_bounds__(256) filtergrad_reduce_kernel(const float2* __restrict__ gpart, const float* __restrict__ gbpart,
                                                float* __restrict__ gw_re, float* __restrict__ gw_im,
                                                float* __restrict__ gb, int B, int D, int F, int k,
                                                float* __restrict__ mc, float* __restrict__ znext) {
    // block (64, 4): threadIdx.x -> pair of bins, threadIdx.y -> every 4th batch element; the four partial sums are combined
    // through shared memory in a fixed order (b-slices 0, 1, 2, 3), so the result does not depend on scheduling
    __shared__ float4 part[3][64];
    griddep_wait();                // the fused backward kernel wrote gpart / gbpart (PDL: this grid may be resident before it ends)
    griddep_launch_dependents();
    const int F2 = (F + 1) / 2;
    const long long idx = (long long)blockIdx.x * 64 + threadIdx.x;
    const bool valid = idx < (long long)D * F2;
    const int d = valid ? (int)(idx / F2) : 0, f = valid ? 2 * (int)(idx - (long long)d * F2) : 0;
    const int by = threadIdx.y;
    float sr0 = 0.f, si0 = 0.f, sr1 = 0.f, si1 = 0.f;
    const size_t stride = (size_t)D * k;
    const float2* p = gpart + (size_t)d * k + f;
    if (valid) {
        if (f + 1 < k && ((((size_t)d * k + f) & 1) == 0) && (stride & 1) == 0) {   // both bins live and 16-byte aligned
#pragma unroll 4
            for (int b = by; b < B; b += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)b * stride));
                sr0 += v.x; si0 += v.y; sr1 += v.z; si1 += v.w;
            }
        } else {
            for (int b = by; b < B; b += 4) {
                if (f < k) { const float2 v = __ldg(p + (size_t)b * stride); sr0 += v.x; si0 += v.y; }
                if (f + 1 < k) { const float2 v = __ldg(p + 1 + (size_t)b * stride); sr1 += v.x; si1 += v.y; }
            }
        }
    }
    if (by > 0) part[by - 1][threadIdx.x] = make_float4(sr0, si0, sr1, si1);
    __syncthreads();
    if (by != 0 || !valid) return;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const float4 q = part[s][threadIdx.x];
        sr0 += q.x; si0 += q.y; sr1 += q.z; si1 += q.w;
    }
    const size_t o = (size_t)d * F + f;
    if (mc == nullptr) {
        gw_re[o] = sr0;
        gw_im[o] = si0;
        if (f + 1 < F) {
            gw_re[o + 1] = sr1;
            gw_im[o + 1] = si1;
        }
    } else {
        const size_t nW = (size_t)D * F;
        if (f < k) { multimem_red_add(mc + o, sr0); multimem_red_add(mc + nW + o, si0); }        // columns >= k stay zero
        if (f + 1 < k) { multimem_red_add(mc + o + 1, sr1); multimem_red_add(mc + nW + o + 1, si1); }
        znext[o] = 0.f;
        znext[nW + o] = 0.f;
        if (f + 1 < F) { znext[o + 1] = 0.f; znext[nW + o + 1] = 0.f; }
    }
    if (f == 0) {
        float sb = 0.f;
        for (int b = 0; b < B; ++b) sb += __ldg(gbpart + (size_t)b * D + d);
        if (mc == nullptr) {
            gb[d] = sb;
        } else {
            multimem_red_add(mc + 2 * (size_t)D * F + d, sb);
            znext[2 * (size_t)D * F + d] = 0.f;
        }
    }
}

// Same reduction, four consecutive bins per thread (F and k multiples of 4): 128-bit loads of the per-batch terms and 128-bit
// stores -- or, for the fused collective, multimem.red.v4: a warp pushes 512 contiguous bytes per instruction, so the switch
// sees full-width packets instead of 4-byte ones (at 8 ranks every GPU receives 8x the gradient: packet efficiency decides).
__device__ __forceinline__ void multimem_red_add4(float* mc_addr, float4 v) {
    asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
static __global__ void __launch_bounds__(256) filtergrad_reduce4_kernel(const float2* __restrict__ gpart, const float* __restrict__ gbpart,
                                                 float* __restrict__ gw_re, float* __restrict__ gw_im,
                                                 float* __restrict__ gb, int B, int D, int F, int k,
                                                 float* __restrict__ mc, float* __restrict__ znext) {
    __shared__ float4 part[3][2][64];
    griddep_wait();
    griddep_launch_dependents();
    const int F4 = F / 4;
    const long long idx = (long long)blockIdx.x * 64 + threadIdx.x;
    const bool valid = idx < (long long)D * F4;
    const int d = valid ? (int)(idx / F4) : 0, f = valid ? 4 * (int)(idx - (long long)d * F4) : 0;
    const int by = threadIdx.y;
    float4 re = make_float4(0.f, 0.f, 0.f, 0.f), im = re;
    const bool live = valid && f < k;      // k % 4 == 0: the four bins are live together
    if (live) {
        const size_t stride = (size_t)D * k;
        const float4* p = reinterpret_cast<const float4*>(gpart + (size_t)d * k + f);
#pragma unroll 4
        for (int b = by; b < B; b += 4) {
            const float4 v0 = __ldg(p + (size_t)b * (stride / 2)), v1 = __ldg(p + (size_t)b * (stride / 2) + 1);
            re.x += v0.x; im.x += v0.y; re.y += v0.z; im.y += v0.w;
            re.z += v1.x; im.z += v1.y; re.w += v1.z; im.w += v1.w;
        }
    }
    if (by > 0) { part[by - 1][0][threadIdx.x] = re; part[by - 1][1][threadIdx.x] = im; }
    __syncthreads();
    if (by != 0 || !valid) return;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const float4 a = part[s][0][threadIdx.x], c = part[s][1][threadIdx.x];
        re.x += a.x; re.y += a.y; re.z += a.z; re.w += a.w;
        im.x += c.x; im.y += c.y; im.z += c.z; im.w += c.w;
    }
    const size_t o = (size_t)d * F + f, nW = (size_t)D * F;
    if (mc == nullptr) {
        *reinterpret_cast<float4*>(gw_re + o) = re;      // columns >= k: zeros
        *reinterpret_cast<float4*>(gw_im + o) = im;
    } else {
        if (live) { multimem_red_add4(mc + o, re); multimem_red_add4(mc + nW + o, im); }
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(znext + o) = z;
        *reinterpret_cast<float4*>(znext + nW + o) = z;
    }
    if (f == 0) {
        float sb = 0.f;
        for (int b = 0; b < B; ++b) sb += __ldg(gbpart + (size_t)b * D + d);
        if (mc == nullptr) {
            gb[d] = sb;
        } else {
            multimem_red_add(mc + 2 * nW + d, sb);
            znext[2 * nW + d] = 0.f;
        }
    }
}

}   // namespace sml
