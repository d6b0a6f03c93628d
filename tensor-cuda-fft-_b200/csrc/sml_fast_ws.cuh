// sml_fast_ws.cuh -- warp-specialised fused SpectralMixingLayer kernel for sm_100a, M = 1024 sub-transforms.
//
// Same math as sml_fast.cuh (read its header first); different machine mapping, built around two measured facts
// (tools/microbench/tma_stream.cu, profiles/):
//   * TMA tiles with 32-byte rows cap a pure load+store stream at ~3.8 TB/s on B200, 64-byte rows reach ~5.9 TB/s,
//     so a CTA must own 8 channel pairs (16 fp32 channels = 64 B per row);
//   * with every warp doing "load, butterflies, exchange, butterflies" in lockstep the FMA pipe and the shared-memory
//     pipe take turns idling; the two halves of the four-step FFT want to run concurrently on different passes.
//
// One 512-thread CTA per SM, two roles of 8 warps each, coupled only through mbarriers:
//   T warps ("time side", thread = (pair p, row residue m2)): own the TMA landing/staging buffer X.
//       analysis : X -> DFT_32 over m1 -> twiddle powers -> Y[c&1]                       (producer of Y)
//       synthesis: Y[c&1] -> iDFT_32 over f1 -> +bias -> X -> TMA store                    (consumer of Y)
//   F warps ("freq side", warp = pair, lane = f1): own the band accumulators (registers) for their pair.
//       analysis : Y[c&1] -> DFT_32 over m2 -> acc += twiddle * (live bins)               (consumer of Y)
//       mid phase: Hermitian split, complex filter (+ X_low save / Wirtinger filter gradient)
//       synthesis: acc -> iDFT_32 over f2 -> twiddle powers -> Y[c&1]                     (producer of Y)
// Buffers: X is double buffered (two TMA landing tiles, loads run two passes ahead; during synthesis one of them is the
// store staging tile while the other already receives pass 0 of the next work item); the exchange tile Y is single
// (full/empty mbarriers, 256 arrivals each): a role only holds it for the copy in or out, the butterflies of pass c+1 (T)
// and pass c (F) still overlap.  Register budgets are rebalanced with setmaxnreg (T: 104, F: 152 registers per thread).
#pragma once

#include "sml_fast.cuh"

namespace sml {

template <typename IO>
struct WsCfg {
    static constexpr int NR = 32, P = 8, M = 1024;
    static constexpr int NT_ROLE = NR * P;   // 256 threads per role
    static constexpr int NT = 2 * NT_ROLE;   // 512
    static constexpr int XS = NR + 2;
    static constexpr int BOXROWS = 256, NBOX = M / BOXROWS;
    static constexpr uint32_t LOAD_BYTES = (uint32_t)M * 2u * P * sizeof(IO);
    static constexpr uint32_t XBUF_BYTES = (LOAD_BYTES + 127u) & ~127u;
    static constexpr uint32_t YBUF_BYTES = ((uint32_t)NT_ROLE * XS * sizeof(cf) + 127u) & ~127u;
    static constexpr uint32_t CJ_BYTES = (uint32_t)P * NR * sizeof(cf);   // one private slot per F warp
    static constexpr size_t SMEM_BYTES = 2u * (size_t)XBUF_BYTES + YBUF_BYTES + CJ_BYTES + 8 * sizeof(uint64_t);
    static constexpr int REGS_T = 104, REGS_F = 152;
};

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ void role_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int KJ, typename IO, bool BWD>
__global__ void __launch_bounds__(512, 1)
    sml_ws_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                  const FastParams prm) {
    using C = WsCfg<IO>;
    constexpr int NR = C::NR, P = C::P, M = C::M, XS = C::XS, NTR = C::NT_ROLE;
    constexpr int NJ = 2 * KJ;
    static_assert(NJ <= NR, "band wider than the sub-transform");

    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* const xbuf0 = smem;                                                       // 2 x landing / staging [M][2P]
    cf* const ybuf = reinterpret_cast<cf*>(smem + 2u * C::XBUF_BYTES);                       // exchange [256][XS]
    cf* const cjbuf = reinterpret_cast<cf*>(smem + 2u * C::XBUF_BYTES + C::YBUF_BYTES);      // [P][NR]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(cjbuf + P * NR);
    uint64_t* const xfull = bars;            // [2] TMA landed (tx)
    uint64_t* const yfull = bars + 2;        // producer role has written Y
    uint64_t* const yempty = bars + 3;       // consumer role has drained Y
    unsigned int* const xdone = reinterpret_cast<unsigned int*>(bars + 4);   // [2] T warps that have drained X[slot]
    auto xbuf = [&](int slot) -> unsigned char* { return xbuf0 + (size_t)slot * C::XBUF_BYTES; };

    const int tid = threadIdx.x;
    const int R = prm.R, T = prm.T, D = prm.D;
    const float2* const gtab = reinterpret_cast<const float2*>(prm.gtab);
    const int my_ntiles = (prm.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        mbar_init(xfull, 1);
        mbar_init(xfull + 1, 1);
        mbar_init(yfull, NTR);
        mbar_init(yempty, NTR);
        xdone[0] = 0u;
        xdone[1] = 0u;
        fence_mbar_init();
    }
    __syncthreads();

    if (tid < NTR) {
        // =====================================================================================================
        // T role
        // =====================================================================================================
        setmaxnreg_dec<C::REGS_T>();
        const int tp = tid % P, tm2 = tid / P;
        const int total_loads = my_ntiles * R;
        // load n = (tile n / R, pass n % R) -> X[n & 1]; completion on xfull[n & 1], phase (n >> 1).
        // Order of issue: 0 and 1 up front; draining load L (not the last pass of its tile) frees its slot for L + 2;
        // the slot of a tile's last load is the store staging tile during synthesis and is re-armed (load + 1 of the next
        // tile) once the tile's last store has left it.
        auto issue_load = [&](int n) {   // one thread
            if (n >= total_loads) return;
            const int it = n / R, r = n - it * R;
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int b = tile / prm.ntd, dt = tile - b * prm.ntd;
            unsigned char* dst = xbuf(n & 1);
            fence_proxy_async();
            mbar_expect_tx(xfull + (n & 1), C::LOAD_BYTES);
#pragma unroll
            for (int bx = 0; bx < C::NBOX; ++bx)
                tma_load_4d(dst + (size_t)bx * C::BOXROWS * 2 * P * sizeof(IO), &tmap_in, xfull + (n & 1), dt * 2 * P, r, bx * C::BOXROWS, b);
        };
        if (tid == 0) {
            issue_load(0);
            issue_load(1);
        }
        int L = 0;          // loads consumed
        unsigned int c = 0; // Y passes so far (analysis and synthesis alike)
        for (int it = 0; it < my_ntiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int b = tile / prm.ntd, dt = tile - b * prm.ntd;
            // ---------------- analysis: X -> DFT over m1 -> twiddle -> Y ----------------
            for (int r = 0; r < R; ++r) {
                const float2 wb = __ldg(gtab + (R * tm2 + r));   // W_T^{R m2 + r}
                const int slot = L & 1;
                mbar_wait(xfull + slot, ((uint32_t)L >> 1) & 1u, prm.dbg, 10u, (uint32_t)L);
                cf v[NR];
                {
                    const IO* src = reinterpret_cast<const IO*>(xbuf(slot)) + tm2 * 2 * P + 2 * tp;
#pragma unroll
                    for (int m1 = 0; m1 < NR; ++m1) v[m1] = PairIO<IO>::load_s(src + m1 * NR * 2 * P);
                }
                if (r + 1 < R) {   // the last T warp to drain X[slot] re-arms it two loads ahead
                    __syncwarp();
                    if ((tid & 31) == 0) {
                        __threadfence_block();
                        if ((atomicAdd(xdone + slot, 1u) % (NTR / 32)) == NTR / 32 - 1) issue_load(L + 2);
                    }
                }
                Dft<NR, -1>::run(v);
                apply_power_twiddles<NR, false, true>(v, cf{1.f, 0.f}, cf{wb.x, wb.y});
                mbar_wait(yempty, (c & 1u) ^ 1u, prm.dbg, 11u, c);
                {
                    float4* xrow = reinterpret_cast<float4*>(ybuf + tid * XS);
#pragma unroll
                    for (int h = 0; h < NR / 2; ++h) xrow[h] = make_float4(v[2 * h].re, v[2 * h].im, v[2 * h + 1].re, v[2 * h + 1].im);
                }
                mbar_arrive(yfull);
                ++c;
                ++L;
            }
            // Role switch: from here on the T warps WAIT on yfull, a barrier they themselves arrived on during analysis.
            // A warp two phases ahead of a slower sibling would see the stale parity and fall through, so all T warps
            // meet here first (every analysis arrival has been made before any synthesis wait starts).
            role_barrier(1, NTR);
            // ---------------- synthesis: Y -> iDFT over f1 -> +bias -> staging tile -> TMA store ----------------
            unsigned char* const stage = xbuf((L - 1) & 1);   // slot of the tile's last load (drained by every T warp)
            for (int r = 0; r < R; ++r) {
                cf v[NR];
                mbar_wait(yfull, c & 1u, prm.dbg, 12u, c);
                {
                    const float4* xrow = reinterpret_cast<const float4*>(ybuf + tid * XS);
#pragma unroll
                    for (int h = 0; h < NR / 2; ++h) {
                        const float4 q = xrow[h];
                        v[2 * h] = cf{q.x, q.y};
                        v[2 * h + 1] = cf{q.z, q.w};
                    }
                }
                mbar_arrive(yempty);
                Dft<NR, +1>::run(v);
                if (tid == 0 && r > 0) tma_store_wait_read();   // previous rows have left the staging tile
                role_barrier(1, NTR);
                {
                    IO* dst = reinterpret_cast<IO*>(stage) + tm2 * 2 * P + 2 * tp;
#pragma unroll
                    for (int m1 = 0; m1 < NR; ++m1) PairIO<IO>::store_s(dst + m1 * NR * 2 * P, v[m1]);
                }
                fence_proxy_async();
                role_barrier(1, NTR);                           // staging tile complete
                if (tid == 0) {
#pragma unroll
                    for (int bx = 0; bx < C::NBOX; ++bx)
                        tma_store_4d(&tmap_out, stage + (size_t)bx * C::BOXROWS * 2 * P * sizeof(IO), dt * 2 * P, r, bx * C::BOXROWS, b);
                    tma_store_commit();
                }
                ++c;
            }
            if (tid == 0) {
                tma_store_wait_read();
                issue_load(L + 1);   // the staging slot's next load: pass 1 of the next tile (pass 0 is already in flight)
            }
        }
        if (tid == 0) tma_store_wait_all();
    } else {
        // =====================================================================================================
        // F role
        // =====================================================================================================
        setmaxnreg_inc<C::REGS_F>();
        const int ft = tid - NTR;
        const int ff1 = ft % NR, fp2 = ft / NR;   // lane = f1, warp = pair
        cf* const cjs = cjbuf + fp2 * NR;         // this warp's private twiddle slot
        auto cj_load = [&](int r) -> float2 {     // lane f2: W_T^{NR r f2s}
            const int f2s = ff1 < NR / 2 ? ff1 : ff1 - NR;
            // |NR r f2s| < 2 NR^2 R = 2T (|f2s| <= 2 NR): reduce mod T (not a power of two in general) without a division
        int idx = NR * r * f2s;
        if (idx < 0) idx += T;
        if (idx < 0) idx += T;
        if (idx >= T) idx -= T;
        return __ldg(gtab + idx);
        };
        unsigned int c = 0;
        for (int it = 0; it < my_ntiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int b = tile / prm.ntd, dt = tile - b * prm.ntd;
            cf acc[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[j] = cf{0.f, 0.f};
            if constexpr (BWD) prefetch_xlow_l2<NR, KJ>(prm, b, dt * 2 * P + 2 * fp2, ff1);
            // ---------------- analysis: Y -> DFT over m2 -> accumulate the band ----------------
            for (int r = 0; r < R; ++r) {
                const float2 cjv = cj_load(r);
                cf v[NR];
                mbar_wait(yfull, c & 1u, prm.dbg, 20u, c);
                {
                    const cf* xb = ybuf + fp2 * XS + ff1;
#pragma unroll
                    for (int m2 = 0; m2 < NR; ++m2) v[m2] = xb[m2 * P * XS];
                }
                mbar_arrive(yempty);
                __syncwarp();
                cjs[ff1] = cf{cjv.x, cjv.y};
                __syncwarp();
                Dft<NR, -1>::run(v);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int f2 = j < KJ ? j : NR - NJ + j;
                    acc[j] = cmac(acc[j], v[f2], cjs[f2]);
                }
                ++c;
            }
            // ---------------- mid phase ----------------
            spectral_mid_phase<NR, KJ, BWD>(acc, prm, b, dt * 2 * P + 2 * fp2, ff1, ft & 31);
            // ---------------- synthesis: band -> iDFT over f2 -> twiddle -> Y ----------------
            for (int r = 0; r < R; ++r) {
                const float2 sr = __ldg(gtab + r * ff1);     // W_T^{r f1}
                const float2 beta = __ldg(gtab + R * ff1);   // W_T^{R f1}
                const float2 cjv = cj_load(r);
                __syncwarp();
                cjs[ff1] = cf{cjv.x, cjv.y};
                __syncwarp();
                cf v[NR];
#pragma unroll
                for (int f2 = 0; f2 < NR; ++f2) v[f2] = cf{0.f, 0.f};
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int f2 = j < KJ ? j : NR - NJ + j;
                    v[f2] = cmulc(acc[j], cjs[f2]);
                }
                Dft<NR, +1>::run(v);
                apply_power_twiddles<NR, true, false>(v, cf{sr.x, sr.y}, cf{beta.x, beta.y});
                mbar_wait(yempty, (c & 1u) ^ 1u, prm.dbg, 21u, c);
                {
                    cf* xb = ybuf + fp2 * XS + ff1;
#pragma unroll
                    for (int m2 = 0; m2 < NR; ++m2) xb[m2 * P * XS] = v[m2];
                }
                mbar_arrive(yfull);
                ++c;
            }
            // Role switch (see the T role): the F warps arrived on yfull during synthesis and wait on it in the next
            // tile's analysis; meet first so that no warp can be two phases ahead of the barrier.
            role_barrier(2, NTR);
        }
    }
}

}   // namespace sml
