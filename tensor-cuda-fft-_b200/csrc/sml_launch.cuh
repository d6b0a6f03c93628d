// sml_launch.cuh -- template definitions of the fused-kernel launchers (one explicit instantiation per sml_inst_*.cu).
#pragma once

#include <atomic>
#include <mutex>

#include "sml_host.h"

namespace sml_host {

#define SML_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)

// The opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute: set it once per (kernel
// instantiation, device) -- a process may drive several GPUs.
template <typename Kern>
int ensure_smem_attr(Kern kern, size_t smem_bytes, std::atomic<unsigned long long>& done) {   // done: one bit per device ordinal
    int dev = 0;
    SML_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return 0;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute(smem=%zu) failed: %s", smem_bytes, cudaGetErrorString(e));
    done.fetch_or(bit, std::memory_order_release);
    return 0;
}

// two landing tiles fit wherever MINB CTAs per SM still fit into 227 KB of shared memory with them
template <int NR, int P, int MINB, typename IO>
constexpr bool xb2_fits() { return (sml::FastCfg<NR, P, IO, 2>::SMEM_BYTES + 1024) * MINB <= 227u * 1024u; }

template <int NR, int KJ, int P, int MINB, typename IO, bool BWD, int XB, bool EXT, bool SPLIT = false>
int launch_fast_xb(const CUtensorMap& map_in, const CUtensorMap& map_out, const CUtensorMap& map_res, const sml::FastParams& prm,
                   int grid, cudaStream_t stream) {
    using C = sml::FastCfg<NR, P, IO, XB>;
    auto kern = sml::sml_fast_kernel<NR, KJ, P, MINB, IO, BWD, XB, EXT, SPLIT>;
    static std::atomic<unsigned long long> attr_done{0};   // per kernel instantiation (this function template)
    if (int rc = ensure_smem_attr(kern, C::SMEM_BYTES, attr_done)) return rc;
    SML_CUDA(launch_pdl(kern, dim3(grid), dim3(C::NT), C::SMEM_BYTES, stream, map_in, map_out, map_res, prm));
    count_launch();
    return 0;
}

template <int NR, int KJ, int P, int MINB, typename IO, bool BWD, bool EXT, bool SPLIT = false>
int launch_fast_inst(const CUtensorMap& map_in, const CUtensorMap& map_out, const CUtensorMap& map_res, const sml::FastParams& prm,
                     int grid, cudaStream_t stream, int xb) {
    if constexpr (xb2_fits<NR, P, MINB, IO>()) {
        if (xb == 2) return launch_fast_xb<NR, KJ, P, MINB, IO, BWD, 2, EXT, SPLIT>(map_in, map_out, map_res, prm, grid, stream);
    }
    return launch_fast_xb<NR, KJ, P, MINB, IO, BWD, 1, EXT, SPLIT>(map_in, map_out, map_res, prm, grid, stream);
}

template <typename IO, bool BWD>
int launch_fast(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm,
                int grid, cudaStream_t stream) {
#define SML_CASE(NR_, KJ_, P_, MINB_) \
    if (p.NR == NR_ && p.KJ == KJ_ && p.P == P_ && p.ctas_per_sm == MINB_) return launch_fast_inst<NR_, KJ_, P_, MINB_, IO, BWD, false>(map_in, map_out, map_in, prm, grid, stream, p.xb);
    SML_CASE(32, 8, 4, 3)
    SML_CASE(32, 12, 4, 3)
    SML_CASE(32, 16, 4, 2)
    SML_CASE(32, 16, 4, 3)
    SML_CASE(32, 24, 4, 2)
    SML_CASE(32, 32, 4, 2)
    SML_CASE(32, 12, 4, 2)
    SML_CASE(16, 4, 8, 3)
    SML_CASE(16, 8, 8, 3)
    SML_CASE(16, 12, 8, 3)
    SML_CASE(16, 16, 8, 3)
    SML_CASE(16, 24, 8, 3)
    SML_CASE(16, 32, 8, 2)
    SML_CASE(8, 4, 32, 2)
    SML_CASE(8, 8, 32, 2)
    SML_CASE(8, 16, 32, 2)
#undef SML_CASE
    return fail("internal: no fast kernel for NR=%d KJ=%d P=%d", p.NR, p.KJ, p.P);
}

// pass-splitting kernels (SPLIT = true, sml_fast.cuh): the largest sub-transform only (T = R * 1024: the long sequences)
template <typename IO, bool BWD>
int launch_fast_split(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm, int grid,
                      cudaStream_t stream) {
#define SML_CASE(NR_, KJ_, P_, MINB_) \
    if (p.NR == NR_ && p.KJ == KJ_ && p.P == P_ && p.ctas_per_sm == MINB_) return launch_fast_inst<NR_, KJ_, P_, MINB_, IO, BWD, false, true>(map_in, map_out, map_in, prm, grid, stream, p.xb);
    SML_CASE(32, 8, 4, 3)
    SML_CASE(32, 12, 4, 3)
    SML_CASE(32, 16, 4, 2)
    SML_CASE(32, 16, 4, 3)
    SML_CASE(32, 24, 4, 2)
    SML_CASE(32, 32, 4, 2)
    SML_CASE(32, 12, 4, 2)
#undef SML_CASE
    return fail("internal: no pass-splitting kernel for NR=%d KJ=%d P=%d", p.NR, p.KJ, p.P);
}

// extended kernels (EXT = true: LayerNorm on load, residual on store, row windows, channel scale, bin T/2; sml_fast.cuh).  The
// largest sub-transform runs with two CTAs per SM here: the row statistics are prefetched into 64 more registers, and two
// landing tiles (the residual rows are prefetched one pass ahead through them) do not fit next to a third fp32 CTA anyway.
template <typename IO, bool BWD>
int launch_fast_ext(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const CUtensorMap& map_res,
                    const sml::FastParams& prm, int grid, cudaStream_t stream) {
#define SML_CASE(NR_, KJ_, P_, MINB_) \
    if (p.NR == NR_ && p.KJ == KJ_ && p.P == P_ && p.ctas_per_sm == MINB_) return launch_fast_inst<NR_, KJ_, P_, MINB_, IO, BWD, true>(map_in, map_out, map_res, prm, grid, stream, p.xb);
    SML_CASE(32, 8, 4, 2)
    SML_CASE(32, 12, 4, 2)
    SML_CASE(32, 8, 4, 3)
    SML_CASE(32, 12, 4, 3)
    SML_CASE(32, 16, 4, 2)
    SML_CASE(32, 24, 4, 2)
    SML_CASE(32, 32, 4, 2)
    SML_CASE(16, 4, 8, 3)
    SML_CASE(16, 8, 8, 3)
    SML_CASE(16, 12, 8, 3)
    SML_CASE(16, 16, 8, 3)
    SML_CASE(16, 24, 8, 3)
    SML_CASE(16, 32, 8, 2)
    SML_CASE(8, 4, 32, 2)
    SML_CASE(8, 8, 32, 2)
    SML_CASE(8, 16, 32, 2)
#undef SML_CASE
    return fail("internal: no extended fast kernel for NR=%d KJ=%d P=%d", p.NR, p.KJ, p.P);
}


}   // namespace sml_host
