// sml_host.h -- host-side declarations shared by the C-ABI translation unit (sml_api.cu) and the per-(IO, direction)
// kernel instantiation units (sml_inst_*.cu), so the fused kernels compile in parallel.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/spectral_mix_b200.h"
#include "sml_fast.cuh"

namespace sml_host {

struct Plan {
    int path = SML_PATH_GENERIC;
    int k = 0;
    int NR = 0, KJ = 0, P = 0, M = 0, R = 0;
    int ctas_per_sm = 1;
    int xb = 1;        // TMA landing tiles (2 only where the kernel variant was built with them)
    bool tunable = false;   // CTAs per SM (2 or 3) decided by a timed first use per shape (sml_api.cu: tuned_ctas)
    bool tc = false;   // tensor-core (tcgen05) kernel of sml_tc.cuh: bf16 I/O, D % 32 == 0, T % 512 == 0, k <= 512
};

// tuning knobs from the environment, read once per process (SML_FAST_CTAS, SML_FAST_XB, SML_TC)
struct Knobs {
    int fast_ctas = 0;   // 2 or 3: CTAs per SM of the NR = 32, KJ = 12 kernel (0 = default)
    int fast_xb = 0;     // 1 or 2: TMA landing tiles (0 = default)
    int tc = -1;         // 0 / 1: tensor-core kernel for eligible bf16 problems (-1 = default)
    int l2_hint = 0;     // SML_L2_HINT=1: evict_first L2 policy on the TMA loads / stores of the activations
    int split = 0;       // SML_SPLIT: CTAs per work item of the pass-splitting schedule (0 = choose by the fill of the grid, 1 = off)
    int ext_ctas = 0;    // 2 or 3: CTAs per SM of the extended NR = 32, KJ <= 12 kernels (0 = default)
    int pdl = 0;         // SML_PDL=1: launch with programmatic dependent launch (measured slower inside the fwd/bwd/reduce chain: off by default)
};
const Knobs& knobs();

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
int get_encode_fn(EncodeTiledFn* out);

// Launch with programmatic stream serialization (PDL): the kernel may become resident and run its prologue (barrier setup,
// table staging) while its predecessor in the stream drains; every kernel of this library executes griddepcontrol.wait
// before it touches global data another kernel may have written, and triggers its own dependents right after that.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_if(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    return launch_pdl_if(knobs().pdl == 1, kern, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

// ---- tensor-core path (sml_inst_tc.cu) ----
struct TcLaunch {
    const float* w_re = nullptr;
    const float* w_im = nullptr;
    const float* bias = nullptr;
    sml::cf* xlow = nullptr;
    float* gw_re = nullptr;
    sml::cf* gpart = nullptr;
    float* gbpart = nullptr;
    int B = 0, T = 0, D = 0, F = 0, k = 0;
    unsigned int* dbg = nullptr;
};
bool tc_eligible(int T, int D, int k, int io_dtype);
template <bool BWD>
int launch_tc(const void* in, void* out, const TcLaunch& a, int sm_count, cudaStream_t stream);
int tc_release_tables();

// records the message returned by sml_last_error() and returns 1
int fail(const char* fmt, ...);
void count_launch(int n = 1);

// launches the fused kernel selected by the plan (defined in sml_launch.cuh, instantiated in sml_inst_*.cu)
template <typename IO, bool BWD>
int launch_fast(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm, int grid,
                cudaStream_t stream);
// the pass-splitting kernels, instantiated in sml_inst_split_*.cu
template <typename IO, bool BWD>
int launch_fast_split(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm, int grid,
                      cudaStream_t stream);
// the extended kernels (block prologue / epilogue fused in), instantiated in sml_inst_ext_*.cu
template <typename IO, bool BWD>
int launch_fast_ext(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const CUtensorMap& map_res,
                    const sml::FastParams& prm, int grid, cudaStream_t stream);

}   // namespace sml_host
