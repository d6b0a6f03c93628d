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
    bool ws = false;   // warp-specialised 512-thread kernel (NR = 32 only)
};

// records the message returned by sml_last_error() and returns 1
int fail(const char* fmt, ...);
void count_launch(int n = 1);

// launches the fused kernel selected by the plan (defined in sml_launch.cuh, instantiated in sml_inst_*.cu)
template <typename IO, bool BWD>
int launch_fast(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm, int grid,
                cudaStream_t stream);

}   // namespace sml_host
