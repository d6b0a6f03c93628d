// explicit instantiation of the pass-splitting fused kernels for IO = __nv_bfloat16, BWD = true (see sml_launch.cuh)
#include "sml_launch.cuh"

namespace sml_host {
template int launch_fast_split<__nv_bfloat16, true>(const Plan&, const CUtensorMap&, const CUtensorMap&, const sml::FastParams&, int, cudaStream_t);
}
