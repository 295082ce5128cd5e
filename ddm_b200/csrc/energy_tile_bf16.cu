// bf16 instantiations of the shared-memory-tile energy kernels (fp32 accumulation).
#include "energy_tile_launch.cuh"

namespace dddm {
template <>
int launch_energy_tile<__nv_bfloat16>(const EnergyParams& p, const TilePlan& plan, cudaStream_t stream) {
    return launch_tile_any<__nv_bfloat16>(p, plan, false, stream);
}
template <>
int launch_energy_bwd_tile<__nv_bfloat16>(const EnergyParams& p, const TilePlan& plan, cudaStream_t stream) {
    return launch_tile_any<__nv_bfloat16>(p, plan, true, stream);
}
}  // namespace dddm
