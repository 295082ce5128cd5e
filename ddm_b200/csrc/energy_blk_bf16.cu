// bf16 instantiations of the blocked (m = 16, 32) packed-fp32 energy kernel (fp32 accumulation).
#include "energy_blk_launch.cuh"

namespace dddm {
template <>
int launch_energy_blk<__nv_bfloat16>(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    return launch_energy_blk_any<__nv_bfloat16>(p, plan, stream);
}
}  // namespace dddm
