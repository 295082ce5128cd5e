// bf16 instantiations of the TMA-staged packed-fp32 energy kernel (fp32 accumulation).
#include "energy_smem_launch.cuh"

namespace dddm {
template <>
int launch_energy_smem<__nv_bfloat16>(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M_SMEM(__nv_bfloat16, p, plan, stream)
}
}  // namespace dddm
