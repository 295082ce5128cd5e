// fp32 instantiations of the register-resident energy kernels.
#include "energy_reg.cuh"

namespace dddm {
template <>
int launch_energy_reg<float>(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M(launch_energy_reg_m, float, p, plan, stream)
}
template <>
int launch_energy_bwd_reg<float>(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M(launch_energy_bwd_reg_m, float, p, plan, stream)
}
}  // namespace dddm
