// bf16 instantiations of the register-resident energy kernels (fp32 accumulation).
#include "energy_reg.cuh"

namespace dddm {
template <>
int launch_energy_reg<__nv_bfloat16>(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M(launch_energy_reg_m, __nv_bfloat16, p, plan, stream)
}
template <>
int launch_energy_bwd_reg<__nv_bfloat16>(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M(launch_energy_bwd_reg_m, __nv_bfloat16, p, plan, stream)
}
}  // namespace dddm
