// bf16 instantiations of the single-wave register-resident energy kernel.
#include "energy_wave_launch.cuh"

namespace dddm {
template <>
int launch_energy_wave<__nv_bfloat16>(const EnergyParams& p, const WavePlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M_WAVE(__nv_bfloat16, p, plan, stream)
}
}  // namespace dddm
