// bf16 instantiations of the row-pipelined cluster kernel.
#include "energy_pipe_launch.cuh"

namespace dddm {
template <>
int launch_energy_pipe<__nv_bfloat16>(const EnergyParams& p, const PipePlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M_PIPE(__nv_bfloat16, p, plan, stream)
}
}  // namespace dddm
