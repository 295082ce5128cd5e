// energy_smem_launch.cuh — host-side launch of the TMA-staged packed-fp32 energy kernel.
#pragma once

#include "energy_smem.cuh"

namespace dddm {

template <typename T, int M>
int launch_energy_smem_m(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    auto kernel = energy_fused_smem_kernel<T, M>;
    static size_t configured = 0;
    if (plan.smem_bytes > 40 * 1024 && plan.smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = plan.smem_bytes;
    }
    return launch_with_attrs(kernel, dim3(plan.cluster, p.B), dim3(plan.threads + 32), plan.smem_bytes, plan.cluster, stream,
                             p, plan.slab_vecs, plan.cluster, plan.chunk_vecs);
}

#define DDDM_DISPATCH_M_SMEM(T, p, plan, stream)                         \
    switch ((p).m) {                                                     \
        case 2: return launch_energy_smem_m<T, 2>(p, plan, stream);      \
        case 3: return launch_energy_smem_m<T, 3>(p, plan, stream);      \
        case 4: return launch_energy_smem_m<T, 4>(p, plan, stream);      \
        case 5: return launch_energy_smem_m<T, 5>(p, plan, stream);      \
        case 6: return launch_energy_smem_m<T, 6>(p, plan, stream);      \
        case 7: return launch_energy_smem_m<T, 7>(p, plan, stream);      \
        case 8: return launch_energy_smem_m<T, 8>(p, plan, stream);      \
        default: return DDDM_ERR_UNSUPPORTED;                            \
    }

}  // namespace dddm
