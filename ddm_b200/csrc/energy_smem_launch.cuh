// energy_smem_launch.cuh — host-side launch of the TMA-staged packed-fp32 energy kernel.
#pragma once

#include "energy_smem.cuh"

namespace dddm {

template <typename T, int M>
int launch_energy_smem_m(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    // fp32: the 110 KB tile limits an SM to 2 CTAs, so 4 columns per step (fewest instructions);
    // bf16: the tile is half as big, registers are the limit -> 2 columns per step, 3 CTAs per SM.
    constexpr int kCols = (sizeof(T) == 4) ? 4 : 2;
    constexpr int kMinCtas = (sizeof(T) == 4) ? 2 : 3;
    auto kernel = energy_fused_smem_kernel<T, M, kCols, kMinCtas>;
    if constexpr (sizeof(T) == 2) {
        if (tuning().ctas == 4) kernel = energy_fused_smem_kernel<T, M, kCols, 4>;  // experiment: tighter register cap
    }
    static size_t configured[2] = {0, 0};
    size_t& conf = configured[tuning().ctas == 4 ? 1 : 0];
    if (plan.smem_bytes > 40 * 1024 && plan.smem_bytes > conf) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        conf = plan.smem_bytes;
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
