// energy_blk_launch.cuh — host-side launch of the blocked (m = 16, 32) packed-fp32 energy kernel.
#pragma once

#include "energy_blk.cuh"

namespace dddm {

template <typename T, int M>
int launch_energy_blk_m(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    // two builds: register cap for 2 CTAs per SM (168) or uncapped for tiles that leave room for only one
    const bool two = plan.smem_bytes <= 104 * 1024 && tuning().ctas != 1;
    auto kernel = two ? energy_fused_blk_kernel<T, M, 2> : energy_fused_blk_kernel<T, M, 1>;
    static SmemOptIn configured[2];  // per instantiation and per device
    if (int e = configured[two ? 1 : 0].ensure(kernel, plan.smem_bytes, 40 * 1024)) return e;
    // the backward needs no cross-CTA sum: same grid, but the D-slabs of a row run as independent CTAs
    const int cluster = (p.mode == kModeBwd) ? 1 : plan.cluster;
    return launch_with_attrs(kernel, dim3(plan.cluster, p.B), dim3(plan.threads + 32), plan.smem_bytes, cluster, stream,
                             p, plan.slab_vecs, cluster, plan.chunk_vecs);
}

template <typename T>
int launch_energy_blk_any(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    switch (p.m) {
        case 16: return launch_energy_blk_m<T, 16>(p, plan, stream);
        case 24: return launch_energy_blk_m<T, 24>(p, plan, stream);
        case 32: return launch_energy_blk_m<T, 32>(p, plan, stream);
        default: return DDDM_ERR_UNSUPPORTED;
    }
}

}  // namespace dddm
