// energy_tile_launch.cuh — host-side launch of the shared-memory-tile energy kernels.
#pragma once

#include "energy_tile.cuh"

namespace dddm {

template <typename T, int VEC>
int launch_tile_impl(const EnergyParams& p, const TilePlan& plan, bool from_dist, cudaStream_t stream) {
    auto kernel = energy_tile_kernel<T, VEC>;
    static SmemOptIn configured;  // per instantiation and per device
    if (int e = configured.ensure(kernel, plan.smem_bytes, 48 * 1024)) return e;
    TileArgs a{};
    a.slab_cols = plan.slab_cols;
    a.chunk_cols = plan.chunk_cols;
    a.bulk = plan.bulk ? 1 : 0;
    a.from_dist = from_dist ? 1 : 0;
    a.cluster = from_dist ? 1 : plan.cluster;  // the backward needs no cross-CTA reduction
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan.cluster, p.B);
    cfg.blockDim = dim3(plan.threads);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    int n = 0;
    if (a.cluster > 1) {
        attrs[n].id = cudaLaunchAttributeClusterDimension;
        attrs[n].val.clusterDim.x = a.cluster;
        attrs[n].val.clusterDim.y = 1;
        attrs[n].val.clusterDim.z = 1;
        ++n;
    }
    if (tuning().pdl) {
        attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p, a);
    count_launch();
    return (int)e;
}

template <typename T>
int launch_tile_any(const EnergyParams& p, const TilePlan& plan, bool from_dist, cudaStream_t stream) {
    if (plan.bulk) return launch_tile_impl<T, Elem<T>::kVec>(p, plan, from_dist, stream);
    return launch_tile_impl<T, 1>(p, plan, from_dist, stream);
}

}  // namespace dddm
