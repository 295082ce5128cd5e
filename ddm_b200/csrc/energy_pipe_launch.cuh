// energy_pipe_launch.cuh — host-side planning and launch of the row-pipelined cluster kernel (energy_pipe.cuh).
#pragma once

#include "energy_pipe.cuh"

namespace dddm {

template <typename T, int M>
int launch_energy_pipe_m(const EnergyParams& p, const PipePlan& plan, cudaStream_t stream) {
    auto kernel = energy_pipe_kernel<T, M, 4>;
    int which = 0;
    if (plan.cols == 2) {
        kernel = energy_pipe_kernel<T, M, 2>;
        which = 1;
    }
    if constexpr (sizeof(T) == 2) {
        if (p.x0_f32) {
            if (plan.cols != 4) return DDDM_ERR_UNSUPPORTED;
            kernel = energy_pipe_kernel<T, M, 4, true>;
            which = 2;
        }
    }
    static SmemOptIn configured[3];
    if (int e = configured[which].ensure(kernel, plan.smem_bytes, 32 * 1024)) return e;
    const int groups = (p.B + plan.cluster - 1) / plan.cluster;
    return launch_with_attrs(kernel, dim3(plan.cluster, groups), dim3(plan.threads + 32), plan.smem_bytes, plan.cluster,
                             stream, p, plan.slab_vecs, plan.cluster, plan.window);
}

#define DDDM_DISPATCH_M_PIPE(T, p, plan, stream)                         \
    switch ((p).m) {                                                     \
        case 4: return launch_energy_pipe_m<T, 4>(p, plan, stream);      \
        case 8: return launch_energy_pipe_m<T, 8>(p, plan, stream);      \
        default: return DDDM_ERR_UNSUPPORTED;                            \
    }

}  // namespace dddm
