// energy_smem_launch.cuh — host-side launch of the TMA-staged packed-fp32 energy kernel.
#pragma once

#include "energy_smem.cuh"

namespace dddm {

template <typename T, int M>
int launch_energy_smem_m(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    // 4 columns per thread step (fewest LDS / conversion / loop instructions per column).  fp32: the 110 KB tile
    // limits an SM to 2 CTAs; bf16: the tile is half as big and the kernel is compiled under the 3-CTA register cap
    // (120 registers, no spills) — measured 0.61 of the HBM peak against 0.54 with 2-column steps.
    constexpr int kCols = 4;
    constexpr int kMinCtas = (sizeof(T) == 4) ? 2 : 3;
    auto kernel = energy_fused_smem_kernel<T, M, kCols, kMinCtas>;
    int which = 0;
    // loader: TMA bulk copies; tuning "energy.loader" = 2 selects cp.async commit groups with a bounded window of
    // chunks in flight (measured equal on a single launch, 8.89 against 8.94 us: the load phase is HBM-bound either way)
    const bool ldgsts = tuning().loader == 2;
    if (ldgsts && p.mode != kModeBwd) {
        kernel = energy_fused_smem_kernel<T, M, kCols, kMinCtas, false, 1>;
        which = 4;
    }
    if (p.mode == kModeBwd) {
        kernel = energy_fused_smem_kernel<T, M, kCols, kMinCtas, true>;
        which = 3;
    }
    if constexpr (sizeof(T) == 2) {
        if (p.x0_f32) {  // mixed entry: fp32 x0 beside bf16 draws (fused loss only; TMA loader)
            if (p.mode != kModeLoss) return DDDM_ERR_UNSUPPORTED;
            kernel = energy_fused_smem_kernel<T, M, kCols, kMinCtas, false, 0, true>;
            which = 5;
        } else if (p.mode == kModeBwd || which == 4) {
        } else if (tuning().ctas == 4) {  // experiment: the 4-CTA register cap (96 registers, ~170 B of spills)
            kernel = energy_fused_smem_kernel<T, M, 4, 4>;
            which = 1;
        } else if (tuning().cols == 2) {  // experiment: 2-column steps under the 3-CTA register cap
            kernel = energy_fused_smem_kernel<T, M, 2, 3>;
            which = 2;
        }
    }
    if (plan.threads > kSmemMaxThreads) {  // 8 compute warps, one CTA per SM (single-wave launches; fused loss / split forward)
        if (which != 0 || p.mode == kModeBwd) return DDDM_ERR_UNSUPPORTED;
        if constexpr (M == 8) {
            kernel = energy_fused_smem_kernel<T, M, kCols, 1, false, 0, false, 8>;
            which = 6;
        } else {
            return DDDM_ERR_UNSUPPORTED;
        }
    }
    static SmemOptIn configured[7];  // per instantiation and per device
    if (int e = configured[which].ensure(kernel, plan.smem_bytes, 40 * 1024)) return e;
    // the backward needs no cross-CTA sum: same grid, but the D-slabs of a row run as independent CTAs
    const int cluster = (p.mode == kModeBwd) ? 1 : plan.cluster;
    return launch_with_attrs(kernel, dim3(plan.cluster, p.B), dim3(plan.threads + 32), plan.smem_bytes, cluster, stream,
                             p, plan.slab_vecs, cluster, plan.chunk_vecs);
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
