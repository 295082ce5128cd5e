// energy_wave_launch.cuh — host-side dispatch of the single-wave register-resident kernel.
#pragma once

#include "energy_wave.cuh"

namespace dddm {

// (threads, nv) instantiations: the register file holds NV * (M+1) * 4 data registers + 72 accumulators per thread
// under the one-CTA-per-SM cap (224 registers at 288 threads, 152 at 416).
template <typename T, int M>
int launch_energy_wave_m(const EnergyParams& p, const WavePlan& plan, cudaStream_t stream) {
    const int ks = plan.ksmem;
    switch (plan.threads * 10 + plan.nv) {
        case 1281: return launch_energy_wave_cfg<T, M, 1, 128>(p, ks, stream);
        case 1282: return launch_energy_wave_cfg<T, M, 2, 128>(p, ks, stream);
        case 1283: return launch_energy_wave_cfg<T, M, 3, 128>(p, ks, stream);
        case 2561: return launch_energy_wave_cfg<T, M, 1, 256>(p, ks, stream);
        case 2562: return launch_energy_wave_cfg<T, M, 2, 256>(p, ks, stream);
        case 2563: return launch_energy_wave_cfg<T, M, 3, 256>(p, ks, stream);
        case 3841: return launch_energy_wave_cfg<T, M, 1, 384>(p, ks, stream);
        case 3842: return launch_energy_wave_cfg<T, M, 2, 384>(p, ks, stream);
        default: return DDDM_ERR_UNSUPPORTED;
    }
}

#define DDDM_DISPATCH_M_WAVE(T, p, plan, stream)                         \
    switch ((p).m) {                                                     \
        case 2: return launch_energy_wave_m<T, 2>(p, plan, stream);      \
        case 3: return launch_energy_wave_m<T, 3>(p, plan, stream);      \
        case 4: return launch_energy_wave_m<T, 4>(p, plan, stream);      \
        case 5: return launch_energy_wave_m<T, 5>(p, plan, stream);      \
        case 6: return launch_energy_wave_m<T, 6>(p, plan, stream);      \
        case 7: return launch_energy_wave_m<T, 7>(p, plan, stream);      \
        case 8: return launch_energy_wave_m<T, 8>(p, plan, stream);      \
        default: return DDDM_ERR_UNSUPPORTED;                            \
    }

}  // namespace dddm
