// fp32 instantiations of the single-wave register-resident energy kernel + its shape planner.
#include "energy_wave_launch.cuh"

namespace dddm {

int device_sm_count();  // api.cu (per-device cache)

WavePlan plan_wave(int B, int m, int D, int elem_size, bool aligned16) {
    WavePlan w{};
    w.ok = false;
    const int vecw = 16 / elem_size;
    if (m < 2 || m > 8 || D < 1 || !aligned16 || D % vecw != 0) return w;
    const Tuning& t = tuning();
    // one wave, one CTA per SM: beyond that the throughput kernels (two CTAs per SM, rows overlapping) win
    if (t.variant != 5 && B > device_sm_count()) return w;
    const long nvec = D / vecw;
    int threads = t.threads, nv = t.nv;
    if (!(threads == 128 || threads == 256 || threads == 384) || nv < 1 || nv > 3) {
        // auto: 256 threads (2 warps per scheduler) x up to 3 vectors; narrow rows take fewer threads
        threads = nvec <= 128 ? 128 : 256;
        nv = (int)((nvec + threads - 1) / threads);
    }
    if ((long)threads * nv < nvec || nv > 3 || (threads == 384 && nv > 2)) return w;
    w.threads = threads;
    w.nv = nv;
    w.ksmem = t.ksmem > 0 ? 1 : 0;
    w.ok = true;
    return w;
}

template <>
int launch_energy_wave<float>(const EnergyParams& p, const WavePlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M_WAVE(float, p, plan, stream)
}
}  // namespace dddm
