// fp32 instantiations of the single-wave register-resident energy kernel + its shape planner.
#include "energy_wave_launch.cuh"

namespace dddm {

WavePlan plan_wave(int B, int m, int D, int elem_size, bool aligned16) {
    WavePlan w{};
    w.ok = false;
    const int vecw = 16 / elem_size;
    if (m < 2 || m > 8 || D < 1 || !aligned16 || D % vecw != 0) return w;
    const Tuning& t = tuning();
    // Opt-in (tuning "energy.variant" = 5).  Measured on B200 at B=128, m=8, D=3072 (profiles/r02_k1_single_launch.md):
    // a single launch is bound by the HBM-saturated load phase (the previous launch's dirty gradient lines are
    // written back while the next inputs are read) and by the serial pass 2, not by how the row is staged: fp32
    // 9.1-9.3 us here against 8.9 us for the TMA-staged kernel, bf16 7.7 against 8.1 us; with several launches in
    // flight the TMA-staged kernel (two CTAs per SM) is far ahead (4.1 against 5.4 us).
    if (t.variant != 5) return w;
    (void)B;
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
    w.ksmem = 2;  // pair-major pass 2
    w.ok = true;
    return w;
}

template <>
int launch_energy_wave<float>(const EnergyParams& p, const WavePlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M_WAVE(float, p, plan, stream)
}
}  // namespace dddm
