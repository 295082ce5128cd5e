// fp32 instantiations of the row-pipelined cluster kernel + its shape planner.
#include "energy_pipe_launch.cuh"

namespace dddm {

// A cluster of C CTAs owns C rows; CTA k holds slab k of each: C * (m + x0_rows) * slab bytes of shared memory.
PipePlan plan_pipe(int B, int m, int D, int elem_size, bool aligned16, int x0_rows) {
    PipePlan s{};
    s.ok = false;
    const int vecw = 16 / elem_size;
    if (!(m == 4 || m == 8) || D < 1 || B < 1 || !aligned16 || D % vecw != 0) return s;
    const long nvec = D / vecw;
    const Tuning& t = tuning();
    int cluster = t.cluster;
    if (!(cluster == 2 || cluster == 4 || cluster == 8)) cluster = 4;
    const long slab = (nvec + cluster - 1) / cluster;
    const size_t smem = (size_t)cluster * (m + x0_rows) * slab * 16;
    if (smem > 200 * 1024) return s;
    int threads = t.threads;
    if (threads < 32 || threads > kPipeMaxThreads || threads % 32) threads = 128;
    // columns per thread step: 4 unless 2-column steps divide the slab evenly over the threads and 4-column steps do not
    int cols = t.cols;
    if (x0_rows == 2) cols = 4;  // the mixed tile (fp32 x0 beside bf16 draws) is instantiated for 4-column steps only
    if (!(cols == 2 || cols == 4)) {
        const long q4 = slab * (16 / elem_size) / 4, q2 = q4 * 2;
        cols = (q4 % threads != 0 && q2 % threads == 0) ? 2 : 4;
    }
    s.cluster = cluster;
    s.threads = threads;
    s.cols = cols;
    s.slab_vecs = (int)slab;
    s.window = t.window > 0 ? t.window : cluster;  // rows requested up front (the rest follow as rows land)
    s.smem_bytes = smem;
    s.ok = true;
    return s;
}

template <>
int launch_energy_pipe<float>(const EnergyParams& p, const PipePlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M_PIPE(float, p, plan, stream)
}
}  // namespace dddm
