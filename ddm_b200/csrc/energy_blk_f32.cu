// fp32 instantiations of the blocked (m = 16, 32) packed-fp32 energy kernel + its shape planner.
#include "energy_blk_launch.cuh"

namespace dddm {

SmemPlan plan_blk(int m, int D, int elem_size, bool aligned16) {
    SmemPlan s{};
    s.ok = false;
    const int vecw = 16 / elem_size;
    if (!(m == 16 || m == 24 || m == 32) || D < 1 || !aligned16 || D % vecw != 0) return s;
    const long nvec = D / vecw;
    const int P = m * (m + 1) / 2;
    const Tuning& t = tuning();
    auto smem_for = [&](int cluster) {
        const long slab = (nvec + cluster - 1) / cluster;
        return (size_t)(m + 1) * slab * 16 + (size_t)kBlkMaxSplit * P * 4;
    };
    int cluster = t.cluster;
    if (!(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8)) {
        // auto: fewest CTAs per row whose tile leaves room for two CTAs per SM (loads of one overlap the
        // arithmetic of the other), but never slabs narrower than one step per compute thread
        cluster = 1;
        // (tuning "energy.ctas" = 1: fattest slabs that fit one CTA per SM instead)
        const size_t budget = (t.ctas == 1) ? 216 * 1024 : 104 * 1024;
        while (cluster < 8 && smem_for(cluster) > budget && nvec / (cluster * 2) >= 64) cluster *= 2;
    }
    if (smem_for(cluster) > 216 * 1024) {
        while (cluster < 8 && smem_for(cluster) > 216 * 1024) cluster *= 2;
        if (smem_for(cluster) > 216 * 1024) return s;  // the chunked tile kernel handles it
    }
    const long slab = (nvec + cluster - 1) / cluster;
    int threads = t.threads;
    if (threads < 32 || threads > kSmemMaxThreads || threads % 32) {
        threads = (int)((slab * (16 / (kBlkCols * elem_size)) + 31) / 32 * 32);
        if (threads > 128) threads = 128;
        if (threads < 32) threads = 32;
    }
    int per_thread = (t.nv >= 1) ? t.nv : 2;
    while ((slab + (long)threads * per_thread - 1) / ((long)threads * per_thread) > kSmemMaxChunks) ++per_thread;
    s.chunk_vecs = threads * per_thread;
    s.cluster = cluster;
    s.threads = threads;
    s.slab_vecs = (int)slab;
    s.smem_bytes = smem_for(cluster);
    s.ok = true;
    return s;
}

template <>
int launch_energy_blk<float>(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    return launch_energy_blk_any<float>(p, plan, stream);
}
}  // namespace dddm
