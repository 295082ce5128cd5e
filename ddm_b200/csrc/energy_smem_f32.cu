// fp32 instantiations of the TMA-staged packed-fp32 energy kernel + its shape planner.
#include "energy_smem_launch.cuh"

namespace dddm {

SmemPlan plan_smem(int m, int D, int elem_size, bool aligned16, int x0_rows, int B) {
    SmemPlan s{};
    s.ok = false;
    const int vecw = 16 / elem_size;
    if (m < 2 || m > 8 || D < 1 || !aligned16 || D % vecw != 0) return s;
    const long nvec = D / vecw;
    const Tuning& t = tuning();
    int cluster = t.cluster;
    if (!(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8)) {
        // auto: whole rows per CTA whenever the (m+1) x D tile leaves room for two CTAs per SM
        // (measured on B200: fewer, fatter CTAs beat D-split clusters — no DSMEM round trip)
        cluster = 1;
        while (cluster < 8 && (size_t)(m + x0_rows) * ((nvec + cluster - 1) / cluster) * 16 > 112 * 1024) cluster *= 2;
        // Small minibatches (the reference's CIFAR recipe is 64 rows per GPU on 4 GPUs, 32 on 8): with fewer rows than half
        // the SMs a row is split along D so that B x cluster CTAs still fit one CTA per SM — each SM then pulls and writes
        // 1/cluster of a row (profiles/r02/k1_small_batch_clusters.log, one launch, fp32 / bf16: B = 32 6.81 -> 5.89 /
        // 6.80 -> 5.75 us with 4 CTAs per row, B = 64 7.36 -> 7.11 / 7.10 -> 6.47 us with 2; B = 96 is slower split).
        // Slabs stay >= 96 vectors per row: bulk copies under ~1.5 KB are bound per request.
        if (B > 0 && t.threads <= kSmemMaxThreads) {  // (the opt-in 8-warp build owns whole rows)
            const long sms = device_sm_count();
            while (cluster < 4 && (long)B * cluster * 2 <= sms && nvec / (cluster * 2) >= 96) cluster *= 2;
        }
    }
    const long slab = (nvec + cluster - 1) / cluster;
    const size_t smem = (size_t)(m + x0_rows) * slab * 16;  // x0_rows = 2: fp32 x0 next to bf16 draws (mixed entry)
    if (smem > 200 * 1024) return s;  // slab too wide: the chunked tile kernel handles it
    int threads = t.threads;
    // up to 256 compute threads on request (m = 8, one CTA per SM: tuning "energy.threads" = 256)
    if (threads < 32 || threads > (m == 8 && cluster == 1 ? 2 * kSmemMaxThreads : kSmemMaxThreads) || threads % 32) {
        threads = (int)((slab + 31) / 32 * 32);  // auto: 128 compute threads, fewer for narrow slabs
        if (threads > 128) threads = 128;
        if (threads < 32) threads = 32;
    }
    // column chunks: nv vectors per thread per chunk (tuning "energy.nv", default 2), at most kSmemMaxChunks chunks
    int per_thread = (t.nv >= 1) ? t.nv : 2;
    while ((slab + (long)threads * per_thread - 1) / ((long)threads * per_thread) > kSmemMaxChunks) ++per_thread;
    s.chunk_vecs = threads * per_thread;
    s.cluster = cluster;
    s.threads = threads;
    s.slab_vecs = (int)slab;
    s.smem_bytes = smem;
    s.ok = true;
    return s;
}

template <>
int launch_energy_smem<float>(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream) {
    DDDM_DISPATCH_M_SMEM(float, p, plan, stream)
}
}  // namespace dddm
