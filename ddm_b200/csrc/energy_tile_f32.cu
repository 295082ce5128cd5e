// fp32 instantiations of the shared-memory-tile energy kernels + the shape planner.
#include "energy_tile_launch.cuh"

namespace dddm {

// Slabs of D per CTA, chunks of a slab per staging step, dynamic shared memory.
TilePlan plan_tile(int m, int D, int elem_size, bool aligned16) {
    TilePlan t{};
    t.ok = false;
    if (m < 2 || m > kTileMaxM || D < 1) return t;
    const int vecw = 16 / elem_size;
    t.bulk = aligned16 && (D % vecw == 0);
    const int vec = t.bulk ? vecw : 1;
    const long nvec = D / vec;
    int cluster = tuning().cluster;
    if (!(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8)) {
        cluster = 8;  // auto: as many CTAs per row as leave >= 32 vectors (one warp-width) per CTA
        while (cluster > 1 && nvec / cluster < 32) cluster /= 2;
    }
    t.cluster = cluster;
    t.threads = kTileThreads;
    long slab_vecs = (nvec + cluster - 1) / cluster;
    t.slab_cols = (int)(slab_vecs * vec);
    const int R = m + 1, RB = (R + 3) & ~3, MB = (m + 3) & ~3;
    const size_t fixed = 16 + (size_t)2 * RB * RB * 4 + (size_t)m * MB * 4 + (size_t)MB * 4;
    const size_t budget = 226 * 1024;  // just below the 227 KB per-CTA opt-in limit
    if (fixed + (size_t)R * vec * elem_size > budget) return t;
    long chunk_vecs = (long)((budget - fixed) / ((size_t)R * vec * elem_size));
    if (chunk_vecs > slab_vecs) chunk_vecs = slab_vecs;
    if (chunk_vecs > 32) chunk_vecs = chunk_vecs / 32 * 32;  // whole warps of vectors per chunk
    if (chunk_vecs < 1) return t;
    t.chunk_cols = (int)(chunk_vecs * vec);
    t.smem_bytes = fixed + (size_t)R * t.chunk_cols * elem_size;
    t.ok = true;
    return t;
}

template <>
int launch_energy_tile<float>(const EnergyParams& p, const TilePlan& plan, cudaStream_t stream) {
    return launch_tile_any<float>(p, plan, false, stream);
}
template <>
int launch_energy_bwd_tile<float>(const EnergyParams& p, const TilePlan& plan, cudaStream_t stream) {
    return launch_tile_any<float>(p, plan, true, stream);
}
}  // namespace dddm
