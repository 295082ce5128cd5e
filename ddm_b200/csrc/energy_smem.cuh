// energy_smem.cuh — TMA-staged, packed-fp32 energy-score kernel for m <= 8 (the headline path).
//
// One thread-block cluster per minibatch row; the CTAs of the cluster split D into slabs.  One
// elected warp stages the CTA's (m+1) x slab tile (m draws + x0) in shared memory with 1-D TMA
// bulk copies (cp.async.bulk, one per row, completion counted on an mbarrier), so the tile is read
// from HBM exactly once, costs no issue slots and no registers, and every thread can walk several
// 16-byte column vectors (amortising the cross-thread reduction).
//   pass 1  m + m(m-1)/2 squared distances; differences and squares in packed fp32 (FADD2/FFMA2:
//           two columns per instruction — the only way to reach B200's fp32 rate), butterfly
//           reduce-scatter across the warp, shared memory across warps, DSMEM across the cluster;
//   coeffs  every CTA evaluates f(d2) and f'(d2) (one powf + one division per distance);
//   pass 2  gradient rows straight from the smem tile, again in packed fp32, streamed out with
//           16-byte stores that bypass L1.
// Reference arithmetic: dddm/losses.py:5-25 (terms), dddm/training.py:84-85 (loss).
#pragma once

#include <cooperative_groups.h>

#include "energy.cuh"
#include "energy_smem_plan.h"
#include "energy_tile.cuh"  // mbarrier / TMA bulk helpers

namespace dddm {

constexpr int kSmemMaxThreads = 256;  // compute threads per CTA; one extra control warp is added at launch
constexpr int kSmemMaxCluster = 8;
constexpr int kSmemMaxChunks = 8;  // column chunks of the tile, each with its own mbarrier

__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

// 16-byte vector in shared memory -> packed fp32 pairs (2 pairs for fp32, 4 for bf16)
template <typename T>
struct Pairs {
    static constexpr int kN = (sizeof(T) == 4) ? 2 : 4;
    float2 v[kN];
};
template <typename T>
__device__ __forceinline__ void lds_pairs(const unsigned char* row, int vec, float2 (&out)[Pairs<T>::kN]) {
    if constexpr (sizeof(T) == 4) {
        const float4 r = reinterpret_cast<const float4*>(row)[vec];
        out[0] = make_float2(r.x, r.y);
        out[1] = make_float2(r.z, r.w);
    } else {
        const uint4 r = reinterpret_cast<const uint4*>(row)[vec];
        out[0] = make_float2(bf16lo(r.x), bf16hi(r.x));
        out[1] = make_float2(bf16lo(r.y), bf16hi(r.y));
        out[2] = make_float2(bf16lo(r.z), bf16hi(r.z));
        out[3] = make_float2(bf16lo(r.w), bf16hi(r.w));
    }
}
template <typename T>
__device__ __forceinline__ void stg_pairs(T* __restrict__ dst, long elem, const float2 (&g)[Pairs<T>::kN]) {
    uint4 r;
    if constexpr (sizeof(T) == 4) {
        r.x = __float_as_uint(g[0].x); r.y = __float_as_uint(g[0].y);
        r.z = __float_as_uint(g[1].x); r.w = __float_as_uint(g[1].y);
    } else {
        r.x = pack_bf16x2(g[0].x, g[0].y); r.y = pack_bf16x2(g[1].x, g[1].y);
        r.z = pack_bf16x2(g[2].x, g[2].y); r.w = pack_bf16x2(g[3].x, g[3].y);
    }
    stg_stream16(dst + elem, r);
}

template <int M>
__host__ __device__ constexpr int pair_slot(int i, int j) {  // i < j
    return M + i * M - i * (i + 1) / 2 + (j - i - 1);
}

template <typename T, int M>
__global__ void __launch_bounds__(kSmemMaxThreads + 32, 1)
energy_fused_smem_kernel(const EnergyParams p, const int slab_vecs, const int cluster_size, const int chunk_vecs) {
    namespace cg = cooperative_groups;
    constexpr int P = M * (M + 1) / 2;
    constexpr int VEC = Elem<T>::kVec;
    constexpr int NP = Pairs<T>::kN;
    using WR = WarpReduce<P>;
    __shared__ __align__(8) uint64_t s_bar[kSmemMaxChunks];
    __shared__ float s_warp[kSmemMaxThreads / 32][P];
    __shared__ float s_cluster[kSmemMaxCluster][P];
    __shared__ float s_coef[P];
    __shared__ float s_val[P];
    extern __shared__ __align__(128) unsigned char s_tile[];  // (M+1) rows x slab_vecs x 16 bytes

    // warps 0 .. nwarps-1 compute; the last warp is the control warp (TMA issue, cross-row reduction)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x >> 5) - 1;
    const int nthr = nwarps * 32;
    const bool control = warp == nwarps;
    const int rank = (cluster_size > 1) ? (int)cg::this_cluster().block_rank() : 0;
    const int b = blockIdx.y;
    if (cluster_size > 1) cluster_arrive_relaxed();  // phase 0: "my shared memory exists"

    const long nvec = p.D / VEC;
    const long v_begin = (long)rank * slab_vecs;
    const int nv = (int)max(0L, min((long)slab_vecs, nvec - v_begin));  // vectors in this CTA's slab
    const int row_bytes = slab_vecs * 16;

    // The slab is staged in column chunks of chunk_vecs vectors (a multiple of the compute-thread
    // count); chunk c signals s_bar[c], so pass 1 starts on chunk 0 while the rest is still in flight.
    const int nchunks = (nv + chunk_vecs - 1) / chunk_vecs;
    if (control && lane == 0) {
        for (int c = 0; c < nchunks; ++c) mbar_init(&s_bar[c], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cudaGridDependencySynchronize();  // PDL: inputs may be produced by the previous kernel in the stream
    if (control) {
        const T* src = (lane < M) ? static_cast<const T*>(p.xhat) + ((long)b * M + lane) * p.D + v_begin * VEC
                                  : static_cast<const T*>(p.x0) + (long)b * p.D + v_begin * VEC;
        for (int c = 0; c < nchunks; ++c) {
            const int c0 = c * chunk_vecs;
            const uint32_t bytes = (uint32_t)min(chunk_vecs, nv - c0) * 16u;
            if (lane == 0) mbar_expect_tx(&s_bar[c], bytes * (uint32_t)(M + 1));
            __syncwarp();
            if (lane <= M)
                tma_bulk_g2s(s_tile + (size_t)lane * row_bytes + (size_t)c0 * 16, src + (long)c0 * VEC, bytes, &s_bar[c]);
        }
    }
    const float W = (p.mode == kModeLoss) ? p.weight_dev[0] * p.weight_scale : 1.0f;
    cudaTriggerProgrammaticLaunchCompletion();

    // ---- pass 1 ----
    float2 acc2[P];
#pragma unroll
    for (int q = 0; q < P; ++q) acc2[q] = make_float2(0.f, 0.f);
    for (int v = control ? nv : tid; v < nv; v += nthr) {
        if ((v - tid) % chunk_vecs == 0) mbar_wait(&s_bar[(v - tid) / chunk_vecs], 0);  // warp-uniform: entering a new chunk
        float2 x[M + 1][NP];
#pragma unroll
        for (int r = 0; r <= M; ++r) lds_pairs<T>(s_tile + (size_t)r * row_bytes, v, x[r]);
#pragma unroll
        for (int h = 0; h < NP; ++h) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const float2 d = sub2(x[i][h], x[M][h]);
                acc2[i] = __ffma2_rn(d, d, acc2[i]);
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = i + 1; j < M; ++j) {
                    const float2 d = sub2(x[i][h], x[j][h]);
                    acc2[pair_slot<M>(i, j)] = __ffma2_rn(d, d, acc2[pair_slot<M>(i, j)]);
                }
        }
    }
    float acc[WR::kPadded];
#pragma unroll
    for (int q = 0; q < WR::kPadded; ++q) acc[q] = (q < P) ? acc2[q < P ? q : 0].x + acc2[q < P ? q : 0].y : 0.f;
    if (!control) WR::run(acc, s_warp[warp], lane);
    __syncthreads();

    // ---- cross-warp and cross-CTA sums, fixed order ----
    if (cluster_size > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster_wait_acquire();  // phase 0 complete: every CTA of the cluster is running
        for (int q = tid; q < P; q += blockDim.x) {
            float t = 0.f;
            for (int w = 0; w < nwarps; ++w) t += s_warp[w][q];
            for (int r = 0; r < cluster_size; ++r) cluster.map_shared_rank(&s_cluster[0][0], r)[rank * P + q] = t;
        }
        cluster_arrive_release();
        cluster_wait_acquire();
    }
    for (int q = tid; q < P; q += blockDim.x) {
        float total = 0.f;
        if (cluster_size > 1) {
            for (int r = 0; r < cluster_size; ++r) total += s_cluster[r][q];
        } else {
            for (int w = 0; w < nwarps; ++w) total += s_warp[w][q];
        }
        // f(d2) and f'(d2) with a single transcendental: f' = (beta/2) f / (d2 + eps)
        float val, der;
        if (p.pw.mode == 2) {
            val = total;
            der = 1.0f;
        } else {
            const float xe = total + kPowEps;
            val = (p.pw.mode == 1) ? sqrtf(xe) : powf(xe, p.pw.half_beta);
            der = p.pw.half_beta * __fdiv_rn(val, xe);
        }
        s_val[q] = val;
        const float cl = p.lam / (2.0f * (float)(M - 1));
        const float nb = (float)p.B * (float)M;
        s_coef[q] = (q < M) ? 2.0f * W / nb * der : -4.0f * W * cl / (nb * (float)(M - 1)) * der;
        if (p.dist != nullptr && rank == 0) p.dist[(long)b * P + q] = total;
    }
    __syncthreads();

    // ---- cross-row reduction: the control warp of the row's first CTA, concurrently with pass 2 ----
    if (control) {
        if (rank == 0) {
            // conf = sum of slots [0, M), inter = 2 * sum of slots [M, P) (ordered pairs), fixed-shape tree
            float c = (lane < M) ? s_val[lane] : 0.f;
            float it = (lane >= M && lane < P) ? s_val[lane] : 0.f;
            if (lane + 32 < P) it += s_val[lane + 32];
            static_assert(P <= 64, "two slots per lane");
            c = warp_sum(c);
            it = 2.0f * warp_sum(it);
            finish_row(p, b, c, it, W, lane);
        }
        return;
    }

    // ---- pass 2 ----
    if (p.grad_xhat != nullptr && nv > 0) {
        float2 K2[P];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const float k = s_coef[q];
            K2[q] = make_float2(k, k);
        }
        T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)b * M * p.D + v_begin * VEC;
        for (int v = tid; v < nv; v += nthr) {
            float2 x[M + 1][NP];
#pragma unroll
            for (int r = 0; r <= M; ++r) lds_pairs<T>(s_tile + (size_t)r * row_bytes, v, x[r]);
#pragma unroll
            for (int i = 0; i < M; ++i) {
                float2 g[NP];
#pragma unroll
                for (int h = 0; h < NP; ++h) g[h] = __fmul2_rn(K2[i], sub2(x[i][h], x[M][h]));
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    if (j == i) continue;
                    const int q = (i < j) ? pair_slot<M>(i, j) : pair_slot<M>(j, i);
#pragma unroll
                    for (int h = 0; h < NP; ++h) g[h] = __ffma2_rn(K2[q], sub2(x[i][h], x[j][h]), g[h]);
                }
                stg_pairs<T>(grow + (long)i * p.D, (long)v * VEC, g);
            }
        }
    }

}

}  // namespace dddm
