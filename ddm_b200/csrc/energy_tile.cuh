// energy_tile.cuh — shared-memory-tile energy-score kernels for any number of draws m (<= 64).
//
// Used when the register-resident kernel does not apply (m > 8, or very wide rows).  One cluster
// per minibatch row, CTAs split D into slabs; each CTA stages its (m+1) x chunk tile (the m draws
// and x0 as row m) in shared memory with TMA bulk copies (cp.async.bulk + mbarrier expect_tx), one
// copy per row, so the tile is read from HBM once and both passes run out of shared memory:
//   pass 1  all pairwise squared distances among the m+1 rows, register-blocked 4x4 per warp with
//           lanes striding the columns, butterfly warp reduction, per-CTA pair matrix in smem,
//           cluster-wide sum through distributed shared memory (fixed order -> deterministic);
//   pass 2  gradient rows g_i = A_i (x_i - x0) + sum_j K_ij (x_i - x_j), register-blocked over
//           4 draws, coefficients broadcast from shared memory, 16-byte streaming stores.
// Slabs wider than the shared-memory budget are processed in chunks; pass 2 then re-stages the
// chunks (they come from L2).  Rows that are not 16-byte aligned fall back to plain loads.
//
// Reference arithmetic: dddm/losses.py:5-25, dddm/training.py:84-85.
#pragma once

#include <cooperative_groups.h>

#include "energy.cuh"

namespace dddm {
namespace cg = cooperative_groups;

constexpr int kTileThreads = 256;
constexpr int kTileMaxM = 64;

struct TileArgs {
    int cluster;
    int slab_cols;
    int chunk_cols;
    int bulk;       // stage with cp.async.bulk
    int from_dist;  // backward-only launch: coefficients from saved distances, no pass 1
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T, int VEC>
__device__ __forceinline__ void lds_pack(const unsigned char* row, int v, float (&out)[VEC]) {
    if constexpr (VEC == 1) {
        out[0] = Elem<T>::to_float(reinterpret_cast<const T*>(row)[v]);
    } else if constexpr (sizeof(T) == 4) {
        const float4 r = reinterpret_cast<const float4*>(row)[v];
        out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    } else {
        const uint4 r = reinterpret_cast<const uint4*>(row)[v];
        out[0] = bf16lo(r.x); out[1] = bf16hi(r.x); out[2] = bf16lo(r.y); out[3] = bf16hi(r.y);
        out[4] = bf16lo(r.z); out[5] = bf16hi(r.z); out[6] = bf16lo(r.w); out[7] = bf16hi(r.w);
    }
}

// Stage rows [0, R) x columns [col0, col0+cols) of this minibatch row into the smem tile.
template <typename T>
__device__ __forceinline__ void stage_chunk(const EnergyParams& p, const TileArgs& a, int b, long col0, int cols,
                                            unsigned char* tile, int row_stride, uint64_t* bar, uint32_t& parity) {
    const int R = p.m + 1;
    const T* xrow = static_cast<const T*>(p.xhat) + (long)b * p.m * p.D;
    const T* crow = static_cast<const T*>(p.x0) + (long)b * p.D;
    if (a.bulk) {
        if (threadIdx.x < 32) {
            const uint32_t bytes = (uint32_t)cols * (uint32_t)sizeof(T);
            if (threadIdx.x == 0) mbar_expect_tx(bar, bytes * (uint32_t)R);
            __syncwarp();
            for (int r = threadIdx.x; r < R; r += 32) {
                const T* src = (r < p.m) ? xrow + (long)r * p.D + col0 : crow + col0;
                tma_bulk_g2s(tile + (size_t)r * row_stride, src, bytes, bar);
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
    } else {
        for (int idx = threadIdx.x; idx < R * cols; idx += blockDim.x) {
            const int r = idx / cols, c = idx - r * cols;
            const T* src = (r < p.m) ? xrow + (long)r * p.D : crow;
            reinterpret_cast<T*>(tile + (size_t)r * row_stride)[c] = src[col0 + c];
        }
        __syncthreads();
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kTileThreads)
energy_tile_kernel(const EnergyParams p, const TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = p.m, R = m + 1;
    const int RB = (R + 3) & ~3;   // pair-matrix stride (rows incl. x0, padded to the 4x4 blocking)
    const int MB = (m + 3) & ~3;   // coefficient-matrix stride
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* pd_local = reinterpret_cast<float*>(smem_raw + 16);
    float* pd_total = pd_local + RB * RB;
    float* Kmat = pd_total + RB * RB;  // [m][MB]: K[j][i]
    float* Avec = Kmat + m * MB;       // [MB]
    unsigned char* tile = reinterpret_cast<unsigned char*>(Avec + MB);
    const int row_stride = a.chunk_cols * (int)sizeof(T);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rank = (a.cluster > 1) ? (int)cg::this_cluster().block_rank() : (int)blockIdx.x;
    const int b = blockIdx.y;
    const long slab0 = (long)rank * a.slab_cols;
    const int slab = (int)max(0L, min((long)a.slab_cols, (long)p.D - slab0));
    const int nchunks = (slab + a.chunk_cols - 1) / a.chunk_cols;

    if (tid == 0 && a.bulk) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int idx = tid; idx < RB * RB; idx += blockDim.x) pd_local[idx] = 0.f;
    __syncthreads();
    uint32_t parity = 0;
    cudaGridDependencySynchronize();

    float W = 1.0f, gc = 0.f, gi = 0.f;
    if (a.from_dist) {
        gc = p.g_conf[0];
        gi = p.g_inter[0];
    } else if (p.mode == kModeLoss) {
        W = p.weight_dev[0] * p.weight_scale;
        gc = W;
        gi = -W * p.lam / (2.0f * (float)(m - 1));
    }

    // ---------------- pass 1: pairwise squared distances ----------------
    if (!a.from_dist) {
        const int nb = RB / 4;
        const int nblk = nb * (nb + 1) / 2;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * a.chunk_cols;
            const int cols = min(a.chunk_cols, slab - c0);
            if (ch > 0) __syncthreads();  // previous chunk fully consumed before it is overwritten
            stage_chunk<T>(p, a, b, slab0 + c0, cols, tile, row_stride, bar, parity);
            const int nv = cols / VEC;
            int bi = 0, bj = 0;  // enumerate blocks (bi <= bj) in order; warp takes every nwarps-th
            for (int blk = 0; blk < nblk; ++blk) {
                if (blk % nwarps == warp) {
                    float acc[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
                    const unsigned char* ri[4];
                    const unsigned char* rj[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        ri[q] = tile + (size_t)min(4 * bi + q, R - 1) * row_stride;
                        rj[q] = tile + (size_t)min(4 * bj + q, R - 1) * row_stride;
                    }
                    for (int v = lane; v < nv; v += 32) {
                        float xi[4][VEC], xj[4][VEC];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            lds_pack<T, VEC>(ri[q], v, xi[q]);
                            lds_pack<T, VEC>(rj[q], v, xj[q]);
                        }
#pragma unroll
                        for (int s = 0; s < 4; ++s)
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int e = 0; e < VEC; ++e) {
                                    const float d = xi[s][e] - xj[u][e];
                                    acc[s * 4 + u] = fmaf(d, d, acc[s * 4 + u]);
                                }
                    }
                    const float tot = reduce_scatter_chunk<16, 0, 16>(acc, lane);
                    if (lane < 16) {
                        const int i = 4 * bi + (lane >> 2), j = 4 * bj + (lane & 3);
                        pd_local[i * RB + j] += tot;  // this warp is the only writer of the block
                    }
                }
                if (++bj == nb) {
                    ++bi;
                    bj = bi;
                }
            }
        }
        __syncthreads();
        // cluster-wide sum of the pair matrices (pull, fixed rank order)
        if (a.cluster > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();
            for (int idx = tid; idx < RB * RB; idx += blockDim.x) {
                float s = 0.f;
                for (int r = 0; r < a.cluster; ++r) s += cluster.map_shared_rank(pd_local, r)[idx];
                pd_total[idx] = s;
            }
            cluster.sync();  // nobody leaves (or reuses pd_local) while peers still read it
        } else {
            for (int idx = tid; idx < RB * RB; idx += blockDim.x) pd_total[idx] = pd_local[idx];
            __syncthreads();
        }
    }

    // ---------------- coefficients, per-row sums, saved distances ----------------
    const long P = (long)m + (long)m * (m - 1) / 2;
    float* vals = pd_local;  // reuse: vals[i*m + j] (j > i) pair values, vals[m*m + i] confinement values
    for (int idx = tid; idx < m * m + m; idx += blockDim.x) {
        if (idx < m * m) {
            const int i = idx / m, j = idx - i * m;
            if (j > i) {
                const long q = (long)m + (long)i * m - (long)i * (i + 1) / 2 + (j - i - 1);
                const float d2 = a.from_dist ? p.dist[(long)b * P + q] : pd_total[i * RB + j];
                const float k = pair_coef(d2, gi, p);
                Kmat[j * MB + i] = k;
                Kmat[i * MB + j] = k;
                vals[idx] = pow_value(d2, p.pw);
                if (!a.from_dist && p.dist != nullptr && rank == 0) p.dist[(long)b * P + q] = d2;
            } else {
                if (j == i) Kmat[i * MB + i] = 0.f;
                vals[idx] = 0.f;
            }
        } else {
            const int i = idx - m * m;
            const float d2 = a.from_dist ? p.dist[(long)b * P + i] : pd_total[i * RB + m];
            Avec[i] = conf_coef(d2, gc, p);
            vals[idx] = pow_value(d2, p.pw);
            if (!a.from_dist && p.dist != nullptr && rank == 0) p.dist[(long)b * P + i] = d2;
        }
    }
    for (int idx = tid; idx < m * (MB - m) + (MB - m); idx += blockDim.x) {  // zero the padding columns
        if (idx < m * (MB - m)) {
            const int j = idx / (MB - m), i = m + idx - j * (MB - m);
            Kmat[j * MB + i] = 0.f;
        } else {
            Avec[m + idx - m * (MB - m)] = 0.f;
        }
    }
    __syncthreads();

    if (!a.from_dist && rank == 0 && warp == 0) {
        float c = 0.f, it = 0.f;
        for (int idx = lane; idx < m * m; idx += 32) it += vals[idx];
        for (int i = lane; i < m; i += 32) c += vals[m * m + i];
        c = warp_sum(c);
        it = 2.0f * warp_sum(it);
        finish_row(p, b, c, it, W, lane);
    }
    if (p.grad_xhat == nullptr) return;

    // ---------------- pass 2: gradient rows ----------------
    T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)b * m * p.D;
    T* __restrict__ g0row = p.grad_x0 ? static_cast<T*>(p.grad_x0) + (long)b * p.D : nullptr;
    const int nib = MB / 4;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * a.chunk_cols;
        const int cols = min(a.chunk_cols, slab - c0);
        if (nchunks > 1 || a.from_dist) {
            __syncthreads();
            stage_chunk<T>(p, a, b, slab0 + c0, cols, tile, row_stride, bar, parity);
        }
        const int nv = cols / VEC;
        const unsigned char* x0row = tile + (size_t)m * row_stride;
        for (int item = tid; item < nv * nib; item += blockDim.x) {
            const int ib = item / nv, v = item - ib * nv;
            const int i0 = ib * 4;
            float x0v[VEC], xi[4][VEC], g[4][VEC];
            lds_pack<T, VEC>(x0row, v, x0v);
            const float4 a4 = *reinterpret_cast<const float4*>(Avec + i0);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                lds_pack<T, VEC>(tile + (size_t)min(i0 + s, m - 1) * row_stride, v, xi[s]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) g[s][e] = av[s] * (xi[s][e] - x0v[e]);
            }
            if (g0row != nullptr) {
                // d/dx0 of the confinement term: -sum_i A_i (x_i - x0); each i-block owns its share
                float g0[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) g0[e] = -(g[0][e] + g[1][e] + g[2][e] + g[3][e]);
                // accumulate across i-blocks in registers is impossible (different threads): use a
                // deterministic two-step — block 0 writes, later blocks are handled below.
                if (nib == 1) store_pack<T, VEC>(g0row, slab0 + c0 + (long)v * VEC, g0);
            }
            for (int j = 0; j < m; ++j) {
                float xj[VEC];
                lds_pack<T, VEC>(tile + (size_t)j * row_stride, v, xj);
                const float4 k4 = *reinterpret_cast<const float4*>(Kmat + j * MB + i0);
                const float kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                for (int s = 0; s < 4; ++s)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) g[s][e] = fmaf(kv[s], xi[s][e] - xj[e], g[s][e]);
            }
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (i0 + s < m) store_pack<T, VEC>(grow + (long)(i0 + s) * p.D, slab0 + c0 + (long)v * VEC, g[s]);
        }
        if (g0row != nullptr && nib > 1) {
            // grad_x0 for m > 4: one thread per column vector sums over all draws in a fixed order
            for (int v = tid; v < nv; v += blockDim.x) {
                float x0v[VEC], g0[VEC];
                lds_pack<T, VEC>(x0row, v, x0v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) g0[e] = 0.f;
                for (int i = 0; i < m; ++i) {
                    float xi1[VEC];
                    lds_pack<T, VEC>(tile + (size_t)i * row_stride, v, xi1);
                    const float ai = Avec[i];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) g0[e] = fmaf(-ai, xi1[e] - x0v[e], g0[e]);
                }
                store_pack<T, VEC>(g0row, slab0 + c0 + (long)v * VEC, g0);
            }
        }
    }
}

}  // namespace dddm
