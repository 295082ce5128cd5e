// energy_reg.cuh — register-resident energy-score kernels for m <= 8 (the headline shapes).
//
// One thread-block CLUSTER per minibatch row; the CTAs of a cluster split D.  Each thread pulls
// NV 16-byte vectors of every draw (and of x0) straight from HBM into registers with streaming
// loads — the m x D draw tile of a row lives in the register files of the cluster (256 KB per SM,
// larger than shared memory) and is read from HBM exactly once.  Pass 1 accumulates the
// m + m(m-1)/2 squared distances per thread, folds them with a butterfly reduce-scatter
// (P-1 shuffles instead of 5P), across warps through shared memory and across the cluster
// through distributed shared memory.  Every CTA then evaluates the beta-power coefficients
// redundantly and pass 2 streams dloss/dxhat out of the registers with 128-bit stores.
//
// Reference arithmetic: dddm/losses.py:5-25 (terms), dddm/training.py:84-85 (loss).
#pragma once

#include <cooperative_groups.h>

#include "energy.cuh"

namespace dddm {
namespace cg = cooperative_groups;

constexpr int kRegMaxThreads = 256;
constexpr int kRegMaxCluster = 8;


// Load the NV vectors of all M draws and of x0 owned by this thread.  Slots past the slab end
// are zero-filled so that they contribute nothing to any distance.
template <typename T, int M, int VEC, int NV>
__device__ __forceinline__ void load_tile_regs(const T* __restrict__ xrow, const T* __restrict__ crow, int D,
                                               long v_first, long v_end, int stride, float (&x)[NV][M + 1][VEC],
                                               bool (&ok)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const long v = v_first + (long)k * stride;
        ok[k] = v < v_end;
        if (ok[k]) {
#pragma unroll
            for (int i = 0; i < M; ++i) load_pack<T, VEC>(xrow + (long)i * D, v * VEC, x[k][i]);
            load_pack<T, VEC>(crow, v * VEC, x[k][M]);
        } else {
#pragma unroll
            for (int i = 0; i <= M; ++i)
#pragma unroll
                for (int e = 0; e < VEC; ++e) x[k][i][e] = 0.f;
        }
    }
}

// Pass 2: gradient rows from the register tile and the P coefficients (K[0..M) confinement,
// K[M..P) pairs (i<j) row-major), written with 16-byte stores.
template <typename T, int M, int VEC, int NV, bool WITH_X0>
__device__ __forceinline__ void store_grad_regs(T* __restrict__ grow, T* __restrict__ g0row, int D, long v_first,
                                                int stride, const float (&x)[NV][M + 1][VEC], const bool (&ok)[NV],
                                                const float* __restrict__ coef_smem) {
    constexpr int P = M * (M + 1) / 2;
    float K[P];
#pragma unroll
    for (int q = 0; q < P; ++q) K[q] = coef_smem[q];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (!ok[k]) continue;
        const long e0 = (v_first + (long)k * stride) * VEC;
        float g[M][VEC];
        float g0[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) g0[e] = 0.f;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                g[i][e] = K[i] * (x[k][i][e] - x[k][M][e]);
                if constexpr (WITH_X0) g0[e] -= g[i][e];
            }
        int q = M;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = i + 1; j < M; ++j, ++q)
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const float d = x[k][i][e] - x[k][j][e];
                    g[i][e] = fmaf(K[q], d, g[i][e]);
                    g[j][e] = fmaf(-K[q], d, g[j][e]);
                }
#pragma unroll
        for (int i = 0; i < M; ++i) store_pack<T, VEC>(grow + (long)i * D, e0, g[i]);
        if constexpr (WITH_X0) store_pack<T, VEC>(g0row, e0, g0);
    }
}

// ---- K1: fused forward (+ backward) ---------------------------------------------------------
template <typename T, int M, int VEC, int NV>
__global__ void __launch_bounds__(kRegMaxThreads, 2)
energy_fused_reg_kernel(const EnergyParams p, const int vec_per_cta, const int cluster_size) {
    constexpr int P = M * (M + 1) / 2;
    using WR = WarpReduce<P>;
    __shared__ float s_warp[kRegMaxThreads / 32][P];
    __shared__ float s_cluster[kRegMaxCluster][P];
    __shared__ float s_coef[P];
    __shared__ float s_val[P];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rank = (cluster_size > 1) ? (int)cg::this_cluster().block_rank() : 0;
    const int b = blockIdx.y;
    if (cluster_size > 1) cluster_arrive_relaxed();  // phase 0: "my shared memory exists"

    const T* __restrict__ xrow = static_cast<const T*>(p.xhat) + (long)b * M * p.D;
    const T* __restrict__ crow = static_cast<const T*>(p.x0) + (long)b * p.D;
    const long nvec = p.D / VEC;
    const long v_begin = (long)rank * vec_per_cta;
    const long v_end = min(v_begin + (long)vec_per_cta, nvec);

    cudaGridDependencySynchronize();  // PDL: inputs may come from the previous kernel in the stream
    float x[NV][M + 1][VEC];
    bool ok[NV];
    load_tile_regs<T, M, VEC, NV>(xrow, crow, p.D, v_begin + tid, v_end, blockDim.x, x, ok);
    const float W = (p.mode == kModeLoss) ? p.weight_dev[0] * p.weight_scale : 1.0f;
    cudaTriggerProgrammaticLaunchCompletion();

    // pass 1: per-thread partial squared distances
    float acc[WR::kPadded];
#pragma unroll
    for (int q = 0; q < WR::kPadded; ++q) acc[q] = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const float d = x[k][i][e] - x[k][M][e];
                acc[i] = fmaf(d, d, acc[i]);
            }
            int q = M;
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = i + 1; j < M; ++j, ++q) {
                    const float d = x[k][i][e] - x[k][j][e];
                    acc[q] = fmaf(d, d, acc[q]);
                }
        }
    WR::run(acc, s_warp[warp], lane);
    __syncthreads();

    // blockDim may be smaller than P (32 threads, 36 slots): every per-slot step strides over the slots
    if (cluster_size > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster_wait_acquire();  // phase 0 complete: every CTA of the cluster is running
        for (int q = tid; q < P; q += blockDim.x) {
            float t = 0.f;
            for (int w = 0; w < nwarps; ++w) t += s_warp[w][q];
            for (int r = 0; r < cluster_size; ++r)  // push my partial into every peer's table
                cluster.map_shared_rank(&s_cluster[0][0], r)[rank * P + q] = t;
        }
        cluster_arrive_release();
        cluster_wait_acquire();
    }
    for (int q = tid; q < P; q += blockDim.x) {
        float total = 0.f;
        if (cluster_size > 1) {
            for (int r = 0; r < cluster_size; ++r) total += s_cluster[r][q];  // fixed order: deterministic
        } else {
            for (int w = 0; w < nwarps; ++w) total += s_warp[w][q];
        }
        s_val[q] = pow_value(total, p.pw);
        if (p.grad_xhat != nullptr) {
            const float cl = p.lam / (2.0f * (float)(M - 1));
            s_coef[q] = (q < M) ? conf_coef(total, W, p) : pair_coef(total, -W * cl, p);
        }
        if (p.dist != nullptr && rank == 0) p.dist[(long)b * P + q] = total;
    }
    __syncthreads();

    if (p.grad_xhat != nullptr) {
        T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)b * M * p.D;
        store_grad_regs<T, M, VEC, NV, false>(grow, nullptr, p.D, v_begin + tid, blockDim.x, x, ok, s_coef);
    }

    if (rank == 0 && warp == 0) {
        float c = 0.f, it = 0.f;
        if (lane == 0) {
            for (int q = 0; q < M; ++q) c += s_val[q];
            for (int q = M; q < P; ++q) it += s_val[q];
            it *= 2.0f;  // ordered pairs (i,j) and (j,i)
        }
        finish_row(p, b, c, it, W, lane);
    }
}

// ---- K1b backward: gradient from saved distances ---------------------------------------------
template <typename T, int M, int VEC, int NV>
__global__ void __launch_bounds__(kRegMaxThreads)
energy_bwd_reg_kernel(const EnergyParams p) {
    constexpr int P = M * (M + 1) / 2;
    __shared__ float s_coef[P];
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const T* __restrict__ xrow = static_cast<const T*>(p.xhat) + (long)b * M * p.D;
    const T* __restrict__ crow = static_cast<const T*>(p.x0) + (long)b * p.D;
    const long nvec = p.D / VEC;
    const long v_begin = (long)blockIdx.x * blockDim.x * NV;

    cudaGridDependencySynchronize();
    float x[NV][M + 1][VEC];
    bool ok[NV];
    load_tile_regs<T, M, VEC, NV>(xrow, crow, p.D, v_begin + tid, nvec, blockDim.x, x, ok);
    for (int q = tid; q < P; q += blockDim.x) {
        const float d2 = p.dist[(long)b * P + q];
        s_coef[q] = (q < M) ? conf_coef(d2, p.g_conf[0], p) : pair_coef(d2, p.g_inter[0], p);
    }
    cudaTriggerProgrammaticLaunchCompletion();
    __syncthreads();
    T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)b * M * p.D;
    if (p.grad_x0 != nullptr) {
        T* __restrict__ g0row = static_cast<T*>(p.grad_x0) + (long)b * p.D;
        store_grad_regs<T, M, VEC, NV, true>(grow, g0row, p.D, v_begin + tid, blockDim.x, x, ok, s_coef);
    } else {
        store_grad_regs<T, M, VEC, NV, false>(grow, nullptr, p.D, v_begin + tid, blockDim.x, x, ok, s_coef);
    }
}

// ---- host-side dispatch ------------------------------------------------------------------
template <typename T, int M>
int launch_energy_reg_m(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream) {
    constexpr int V = Elem<T>::kVec;
    const long nvec = p.D / plan.vec;
    const int vec_per_cta = (int)((nvec + plan.cluster - 1) / plan.cluster);
    dim3 grid(plan.cluster, p.B), block(plan.threads);
    if (plan.vec == 1)
        return launch_with_attrs(energy_fused_reg_kernel<T, M, 1, 1>, grid, block, 0, plan.cluster, stream, p, vec_per_cta,
                                 plan.cluster);
    if (plan.nv == 1)
        return launch_with_attrs(energy_fused_reg_kernel<T, M, V, 1>, grid, block, 0, plan.cluster, stream, p, vec_per_cta,
                                 plan.cluster);
    if constexpr (sizeof(T) == 4) {  // two vectors per thread only pay off (and fit) for fp32
        return launch_with_attrs(energy_fused_reg_kernel<T, M, V, 2>, grid, block, 0, plan.cluster, stream, p,
                                 vec_per_cta, plan.cluster);
    }
    return DDDM_ERR_UNSUPPORTED;
}

template <typename T, int M>
int launch_energy_bwd_reg_m(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream) {
    constexpr int V = Elem<T>::kVec;
    const long nvec = p.D / plan.vec;
    const long per_cta = (long)plan.threads * plan.nv;
    dim3 grid((unsigned)((nvec + per_cta - 1) / per_cta), p.B), block(plan.threads);
    if (plan.vec == 1) return launch_with_attrs(energy_bwd_reg_kernel<T, M, 1, 1>, grid, block, 0, 1, stream, p);
    if (plan.nv == 1) return launch_with_attrs(energy_bwd_reg_kernel<T, M, V, 1>, grid, block, 0, 1, stream, p);
    if constexpr (sizeof(T) == 4) {
        return launch_with_attrs(energy_bwd_reg_kernel<T, M, V, 2>, grid, block, 0, 1, stream, p);
    }
    return DDDM_ERR_UNSUPPORTED;
}

#define DDDM_DISPATCH_M(FN, T, p, plan, stream)            \
    switch ((p).m) {                                       \
        case 2: return FN<T, 2>(p, plan, stream);          \
        case 3: return FN<T, 3>(p, plan, stream);          \
        case 4: return FN<T, 4>(p, plan, stream);          \
        case 5: return FN<T, 5>(p, plan, stream);          \
        case 6: return FN<T, 6>(p, plan, stream);          \
        case 7: return FN<T, 7>(p, plan, stream);          \
        case 8: return FN<T, 8>(p, plan, stream);          \
        default: return DDDM_ERR_UNSUPPORTED;              \
    }

}  // namespace dddm
