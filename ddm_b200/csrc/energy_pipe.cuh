// energy_pipe.cuh — row-pipelined cluster kernel for m <= 8: the single-launch latency path.
//
// One launch of B = 128 rows gives the one-CTA-per-row kernel (energy_smem.cuh) exactly one wave: every SM loads its
// row (HBM busy, fp32 pipe idle), then every SM runs pass 2 (fp32 pipe busy, HBM idle) — the two resources are never
// used together (profiles/r02_k1_single_launch.md).  Pass 2 of a row needs the row's complete distances, so the only
// way to overlap the two inside ONE launch is to give an SM pieces of SEVERAL rows:
//
//   a cluster of C CTAs owns C consecutive rows; CTA k holds column slab k (D / C columns) of every one of them.
//   Row r's slabs are requested first, row r+1's next, ...; as soon as row r has landed the C CTAs take its partial
//   distances, send them to every peer (st.async into distributed shared memory, completion counted on the peer's
//   mbarrier — no cluster-wide barrier on the data path), go on with pass 1 of row r+1 and then run pass 2 of row r
//   on their slab while rows r+2.. are still in flight.  Only the last row's coefficient evaluation and 1/C of one
//   pass 2 are left after the last byte has arrived.
//
// Arithmetic is the packed-fp32 direct-difference form of energy_smem.cuh; the cross-CTA sums are taken in rank
// order (deterministic).  Reference: dddm/losses.py:5-25, dddm/training.py:84-85.
#pragma once

#include <cooperative_groups.h>

#include <type_traits>

#include "energy_smem.cuh"

namespace dddm {

constexpr int kPipeMaxThreads = 256;  // compute threads; a control warp is added at launch
constexpr int kPipeMaxRows = 8;       // rows per cluster == CTAs per cluster

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
// 4-byte store into a peer CTA's shared memory; the peer's mbarrier counts the bytes.
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_LOOP:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAITC_DONE;\n"
        "bra WAITC_LOOP;\n"
        "WAITC_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

#define DDDM_PTRACE(slot)                                                                                     \
    do {                                                                                                      \
        if (p.trace != nullptr) p.trace[((long)blockIdx.y * C + rank) * 16 + (slot)] = globaltimer_ns();      \
    } while (0)

template <typename T, int M, int COLS, bool X0F32 = false>
__global__ void __launch_bounds__(kPipeMaxThreads + 32, 1)
energy_pipe_kernel(const EnergyParams p, const int slab_vecs, const int C, const int window) {
    namespace cg = cooperative_groups;
    constexpr int P = M * (M + 1) / 2;
    constexpr int VEC = Elem<T>::kVec;
    constexpr int X0S = X0F32 ? 2 : 1;
    using T0 = typename std::conditional<X0F32, float, T>::type;
    constexpr int U = Step<T, COLS>::kPerVec;
    constexpr int NP = Step<T, COLS>::kPairs;
    constexpr int NWMAX = kPipeMaxThreads / 32;
    using WR = WarpReduce<P>;
    __shared__ __align__(8) uint64_t s_full[kPipeMaxRows];  // row r's slab has landed (TMA byte count)
    __shared__ __align__(8) uint64_t s_xbar[kPipeMaxRows];  // row r's partial distances from all C CTAs have landed
    __shared__ float s_part[kPipeMaxRows][kPipeMaxRows][P];  // [row][source rank][distance]
    __shared__ float s_warp[2][NWMAX][P];
    __shared__ float s_coef[2][P];
    extern __shared__ __align__(128) unsigned char s_tile[];  // [row][M + X0S tile rows][slab_vecs * 16 bytes]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x >> 5) - 1;
    const int nthr = nwarps * 32;
    const bool control = warp == nwarps;
    const int rank = (int)cg::this_cluster().block_rank();
    const int b0 = blockIdx.y * C;
    const int R = min(C, p.B - b0);  // rows of this cluster (cluster-uniform)
    if (tid == 0) DDDM_PTRACE(0);

    const long nvec = p.D / VEC;
    const long v_begin = (long)rank * slab_vecs;
    const int nv = (int)max(0L, min((long)slab_vecs, nvec - v_begin));
    const int nq = nv * U;
    const int row_bytes = slab_vecs * 16;
    const size_t rtile_bytes = (size_t)(M + X0S) * row_bytes;  // one row's (M draws + x0) slab tile

    if (control && lane == 0) {
        for (int r = 0; r < R; ++r) {
            mbar_init(&s_full[r], 1);
            mbar_init(&s_xbar[r], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int r = 0; r < R; ++r) {
            mbar_expect_tx(&s_full[r], (uint32_t)nv * 16u * (uint32_t)(M + X0S));
            mbar_expect_tx(&s_xbar[r], (uint32_t)C * P * 4u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    cluster_arrive_release();  // "my barriers are armed": peers wait for this before their first st.async
    cudaGridDependencySynchronize();
    if (tid == 0) DDDM_PTRACE(1);

    // ---- TMA issue: row-major, dealt to lane 0 of `nw_issue` warps so that the requests of a row leave together ----
    auto issue_row = [&](int r, int w, int nw_issue) {
        const uint32_t bytes = (uint32_t)nv * 16u;
        if (bytes == 0) return;
        unsigned char* dst = s_tile + (size_t)r * rtile_bytes;
        for (int t = w; t <= M; t += nw_issue) {
            if (t < M) {
                const T* src = static_cast<const T*>(p.xhat) + ((long)(b0 + r) * M + t) * p.D + v_begin * VEC;
                tma_bulk_g2s(dst + (size_t)t * row_bytes, src, bytes, &s_full[r]);
            } else {
                const T0* src = static_cast<const T0*>(p.x0) + (long)(b0 + r) * p.D + v_begin * VEC;
                tma_bulk_g2s(dst + (size_t)M * row_bytes, src, bytes * X0S, &s_full[r]);
            }
        }
    };
    const int win = min(max(window, 1), R);
    if (lane == 0) {
        for (int r = 0; r < win; ++r) issue_row(r, warp, nwarps + 1);
        if (control) DDDM_PTRACE(6);
    }
    __syncwarp();
    const float W = (p.mode == kModeLoss) ? p.weight_dev[0] * p.weight_scale : 1.0f;
    cudaTriggerProgrammaticLaunchCompletion();
    const float nb = (float)p.B * (float)M;
    const float pre_conf = 2.0f * W / nb;
    const float pre_pair = -4.0f * W * (p.lam / (2.0f * (float)(M - 1))) / (nb * (float)(M - 1));
    cluster_wait_acquire();  // every peer's barriers are armed

    // ---- control warp: row sums of the rows this CTA publishes (row r -> rank r mod C), concurrently with the rest ----
    if (control) {
        for (int r = rank; r < R; r += C) {
            mbar_wait_cluster(&s_xbar[r], 0);
            float c = 0.f, it = 0.f;
            for (int s = lane; s < P; s += 32) {
                float total = 0.f;
                for (int k = 0; k < C; ++k) total += s_part[r][k][s];
                if (p.dist != nullptr) p.dist[(long)(b0 + r) * P + s] = total;
                const float val = pow_value(total, p.pw);
                if (s < M) c += val; else it += val;
            }
            c = warp_sum(c);
            it = 2.0f * warp_sum(it);
            finish_row(p, b0 + r, c, it, W, lane);
        }
        if (lane == 0) DDDM_PTRACE(7);
        return;
    }

    const uint32_t part_base = smem_u32(&s_part[0][0][0]);
    const uint32_t xbar_base = smem_u32(&s_xbar[0]);

    // pass 1 of row r + publication of its partial distances
    auto pass1 = [&](int r) {
        const unsigned char* tile = s_tile + (size_t)r * rtile_bytes;
        mbar_wait(&s_full[r], 0);
        if (lane == 0 && r + win < R) issue_row(r + win, warp, nwarps);  // bounded window: next row's requests
        if (tid == 0 && r == 0) DDDM_PTRACE(2);
        if (tid == 0 && r == R - 1) DDDM_PTRACE(8);
        float2 acc2[P];
#pragma unroll
        for (int s = 0; s < P; ++s) acc2[s] = make_float2(0.f, 0.f);
        for (int q = tid; q < nq; q += nthr) {
            float2 x[M + 1][NP];
#pragma unroll
            for (int t = 0; t < M; ++t) lds_step<T, COLS>(tile + (size_t)t * row_bytes, q, x[t]);
            lds_step<T0, COLS>(tile + (size_t)M * row_bytes, q, x[M]);
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
        for (int s = 0; s < WR::kPadded; ++s) acc[s] = (s < P) ? acc2[s < P ? s : 0].x + acc2[s < P ? s : 0].y : 0.f;
        WR::run(acc, s_warp[r & 1][warp], lane);
        bar_sync_named(1, nthr);
        if (tid < P) {
            float t = 0.f;
            for (int w = 0; w < nwarps; ++w) t += s_warp[r & 1][w][tid];
            const uint32_t off = (uint32_t)(((r * kPipeMaxRows + rank) * P + tid) * 4);
            for (int k = 0; k < C; ++k)
                st_async_f32(mapa_u32(part_base + off, k), t, mapa_u32(xbar_base + (uint32_t)r * 8u, k));
        }
        if (tid == 0 && r == 0) DDDM_PTRACE(3);
        if (tid == 0 && r == R - 1) DDDM_PTRACE(9);
    };

    // coefficients + pass 2 of row r on this CTA's slab
    auto pass2 = [&](int r) {
        if (tid < P) {
            mbar_wait_cluster(&s_xbar[r], 0);
            float total = 0.f;
            for (int k = 0; k < C; ++k) total += s_part[r][k][tid];
            float val, der;
            pow_value_deriv(total, p.pw, val, der);
            s_coef[r & 1][tid] = ((tid < M) ? pre_conf : pre_pair) * der;
        }
        bar_sync_named(1, nthr);
        if (tid == 0 && r == 0) DDDM_PTRACE(4);
        if (tid == 0 && r == R - 1) DDDM_PTRACE(10);
        if (p.grad_xhat == nullptr || nq == 0) return;
        const unsigned char* tile = s_tile + (size_t)r * rtile_bytes;
        float2 K2[P];
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const float k = s_coef[r & 1][s];
            K2[s] = make_float2(k, k);
        }
        T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)(b0 + r) * M * p.D + v_begin * VEC;
        for (int q = tid; q < nq; q += nthr) {
            float2 x[M + 1][NP], g[M][NP];
#pragma unroll
            for (int t = 0; t < M; ++t) lds_step<T, COLS>(tile + (size_t)t * row_bytes, q, x[t]);
            lds_step<T0, COLS>(tile + (size_t)M * row_bytes, q, x[M]);
#pragma unroll
            for (int h = 0; h < NP; ++h) {
#pragma unroll
                for (int i = 0; i < M; ++i) g[i][h] = __fmul2_rn(K2[i], sub2(x[i][h], x[M][h]));
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int j = i + 1; j < M; ++j) {
                        const float2 d = sub2(x[i][h], x[j][h]);
                        const float2 k = K2[pair_slot<M>(i, j)];
                        g[i][h] = __ffma2_rn(k, d, g[i][h]);
                        g[j][h] = __ffma2_rn(make_float2(-k.x, -k.y), d, g[j][h]);
                    }
            }
#pragma unroll
            for (int i = 0; i < M; ++i) stg_step<T, COLS>(grow + (long)i * p.D + (long)q * COLS, g[i]);
        }
    };

    // Software pipeline over the rows.  After pass 1 of row r the exchange of its partial distances is in flight;
    // it is consumed one step later, so its latency hides behind pass 1 of row r+1.  If row r+1 has not landed yet
    // but row r's distances are complete, pass 2 of row r goes first (decided by thread 0, CTA-uniform).
    __shared__ int s_order;
    int done2 = 0;  // rows whose pass 2 has run
    for (int r = 0; r < R; ++r) {
        if (done2 < r) {
            if (tid == 0) s_order = (!mbar_test(&s_full[r], 0) && mbar_test(&s_xbar[done2], 0)) ? 1 : 0;
            bar_sync_named(1, nthr);
            const int early = s_order;
            bar_sync_named(1, nthr);
            if (early) pass2(done2++);
        }
        pass1(r);
        if (done2 < r) pass2(done2++);
    }
    while (done2 < R) pass2(done2++);
    if (tid == 0) DDDM_PTRACE(5);
}

}  // namespace dddm
