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

#include <type_traits>

#include "energy.cuh"
#include "energy_smem_plan.h"
#include "energy_tile.cuh"  // mbarrier / TMA bulk helpers

namespace dddm {

constexpr int kSmemMaxThreads = 128;  // compute threads per CTA (one warp per SM sub-partition); a control warp is added at launch
constexpr int kSmemMaxCluster = 8;
constexpr int kSmemMaxChunks = 8;  // column chunks of the tile, each with its own mbarrier

__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
constexpr float kCentredTau = 1.0f / 16.0f;  // rows with a pair closer than this (relative to the draws' distances to x0) go direct

// The unit of work of a thread is a "step" of COLS consecutive columns = COLS/2 packed-fp32 pairs.
// COLS = 4 keeps instruction count lowest (one LDS.128 per row for fp32); COLS = 2 halves the live
// registers so that more CTAs fit an SM (used for bf16, whose small tile leaves registers as the limit).
template <typename T, int COLS>
struct Step {
    static_assert(COLS == 2 || COLS == 4, "2 or 4 columns per step");
    static constexpr int kPairs = COLS / 2;
    static constexpr int kBytes = COLS * (int)sizeof(T);
    static constexpr int kPerVec = 16 / kBytes;  // steps per 16-byte vector
};
template <typename T, int COLS>
__device__ __forceinline__ void lds_step(const unsigned char* row, int q, float2 (&out)[COLS / 2]) {
    if constexpr (sizeof(T) == 4 && COLS == 4) {
        const float4 r = reinterpret_cast<const float4*>(row)[q];
        out[0] = make_float2(r.x, r.y);
        out[1] = make_float2(r.z, r.w);
    } else if constexpr (sizeof(T) == 4) {
        out[0] = reinterpret_cast<const float2*>(row)[q];
    } else if constexpr (COLS == 4) {
        const uint2 r = reinterpret_cast<const uint2*>(row)[q];
        out[0] = make_float2(bf16lo(r.x), bf16hi(r.x));
        out[1] = make_float2(bf16lo(r.y), bf16hi(r.y));
    } else {
        const uint32_t r = reinterpret_cast<const uint32_t*>(row)[q];
        out[0] = make_float2(bf16lo(r), bf16hi(r));
    }
}
template <typename T, int COLS>
__device__ __forceinline__ void stg_step(T* __restrict__ dst, const float2 (&g)[COLS / 2]) {
    if constexpr (sizeof(T) == 4 && COLS == 4) {
        uint4 r;
        r.x = __float_as_uint(g[0].x); r.y = __float_as_uint(g[0].y);
        r.z = __float_as_uint(g[1].x); r.w = __float_as_uint(g[1].y);
        stg_stream16(dst, r);
    } else if constexpr (sizeof(T) == 4) {
        stg_stream8(dst, make_uint2(__float_as_uint(g[0].x), __float_as_uint(g[0].y)));
    } else if constexpr (COLS == 4) {
        stg_stream8(dst, make_uint2(pack_bf16x2(g[0].x, g[0].y), pack_bf16x2(g[1].x, g[1].y)));
    } else {
        stg_stream4(dst, pack_bf16x2(g[0].x, g[0].y));
    }
}

template <int M>
__host__ __device__ constexpr int pair_slot(int i, int j) {  // i < j
    return M + i * M - i * (i + 1) / 2 + (j - i - 1);
}

// Optional in-kernel timeline (diagnostics; tools/trace_energy.py): 16 globaltimer stamps per CTA.
#define DDDM_TRACE(slot)                                                                              \
    do {                                                                                              \
        if (p.trace != nullptr) p.trace[((long)b * cluster_size + rank) * 16 + (slot)] = globaltimer_ns(); \
    } while (0)

// cp.async (LDGSTS) of one thread step; completion is tracked per thread in commit groups.
template <int BYTES>
__device__ __forceinline__ void cp_async_step(void* dst, const void* src) {
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of this thread's most recent commit groups are still in flight
__device__ __forceinline__ void cp_async_wait_pending(int pending) {
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}

// LOADER 0: the tile is staged by 1-D TMA bulk copies, one per tile row and column chunk, completion per chunk on an
//           mbarrier.  Costs no issue slots: the throughput path (two CTAs of different launches per SM).
// LOADER 1: every compute thread copies exactly the steps it will read itself with cp.async, one commit group per
//           column chunk.  On a single launch the bulk copies of a CTA all complete together at the END of the load
//           phase (measured: first chunk after 2.6 us of 3.7), so pass 1 could not start before the whole tile had
//           landed; commit groups complete in issue order, so pass 1 follows the loads chunk by chunk and hides inside
//           the HBM-bound load phase.  The slots are thread-private: no barrier between the copy and its use.
// X0F32 (bf16 draws only): x0 stays fp32 in memory and in the tile — the mixed entry point a bf16 backbone uses
//           (bf16 xhat in, fp32 data, bf16 gradient out), so that neither xhat is up-converted nor x0 rounded.
// NW: compute warps the instantiation is built for (4 = one per SM sub-partition, two CTAs of different launches per SM;
// 8 = one CTA per SM, for launches whose rows fit one wave: pass 1 runs in the shadow of the load phase only while data
// keeps arriving, and what is left of it after the last chunk has landed is halved).
template <typename T, int M, int COLS, int MIN_CTAS, bool BWD = false, int LOADER = 0, bool X0F32 = false, int NW = 4>
__global__ void __launch_bounds__(NW * 32 + 32, MIN_CTAS)
energy_fused_smem_kernel(const EnergyParams p, const int slab_vecs, const int cluster_size, const int chunk_vecs) {
    namespace cg = cooperative_groups;
    constexpr int P = M * (M + 1) / 2;
    constexpr int VEC = Elem<T>::kVec;
    static_assert(!X0F32 || (sizeof(T) == 2 && !BWD), "mixed entry: bf16 draws, fused loss only");
    constexpr int X0S = X0F32 ? 2 : 1;  // bytes of the x0 tile row relative to a draw row
    using T0 = typename std::conditional<X0F32, float, T>::type;
    constexpr int U = Step<T, COLS>::kPerVec;
    constexpr int NP = Step<T, COLS>::kPairs;
    using WR = WarpReduce<P>;
    __shared__ __align__(8) uint64_t s_bar[kSmemMaxChunks];
    __shared__ float s_warp[NW][P];
    __shared__ float s_cluster[kSmemMaxCluster][P];
    __shared__ float s_coef[P];
    __shared__ float s_val[P];
    __shared__ unsigned char s_pi[P], s_pj[P];  // pair slot -> (i, j), filled in the prologue (before the inputs are waited for)
    __shared__ int s_close;    // set when some pair of draws is much closer to each other than to x0 (direct pass 2)
    extern __shared__ __align__(128) unsigned char s_tile[];  // (M+1) rows x slab_vecs x 16 bytes

    // warps 0 .. nwarps-1 compute; the last warp is the control warp (TMA issue, cross-row reduction)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x >> 5) - 1;
    const int nthr = nwarps * 32;
    const bool control = warp == nwarps;
    const int rank = (cluster_size > 1) ? (int)cg::this_cluster().block_rank() : (int)blockIdx.x;
    const int b = blockIdx.y;
    constexpr bool bwd = BWD;  // kModeBwd, its own instantiation; launched without a cluster (independent D-slabs)
    if (tid == 0) DDDM_TRACE(0);
    if (cluster_size > 1) cluster_arrive_relaxed();  // phase 0: "my shared memory exists"

    const long nvec = p.D / VEC;
    const long v_begin = (long)rank * slab_vecs;
    const int nv = (int)max(0L, min((long)slab_vecs, nvec - v_begin));  // 16-byte vectors in this CTA's slab
    const int nq = nv * U;                                              // steps (COLS columns each) in this CTA's slab
    const int row_bytes = slab_vecs * 16;

    // The slab is staged in column chunks of chunk_vecs vectors (a multiple of the compute-thread
    // count); chunk c signals s_bar[c], so pass 1 starts on chunk 0 while the rest is still in flight.
    const int nchunks = (nv + chunk_vecs - 1) / chunk_vecs;
    const int chunk_q = chunk_vecs * U;
    if (tid == 0) s_close = 0;
    if (!BWD && tid >= M && tid < P) {
        int i = 0, r = tid - M;
        while (r >= M - 1 - i) {
            r -= M - 1 - i;
            ++i;
        }
        s_pi[tid] = (unsigned char)i;
        s_pj[tid] = (unsigned char)(i + 1 + r);
    }
    if (control) {  // rows of absent compute warps stay zero: the cross-warp sum always adds all 4 rows
        for (int w = nwarps; w < NW; ++w)
            for (int s = lane; s < P; s += 32) s_warp[w][s] = 0.f;
    }
    if (LOADER == 0 && control && lane == 0) {
        for (int c = 0; c < nchunks; ++c) {
            mbar_init(&s_bar[c], 1);
            mbar_expect_tx(&s_bar[c], (uint32_t)min(chunk_vecs, nv - c * chunk_vecs) * 16u * (uint32_t)(M + X0S));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the barriers visible to the TMA (async proxy)
    }
    __syncthreads();
    cudaGridDependencySynchronize();  // PDL: inputs may be produced by the previous kernel in the stream
    if (tid == 0) DDDM_TRACE(1);
    // One bulk copy costs its issuing thread ~45 ns (tools/trace_energy.py), so the (M+1) x nchunks copies are dealt
    // to ALL warps (row r -> warp r mod nwarps+1, chunk-major): the tile is requested 5x sooner than from one warp.
    // cp.async loader: at most `window` column chunks of the CTA are in flight.  All 128 rows of a launch start
    // together; with every request of every SM queued at once the memory system serves them in no particular order
    // and the FIRST chunk lands after 60-100 % of the load phase (measured) — with a bounded window the queues stay
    // short, chunks land in order and pass 1 follows them.
    constexpr int SB = Step<T, COLS>::kBytes;
    const unsigned char* xsrc = reinterpret_cast<const unsigned char*>(static_cast<const T*>(p.xhat) + (long)b * M * p.D + v_begin * VEC);
    const unsigned char* csrc = reinterpret_cast<const unsigned char*>(static_cast<const T0*>(p.x0) + (long)b * p.D + v_begin * VEC);
    const size_t grow_bytes = (size_t)p.D * sizeof(T);
    const int window = min(max(p.window, 1), 8);
    auto issue_chunk = [&](int c) {  // every thread commits one group per call (an empty one past the last chunk)
        if (c < nchunks) {
            const int q_end = min(nq, (c + 1) * chunk_q);
            for (int q = c * chunk_q + tid; q < q_end; q += nthr) {
#pragma unroll
                for (int r = 0; r < M; ++r)
                    cp_async_step<SB>(s_tile + (size_t)r * row_bytes + (size_t)q * SB, xsrc + (size_t)r * grow_bytes + (size_t)q * SB);
                cp_async_step<SB * X0S>(s_tile + (size_t)M * row_bytes + (size_t)q * SB * X0S, csrc + (size_t)q * SB * X0S);
            }
        }
        cp_async_commit();
    };
    if constexpr (LOADER == 1) {
        if (!control) {
            for (int c = 0; c < window; ++c) issue_chunk(c);
        }
        if (control && lane == 0) DDDM_TRACE(6);
    } else if (lane == 0) {
        for (int c = 0; c < nchunks; ++c) {
            const int c0 = c * chunk_vecs;
            const uint32_t bytes = (uint32_t)min(chunk_vecs, nv - c0) * 16u;
            for (int r = warp; r <= M; r += nwarps + 1) {
                if (r < M) {
                    const T* src = static_cast<const T*>(p.xhat) + ((long)b * M + r) * p.D + v_begin * VEC;
                    tma_bulk_g2s(s_tile + (size_t)r * row_bytes + (size_t)c0 * 16, src + (long)c0 * VEC, bytes, &s_bar[c]);
                } else {
                    const T0* src = static_cast<const T0*>(p.x0) + (long)b * p.D + v_begin * VEC;
                    tma_bulk_g2s(s_tile + (size_t)M * row_bytes + (size_t)c0 * 16 * X0S, src + (long)c0 * VEC, bytes * X0S,
                                 &s_bar[c]);
                }
            }
        }
        if (control) DDDM_TRACE(6);
    }
    __syncwarp();
    const float W = (p.mode == kModeLoss) ? p.weight_dev[0] * p.weight_scale : 1.0f;
    cudaTriggerProgrammaticLaunchCompletion();
    // gradient prefactors (independent of the distances: computed while the tile is in flight);
    // kModeBwd: upstream gradients of conf / inter are device scalars (SURVEY.md §8a closed form)
    const float nb = (float)p.B * (float)M;
    const float pre_conf = bwd ? 2.0f * p.g_conf[0] / nb : 2.0f * W / nb;
    const float pre_pair = bwd ? 4.0f * p.g_inter[0] / (nb * (float)(M - 1))
                               : -4.0f * W * (p.lam / (2.0f * (float)(M - 1))) / (nb * (float)(M - 1));

    // ---- pass 1: squared distances, COLS columns per thread per step.  Two forms:
    //   DIRECT   (x_i - x_j)^2 summed: 36 differences + 36 FMAs = 72 packed operations per column pair;
    //   CENTRED  with z_i = x_i - x0 (exact): the 36 inner products z_i . z_j (8 differences + 36 FMAs = 44), and
    //            d2_ij = |z_i|^2 + |z_j|^2 - 2 z_i . z_j afterwards.  Centred on x0 this cancels only when two draws are much
    //            closer to each other than to the data; such rows (d2_ij < 1/16 (d2_i0 + d2_j0): error amplification <= 16,
    //            2e-6 in fp32) are detected from the result and REDONE in the direct form from the tile in shared memory.
    constexpr bool kCentredCapable = !BWD && LOADER == 0;
    const bool centred1 = kCentredCapable && cluster_size == 1 && nthr >= 64;
    float2 acc2[P];
    auto pass1 = [&](auto centred_tag) {
        constexpr bool CENTRED = decltype(centred_tag)::value;
#pragma unroll
        for (int s = 0; s < P; ++s) acc2[s] = make_float2(0.f, 0.f);
        for (int q = (control || bwd) ? nq : tid; q < nq; q += nthr) {
            if ((q - tid) % chunk_q == 0) {  // warp-uniform: entering a new chunk
                if constexpr (LOADER == 1) {
                    cp_async_wait_pending(window - 1);  // chunk c has landed (window groups were committed after c - 1)
                    issue_chunk((q - tid) / chunk_q + window);
                } else {
                    mbar_wait(&s_bar[(q - tid) / chunk_q], 0);
                }
                if (tid == 0 && q == 0) DDDM_TRACE(2);
            }
            float2 x[M + 1][NP];
#pragma unroll
            for (int r = 0; r < M; ++r) lds_step<T, COLS>(s_tile + (size_t)r * row_bytes, q, x[r]);
            lds_step<T0, COLS>(s_tile + (size_t)M * row_bytes, q, x[M]);
#pragma unroll
            for (int h = 0; h < NP; ++h) {
                if constexpr (CENTRED) {
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        x[i][h] = sub2(x[i][h], x[M][h]);  // z_i
                        acc2[i] = __ffma2_rn(x[i][h], x[i][h], acc2[i]);
                    }
#pragma unroll
                    for (int i = 0; i < M; ++i)
#pragma unroll
                        for (int j = i + 1; j < M; ++j)
                            acc2[pair_slot<M>(i, j)] = __ffma2_rn(x[i][h], x[j][h], acc2[pair_slot<M>(i, j)]);
                } else {
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
        }
    };
    auto reduce_to_smem = [&]() {  // per-lane partial sums -> the warp's row of s_warp
        float acc[WR::kPadded];
#pragma unroll
        for (int s = 0; s < WR::kPadded; ++s) acc[s] = (s < P) ? acc2[s < P ? s : 0].x + acc2[s < P ? s : 0].y : 0.f;
        if (!control && !bwd) WR::run(acc, s_warp[warp], lane);
    };
    if constexpr (kCentredCapable) {
        if (centred1) pass1(std::true_type{}); else pass1(std::false_type{});
    } else {
        pass1(std::false_type{});
    }
    if (tid == 0) DDDM_TRACE(3);
    if (tid == nthr - 32) DDDM_TRACE(11);
    reduce_to_smem();
    if (tid == 0) DDDM_TRACE(8);
    __syncthreads();
    if (tid == 0) DDDM_TRACE(9);

    // ---- cross-warp and cross-CTA sums, fixed order; one thread per distance ----
    static_assert(NW == 4 || NW == 8, "fixed-shape cross-warp sum");
    auto rows_sum = [&](int s) {  // the warps' partial sums of slot s, in a fixed tree
        float t = (s_warp[0][s] + s_warp[1][s]) + (s_warp[2][s] + s_warp[3][s]);
        if constexpr (NW == 8) t += (s_warp[4][s] + s_warp[5][s]) + (s_warp[6][s] + s_warp[7][s]);
        return t;
    };
    if (cluster_size > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster_wait_acquire();  // phase 0 complete: every CTA of the cluster is running
        if (tid < P) {
            const float t = rows_sum(tid);
            for (int r = 0; r < cluster_size; ++r) cluster.map_shared_rank(&s_cluster[0][0], r)[rank * P + tid] = t;
        }
        cluster_arrive_release();
        cluster_wait_acquire();
    }
    auto coef_step = [&](bool from_centred) {  // executed by the threads tid < P: one distance each
        const int s = tid;
        float total = 0.f;
        if (bwd) {
            total = p.dist[(long)b * P + s];
        } else if (cluster_size > 1) {
            for (int r = 0; r < cluster_size; ++r) total += s_cluster[r][s];
        } else {
            total = rows_sum(s);
        }
        if constexpr (!BWD) {
            // A pair's thread re-derives the confinement distances of its two draws from the warp partials (8 loads)
            // instead of waiting for their threads: they turn the inner product into a distance (centred pass 1) and
            // decide which forms the row takes (see pass 1 / pass 2).
            if (s >= M) {
                const int i = s_pi[s], j = s_pj[s];
                float di = 0.f, dj = 0.f;
                if (cluster_size > 1) {  // D-split rows: the same sums in every CTA of the cluster, hence the same verdict
                    for (int r = 0; r < cluster_size; ++r) {
                        di += s_cluster[r][i];
                        dj += s_cluster[r][j];
                    }
                } else {
                    di = rows_sum(i);
                    dj = rows_sum(j);
                }
                if (from_centred) total = fmaxf((di + dj) - 2.0f * total, 0.f);
                if (!(total >= kCentredTau * (di + dj))) s_close = 1;
            }
        }
        float val, der;
        pow_value_deriv(total, p.pw, val, der);
        s_val[s] = val;
        s_coef[s] = ((s < M) ? pre_conf : pre_pair) * der;
        if (!bwd && p.dist != nullptr && rank == 0) p.dist[(long)b * P + s] = total;
    };
    if (tid < P) coef_step(centred1);
    if (tid == 0) DDDM_TRACE(10);
    __syncthreads();
    if constexpr (kCentredCapable) {
        if (centred1 && s_close != 0) {  // CTA-uniform: a near-duplicate pair — redo the distances in the direct form
            if (!control) {
                pass1(std::false_type{});
                reduce_to_smem();
                bar_sync_named(1, nthr);
                if (tid < P) coef_step(false);  // nthr >= 64 > P: these are compute threads
            }
            bar_sync_named(2, nthr + 32);  // the control warp reads the row sums after this
        }
    }
    if (tid == 0) DDDM_TRACE(4);

    // ---- cross-row reduction: the control warp of the row's first CTA, concurrently with pass 2 ----
    if (control) {
        if (rank == 0 && !bwd) {
            // conf = sum of slots [0, M), inter = 2 * sum of slots [M, P) (ordered pairs), fixed-shape tree
            float c = (lane < M) ? s_val[lane] : 0.f;
            float it = (lane >= M && lane < P) ? s_val[lane] : 0.f;
            if (lane + 32 < P) it += s_val[lane + 32];
            static_assert(P <= 64, "two slots per lane");
            c = warp_sum(c);
            it = 2.0f * warp_sum(it);
            if (p.finish == 1) finish_row(p, b, c, it, W, lane); else finish_row_polled(p, b, c, it, W, lane);
            if (lane == 0) DDDM_TRACE(7);
        }
        return;
    }

    // ---- pass 2: gradient rows from the same tile.  Every pair difference x_i - x_j is formed once and
    //      feeds both rows (g_i += k d, g_j -= k d; the negation is an operand modifier of FFMA2). ----
    if (p.grad_xhat != nullptr && nq > 0) {
        float2 K2[P];
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const float k = s_coef[s];
            K2[s] = make_float2(k, k);
        }
        T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)b * M * p.D + v_begin * VEC;
        // grad_x0 exists only in the backward instantiation (the fused loss never differentiates w.r.t. the data)
        T* __restrict__ g0row = (BWD && p.grad_x0 != nullptr) ? static_cast<T*>(p.grad_x0) + (long)b * p.D + v_begin * VEC : nullptr;
        // Pass 2 has two forms.  CENTRED (the fast one): with z_i = x_i - x0,
        //     g_i = c_i z_i + sum_j k_ij (z_i - z_j) = (c_i + sum_j k_ij) z_i - sum_j k_ij z_j,
        // 8 differences + 8 products + 56 FMAs = 72 packed operations per column pair instead of 100.  It cancels against
        // |z| only (never against |x|: the differences to x0 are exact), i.e. it loses accuracy only where two draws are much
        // closer to each other than to the data — duplicates, the identical draws of a zero-initialised output layer, where
        // for beta < 1 the coefficient f' also explodes.  The threads that evaluate the coefficients compare every pair
        // distance with 1/16 (d2_i0 + d2_j0) and raise s_close; such rows take the DIRECT form below (every difference
        // x_i - x_j formed explicitly), as do the backward-only launches.
        const bool direct = BWD || s_close != 0;
        if (!direct) {
#pragma unroll
            for (int i = 0; i < M; ++i) {  // diagonal: c_i + sum_j k_ij (takes the slot of c_i)
                float a = K2[i].x;
#pragma unroll
                for (int j = 0; j < M; ++j)
                    if (j != i) a += K2[pair_slot<M>(i < j ? i : j, i < j ? j : i)].x;
                K2[i] = make_float2(a, a);
            }
            for (int q = tid; q < nq; q += nthr) {
                float2 x[M + 1][NP], g[M][NP];
#pragma unroll
                for (int r = 0; r < M; ++r) lds_step<T, COLS>(s_tile + (size_t)r * row_bytes, q, x[r]);
                lds_step<T0, COLS>(s_tile + (size_t)M * row_bytes, q, x[M]);
#pragma unroll
                for (int h = 0; h < NP; ++h) {
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        x[i][h] = sub2(x[i][h], x[M][h]);  // z_i
                        g[i][h] = __fmul2_rn(K2[i], x[i][h]);
                    }
#pragma unroll
                    for (int i = 0; i < M; ++i)
#pragma unroll
                        for (int j = i + 1; j < M; ++j) {
                            const float2 k = K2[pair_slot<M>(i, j)];
                            const float2 nk = make_float2(-k.x, -k.y);  // folded into FFMA2's operand modifier
                            g[i][h] = __ffma2_rn(nk, x[j][h], g[i][h]);
                            g[j][h] = __ffma2_rn(nk, x[i][h], g[j][h]);
                        }
                }
#pragma unroll
                for (int i = 0; i < M; ++i) stg_step<T, COLS>(grow + (long)i * p.D + (long)q * COLS, g[i]);
            }
        } else {
            for (int q = tid; q < nq; q += nthr) {
                if (bwd && (q - tid) % chunk_q == 0) {  // no pass 1 waited for it
                    mbar_wait(&s_bar[(q - tid) / chunk_q], 0);
                }
                float2 x[M + 1][NP], g[M][NP];
#pragma unroll
                for (int r = 0; r < M; ++r) lds_step<T, COLS>(s_tile + (size_t)r * row_bytes, q, x[r]);
                lds_step<T0, COLS>(s_tile + (size_t)M * row_bytes, q, x[M]);
#pragma unroll
                for (int h = 0; h < NP; ++h) {
#pragma unroll
                    for (int i = 0; i < M; ++i) g[i][h] = __fmul2_rn(K2[i], sub2(x[i][h], x[M][h]));
                    if (BWD && g0row != nullptr) {  // d/dx0 of the confinement term: -sum_i k_i (x_i - x0), fixed order
                        float2 s0 = g[0][h];
#pragma unroll
                        for (int i = 1; i < M; ++i) s0 = __fadd2_rn(s0, g[i][h]);
                        x[M][h] = make_float2(-s0.x, -s0.y);  // x0 is not needed below: reuse its registers
                    }
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
                if (BWD && g0row != nullptr) stg_step<T, COLS>(g0row + (long)q * COLS, x[M]);
            }
        }
    }
    if (tid == 0) DDDM_TRACE(5);
}

}  // namespace dddm
