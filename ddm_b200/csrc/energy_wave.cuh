// energy_wave.cuh — single-wave, register-resident energy-score kernel for m <= 8 (K1, latency path).
//
// When the whole minibatch fits ONE wave of CTAs (B <= number of SMs; the headline shape B = 128 on 148 SMs)
// nothing can overlap across rows, so the launch is as long as one row's critical path:
//   inputs ready -> row in the SM -> pass 1 -> coefficients -> pass 2 -> last store.
// The TMA-staged kernel (energy_smem.cuh) is built for throughput (two CTAs per SM, rows of different
// launches overlapping); on a single launch its bulk copies all complete together at the END of the load phase
// (the copy engine interleaves them), so pass 1 starts when the whole tile has landed and the chain is strictly
// serial.  Here one CTA owns the SM:
//   * every compute thread issues its (m+1) x NV 16-byte streaming loads up front, vector-major, and keeps the
//     row in REGISTERS (a 110 KB row is 108 registers x 256 threads); the loads return in issue order, so pass 1
//     runs on vector k while vectors k+1.. are still in flight — it hides inside the HBM-bound load phase;
//   * 2-3 warps per scheduler (256/384 compute threads) instead of one: the packed-fp32 pipe is issued every
//     2.2 cycles instead of 2.6 (tools/ubench/fp32_pipes.cu), and there are no shared-memory reads in either pass;
//   * pass 2 forms the gradient rows from the same registers and streams them out with 16-byte stores;
//   * a control warp publishes the row sums and runs the deterministic cross-row reduction concurrently with
//     pass 2 (finish_row, energy.cuh).
// Arithmetic is identical to the TMA-staged kernel (direct differences, packed fp32, fixed summation shapes).
// Reference arithmetic: dddm/losses.py:5-25 (terms), dddm/training.py:84-85 (loss).
#pragma once

#include "energy.cuh"
#include "energy_smem.cuh"  // pair_slot, sub2, DDDM_TRACE

namespace dddm {

// Shared-memory read the compiler may neither hoist nor merge (pass 2 re-reads the coefficient pairs instead of
// pinning 72 registers).
__device__ __forceinline__ float2 lds64_volatile(const float2* p) {
    float2 r;
    asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(smem_u32(p)));
    return r;
}
// Opaque identity on a 16-byte register group: the compiler must treat the value as new.  Between the two passes it
// stops the pass-1 differences (36 register pairs per column pair) from being kept alive for pass 2.
__device__ __forceinline__ void opaque(uint4& r) { asm volatile("" : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)); }

// Column pair h of a 16-byte vector as two fp32 lanes.
template <typename T>
__device__ __forceinline__ float2 unpack_pair(const uint4& r, int h) {
    if constexpr (sizeof(T) == 4) {
        return h == 0 ? make_float2(__uint_as_float(r.x), __uint_as_float(r.y))
                      : make_float2(__uint_as_float(r.z), __uint_as_float(r.w));
    } else {
        const uint32_t w = h == 0 ? r.x : (h == 1 ? r.y : (h == 2 ? r.z : r.w));
        return make_float2(bf16lo(w), bf16hi(w));
    }
}

// finish_row (energy.cuh) cut into steps that never wait: the 8 compute warps fill the register file of the SM
// (a ninth warp would cost every thread a quarter of its registers: the file is split per scheduler), so warp 0
// interleaves the row publication with its pass-2 sections and consumes each atomic's result one section later.
struct RowTicket {
    unsigned long long prev;
    unsigned old;
    float2 part[4];
};
__device__ __forceinline__ void ticket_publish(RowTicket& t, const EnergyParams& p, int b, float conf_row, float inter_row) {
    const unsigned long long packed = (unsigned long long)(__float_as_uint(conf_row) & 0x7fffffffu) |
                                      ((unsigned long long)(__float_as_uint(inter_row) & 0x7fffffffu) << 32);
    asm volatile("atom.global.exch.b64 %0, [%1], %2;"
                 : "=l"(t.prev)
                 : "l"(reinterpret_cast<unsigned long long*>(p.row_partials) + b), "l"(packed)
                 : "memory");
}
__device__ __forceinline__ void ticket_take(RowTicket& t, const EnergyParams& p) {
    // consumes the exchange's return value: cannot be performed before the row's sums are in L2 (see finish_row)
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(t.old) : "l"(p.ticket), "r"(1u + (unsigned)(t.prev >> 63)) : "memory");
}
// warp-uniform: is this the last row?  If so, start reading every row's sums (4 rows per lane and round trip).
__device__ __forceinline__ bool ticket_is_last(RowTicket& t, const EnergyParams& p, int lane) {
    return __shfl_sync(0xffffffffu, t.old, 0) == (unsigned)(p.B - 1);
}
__device__ __forceinline__ void ticket_finalize(const EnergyParams& p, float W, int lane) {
    float c = 0.f, i = 0.f;
    for (int r0 = 0; r0 < p.B; r0 += 128) {  // fixed order: lane-strided, 4 independent loads in flight
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * 32 + lane;
            v[u] = (r < p.B) ? __ldcg(reinterpret_cast<const float2*>(p.row_partials) + r) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            c += v[u].x;
            i += v[u].y;
        }
    }
    c = warp_sum(c);
    i = warp_sum(i);
    if (lane == 0) {
        const float conf = c / ((float)p.B * (float)p.m);
        const float inter = i / ((float)p.B * (float)p.m * (float)(p.m - 1));
        if (p.mode == kModeLoss) {
            const float cl = p.lam / (2.0f * (float)(p.m - 1));
            p.out[0] = W * (conf - cl * inter);
            p.out[1] = conf;
            p.out[2] = inter;
            p.out[3] = W;
        } else {
            p.out[0] = conf;
            p.out[1] = inter;
        }
        *p.ticket = 0u;  // leave the workspace reusable
    }
}

template <typename T, int M, int NV, int THREADS, bool KSMEM>
__global__ void __launch_bounds__(THREADS, 1) energy_fused_wave_kernel(const EnergyParams p) {
    constexpr int P = M * (M + 1) / 2;
    constexpr int VEC = Elem<T>::kVec;
    constexpr int NH = VEC / 2;  // column pairs per vector
    constexpr int NW = THREADS / 32;
    using WR = WarpReduce<P>;
    __shared__ float s_warp[NW][P];
    __shared__ __align__(8) float2 s_coef2[P];  // (k, k): the packed operand of FFMA2
    __shared__ float s_val[P];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    constexpr int cluster_size = 1, rank = 0;  // for DDDM_TRACE
    if (tid == 0) DDDM_TRACE(0);

    const long nvec = p.D / VEC;
    const T* __restrict__ xrow = static_cast<const T*>(p.xhat) + (long)b * M * p.D;
    const T* __restrict__ crow = static_cast<const T*>(p.x0) + (long)b * p.D;

    cudaGridDependencySynchronize();  // PDL: the inputs may be produced by the previous kernel in the stream
    if (tid == 0) DDDM_TRACE(1);

    // ---- the row, vector-major: vector k of every tile row before vector k+1 (arrival order = use order) ----
    uint4 raw[NV][M + 1];
    bool ok[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const long v = tid + (long)k * THREADS;
        ok[k] = v < nvec;
        if (ok[k]) {
#pragma unroll
            for (int r = 0; r < M; ++r) raw[k][r] = ldg_stream16(xrow + (long)r * p.D + v * VEC);
            raw[k][M] = ldg_stream16(crow + v * VEC);
        } else {  // past the row end: zeros contribute nothing to any distance
#pragma unroll
            for (int r = 0; r <= M; ++r) raw[k][r] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (tid == 0) DDDM_TRACE(6);
    const float W = (p.mode == kModeLoss) ? p.weight_dev[0] * p.weight_scale : 1.0f;
    cudaTriggerProgrammaticLaunchCompletion();
    const float nb = (float)p.B * (float)M;
    const float pre_conf = 2.0f * W / nb;
    const float pre_pair = -4.0f * W * (p.lam / (2.0f * (float)(M - 1))) / (nb * (float)(M - 1));

    // ---- pass 1: squared distances, packed fp32, in arrival order ----
    {
        float2 acc2[P];
#pragma unroll
        for (int s = 0; s < P; ++s) acc2[s] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float2 x[M + 1];
#pragma unroll
                for (int r = 0; r <= M; ++r) x[r] = unpack_pair<T>(raw[k][r], h);
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const float2 d = sub2(x[i], x[M]);
                    acc2[i] = __ffma2_rn(d, d, acc2[i]);
                }
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int j = i + 1; j < M; ++j) {
                        const float2 d = sub2(x[i], x[j]);
                        acc2[pair_slot<M>(i, j)] = __ffma2_rn(d, d, acc2[pair_slot<M>(i, j)]);
                    }
            }
            if (k == 0 && tid == 0) DDDM_TRACE(2);
        }
        if (tid == 0) DDDM_TRACE(3);
        if (tid == THREADS - 32) DDDM_TRACE(11);
        float acc[WR::kPadded];
#pragma unroll
        for (int s = 0; s < WR::kPadded; ++s) acc[s] = (s < P) ? acc2[s < P ? s : 0].x + acc2[s < P ? s : 0].y : 0.f;
        WR::run(acc, s_warp[warp], lane);
    }
    if (tid == 0) DDDM_TRACE(8);
    __syncthreads();
    if (tid == 0) DDDM_TRACE(9);

    // ---- cross-warp sum (fixed-shape tree), beta-power and its derivative: one thread per distance ----
    if (tid < P) {
        const int s = tid;
        float part[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) part[w] = s_warp[w][s];
#pragma unroll
        for (int span = 1; span < NW; span *= 2)
#pragma unroll
            for (int w = 0; w + span < NW; w += 2 * span) part[w] += part[w + span];
        const float total = part[0];
        float val, der;
        pow_value_deriv(total, p.pw, val, der);
        s_val[s] = val;
        const float k = ((s < M) ? pre_conf : pre_pair) * der;
        s_coef2[s] = make_float2(k, k);
        if (p.dist != nullptr) p.dist[(long)b * P + s] = total;
    }
    if (tid == 0) DDDM_TRACE(10);
    __syncthreads();
    if (tid == 0) DDDM_TRACE(4);

    // ---- row sums: warp 0 publishes them and takes the ticket while it works through pass 2 ----
    RowTicket ticket;
    bool last_row = false;
    if (warp == 0) {
        // conf = sum of slots [0, M), inter = 2 * sum of slots [M, P) (ordered pairs), fixed-shape tree
        float c = (lane < M) ? s_val[lane] : 0.f;
        float it = (lane >= M && lane < P) ? s_val[lane] : 0.f;
        if (lane + 32 < P) it += s_val[lane + 32];
        static_assert(P <= 64, "two slots per lane");
        c = warp_sum(c);
        it = 2.0f * warp_sum(it);
        if (lane == 0) ticket_publish(ticket, p, b, c, it);
    }

    // ---- pass 2: gradient rows from the registers.  Every pair difference is formed once and feeds both rows
    //      (g_i += k d, g_j -= k d; the negation is an operand modifier of FFMA2). ----
    const bool with_grad = p.grad_xhat != nullptr;
    float2 K2[KSMEM ? 1 : P];
    if constexpr (!KSMEM) {
#pragma unroll
        for (int s = 0; s < P; ++s) K2[s] = s_coef2[s];
    }
    T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)b * M * p.D;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (with_grad && ok[k]) {
#pragma unroll
            for (int r = 0; r <= M; ++r) opaque(raw[k][r]);
            const long e0 = (tid + (long)k * THREADS) * VEC;
            uint32_t outw[M][4];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float2 x[M + 1], g[M];
#pragma unroll
                for (int r = 0; r <= M; ++r) x[r] = unpack_pair<T>(raw[k][r], h);
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const float2 kc = KSMEM ? lds64_volatile(&s_coef2[i]) : K2[KSMEM ? 0 : i];
                    g[i] = __fmul2_rn(kc, sub2(x[i], x[M]));
                }
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int j = i + 1; j < M; ++j) {
                        const float2 d = sub2(x[i], x[j]);
                        const float2 kc =
                            KSMEM ? lds64_volatile(&s_coef2[pair_slot<M>(i, j)]) : K2[KSMEM ? 0 : pair_slot<M>(i, j)];
                        g[i] = __ffma2_rn(kc, d, g[i]);
                        g[j] = __ffma2_rn(make_float2(-kc.x, -kc.y), d, g[j]);
                    }
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    if constexpr (sizeof(T) == 4) {
                        outw[i][2 * h] = __float_as_uint(g[i].x);
                        outw[i][2 * h + 1] = __float_as_uint(g[i].y);
                    } else {
                        outw[i][h] = pack_bf16x2(g[i].x, g[i].y);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
                stg_stream16(grow + (long)i * p.D + e0, make_uint4(outw[i][0], outw[i][1], outw[i][2], outw[i][3]));
        }
        if (warp == 0) {  // one step of the row publication per section; each result was requested a section ago
            if (k == 0 && lane == 0) ticket_take(ticket, p);
            if (k == (NV > 1 ? 1 : 0)) last_row = ticket_is_last(ticket, p, lane);
        }
    }
    if (tid == 0) DDDM_TRACE(5);
    if (warp == 0 && last_row) {
        ticket_finalize(p, W, lane);
        if (lane == 0) DDDM_TRACE(7);
    }
}

// ---- host side ------------------------------------------------------------------------------------------
template <typename T, int M, int NV, int THREADS>
int launch_energy_wave_cfg(const EnergyParams& p, bool ksmem, cudaStream_t stream) {
    if (ksmem)
        return launch_with_attrs(energy_fused_wave_kernel<T, M, NV, THREADS, true>, dim3(p.B), dim3(THREADS), 0, 1,
                                 stream, p);
    return launch_with_attrs(energy_fused_wave_kernel<T, M, NV, THREADS, false>, dim3(p.B), dim3(THREADS), 0, 1,
                             stream, p);
}

}  // namespace dddm
