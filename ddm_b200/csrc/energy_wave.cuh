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

// L2 eviction-priority policy chosen at run time (tuning "energy.ldhint" / "energy.sthint"):
// 0 evict_normal (the default), 1 evict_first, 2 evict_last, 3 evict_unchanged.
__device__ __forceinline__ uint64_t make_l2_policy(int kind) {
    uint64_t pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 3) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_stream16_hint(const void* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg_stream16_hint(void* p, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}
// Pins a section boundary for the in-kernel timeline: the listed accumulators must have been computed here.
template <int N>
__device__ __forceinline__ void pin_pairs(float2 (&a)[N]) {
#pragma unroll
    for (int s = 0; s + 1 < N; s += 2)
        asm volatile("" : "+f"(a[s].x), "+f"(a[s].y), "+f"(a[s + 1].x), "+f"(a[s + 1].y));
}

// Round-robin ("circle method") schedule of the M(M-1)/2 pairs: in every round each row meets exactly one partner,
// so consecutive FFMA2s of pass 2 update different gradient rows and no update waits for the previous one of its row.
template <int M>
struct PairSchedule {
    static constexpr int kN = (M % 2 == 0) ? M : M + 1;  // odd M: one bye per round
    static constexpr int kRounds = kN - 1;
    static constexpr int kPerRound = kN / 2;
    // pair k of round r -> (a, b); a or b == M (only for odd M) means a bye
    __host__ __device__ static constexpr int first(int r, int k) {
        return k == 0 ? kN - 1 : (r + k) % (kN - 1);
    }
    __host__ __device__ static constexpr int second(int r, int k) {
        return k == 0 ? r % (kN - 1) : (r + kN - 1 - k) % (kN - 1);
    }
};

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
// The last row requests the first 128 rows' sums as soon as it knows it is last ...
__device__ __forceinline__ void ticket_prefetch(RowTicket& t, const EnergyParams& p, int lane) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int r = u * 32 + lane;
        t.part[u] = make_float2(0.f, 0.f);
        if (r < p.B)
            asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];"
                         : "=f"(t.part[u].x), "=f"(t.part[u].y)
                         : "l"(reinterpret_cast<const float2*>(p.row_partials) + r)
                         : "memory");
    }
}
// ... and folds them after its last pass-2 section (fixed order: lane-strided, then the shuffle tree).
__device__ __forceinline__ void ticket_finalize(const RowTicket& t, const EnergyParams& p, float W, int lane) {
    float c = 0.f, i = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        c += t.part[u].x;
        i += t.part[u].y;
    }
    for (int r0 = 128; r0 < p.B; r0 += 128) {  // more than 128 rows (only when the variant is forced)
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

template <typename T, int M, int NV, int THREADS, int P2, bool NOSTORE = false>
__global__ void __launch_bounds__(THREADS, 1) energy_fused_wave_kernel(const EnergyParams p) {
    constexpr int P = M * (M + 1) / 2;
    constexpr int VEC = Elem<T>::kVec;
    constexpr int NH = VEC / 2;  // column pairs per vector
    constexpr int NW = THREADS / 32;
    using WR = WarpReduce<P>;
    __shared__ float s_warp[NW][P];
    __shared__ __align__(8) float2 s_coef2[P];  // (k, k): the packed operand of FFMA2
    __shared__ float s_val[P];

    // pass-2 form: 0 column-major, the 36 coefficient pairs in registers; 1 column-major, coefficients re-read from
    // shared memory; 2 pair-major (one coefficient at a time over all the thread's columns)
    constexpr bool KSMEM = P2 == 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    constexpr int cluster_size = 1, rank = 0;  // for DDDM_TRACE
    if (tid == 0) DDDM_TRACE(0);

    const long nvec = p.D / VEC;
    const T* __restrict__ xrow = static_cast<const T*>(p.xhat) + (long)b * M * p.D;
    const T* __restrict__ crow = static_cast<const T*>(p.x0) + (long)b * p.D;

    const uint64_t ld_pol = make_l2_policy(p.ld_hint), st_pol = make_l2_policy(p.st_hint);
    cudaGridDependencySynchronize();  // PDL: the inputs may be produced by the previous kernel in the stream
    if (tid == 0) {
        DDDM_TRACE(1);
        if (p.trace != nullptr) p.trace[(long)b * 16 + 12] = (unsigned long long)clock64();
    }

    // The weight is requested BEFORE the row: the memory pipeline is in-order per SM, and behind the 27 row loads
    // of every warp it would arrive last and hold back pass 1 (measured: the whole pass waited for it).
    float W = 1.0f;
    if (p.mode == kModeLoss) asm volatile("ld.global.f32 %0, [%1];" : "=f"(W) : "l"(p.weight_dev));

    // ---- the row, vector-major: vector k of every tile row before vector k+1 (arrival order = use order) ----
    uint4 raw[NV][M + 1];
    bool ok[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const long v = tid + (long)k * THREADS;
        ok[k] = v < nvec;
        if (ok[k]) {
#pragma unroll
            for (int r = 0; r < M; ++r) raw[k][r] = ldg_stream16_hint(xrow + (long)r * p.D + v * VEC, ld_pol);
            raw[k][M] = ldg_stream16_hint(crow + v * VEC, ld_pol);
        } else {  // past the row end: zeros contribute nothing to any distance
#pragma unroll
            for (int r = 0; r <= M; ++r) raw[k][r] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (tid == 0) DDDM_TRACE(6);
    cudaTriggerProgrammaticLaunchCompletion();

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
            if (p.trace != nullptr) {  // timeline only: section k ends here (otherwise the sections may interleave)
                pin_pairs(acc2);
                if (k == 0 && tid == 0) DDDM_TRACE(2);
            }
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
    if (p.mode == kModeLoss) W *= p.weight_scale;
    if (tid < P) {
        const float nb = (float)p.B * (float)M;
        const float pre_conf = 2.0f * W / nb;
        const float pre_pair = -4.0f * W * (p.lam / (2.0f * (float)(M - 1))) / (nb * (float)(M - 1));
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

    const bool with_grad = p.grad_xhat != nullptr;
    T* __restrict__ grow = static_cast<T*>(p.grad_xhat) + (long)b * M * p.D;
    if constexpr (P2 == 2) {
        // ---- pass 2, pair-major: ONE coefficient at a time, applied to all NC column pairs of the thread.  The NC
        //      difference/update triples of a pair are independent (no FFMA2 waits for the one before it), a
        //      coefficient costs one shared-memory read per thread instead of a register pair held throughout, and
        //      row i is complete — and stored — as soon as pair (i, M-1) is done, so the stores spread over the pass.
        constexpr int NC = NV * NH;
        constexpr int kCheck = (M - 2 < 2) ? M - 2 : 2;  // row after which the ticket's answer is looked at
        if (with_grad) {
            float2 xf[M + 1][NC], g[M][NC];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
#pragma unroll
                for (int r = 0; r <= M; ++r) {
                    opaque(raw[k][r]);
#pragma unroll
                    for (int h = 0; h < NH; ++h) xf[r][k * NH + h] = unpack_pair<T>(raw[k][r], h);
                }
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const float2 kc = lds64_volatile(&s_coef2[i]);
#pragma unroll
                for (int c = 0; c < NC; ++c) g[i][c] = __fmul2_rn(kc, sub2(xf[i][c], xf[M][c]));
            }
            auto store_row = [&](int i) {
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    uint4 o;
                    if constexpr (sizeof(T) == 4) {
                        o = make_uint4(__float_as_uint(g[i][k * NH].x), __float_as_uint(g[i][k * NH].y),
                                       __float_as_uint(g[i][k * NH + 1].x), __float_as_uint(g[i][k * NH + 1].y));
                    } else {
                        o = make_uint4(pack_bf16x2(g[i][k * NH].x, g[i][k * NH].y),
                                       pack_bf16x2(g[i][k * NH + 1].x, g[i][k * NH + 1].y),
                                       pack_bf16x2(g[i][k * NH + 2].x, g[i][k * NH + 2].y),
                                       pack_bf16x2(g[i][k * NH + 3].x, g[i][k * NH + 3].y));
                    }
                    if (ok[k] && (!NOSTORE || o.x == 0x7fc12345u))
                        stg_stream16_hint(grow + (long)i * p.D + (tid + (long)k * THREADS) * VEC, o, st_pol);
                }
            };
#pragma unroll
            for (int i = 0; i < M - 1; ++i) {
#pragma unroll
                for (int j = i + 1; j < M; ++j) {
                    const float2 kc = lds64_volatile(&s_coef2[pair_slot<M>(i, j)]);
                    const float2 nkc = make_float2(-kc.x, -kc.y);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const float2 d = sub2(xf[i][c], xf[j][c]);
                        g[i][c] = __ffma2_rn(kc, d, g[i][c]);
                        g[j][c] = __ffma2_rn(nkc, d, g[j][c]);
                    }
                }
                store_row(i);
                if (warp == 0) {  // one step of the row publication per finished row; each result was requested rows ago
                    if (i == 0 && lane == 0) ticket_take(ticket, p);
                    if (i == kCheck) {
                        last_row = ticket_is_last(ticket, p, lane);
                        if (last_row) ticket_prefetch(ticket, p, lane);
                    }
                }
            }
            store_row(M - 1);
        } else if (warp == 0) {
            if (lane == 0) ticket_take(ticket, p);
            last_row = ticket_is_last(ticket, p, lane);
            if (last_row) ticket_prefetch(ticket, p, lane);
        }
    } else {
    // ---- pass 2: gradient rows from the registers.  Every pair difference is formed once and feeds both rows
    //      (g_i += k d, g_j -= k d; the negation is an operand modifier of FFMA2). ----
    float2 K2[KSMEM ? 1 : P];
    if constexpr (!KSMEM) {
#pragma unroll
        for (int s = 0; s < P; ++s) K2[s] = s_coef2[s];
    }
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
                using PS = PairSchedule<M>;
#pragma unroll
                for (int rd = 0; rd < PS::kRounds; ++rd)
#pragma unroll
                    for (int q = 0; q < PS::kPerRound; ++q) {
                        const int a = PS::first(rd, q), c = PS::second(rd, q);
                        const int i = a < c ? a : c, j = a < c ? c : a;
                        if (j < M) {  // (odd M: the bye of this round)
                            const float2 d = sub2(x[i], x[j]);
                            const int slot = M + i * M - i * (i + 1) / 2 + (j - i - 1);
                            const float2 kc = KSMEM ? lds64_volatile(&s_coef2[slot]) : K2[KSMEM ? 0 : slot];
                            g[i] = __ffma2_rn(kc, d, g[i]);
                            g[j] = __ffma2_rn(make_float2(-kc.x, -kc.y), d, g[j]);
                        }
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
            for (int i = 0; i < M; ++i) {
                // NOSTORE (diagnostics): pass 2 computes everything but only a never-occurring bit pattern is stored
                if (!NOSTORE || outw[i][0] == 0x7fc12345u)
                    stg_stream16_hint(grow + (long)i * p.D + e0, make_uint4(outw[i][0], outw[i][1], outw[i][2], outw[i][3]),
                                      st_pol);
            }
        }
        if (warp == 0) {  // one step of the row publication per section; each result was requested a section ago
            if (k == 0 && lane == 0) ticket_take(ticket, p);
            if (k == (NV >= 2 ? 1 : 0)) {  // NV = 3: the last row's reads fly during its last section
                last_row = ticket_is_last(ticket, p, lane);
                if (last_row) ticket_prefetch(ticket, p, lane);
            }
        }
    }
    }
    if (tid == 0) {
        DDDM_TRACE(5);
        if (p.trace != nullptr) p.trace[(long)b * 16 + 13] = (unsigned long long)clock64();
    }
    if (warp == 0 && last_row) {
        ticket_finalize(ticket, p, W, lane);
        if (lane == 0) DDDM_TRACE(7);
    }
}

// ---- host side ------------------------------------------------------------------------------------------
template <typename T, int M, int NV, int THREADS>
int launch_energy_wave_cfg(const EnergyParams& p, int ksmem, cudaStream_t stream) {
    if constexpr (M == 8 && NV == 3 && THREADS == 256) {
        if (tuning().nostore)
            return launch_with_attrs(energy_fused_wave_kernel<T, M, NV, THREADS, 2, true>, dim3(p.B), dim3(THREADS), 0, 1,
                                     stream, p);
    }
    (void)ksmem;  // the column-major pass-2 forms (P2 = 0, 1) measured no faster than pair-major and are not instantiated
    return launch_with_attrs(energy_fused_wave_kernel<T, M, NV, THREADS, 2>, dim3(p.B), dim3(THREADS), 0, 1, stream, p);
}

}  // namespace dddm
