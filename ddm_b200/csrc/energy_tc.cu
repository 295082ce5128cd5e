// energy_tc.cu — the energy-score loss for m = 16 / 32 bf16 draws on the 5th-generation tensor cores (BASELINE config 3).
//
// At m >= 16 the direct-difference kernels are bound by the fp32 pipe (m(m-1)/2 = 120 / 496 pair distances per column,
// profiles/r01_k1_blk_m32.*).  Under the bf16 tolerance of the path (1e-2) the two pairwise contractions are GEMMs:
//
//   pass 1   Gram  G = X X^T  over D  (X = the row's m draws, bf16, K-major)      d2_ij = G_ii + G_jj - 2 G_ij
//   pass 2   grad^T [D x m] = X^T [D x m] * C^T [m x m],   C_ii = c_i + sum_j k_ij,  C_ij = -k_ij
//            (g_i = c_i (x_i - x0) + sum_j k_ij (x_i - x_j), losses.py:16-24 differentiated; the -c_i x0 term is added
//            by the epilogue)
//
// Both run as tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in tensor memory).  The SAME 128-byte-swizzled
// shared-memory tile serves both: read K-major (128-byte row = 64 columns of one draw) for the Gram, MN-major (the
// 128-byte row = 64 consecutive output rows of ONE k = draw) for the gradient.  The coefficient matrix C is split
// into two bf16 terms (hi + lo, 2^-17 relative) so that the cancellation S_i x_i - sum_j k_ij x_j keeps ~1e-5.
//
// One CTA per minibatch row, persistent.  Warp 0 = TMA producer (streams the row's draws twice: HBM, then L2),
// warp 1 = MMA issuer / TMEM owner, warps 2-5 = workers: confinement distances ||x_i - x0||^2 by direct fp32
// differences from the same tiles while the Gram accumulates (x0 may be fp32: the mixed entry point), then
// distances -> f, f' -> coefficient matrices, then the epilogue (TMEM -> registers -> -c_i x0 -> bf16 -> HBM).
//
// Near-duplicate draws (d2 < 2^-8 (|a|^2 + |b|^2), e.g. the identical draws a zero-initialised output layer emits):
// the Gram form loses d2 and the mixing form loses the gradient there, so such pairs are flagged, their distance is
// recomputed by direct differences, their coefficients are left out of C and their gradient contribution is added by
// a direct-difference post-pass — exact duplicates give (1e-12)^(beta/2) and zero gradient like the reference.
//
// Backward-only launches (kModeBwd: the autograd backward of `generalized_energy_terms`, losses.py:5-25) reuse the kernel:
// the distances come from the forward's saved `dist`, pass 1 only measures |x_i|^2 for the flags (no Gram), the prefactors
// are the two upstream gradients, and grad_x0 = - sum_i grad_xhat_i falls out of the epilogue (the pair terms are
// antisymmetric: their sum over the draws vanishes, what is left is minus the confinement gradients).
//
// Reference: dddm/losses.py:5-25, dddm/training.py:84-85.
#include "energy.cuh"
#include "energy_smem_plan.h"
#include "umma.cuh"

namespace dddm {

namespace {
constexpr int kTcWorkerWarps = 8;
constexpr int kTcWorkers = kTcWorkerWarps * 32;  // warps 2..9
constexpr int kTcThreads = 64 + kTcWorkers;      // + TMA producer warp + MMA warp
constexpr int kTcStageKB = 8;                    // 64-column K-blocks per pipeline stage (512 columns): few, fat stages —
                                                 // the workers' per-stage wait / fence / arrive round trip is latency, not work
constexpr int kTcOutTiles = 3;                   // epilogue staging tiles per group (a TMA store holds its tile ~1.5 us)
constexpr int kTcGramAccs = 2;                   // Gram accumulators (128 x 128 each), alternating so that consecutive MMAs never
                                                 // wait for their predecessor's accumulator; summed at read-out
constexpr int kTcAccBufs = 4;                    // gradient accumulators in TMEM (64 columns each: hi and lo products)
constexpr int kTcGradCol0 = kTcGramAccs * 128;   // first TMEM column of the gradient accumulators
constexpr int kTcTmemCols = 512;                 // 2 x 128 (Gram) + 4 x 64 (gradient): the whole tensor memory (1 CTA per SM)
constexpr float kTcTauDist = 1.0f / 256.0f;      // z-space: below this the Gram form of d2 is replaced by direct differences
constexpr float kTcTauGrad = 1.0f / 65536.0f;    // x-space: below this the mixing form of the gradient is replaced too

template <int M>
struct TcCfg {
    static constexpr int kSub = M * 128;               // bytes of one K-block sub-tile (M rows x 64 bf16)
    static constexpr int kStageBytes = kTcStageKB * kSub;
    static constexpr int kMaxStages = 6;                 // ring depth is chosen at launch from the shared memory left (3..6)
    static constexpr int kOutTile = M * 256;             // one epilogue staging tile: M draws x 128 columns bf16
    static constexpr int kOut = 2 * kTcOutTiles * kOutTile;  // two epilogue groups x kTcOutTiles; doubles as the slack the
                                                         // M = 128 Gram descriptor reads past the last sub-tile (16 KB)
    static constexpr int kPairs = M * (M - 1) / 2;
    static constexpr int kP = M + kPairs;
    static constexpr int kCoefBytes = M * M * 2;         // one bf16 coefficient matrix in core-matrix layout
    static constexpr int kItems = kTcStageKB * M * 8 / kTcWorkers;  // 16-byte chunks per worker thread and stage
    static constexpr int kGroupKB = 128 / M;             // K-blocks one Gram instruction covers (see the MMA issuer)
    static constexpr int kGroups = kTcStageKB / kGroupKB;  // such groups per stage
    static_assert(kTcStageKB % kGroupKB == 0, "a stage is a whole number of Gram groups");
    static_assert(kOut >= kGroupKB * M * (M + 1) * 4, "the partial Grams are parked in the epilogue staging tiles");
    static_assert(kOut >= 128 * 128, "slack for the 128-row operand descriptor");
};

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int M>
__device__ __forceinline__ int pair_index(int i, int j) {  // i < j, row-major upper triangle
    return i * M - i * (i + 1) / 2 + (j - i - 1);
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
    v[0] = bf16lo(r.x); v[1] = bf16hi(r.x); v[2] = bf16lo(r.y); v[3] = bf16hi(r.y);
    v[4] = bf16lo(r.z); v[5] = bf16hi(r.z); v[6] = bf16lo(r.w); v[7] = bf16hi(r.w);
}
}  // namespace

// in-kernel timeline (diagnostics, tools/trace_energy.py): 16 globaltimer stamps per CTA, first row of the CTA only
#define TC_TRACE(slot)                                                                                  \
    do {                                                                                                \
        if (p.trace != nullptr && b == (int)blockIdx.x) p.trace[(long)blockIdx.x * 16 + (slot)] = globaltimer_ns(); \
    } while (0)

template <int M>
__global__ void __launch_bounds__(kTcThreads, 1)
energy_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const EnergyParams p,
                 const int nstages) {
    using namespace umma;
    using C = TcCfg<M>;
    constexpr int P2 = C::kPairs, P = C::kP;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* outs = ring + (size_t)nstages * C::kStageBytes;   // epilogue staging tiles [group][buffer]
    float* x0s = reinterpret_cast<float*>(outs + C::kOut);           // [D] fp32 copy of the row's x0

    __shared__ __align__(8) uint64_t full_bar[C::kMaxStages], zfull_bar[C::kMaxStages], empty_bar[C::kMaxStages];
    __shared__ __align__(8) uint64_t acc_full[kTcAccBufs], acc_empty[kTcAccBufs];
    __shared__ __align__(8) uint64_t gram_full, gram_empty, coef_ready;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(128) unsigned char s_cmat[2 * C::kCoefBytes];  // [hi; lo] stacked along N: ONE MMA forms both products
    unsigned char* s_chi = s_cmat;
    unsigned char* s_clo = s_cmat + C::kCoefBytes;
    __shared__ float s_G[M][M + 1];                 // Gram, then the symmetric coefficient matrix k_ij (0 where flagged)
    __shared__ float s_d2[P], s_val[P], s_coef[P];  // slot s < M: confinement of draw s; M + pair_index(i, j): pair
    __shared__ float s_c[M];                        // confinement coefficient the epilogue applies to x0
    __shared__ float s_n2[M + 1];                   // |x_i|^2 (x-space scale of the gradient flags); [M] = |x0|^2
    __shared__ uint32_t s_fmask[M];                 // bit j: gradient of pair (i, j) by direct differences
    __shared__ uint32_t s_cflag;                    // bit i: gradient of the confinement term of draw i by direct differences
    __shared__ unsigned short s_flag[P2];           // pairs whose d2 is recomputed by direct differences
    __shared__ int s_nflag, s_npost;
    __shared__ unsigned char s_pi[P2], s_pj[P2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.D;
    const int nkb = D / 64;                                  // K-blocks per row
    const int nfill = (nkb + kTcStageKB - 1) / kTcStageKB;   // stage fills per pass
    const bool want_grad = p.grad_xhat != nullptr;
    const bool bwd = p.mode == kModeBwd;  // distances given, no loss outputs, grad_x0 on request
    const int dbg = p.ld_hint;  // diagnostics (tuning "energy.ldhint"): 1 no transform, 2 no Gram MMAs, 4 no gradient MMAs, 8 no stores

    if (threadIdx.x == 0) {
        tma_prefetch_descriptor(&map_x);
        tma_prefetch_descriptor(&map_g);
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&zfull_bar[s], kTcWorkerWarps);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < kTcAccBufs; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kTcWorkerWarps / 2);
        }
        mbar_init(&gram_full, 1);
        mbar_init(&gram_empty, 4);  // the four warps that read the Gram out of tensor memory
        mbar_init(&coef_ready, 1);
        fence_barrier_init();
    }
    for (int s = threadIdx.x; s < P2; s += kTcThreads) {  // pair slot -> (i, j)
        int i = 0, r = s;
        while (r >= M - 1 - i) {
            r -= M - 1 - i;
            ++i;
        }
        s_pi[s] = (unsigned char)i;
        s_pj[s] = (unsigned char)(i + 1 + r);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, kTcTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0 && p.trace != nullptr) p.trace[(long)blockIdx.x * 16 + 0] = globaltimer_ns();
    cudaGridDependencySynchronize();  // PDL: the draws may come from the previous kernel in the stream
    cudaTriggerProgrammaticLaunchCompletion();
    if (threadIdx.x == 0 && p.trace != nullptr) p.trace[(long)blockIdx.x * 16 + 1] = globaltimer_ns();

    if (warp == 0) {
        // ===================== TMA producer: every row is streamed twice (Gram pass, gradient pass) =====================
        if (lane == 0) {
            uint32_t n = 0;  // running fill counter -> ring slot and phase
            for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
                for (int pass = 0; pass < (want_grad ? 2 : 1); ++pass) {
                    for (int f = 0; f < nfill; ++f, ++n) {
                        const int slot = n % nstages;
                        const uint32_t phase = (n / nstages) & 1;
                        mbar_wait(&empty_bar[slot], phase ^ 1);
                        const int kb0 = f * kTcStageKB, cnt = min(kTcStageKB, nkb - kb0);
                        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)cnt * C::kSub);
                        unsigned char* st = ring + (size_t)slot * C::kStageBytes;
                        for (int k = 0; k < cnt; ++k)
                            tma_load_2d(st + (size_t)k * C::kSub, &map_x, &full_bar[slot], (kb0 + k) * 64, b * M);
                    }
                    TC_TRACE(2 + pass);  // all requests of this pass issued
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_gram = make_instr_desc(kFmtBF16, kFmtBF16, 128, 128, false, false);
            constexpr uint32_t idesc_grad = make_instr_desc(kFmtBF16, kFmtBF16, 128, 2 * M, true, false);
            const uint32_t cmat = smem_addr(s_cmat);
            constexpr uint32_t kCoefLbo = 128, kCoefSbo = (M / 8) * 128;  // core matrices: next 8 k / next 8 rows
            uint32_t n = 0, g = 0, row_it = 0, zpar = 0;  // zpar: per-slot phase of zfull_bar (completed by Gram-pass fills only)
            for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++row_it) {
                // ---- pass 1: Gram of z = bf16(x - x0) over D into TMEM columns [0, M) ----
                mbar_wait(&gram_empty, (row_it & 1) ^ 1);  // the previous row's Gram has been read
                tc_fence_after_sync();
                for (int f = 0; f < nfill; ++f, ++n) {
                    const int slot = n % nstages;
                    mbar_wait(&zfull_bar[slot], (zpar >> slot) & 1u);  // the workers have replaced x by z in this slot
                    zpar ^= 1u << slot;
                    tc_fence_after_sync();
                    // Block-diagonal batching.  A 64 x m x 16 instruction costs ~90 cycles whatever its shape (it is latency-,
                    // not throughput-bound on the tensor pipe), and a row needs D / 16 of them.  The sub-tiles of consecutive
                    // K-blocks are contiguous in shared memory, so a 128-row operand starting at sub-tile k covers the SAME
                    // 16 columns-within-the-block of 128 / m consecutive K-blocks: ONE 128 x 128 x 16 instruction (A = B =
                    // that operand) forms 128 / m Gram contributions at once on its diagonal m x m blocks (the off-diagonal
                    // blocks mix different K-blocks and are never read).  The read-out sums the diagonal blocks.
                    const uint64_t d0 = make_desc_kmajor_sw128(smem_addr(ring + (size_t)slot * C::kStageBytes));
                    const int cnt = ((dbg & 2) || bwd) ? 0 : min(kTcStageKB, nkb - f * kTcStageKB);
#pragma unroll
                    for (int gq = 0; gq < C::kGroups; ++gq) {
                        if (gq * C::kGroupKB < cnt) {  // (a partial group's missing sub-tiles were zeroed by the workers)
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {  // 16 columns = 32 bytes of K per instruction
                                const uint64_t d = d0 + (uint64_t)((gq * C::kGroupKB * C::kSub + k4 * 32) >> 4);
                                const uint32_t acc = tmem_base + (uint32_t)(k4 % kTcGramAccs) * 128u;
                                if (gq == 0 && k4 < kTcGramAccs)
                                    mma_f16_ss(acc, d, d, idesc_gram, f != 0);
                                else
                                    mma_f16_ss(acc, d, d, idesc_gram, 1);
                            }
                        }
                    }
                    mma_commit(&empty_bar[slot]);
                }
                mma_commit(&gram_full);
                TC_TRACE(4);
                if (!want_grad) continue;
                // ---- pass 2: 128 output columns per accumulator, K = the m draws (the x tiles as they come from TMA) ----
                mbar_wait(&coef_ready, row_it & 1);
                tc_fence_after_sync();
                uint64_t dcm[M / 16];  // the coefficient operand never moves
#pragma unroll
                for (int ks = 0; ks < M / 16; ++ks) dcm[ks] = make_desc_kmajor_core(cmat + ks * 2 * kCoefLbo, kCoefLbo, kCoefSbo);
                for (int f = 0; f < nfill; ++f, ++n) {
                    const int slot = n % nstages;
                    mbar_wait(&full_bar[slot], (n / nstages) & 1);
                    const uint64_t dst0 = make_desc_mnmajor_sw128(smem_addr(ring + (size_t)slot * C::kStageBytes), C::kSub);
                    const int nblk_stage = min(kTcStageKB, nkb - f * kTcStageKB) / 2;  // blocks of 128 output columns
                    for (int h = 0; h < nblk_stage; h += 2) {  // two blocks at a time, their instructions alternating
                        const int nblk = min(2, nblk_stage - h);
                        const uint64_t da0 = dst0 + (uint64_t)((h * 2 * C::kSub) >> 4);
                        const int buf0 = g % kTcAccBufs, buf1 = (g + 1) % kTcAccBufs;
                        mbar_wait(&acc_empty[buf0], ((g / kTcAccBufs) & 1) ^ 1);
                        if (nblk > 1) mbar_wait(&acc_empty[buf1], (((g + 1) / kTcAccBufs) & 1) ^ 1);
                        tc_fence_after_sync();
                        const uint32_t acc0 = tmem_base + kTcGradCol0 + (uint32_t)buf0 * 64u;
                        const uint32_t acc1 = tmem_base + kTcGradCol0 + (uint32_t)buf1 * 64u;
                        if (!(dbg & 4)) {
#pragma unroll
                            for (int ks = 0; ks < M / 16; ++ks) {  // 16 draws per instruction = 2 groups of 8 tile rows
                                const uint64_t dA = da0 + (uint64_t)((ks * 2048) >> 4);
                                const uint64_t dB = da0 + (uint64_t)((2 * C::kSub + ks * 2048) >> 4);
                                mma_f16_ss(acc0, dA, dcm[ks], idesc_grad, ks != 0);
                                if (nblk > 1) mma_f16_ss(acc1, dB, dcm[ks], idesc_grad, ks != 0);
                            }
                        }
                        mma_commit(&acc_full[buf0]);
                        if (nblk > 1) mma_commit(&acc_full[buf1]);
                        g += nblk;
                    }
                    mma_commit(&empty_bar[slot]);
                }
                TC_TRACE(5);
            }
        }
    } else {
        // ===================== workers (8 warps) =====================
        const int wt = threadIdx.x - 64;  // 0..255
        const int ww = warp - 2;          // 0..7
        const int quarter = warp & 3;     // TMEM lanes 32 * quarter ..
        const int egroup = ww >> 2;       // epilogue: group 0 takes even 128-column blocks, group 1 odd ones
        const bool eleader = (ww & 3) == 0 && lane == 0;  // issues the group's TMA stores
        // transform pass: a thread owns one 16-byte chunk position (draw ti, chunk tc) of every K-block it touches
        constexpr int ITEMS = C::kItems;                 // 4 (m = 32) or 2 (m = 16) K-blocks of a stage per thread
        const int ti = (wt % (M * 8)) >> 3, tc = wt & 7;
        const int tkb0 = wt / (M * 8);                   // first K-block of the stage this thread touches (0 for m = 32)
        constexpr int KBSTEP = kTcWorkers / (M * 8);     // K-block stride between a thread's items (1 or 2)
        const uint32_t toff = (uint32_t)ti * 128u + (uint32_t)((tc ^ (ti & 7)) << 4);
        const __nv_bfloat16* xrow0 = static_cast<const __nv_bfloat16*>(p.xhat);
        const bool x0f = p.x0_f32 != 0;
        uint32_t n = 0, g = 0, row_it = 0, eblk = 0;  // eblk: blocks this epilogue group has staged (staging-buffer parity)
        for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++row_it) {
            const long x0_base = (long)b * D;
            // ---- stage x0 as fp32 (bf16 or fp32 in memory) ----
            if (x0f) {
                const float* src = static_cast<const float*>(p.x0) + (long)b * D;
                for (int d = wt * 4; d < D; d += kTcWorkers * 4) *reinterpret_cast<float4*>(x0s + d) = *reinterpret_cast<const float4*>(src + d);
            } else {
                const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.x0) + (long)b * D;
                for (int d = wt * 8; d < D; d += kTcWorkers * 8) {
                    float v[8];
                    unpack8(*reinterpret_cast<const uint4*>(src + d), v);
                    *reinterpret_cast<float4*>(x0s + d) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(x0s + d + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            if (wt == 0) {
                s_nflag = 0;
                s_npost = 0;
                s_cflag = 0;
            }
            if (wt < M) s_fmask[wt] = 0;
            if (wt <= M) s_n2[wt] = 0.f;
            named_bar(1, kTcWorkers);
            const float W = (p.mode == kModeLoss) ? p.weight_dev[0] * p.weight_scale : 1.0f;
            const float nb = (float)p.B * (float)M;
            const float pre_conf = bwd ? 2.0f * p.g_conf[0] / nb : 2.0f * W / nb;
            const float pre_pair = bwd ? 4.0f * p.g_inter[0] / (nb * (float)(M - 1))
                                       : -4.0f * W * (p.lam / (2.0f * (float)(M - 1))) / (nb * (float)(M - 1));

            // ---- pass 1 (workers): x -> z = bf16(x - x0) in place (same swizzled position): centred on x0 the Gram keeps
            //      the pairwise distances without cancellation against |x|^2, and its diagonal IS the confinement term ----
            float acc_n2 = 0.f, acc_n0 = 0.f;
            for (int f = 0; f < nfill; ++f, ++n) {
                const int slot = n % nstages;
                const int kb0 = f * kTcStageKB, cnt = min(kTcStageKB, nkb - kb0);
                mbar_wait(&full_bar[slot], (n / nstages) & 1);
                unsigned char* st = ring + (size_t)slot * C::kStageBytes;
#pragma unroll
                for (int r = 0; r < ITEMS; ++r) {
                    const int kb = tkb0 + r * KBSTEP;
                    if (kb < cnt && !(dbg & 1)) {
                        const uint32_t off = (uint32_t)kb * C::kSub + toff;
                        float x[8];
                        unpack8(*reinterpret_cast<const uint4*>(st + off), x);
                        const float* zp = x0s + (kb0 + kb) * 64 + tc * 8;
                        const float4 za = *reinterpret_cast<const float4*>(zp), zb = *reinterpret_cast<const float4*>(zp + 4);
                        const float z0[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
                        uint32_t hw[4];
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            acc_n2 = fmaf(x[e], x[e], acc_n2);
                            acc_n2 = fmaf(x[e + 1], x[e + 1], acc_n2);
                            hw[e >> 1] = pack_bf16x2(x[e] - z0[e], x[e + 1] - z0[e + 1]);
                        }
                        if (ti == 0) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc_n0 = fmaf(z0[e], z0[e], acc_n0);
                        }
                        if (!bwd) *reinterpret_cast<uint4*>(st + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    } else if (kb >= cnt && kb < (cnt + C::kGroupKB - 1) / C::kGroupKB * C::kGroupKB) {
                        // last stage of a row whose K-blocks do not fill the Gram group: its diagonal blocks must add zero
                        *reinterpret_cast<uint4*>(st + (uint32_t)kb * C::kSub + toff) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                fence_async_smem();  // generic-proxy writes -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&zfull_bar[slot]);
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {  // the 8 chunk positions of a draw sit in 8 consecutive lanes
                acc_n2 += __shfl_xor_sync(0xffffffffu, acc_n2, o);
                acc_n0 += __shfl_xor_sync(0xffffffffu, acc_n0, o);
            }
            if (tc == 0) {
                atomicAdd(&s_n2[ti], acc_n2);  // m = 16: two threads per draw (a flag scale: the order does not matter)
                if (ti == 0) atomicAdd(&s_n2[M], acc_n0);
            }
            if (wt == 0) TC_TRACE(6);

            // ---- Gram: the diagonal m x m blocks of the two 128 x 128 accumulators -> shared memory -> their sum ----
            float* s_Gp = reinterpret_cast<float*>(outs);  // [128 / m blocks][m][m + 1], parked in the (idle) staging tiles
            if (ww < 4) {  // warps 2..5 own TMEM lane quarters 2, 3, 0, 1: accumulator rows 32 q .. 32 q + 31
                mbar_wait(&gram_full, row_it & 1);
                tc_fence_after_sync();
                float gs[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) gs[j] = 0.f;
#pragma unroll
                for (int a = 0; a < kTcGramAccs; ++a) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * 128u + (uint32_t)quarter * 32u, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) gs[j] += __uint_as_float(v[j]);
                }
                tc_fence_before_sync();
                // lane l of quarter q holds accumulator row 32 q + l, columns 32 q .. 32 q + 31: block (32 q + l) / m, row
                // (32 q + l) % m of it, and the block's columns start at column (block * m) - 32 q of what was loaded
                const int arow = quarter * 32 + lane, blk = arow / M, brow = arow % M, c0 = blk * M - quarter * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j >= c0 && j < c0 + M) s_Gp[(blk * M + brow) * (M + 1) + (j - c0)] = gs[j];
                __syncwarp();
                if (lane == 0) mbar_arrive(&gram_empty);
            }
            named_bar(1, kTcWorkers);
            for (int e = wt; e < M * M; e += kTcWorkers) {  // fixed order: deterministic
                const int i = e / M, j = e % M;
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < C::kGroupKB; ++k) t += s_Gp[(k * M + i) * (M + 1) + j];
                s_G[i][j] = t;
            }
            named_bar(1, kTcWorkers);
            if (wt == 0) TC_TRACE(7);

            // ---- distances in z-space: conf = diagonal; pairs G_ii + G_jj - 2 G_ij, near-duplicates by direct differences
            //      (backward-only launches take the forward's saved distances) ----
            if (bwd) {
                for (int s = wt; s < P; s += kTcWorkers) s_d2[s] = p.dist[(long)b * P + s];
            } else {
                if (wt < M) s_d2[wt] = fmaxf(s_G[wt][wt], 0.f);
                for (int s = wt; s < P2; s += kTcWorkers) {
                    const int i = s_pi[s], j = s_pj[s];
                    const float nn = s_G[i][i] + s_G[j][j];
                    const float d2 = nn - 2.0f * s_G[i][j];
                    if (!(d2 >= kTcTauDist * nn)) {  // also catches NaN
                        const int k = atomicAdd(&s_nflag, 1);
                        s_flag[k] = (unsigned short)s;
                    } else {
                        s_d2[M + s] = d2;
                    }
                }
            }
            named_bar(1, kTcWorkers);
            const int nflag = s_nflag;
            for (int fidx = ww; fidx < nflag; fidx += kTcWorkerWarps) {
                const int s = s_flag[fidx];
                const __nv_bfloat16* xi = xrow0 + ((long)b * M + s_pi[s]) * D;
                const __nv_bfloat16* xj = xrow0 + ((long)b * M + s_pj[s]) * D;
                float a = 0.f;
                for (int d0 = lane * 8; d0 < D; d0 += 1024) {  // four independent 16-byte loads per row in flight
                    uint4 u[4], w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (d0 + q * 256 < D) {
                            u[q] = *reinterpret_cast<const uint4*>(xi + d0 + q * 256);
                            w[q] = *reinterpret_cast<const uint4*>(xj + d0 + q * 256);
                        }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (d0 + q * 256 < D) {
                            float xa[8], xb[8];
                            unpack8(u[q], xa);
                            unpack8(w[q], xb);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float t = xa[e] - xb[e];
                                a = fmaf(t, t, a);
                            }
                        }
                }
                a = warp_sum(a);
                if (lane == 0) s_d2[M + s] = a;
            }
            named_bar(1, kTcWorkers);  // also: every thread is done with the Gram values in s_G

            // ---- f(d2), f'(d2); x-space flags of the gradient's mixing form; symmetric coefficient matrix into s_G ----
            for (int s = wt; s < P; s += kTcWorkers) {
                float val, der;
                const float d2 = s_d2[s];
                pow_value_deriv(d2, p.pw, val, der);
                s_val[s] = val;
                const float coef = ((s < M) ? pre_conf : pre_pair) * der;
                s_coef[s] = coef;
                if (p.dist != nullptr && !bwd) p.dist[(long)b * P + s] = d2;
                if (want_grad) {
                    if (s < M) {
                        const bool fl = !(d2 >= kTcTauGrad * (s_n2[s] + s_n2[M]));
                        if (fl) {
                            atomicOr(&s_cflag, 1u << s);
                            atomicAdd(&s_npost, 1);
                        }
                        s_c[s] = fl ? 0.f : coef;
                    } else {
                        const int i = s_pi[s - M], j = s_pj[s - M];
                        const bool fl = !(d2 >= kTcTauGrad * (s_n2[i] + s_n2[j]));
                        if (fl) {
                            atomicOr(&s_fmask[i], 1u << j);
                            atomicOr(&s_fmask[j], 1u << i);
                            atomicAdd(&s_npost, 1);
                        }
                        s_G[i][j] = s_G[j][i] = fl ? 0.f : coef;
                    }
                }
            }
            named_bar(1, kTcWorkers);

            // ---- coefficient matrices C = hi + lo (bf16, K-major core-matrix layout): one warp per row of C ----
            if (want_grad) {
                constexpr int kSbo = (M / 8) * 128;
                for (int i = ww; i < M; i += kTcWorkerWarps) {
                    const float k = (lane < M && lane != i) ? s_G[i][lane] : 0.f;
                    const float S = warp_sum(k) + s_c[i];  // c_i + sum_j k_ij over the terms the tensor core handles
                    if (lane < M) {
                        const float v = (lane == i) ? S : -k;
                        const __nv_bfloat16 h = __float2bfloat16_rn(v);
                        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
                        const int off = (i >> 3) * kSbo + (lane >> 3) * 128 + (i & 7) * 16 + (lane & 7) * 2;
                        *reinterpret_cast<__nv_bfloat16*>(s_chi + off) = h;
                        *reinterpret_cast<__nv_bfloat16*>(s_clo + off) = l;
                    }
                }
                fence_async_smem();
                named_bar(1, kTcWorkers);
                if (wt == 0) mbar_arrive(&coef_ready);
            }
            if (wt == 0) TC_TRACE(8);

            // ---- row sums -> loss (one warp; the others go on) ----
            if (ww == 7 && !bwd) {
                float c = 0.f, it = 0.f;
                for (int s = lane; s < P; s += 32) {
                    if (s < M) c += s_val[s]; else it += s_val[s];
                }
                c = warp_sum(c);
                it = 2.0f * warp_sum(it);
                finish_row(p, b, c, it, W, lane);
                if (lane == 0) TC_TRACE(11);
            }

            if (want_grad) {
                // ---- pass 2 epilogue: accumulator (128 output columns x m draws) -> -c_i x0 -> bf16 -> staging tile
                //      [draw][128 columns] -> ONE TMA tensor store per block (256-byte row segments) ----
                float cneg[M];
#pragma unroll
                for (int i = 0; i < M; ++i) cneg[i] = -s_c[i];
                for (int f = 0; f < nfill; ++f) {
                    const int cnt = min(kTcStageKB, nkb - f * kTcStageKB);
                    for (int blk = 0; blk < cnt / 2; ++blk, ++g) {
                        if ((int)(g & 1) != egroup) continue;
                        const int buf = g % kTcAccBufs;
                        const int d0 = (f * kTcStageKB + 2 * blk) * 64;
                        const int dl = quarter * 32 + lane;  // column within the block
                        const float z = x0s[d0 + dl];
                        unsigned char* ot = outs + (size_t)(egroup * kTcOutTiles + (eblk % kTcOutTiles)) * C::kOutTile;
                        // accumulator first (its TMEM read overlaps the wait for the staging tile below)
                        mbar_wait(&acc_full[buf], (g / kTcAccBufs) & 1);
                        tc_fence_after_sync();
                        // columns [0, M): x . C_hi, [M, 2M): x . C_lo
                        uint32_t v[32], w[32];
                        const uint32_t tad = tmem_base + ((uint32_t)(quarter * 32) << 16) + kTcGradCol0 + (uint32_t)buf * 64u;
                        tmem_ld_32x32(tad, v);
                        if (M == 32) tmem_ld_32x32(tad + 32u, w);
                        if (eleader) tma_store_wait_read<kTcOutTiles - 1>();  // the store that last read this staging tile is done with it
                        tmem_ld_wait();
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);
                        named_bar(2 + egroup, kTcWorkers / 2);
                        float gsum = 0.f;
#pragma unroll
                        for (int i = 0; i < M; ++i) {
                            const float lo = (M == 32) ? __uint_as_float(w[i]) : __uint_as_float(v[(i + 16) & 31]);
                            const float o = fmaf(cneg[i], z, __uint_as_float(v[i]) + lo);
                            gsum += o;
                            *reinterpret_cast<__nv_bfloat16*>(ot + i * 256 + dl * 2) = __float2bfloat16_rn(o);
                        }
                        // d/dx0 = - sum_i d/dxhat_i: the pair terms cancel in the sum, the confinement terms change sign
                        if (bwd && p.grad_x0 != nullptr)
                            static_cast<__nv_bfloat16*>(p.grad_x0)[x0_base + d0 + dl] = __float2bfloat16_rn(-gsum);
                        fence_async_smem();
                        named_bar(2 + egroup, kTcWorkers / 2);
                        if (eleader && !(dbg & 8)) {
                            tma_store_2d(&map_g, ot, d0, b * M);
                            tma_store_commit();
                        }
                        ++eblk;
                    }
                }
                n += nfill;  // the gradient pass' fills are consumed by the MMA warp alone
                if (eleader) tma_store_wait<0>();  // this row's gradient is in memory (the post-pass below reads it back)
                named_bar(1, kTcWorkers);
                if (wt == 0) TC_TRACE(9);

                // ---- direct-difference post-pass for the flagged terms (read-modify-write of the stored gradient) ----
                if (s_npost > 0) {
                    __nv_bfloat16* grow = static_cast<__nv_bfloat16*>(p.grad_xhat) + (long)b * M * D;
                    for (int i = ww; i < M; i += kTcWorkerWarps) {
                        const uint32_t fm = s_fmask[i];
                        const bool cf = (s_cflag >> i) & 1u;
                        if (fm == 0 && !cf) continue;
                        const __nv_bfloat16* xi = xrow0 + ((long)b * M + i) * D;
                        __nv_bfloat16* gi = grow + (long)i * D;
                        for (int d = lane * 8; d < D; d += 256) {
                            float xv[8], a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            unpack8(*reinterpret_cast<const uint4*>(xi + d), xv);
                            uint32_t mm = fm;
                            while (mm) {  // partners four at a time: their loads are independent
                                int js[4];
                                uint4 w[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    js[q] = mm ? __ffs(mm) - 1 : -1;
                                    if (mm) mm &= mm - 1;
                                    if (js[q] >= 0) w[q] = *reinterpret_cast<const uint4*>(xrow0 + ((long)b * M + js[q]) * D + d);
                                }
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (js[q] < 0) continue;
                                    const int j = js[q];
                                    const float k = s_coef[M + (j > i ? pair_index<M>(i, j) : pair_index<M>(j, i))];
                                    float yv[8];
                                    unpack8(w[q], yv);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) a[e] = fmaf(k, xv[e] - yv[e], a[e]);
                                }
                            }
                            if (cf) {
                                const float k = s_coef[i];
#pragma unroll
                                for (int e = 0; e < 8; ++e) a[e] = fmaf(k, xv[e] - x0s[d + e], a[e]);
                            }
                            uint4 gq;
                            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(gq.x), "=r"(gq.y), "=r"(gq.z), "=r"(gq.w) : "l"(gi + d));
                            gq.x = pack_bf16x2(bf16lo(gq.x) + a[0], bf16hi(gq.x) + a[1]);
                            gq.y = pack_bf16x2(bf16lo(gq.y) + a[2], bf16hi(gq.y) + a[3]);
                            gq.z = pack_bf16x2(bf16lo(gq.z) + a[4], bf16hi(gq.z) + a[5]);
                            gq.w = pack_bf16x2(bf16lo(gq.w) + a[6], bf16hi(gq.w) + a[7]);
                            *reinterpret_cast<uint4*>(gi + d) = gq;
                        }
                    }
                }
            }
            if (want_grad && bwd && p.grad_x0 != nullptr && s_cflag != 0) {
                // confinement terms that went through the direct post-pass also belong in grad_x0 (one warp, all of them)
                named_bar(1, kTcWorkers);
                if (ww == 0) {
                    __nv_bfloat16* g0 = static_cast<__nv_bfloat16*>(p.grad_x0) + x0_base;
                    for (int d = lane * 8; d < D; d += 256) {
                        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        for (int i = 0; i < M; ++i) {
                            if (!((s_cflag >> i) & 1u)) continue;
                            const float k = s_coef[i];
                            float xv[8];
                            unpack8(*reinterpret_cast<const uint4*>(xrow0 + ((long)b * M + i) * D + d), xv);
#pragma unroll
                            for (int e = 0; e < 8; ++e) a[e] = fmaf(k, xv[e] - x0s[d + e], a[e]);
                        }
                        uint4 gq;
                        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(gq.x), "=r"(gq.y), "=r"(gq.z), "=r"(gq.w) : "l"(g0 + d));
                        gq.x = pack_bf16x2(bf16lo(gq.x) - a[0], bf16hi(gq.x) - a[1]);
                        gq.y = pack_bf16x2(bf16lo(gq.y) - a[2], bf16hi(gq.y) - a[3]);
                        gq.z = pack_bf16x2(bf16lo(gq.z) - a[4], bf16hi(gq.z) - a[5]);
                        gq.w = pack_bf16x2(bf16lo(gq.w) - a[6], bf16hi(gq.w) - a[7]);
                        *reinterpret_cast<uint4*>(g0 + d) = gq;
                    }
                }
            }
            if (wt == 0) TC_TRACE(10);
            named_bar(1, kTcWorkers);  // x0s, s_* are rewritten by the next row
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTcTmemCols);
    }
}

// Shapes the tensor-core kernel covers: bf16 draws, m = 16 or 32, D a multiple of 128 (one accumulator = 128 output
// columns), 16-byte aligned rows.
TcPlan plan_tc(int B, int m, int D, int elem_size, bool aligned16) {
    TcPlan t{};
    t.ok = false;
    if (elem_size != 2 || !(m == 16 || m == 32) || B < 1 || D < 128 || D % 128 != 0 || !aligned16) return t;
    const size_t stage = (m == 32) ? (size_t)TcCfg<32>::kStageBytes : (size_t)TcCfg<16>::kStageBytes;
    const size_t fixed = 1024 + ((m == 32) ? (size_t)TcCfg<32>::kOut : (size_t)TcCfg<16>::kOut) + (size_t)D * 4;  // + fp32 copy of x0
    const size_t budget = 232448 - 20 * 1024;  // + ~18 KB of static shared memory
    if (fixed + 3 * stage > budget) return t;
    int stages = (int)((budget - fixed) / stage);
    const int want = tuning().nv > 0 ? tuning().nv : 4;  // ("energy.nv" doubles as the ring depth knob of this kernel)
    if (stages > want) stages = want;
    if (stages > TcCfg<32>::kMaxStages) stages = TcCfg<32>::kMaxStages;
    t.stages = stages;
    t.smem_bytes = fixed + (size_t)stages * stage;
    t.ok = true;
    return t;
}

int launch_energy_tc(const EnergyParams& p, const TcPlan& plan, cudaStream_t stream) {
    if (p.mode == kModeBwd && (p.x0_f32 || p.dist == nullptr || p.grad_xhat == nullptr)) return DDDM_ERR_UNSUPPORTED;
    CUtensorMap map, map_g;
    if (umma::make_tensor_map_bf16_rows(&map, p.xhat, (uint64_t)p.B * p.m, (uint64_t)p.D, (uint32_t)p.m)) return DDDM_ERR_UNSUPPORTED;
    // the gradient leaves through TMA tensor stores of [m draws x 128 columns] tiles (a forward-only launch never stores)
    const void* gbase = p.grad_xhat ? p.grad_xhat : p.xhat;
    if (umma::make_tensor_map_bf16_dense(&map_g, gbase, (uint64_t)p.B * p.m, (uint64_t)p.D, (uint32_t)p.m, 128)) return DDDM_ERR_UNSUPPORTED;
    const int sms = device_sm_count();
    const int grid = p.B < sms ? p.B : sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = tuning().pdl ? 1 : 0;
    cudaError_t e;
    if (p.m == 32) {
        static SmemOptIn configured;
        if (int r = configured.ensure(energy_tc_kernel<32>, plan.smem_bytes, 0)) return r;
        e = cudaLaunchKernelEx(&cfg, energy_tc_kernel<32>, map, map_g, p, plan.stages);
    } else {
        static SmemOptIn configured;
        if (int r = configured.ensure(energy_tc_kernel<16>, plan.smem_bytes, 0)) return r;
        e = cudaLaunchKernelEx(&cfg, energy_tc_kernel<16>, map, map_g, p, plan.stages);
    }
    count_launch();
    return (int)e;
}

}  // namespace dddm
