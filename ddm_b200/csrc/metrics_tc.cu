// metrics_tc.cu — rbf_mmd2's pairwise kernel sums as ONE fused tensor-core kernel (dddm/metrics.py:140-163).
//
//   sum_{i,j} w_ij exp(-gamma (a2_i + b2_j - 2 a_i . b_j))
//
// The n x n Gram tile never reaches HBM: a persistent kernel walks 128 x 256 tiles, accumulates a_i . b_j on the
// 5th-generation tensor cores (tcgen05.mma, kind::f16, fp32 accumulators in tensor memory) and its epilogue warps read
// the accumulator straight from TMEM, form the distance, exponentiate, mask and sum.  Roles (one warp each, as the
// tcgen05 model wants): warp 0 = TMA producer (2-D tensor loads into 128-byte-swizzled tiles, two-stage ring),
// warp 1 = MMA issuer (one elected thread) and TMEM owner, warps 2-5 = epilogue (TMEM double-buffered: the epilogue of
// tile k overlaps the MMAs of tile k+1).
//
// Precision: the reference computes the Gram in fp32 (metrics.py:146).  Here every fp32 input is split ONCE, by a
// streaming pre-pass, into two bf16 terms x = hi + lo (+ a remainder below 2^-17 |x|), and a tile accumulates
// hi.hi + hi.lo + lo.hi in fp32: products of bf16 are exact in fp32, the dropped lo.lo and remainder terms are
// ~2^-17 relative per product with random sign.  a2 / b2 come from the fp32 inputs as in the reference.
// Symmetric terms (x against x) only visit tiles that touch the strict upper triangle and count it twice.
#include "common.cuh"
#include "umma.cuh"

namespace dddm {

namespace {
constexpr int kTileM = 128, kTileN = 256, kTileK = 64;  // K in bf16 elements: one 128-byte swizzled row
constexpr int kStages = 2;
constexpr int kABytes = kTileM * 128, kBBytes = kTileN * 128;            // one bf16 operand tile
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;                    // a_hi, a_lo, b_hi, b_lo
constexpr int kEpiWarps = 4, kThreads = (2 + kEpiWarps) * 32;
constexpr int kTmemCols = 512;                                            // two 256-column accumulators
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /* alignment slack */;

struct TileWalk {  // linear tile id -> (row tile, column tile); symmetric: only tiles touching j > i
    long tiles_m, tiles_n;
    int symmetric;
    __device__ long first_n(long tm) const { return symmetric ? (tm * kTileM) / kTileN : 0; }
    __device__ long total() const {
        long t = 0;
        for (long tm = 0; tm < tiles_m; ++tm) t += tiles_n - first_n(tm);
        return t;
    }
    __device__ void locate(long id, long& tm, long& tn) const {
        for (tm = 0; tm < tiles_m; ++tm) {
            const long cnt = tiles_n - first_n(tm);
            if (id < cnt) break;
            id -= cnt;
        }
        tn = first_n(tm) + id;
    }
};
}  // namespace

// x fp32 [n, D] -> hi, lo bf16 [n, Dp] (Dp = D rounded up to 64, zero padded): x ~= hi + lo to 2^-17 relative.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, long n, long D, long Dp) {
    const long total = n * Dp;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / Dp, c = idx - r * Dp;
        const float v = c < D ? x[r * D + c] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[idx] = h;
        lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

__global__ void __launch_bounds__(kThreads, 1)
rbf_gram_sum_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                       const float* __restrict__ a2, const float* __restrict__ b2, long rows_a, long rows_b, int k_blocks,
                       float neg_gamma_log2e, int symmetric, double* __restrict__ part) {
    using namespace umma;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_b2[2][kTileN];
    __shared__ double s_red[kEpiWarps];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TileWalk walk{(rows_a + kTileM - 1) / kTileM, (rows_b + kTileN - 1) / kTileN, symmetric};
    const long ntiles = walk.total();

    if (warp == 0 && lane == 0) {
        tma_prefetch_descriptor(&map_a_hi);
        tma_prefetch_descriptor(&map_a_lo);
        tma_prefetch_descriptor(&map_b_hi);
        tma_prefetch_descriptor(&map_b_lo);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                long tm, tn;
                walk.locate(t, tm, tn);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);  // the MMAs that read this slot have completed
                    unsigned char* st = smem + (size_t)stage * kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                    tma_load_2d(st, &map_a_hi, &full_bar[stage], kb * kTileK, (int)(tm * kTileM));
                    tma_load_2d(st + kABytes, &map_a_lo, &full_bar[stage], kb * kTileK, (int)(tm * kTileM));
                    tma_load_2d(st + 2 * kABytes, &map_b_hi, &full_bar[stage], kb * kTileK, (int)(tn * kTileN));
                    tma_load_2d(st + 2 * kABytes + kBBytes, &map_b_lo, &full_bar[stage], kb * kTileK, (int)(tn * kTileN));
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_instr_desc(kFmtBF16, kFmtBF16, kTileM, kTileN, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // the epilogue has drained this accumulator
                tc_fence_after_sync();
                const uint32_t d = tmem_base + (uint32_t)acc * kTileN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t st = smem_addr(smem + (size_t)stage * kStageBytes);
                    const uint32_t a_hi = st, a_lo = st + kABytes, b_hi = st + 2 * kABytes, b_lo = st + 2 * kABytes + kBBytes;
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k) {  // 16 bf16 = 32 bytes of K per instruction
                        const uint32_t off = k * 32;
                        const uint64_t dah = make_desc_kmajor_sw128(a_hi + off), dal = make_desc_kmajor_sw128(a_lo + off);
                        const uint64_t dbh = make_desc_kmajor_sw128(b_hi + off), dbl = make_desc_kmajor_sw128(b_lo + off);
                        mma_f16_ss(d, dah, dbh, idesc, (kb | k) != 0);
                        mma_f16_ss(d, dah, dbl, idesc, 1);
                        mma_f16_ss(d, dal, dbh, idesc, 1);
                    }
                    mma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs are done
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                mma_commit(&tmem_full[acc]);  // the accumulator of this tile is complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> distance -> exp -> masked sum =====
        const int ew = warp - 2;          // 0..3
        const int quarter = warp & 3;     // the TMEM lanes this warp may read: 32 * (warp % 4) ..
        const int et = ew * 32 + lane;    // 0..127 among the epilogue threads
        double total = 0.0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            long tm, tn;
            walk.locate(t, tm, tn);
            const long gi = tm * kTileM + quarter * 32 + lane;
            const long gj0 = tn * kTileN;
            // column norms of this tile (double-buffered with the accumulator; the barrier orders reuse)
            for (int c = et; c < kTileN; c += kEpiWarps * 32) s_b2[acc][c] = (gj0 + c < rows_b) ? b2[gj0 + c] : 0.f;
            const float a2i = gi < rows_a ? a2[gi] : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * kTileN;
#pragma unroll 1
            for (int c0 = 0; c0 < kTileN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c0, v);
                tmem_ld_wait();
                float part_sum = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const long gj = gj0 + c0 + c;
                    const float d2 = (a2i + s_b2[acc][c0 + c]) - 2.0f * __uint_as_float(v[c]);  // metrics.py:146
                    const float kv = exp2f(neg_gamma_log2e * d2);
                    const bool valid = gi < rows_a && gj < rows_b;
                    const float w = symmetric ? (gj > gi ? 2.0f : 0.0f) : 1.0f;
                    part_sum += valid ? w * kv : 0.f;
                }
                total += (double)part_sum;
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (lane == 0) s_red[ew] = total;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (et == 0) part[blockIdx.x] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

__global__ void __launch_bounds__(256) fold_double_tc_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    __shared__ double s_red[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) t += part[i];
    s_red[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s_red[0];
}

}  // namespace dddm

using namespace dddm;

extern "C" {

long dddm_rbf_tc_padded_cols(long D) { return D < 1 ? 0 : (D + kTileK - 1) / kTileK * kTileK; }

int dddm_rbf_split_bf16(const float* x, dddm_bf16* hi, dddm_bf16* lo, long n, long D, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K5 rbf split bf16");
    if (!x || !hi || !lo) return DDDM_ERR_NULL_POINTER;
    if (n < 1 || D < 1) return DDDM_ERR_BAD_SHAPE;
    const long Dp = dddm_rbf_tc_padded_cols(D);
    const long total = n * Dp;
    long blocks = (total + 255) / 256;
    const long cap = (long)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n, D, Dp);
    count_launch();
    return (int)cudaGetLastError();
}

size_t dddm_rbf_tc_scratch_bytes(void) { return (size_t)device_sm_count() * sizeof(double); }

int dddm_rbf_kernel_sum_tc(const dddm_bf16* a_hi, const dddm_bf16* a_lo, const dddm_bf16* b_hi, const dddm_bf16* b_lo,
                           const float* a2, const float* b2, long rows_a, long rows_b, long D, float gamma, int symmetric,
                           double* scratch, size_t scratch_bytes, double* out, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K5 rbf_kernel_sum tcgen05");
    if (!a_hi || !a_lo || !b_hi || !b_lo || !a2 || !b2 || !scratch || !out) return DDDM_ERR_NULL_POINTER;
    if (rows_a < 1 || rows_b < 1 || D < 1 || rows_a > 2000000000L || rows_b > 2000000000L) return DDDM_ERR_BAD_SHAPE;
    if (symmetric && rows_a != rows_b) return DDDM_ERR_BAD_ARGUMENT;
    const long Dp = dddm_rbf_tc_padded_cols(D);
    const int sms = device_sm_count();
    if (scratch_bytes < (size_t)sms * sizeof(double)) return DDDM_ERR_BAD_ARGUMENT;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    if (umma::make_tensor_map_bf16_rows(&ma_hi, a_hi, (uint64_t)rows_a, (uint64_t)Dp, kTileM) ||
        umma::make_tensor_map_bf16_rows(&ma_lo, a_lo, (uint64_t)rows_a, (uint64_t)Dp, kTileM) ||
        umma::make_tensor_map_bf16_rows(&mb_hi, b_hi, (uint64_t)rows_b, (uint64_t)Dp, kTileN) ||
        umma::make_tensor_map_bf16_rows(&mb_lo, b_lo, (uint64_t)rows_b, (uint64_t)Dp, kTileN))
        return DDDM_ERR_UNSUPPORTED;
    static SmemOptIn configured;
    if (int e = configured.ensure(rbf_gram_sum_tc_kernel, kSmemBytes, 0)) return e;
    const long tiles_m = (rows_a + kTileM - 1) / kTileM, tiles_n = (rows_b + kTileN - 1) / kTileN;
    long ntiles = 0;
    for (long tm = 0; tm < tiles_m; ++tm) ntiles += tiles_n - (symmetric ? (tm * kTileM) / kTileN : 0);
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const float ngl2e = -gamma * 1.4426950408889634f;
    rbf_gram_sum_tc_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, a2, b2, rows_a, rows_b, (int)(Dp / kTileK), ngl2e, symmetric ? 1 : 0, scratch);
    count_launch();
    fold_double_tc_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scratch, grid, out);
    count_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
