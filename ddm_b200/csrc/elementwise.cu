// elementwise.cu — the HBM-bound elementwise kernels around the energy score.
//   K2  forward marginal + m-fold expansion   (dddm/schedules.py:17-25, dddm/training.py:70)
//   K3  Gaussian bridge + Algorithm-2 update  (dddm/schedules.py:28-78, dddm/sampling.py:29-31)
//   K4  logistic weight w(t) and its batch sum (dddm/losses.py:28-35, dddm/training.py:84)
//   scale-in-place (hand-off of the pre-multiplied gradient of K1 to autograd)
// All are pure streaming kernels: 16-byte vector loads/stores, grids sized as a multiple of the
// SM count, no shared-memory staging (nothing is reused).
#include <type_traits>

#include "common.cuh"
#include "energy.cuh"  // launch_with_attrs (programmatic dependent launch)

namespace dddm {

static int num_sms() { return device_sm_count(); }  // per-device cache in api.cu

// Every streaming kernel below starts with pdl_prologue(): launched with the PDL attribute (tuning().pdl, like K1) its CTAs
// may become resident while the previous kernel of the stream drains, and wait here — before their first global access —
// until that kernel has completed and its writes are visible.  Without the attribute both calls are no-ops.
__device__ __forceinline__ void pdl_prologue() {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
}
template <typename K, typename... Args>
static int launch_streaming(K kernel, dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
    return launch_with_attrs(kernel, grid, block, 0, 1, stream, args...);  // counts the launch
}
// A/B on one box, back-to-back launches with HBM-cold inputs and outputs (profiles/r02/elementwise_pdl_ab.log):
// K2c 7.51 / 7.67 -> 7.34 / 7.35 us and K3 9.86 / 9.91 -> 9.44 / 9.26 us with the attribute, but K2 (almost only stores:
// 3 MB in, 12.6 MB out) 3.99 / 4.22 -> 4.69 / 4.78 us, so K2 keeps the plain launch.
template <typename K, typename... Args>
static int launch_plain(K kernel, dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
    kernel<<<grid, block, 0, stream>>>(args...);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename T>
static bool aligned16(const T* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// ---- K2 ------------------------------------------------------------------------------------
// grid: (column tiles, rows).  Each thread owns one VEC-wide column group of one row b, computes
// xt once and writes it to xt (optional) and to the m expanded rows b*m .. b*m+m-1.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
forward_marginal_expand_kernel(const T* __restrict__ x0, const float* __restrict__ t, const T* __restrict__ eps,
                               T* __restrict__ xt, T* __restrict__ xt_rep, int m, long D) {
    const long nvec = D / VEC;
    const int b = blockIdx.y;
    const float tb = t[b];
    const float ab = 1.0f - tb;  // alpha_sigma, schedules.py:5-14
    const T* x0r = x0 + (long)b * D;
    const T* er = eps + (long)b * D;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
        float a[VEC], e[VEC], o[VEC];
        load_pack<T, VEC>(x0r, v * VEC, a);
        load_pack<T, VEC>(er, v * VEC, e);
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = __fadd_rn(__fmul_rn(ab, a[k]), __fmul_rn(tb, e[k]));  // no FMA: matches eager
        if (xt != nullptr) store_pack<T, VEC>(xt + (long)b * D, v * VEC, o);
        if (xt_rep != nullptr) {
            T* base = xt_rep + (long)b * m * D;
            for (int i = 0; i < m; ++i) store_pack<T, VEC>(base + (long)i * D, v * VEC, o);
        }
    }
}

template <typename T>
static int forward_marginal_expand(const T* x0, const float* t, const T* eps, T* xt, T* xt_rep, int B, int m, long D,
                                   cudaStream_t stream) {
    if (!x0 || !t || !eps || (!xt && !xt_rep)) return DDDM_ERR_NULL_POINTER;
    if (B < 0 || D < 0 || (xt_rep && m < 1) || B > 65535 * 32) return DDDM_ERR_BAD_SHAPE;
    if (B == 0 || D == 0) return DDDM_OK;
    constexpr int V = Elem<T>::kVec;
    const bool vec = (D % V == 0) && aligned16(x0) && aligned16(eps) && (!xt || aligned16(xt)) &&
                     (!xt_rep || aligned16(xt_rep));
    const long nvec = vec ? D / V : D;
    const int threads = nvec >= 256 ? 256 : (int)((nvec + 31) / 32 * 32);
    long tiles = (nvec + threads - 1) / threads;
    // enough CTAs to fill the machine, no more than ~8 waves of work items per CTA column
    const long max_tiles = ((long)num_sms() * 8 + B - 1) / B;
    if (tiles > max_tiles) tiles = max_tiles < 1 ? 1 : max_tiles;
    if (B > 65535) return DDDM_ERR_BAD_SHAPE;
    dim3 grid((unsigned)tiles, (unsigned)B);
    return vec ? launch_plain(forward_marginal_expand_kernel<T, V>, grid, dim3(threads), stream, x0, t, eps, xt, xt_rep, m, D)
               : launch_plain(forward_marginal_expand_kernel<T, 1>, grid, dim3(threads), stream, x0, t, eps, xt, xt_rep, m, D);
}

// ---- K2c -----------------------------------------------------------------------------------
// Forward marginal written m-fold straight into the channel-concatenated backbone input
//   x6[b*m+i] = cat(x_t[b], xi[b,i])   ([B*m, 2C, H, W]; dddm/model.py:236 builds it with torch.cat),
// optionally down-converted to bf16 on the way out, plus (optional) x0 re-ordered into the patch-token
// layout PatchUnembed produces (dddm/model.py:125-128) so that the loss can consume the backbone's tokens
// without the unpatchify copy: the energy score is invariant to a common permutation of the D axis.
// One thread owns 4 consecutive pixels of one image row (W % 4 == 0, patch % 4 == 0).
template <typename TI, typename TO>
__device__ __forceinline__ void store4(TO* dst, const float (&v)[4]) {
    if constexpr (sizeof(TO) == 4) {
        stg_stream16(dst, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
    } else {
        stg_stream8(dst, make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3])));
    }
}
template <typename TI>
__device__ __forceinline__ void load4(const TI* src, float (&v)[4]) {
    if constexpr (sizeof(TI) == 4) {
        const uint4 r = ldg_stream16(src);
        v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
    } else {
        const uint2 r = *reinterpret_cast<const uint2*>(src);
        v[0] = bf16lo(r.x); v[1] = bf16hi(r.x); v[2] = bf16lo(r.y); v[3] = bf16hi(r.y);
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256, 3)  // 3 CTAs per SM: the B x 3 CTAs of the CIFAR shape are resident in one wave
forward_marginal_concat_kernel(const TI* __restrict__ x0, const float* __restrict__ t, const TI* __restrict__ eps,
                               const TI* __restrict__ xi, TO* __restrict__ x6, TI* __restrict__ x0_tok, int m, int C, int H,
                               int W, int patch) {
    const long D = (long)C * H * W;
    const long nquad = D / 4;
    const int b = blockIdx.y;
    pdl_prologue();
    const float tb = t[b];
    const float ab = 1.0f - tb;
    const TI* x0r = x0 + (long)b * D;
    const TI* er = eps + (long)b * D;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nquad; v += (long)gridDim.x * blockDim.x) {
        const long e = v * 4;
        float a[4], n[4], o[4];
        load4<TI>(x0r + e, a);
        load4<TI>(er + e, n);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = __fadd_rn(__fmul_rn(ab, a[k]), __fmul_rn(tb, n[k]));  // no FMA: matches eager
        // xi of up to 8 draws is requested before the first of them is stored: a thread owns ONE quad, so with one
        // load -> store round trip per draw the launch was a chain of m HBM latencies (8.35 us for 29.9 MB at m = 8)
        constexpr int U = 8;
        auto draws = [&](int i0, auto full_tag) {  // FULL: all U draws exist (m = 8: no guards, no branches)
            constexpr bool FULL = decltype(full_tag)::value;
            const TI* src = xi + ((long)b * m + i0) * D + e;
            TO* dst = x6 + ((long)b * m + i0) * 2 * D + e;
            float z[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (FULL || i0 + u < m) load4<TI>(src + (long)u * D, z[u]);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (FULL || i0 + u < m) {
                    store4<TI, TO>(dst + (long)u * 2 * D, o);
                    store4<TI, TO>(dst + (long)u * 2 * D + D, z[u]);
                }
        };
        int i0 = 0;
        for (; i0 + U <= m; i0 += U) draws(i0, std::true_type{});
        if (i0 < m) draws(i0, std::false_type{});
        if (x0_tok != nullptr) {
            // pixel (c, y, x) -> token ((gy, gx), c, py, px) in 32-bit unsigned arithmetic (D < 2^31 is checked by the
            // launcher): the 64-bit divisions this used to be were more than half of the kernel's instructions
            const unsigned ue = (unsigned)e, uW = (unsigned)W, uH = (unsigned)H, up = (unsigned)patch;
            const unsigned row = ue / uW, x = ue - row * uW;
            const unsigned c = row / uH, y = row - c * uH;
            const unsigned gx = x / up, px = x - gx * up, gy = y / up, py = y - gy * up;
            const unsigned tok = ((gy * (uW / up) + gx) * (unsigned)C + c) * up * up + py * up + px;
            store4<TI, TI>(x0_tok + (long)b * D + tok, a);
        }
    }
}

template <typename TI, typename TO>
static int forward_marginal_concat(const TI* x0, const float* t, const TI* eps, const TI* xi, TO* x6, TI* x0_tok, int B,
                                   int m, int C, int H, int W, int patch, cudaStream_t stream) {
    if (!x0 || !t || !eps || !xi || !x6) return DDDM_ERR_NULL_POINTER;
    if (B < 0 || m < 1 || C < 1 || H < 1 || W < 1 || B > 65535 || (long)C * H * W >= (1L << 31)) return DDDM_ERR_BAD_SHAPE;
    if (W % 4 != 0) return DDDM_ERR_UNSUPPORTED;
    if (x0_tok && (patch < 4 || patch % 4 != 0 || W % patch != 0 || H % patch != 0)) return DDDM_ERR_BAD_SHAPE;
    if (!aligned16(x0) || !aligned16(eps) || !aligned16(xi) || !aligned16(x6) || (x0_tok && !aligned16(x0_tok)))
        return DDDM_ERR_BAD_ALIGNMENT;
    if (B == 0) return DDDM_OK;
    const long nquad = (long)C * H * W / 4;
    const int threads = nquad >= 256 ? 256 : (int)((nquad + 31) / 32 * 32);
    long tiles = (nquad + threads - 1) / threads;
    const long max_tiles = ((long)num_sms() * 8 + B - 1) / B;
    if (tiles > max_tiles) tiles = max_tiles < 1 ? 1 : max_tiles;
    return launch_streaming(forward_marginal_concat_kernel<TI, TO>, dim3((unsigned)tiles, (unsigned)B), dim3(threads), stream,
                            x0, t, eps, xi, x6, x0_tok, m, C, H, W, patch);
}

// ---- K4 ------------------------------------------------------------------------------------
__device__ __forceinline__ float logistic_weight(float t, float bias) {
    const float a = 1.0f - t;
    const float ratio = __fdiv_rn(a * a, t * t + kWeightEps);
    const float z = logf(ratio + kWeightEps) - bias;
    return 1.0f / (1.0f + expf(-z));
}

// Single CTA (B is a minibatch size): per-thread strided partial sums, fixed-shape tree -> deterministic.
__global__ void __launch_bounds__(1024)
sigmoid_weight_sum_kernel(const float* __restrict__ t, float bias, float* __restrict__ w, float* __restrict__ w_sum,
                          int B) {
    __shared__ float s_part[32];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float wb = logistic_weight(t[b], bias);
        if (w != nullptr) w[b] = wb;
        acc += wb;
    }
    if (w_sum == nullptr) return;
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? s_part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) w_sum[0] = v;
    }
}

// ---- K3 ------------------------------------------------------------------------------------
struct BridgeCoef {
    float c_xt, c_x0, std;
};
// schedules.py:45-77 in fp32, operation for operation (so the scalar-coefficient form reproduces
// the reference's eager result; SURVEY.md §8c "bit-exactly on CPU").
// e2 = eps_churn^2 and ome2 = 1 - eps_churn^2 are formed in double on the host and rounded once,
// exactly like the Python scalars the reference multiplies its fp32 tensors with.
__device__ __forceinline__ BridgeCoef bridge_coef(float s, float t, float e2, float ome2) {
    const float a_s = 1.0f - s, a_t = 1.0f - t;
    const float ratio = __fdiv_rn(s, t + kBridgeEps);
    const float alpha_ratio = __fdiv_rn(a_t, a_s + kBridgeEps);
    const float r11 = __fmul_rn(alpha_ratio, ratio);
    const float r12 = __fmul_rn(alpha_ratio, __fmul_rn(ratio, ratio));
    BridgeCoef c;
    c.c_xt = __fadd_rn(__fmul_rn(e2, r12), __fmul_rn(ome2, ratio));
    c.c_x0 = __fmul_rn(a_s, __fsub_rn(__fsub_rn(1.0f, __fmul_rn(e2, r12)), __fmul_rn(ome2, r11)));
    const float inner = __fadd_rn(__fmul_rn(e2, r11), ome2);
    const float var = __fmul_rn(__fmul_rn(s, s), fmaxf(__fsub_rn(1.0f, __fmul_rn(inner, inner)), 0.0f));
    c.std = sqrtf(fmaxf(var, 0.0f));
    return c;
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
bridge_step_kernel(T* __restrict__ x_out, const T* __restrict__ x, const T* __restrict__ xhat0,
                   const T* __restrict__ z, const float* __restrict__ s, const float* __restrict__ t,
                   int st_is_vector, float e2, float ome2, T* __restrict__ mu_out, float* __restrict__ std_out, long N,
                   long D) {
    const long nvec = D / VEC;
    pdl_prologue();
    for (long n = blockIdx.y; n < N; n += gridDim.y) {
        const BridgeCoef c = st_is_vector ? bridge_coef(s[n], t[n], e2, ome2) : bridge_coef(s[0], t[0], e2, ome2);
        if (std_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && (st_is_vector || n == 0)) std_out[n] = c.std;
        const long base = n * D;
        for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
            float xv[VEC], hv[VEC], zv[VEC], mu[VEC], o[VEC];
            load_pack<T, VEC>(x + base, v * VEC, xv);
            load_pack<T, VEC>(xhat0 + base, v * VEC, hv);
#pragma unroll
            for (int k = 0; k < VEC; ++k) mu[k] = __fadd_rn(__fmul_rn(c.c_xt, xv[k]), __fmul_rn(c.c_x0, hv[k]));
            if (mu_out != nullptr) store_pack<T, VEC>(mu_out + base, v * VEC, mu);
            if (x_out != nullptr) {
                if (z != nullptr) {
                    load_pack<T, VEC>(z + base, v * VEC, zv);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) o[k] = __fadd_rn(mu[k], __fmul_rn(c.std, zv[k]));
                    store_pack<T, VEC>(x_out + base, v * VEC, o);
                } else {
                    store_pack<T, VEC>(x_out + base, v * VEC, mu);
                }
            }
        }
    }
}

// Flat form for scalar (s, t) — the sampler's case: the [N, D] tensors are one array of 16-byte vectors and every
// thread owns UNROLL of them, all loads issued before the first use (3 x UNROLL x 16 bytes in flight per thread: a
// 50 MB update is ~10 us long, so memory-level parallelism per SM, not occupancy, decides its bandwidth).
template <typename T, int VEC, int UNROLL>
__global__ void __launch_bounds__(256)
bridge_step_flat_kernel(T* __restrict__ x_out, const T* __restrict__ x, const T* __restrict__ xhat0,
                        const T* __restrict__ z, const float* __restrict__ s, const float* __restrict__ t, float e2,
                        float ome2, long nvec) {
    pdl_prologue();
    const BridgeCoef c = bridge_coef(s[0], t[0], e2, ome2);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long v0 = (long)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * UNROLL) {
        float xv[UNROLL][VEC], hv[UNROLL][VEC], zv[UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long v = v0 + u * stride;
            if (v < nvec) {
                load_pack<T, VEC>(x, v * VEC, xv[u]);
                load_pack<T, VEC>(xhat0, v * VEC, hv[u]);
                load_pack<T, VEC>(z, v * VEC, zv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long v = v0 + u * stride;
            if (v < nvec) {
                float o[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float mu = __fadd_rn(__fmul_rn(c.c_xt, xv[u][k]), __fmul_rn(c.c_x0, hv[u][k]));
                    o[k] = __fadd_rn(mu, __fmul_rn(c.std, zv[u][k]));
                }
                store_pack<T, VEC>(x_out, v * VEC, o);
            }
        }
    }
}

template <typename T>
static int bridge_step(T* x_out, const T* x, const T* xhat0, const T* z, const float* s, const float* t,
                       int st_is_vector, double eps_churn, T* mu_out, float* std_out, long N, long D,
                       cudaStream_t stream) {
    if (!x || !xhat0 || !s || !t || (!x_out && !mu_out)) return DDDM_ERR_NULL_POINTER;
    if (N < 0 || D < 0) return DDDM_ERR_BAD_SHAPE;
    if (N == 0 || D == 0) return DDDM_OK;
    constexpr int V = Elem<T>::kVec;
    const bool vec = (D % V == 0) && aligned16(x) && aligned16(xhat0) && (!z || aligned16(z)) &&
                     (!x_out || aligned16(x_out)) && (!mu_out || aligned16(mu_out));
    const float e2 = (float)(eps_churn * eps_churn), ome2 = (float)(1.0 - eps_churn * eps_churn);
    if (vec && !st_is_vector && x_out && z && !mu_out && !std_out && x_out != x && x_out != xhat0 && x_out != z) {
        constexpr int kUnroll = 4;
        const long total = N * (D / V);
        long blocks = (total + 256L * kUnroll - 1) / (256L * kUnroll);
        const long cap = (long)num_sms() * 16;
        if (blocks > cap) blocks = cap;
        return launch_streaming(bridge_step_flat_kernel<T, V, kUnroll>, dim3((unsigned)blocks), dim3(256), stream, x_out, x, xhat0, z,
                                s, t, e2, ome2, total);
    }
    const long nvec = vec ? D / V : D;
    const int threads = nvec >= 256 ? 256 : (int)((nvec + 31) / 32 * 32);
    long tiles = (nvec + threads - 1) / threads;
    const long rows = N < 65535 ? N : 65535;
    const long max_tiles = ((long)num_sms() * 8 + rows - 1) / rows;
    if (tiles > max_tiles) tiles = max_tiles < 1 ? 1 : max_tiles;
    dim3 grid((unsigned)tiles, (unsigned)rows);
    return vec ? launch_streaming(bridge_step_kernel<T, V>, grid, dim3(threads), stream, x_out, x, xhat0, z, s, t, st_is_vector, e2,
                                  ome2, mu_out, std_out, N, D)
               : launch_streaming(bridge_step_kernel<T, 1>, grid, dim3(threads), stream, x_out, x, xhat0, z, s, t, st_is_vector, e2,
                                  ome2, mu_out, std_out, N, D);
}

// ---- K3 with the Gaussian draws fused in (dddm/sampling.py:27,30) ----------------------------------------------
// The reference draws xi = randn_like(x) before the denoiser call and z = randn_like(x) before the update: two more
// launches per step, and z costs a write and a read of N*D elements that never needs to exist.  Here z is generated
// in registers and the NEXT step's xi is written by the same kernel.  The values are bit-identical to what
// torch.randn_like returns for the same (seed, offset): ATen's normal kernel (distribution_elementwise_grid_stride_kernel
// + curand_normal4) assigns element li = v + G*k + 4*G*it  (G = its total thread count, k = 0..3) component k of the
// it-th Philox4x32-10 block of subsequence v at counter offset/4 + it, turned into normals by Box-Muller on the pairs
// (x, y) and (z, w).  The kernel walks the same (v, it) lattice — four elements G apart per work item, each access
// coalesced across the warp — so one Philox block serves four elements exactly as in ATen.
struct Philox {
    static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
    __device__ static __forceinline__ uint4 round(uint4 c, uint2 k) {
        const uint32_t hi0 = __umulhi(kM0, c.x), lo0 = kM0 * c.x;
        const uint32_t hi1 = __umulhi(kM1, c.z), lo1 = kM1 * c.z;
        return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    }
    __device__ static __forceinline__ uint4 block(uint4 c, uint2 k) {  // Philox4x32-10
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            c = round(c, k);
            k.x += kW0;
            k.y += kW1;
        }
        return round(c, k);
    }
};
// curand's _curand_box_muller (curand_normal.h), same expression shapes so that the compiler contracts them alike
__device__ __forceinline__ void box_muller(uint32_t x, uint32_t y, float& a, float& b) {
    constexpr float k2Pow32Inv = 2.3283064e-10f, k2Pow32Inv2Pi = 2.3283064e-10f * 6.2831855f;
    const float u = x * k2Pow32Inv + (k2Pow32Inv / 2);
    const float v = y * k2Pow32Inv2Pi + (k2Pow32Inv2Pi / 2);
    const float s = sqrtf(-2.0f * logf(u));
    float sn, cs;
    __sincosf(v, &sn, &cs);
    a = sn * s;
    b = cs * s;
}
__device__ __forceinline__ void normal4(unsigned long long seed, unsigned long long block_index, unsigned long long subseq,
                                        float (&n)[4]) {
    const uint4 c = make_uint4((uint32_t)block_index, (uint32_t)(block_index >> 32), (uint32_t)subseq, (uint32_t)(subseq >> 32));
    const uint4 r = Philox::block(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    box_muller(r.x, r.y, n[0], n[1]);
    box_muller(r.z, r.w, n[2], n[3]);
}

// philox = {seed, offset of the z draw, offset of the xi draw} (device memory, so that a captured graph can be
// replayed with new offsets), or null with the three values passed by value.
// IDX = unsigned (the whole lattice below 2^31 work items and elements: one 32-bit division per work item) or long.
template <typename T, typename IDX>
__global__ void __launch_bounds__(256)
bridge_step_philox_kernel(T* x_out, const T* x, const T* xhat0, T* xi_out, const float* __restrict__ s,
                          const float* __restrict__ t, float e2, float ome2, const unsigned long long* __restrict__ philox,
                          unsigned long long seed, unsigned long long off_z, unsigned long long off_xi, IDX G, int iters,
                          IDX numel) {
    if (philox != nullptr) {
        seed = philox[0];
        off_z = philox[1];
        off_xi = philox[2];
    }
    const BridgeCoef c = bridge_coef(s[0], t[0], e2, ome2);
    const IDX total = G * (IDX)iters;
    for (IDX w = (IDX)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (IDX)gridDim.x * blockDim.x) {
        const IDX it = w / G, v = w - it * G;
        const IDX li0 = v + 4 * G * it;
        float xv[4], hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const IDX li = li0 + G * k;
            if (li < numel) {
                xv[k] = Elem<T>::to_float(x[li]);
                hv[k] = Elem<T>::to_float(xhat0[li]);
            }
        }
        float z[4];
        normal4(seed, off_z / 4 + (unsigned long long)it, (unsigned long long)v, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const IDX li = li0 + G * k;
            if (li < numel) {
                const float mu = __fadd_rn(__fmul_rn(c.c_xt, xv[k]), __fmul_rn(c.c_x0, hv[k]));
                const float zk = Elem<T>::to_float(Elem<T>::from_float(z[k]));  // randn_like(x) has x's dtype
                x_out[li] = Elem<T>::from_float(__fadd_rn(mu, __fmul_rn(c.std, zk)));
            }
        }
        if (xi_out != nullptr) {
            float xi[4];
            normal4(seed, off_xi / 4 + (unsigned long long)it, (unsigned long long)v, xi);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const IDX li = li0 + G * k;
                if (li < numel) xi_out[li] = Elem<T>::from_float(xi[k]);
            }
        }
    }
}

// ATen's launch geometry for a normal_ over numel elements on the current device (calc_execution_policy):
// G threads in total, `iters` Philox blocks per thread, the generator advances by 4 * iters.
static void torch_philox_plan(long numel, long* G, long* iters) {
    int dev = 0, max_threads = 2048;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_threads, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    const long block = 256;
    long grid = (numel + block - 1) / block;
    const long cap = (long)num_sms() * (max_threads / block);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    *G = grid * block;
    *iters = numel > 0 ? (numel - 1) / (*G * 4) + 1 : 0;
}

template <typename T>
static int bridge_step_philox(T* x_out, const T* x, const T* xhat0, T* xi_out, const float* s, const float* t,
                              double eps_churn, const unsigned long long* philox_dev, unsigned long long seed,
                              unsigned long long off_z, unsigned long long off_xi, long N, long D, cudaStream_t stream) {
    if (!x_out || !x || !xhat0 || !s || !t) return DDDM_ERR_NULL_POINTER;
    if (N < 0 || D < 0) return DDDM_ERR_BAD_SHAPE;
    const long numel = N * D;
    if (numel == 0) return DDDM_OK;
    if (!philox_dev && ((off_z | off_xi) & 3ull)) return DDDM_ERR_BAD_ARGUMENT;  // torch's offsets are multiples of 4
    long G, iters;
    torch_philox_plan(numel, &G, &iters);
    const long total = G * iters;
    // one resident wave: the kernel is bound by the Philox rounds and Box-Muller, not by memory, so tails cost
    static int resident_per_device[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    if (resident_per_device[dev] == 0) {
        int r = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, bridge_step_philox_kernel<T, unsigned>, 256, 0);
        resident_per_device[dev] = r > 0 ? r : 4;
    }
    long blocks = (total + 255) / 256;
    const long cap = (long)num_sms() * resident_per_device[dev];
    if (blocks > cap) blocks = cap;
    const float e2 = (float)(eps_churn * eps_churn), ome2 = (float)(1.0 - eps_churn * eps_churn);
    if (total * 5 < (1L << 31) && numel < (1L << 31))  // 4 G it + 3 G stays below 2^31 as well
        bridge_step_philox_kernel<T, unsigned><<<(unsigned)blocks, 256, 0, stream>>>(
            x_out, x, xhat0, xi_out, s, t, e2, ome2, philox_dev, seed, off_z, off_xi, (unsigned)G, (int)iters, (unsigned)numel);
    else
        bridge_step_philox_kernel<T, long><<<(unsigned)blocks, 256, 0, stream>>>(x_out, x, xhat0, xi_out, s, t, e2, ome2, philox_dev,
                                                                                  seed, off_z, off_xi, G, (int)iters, numel);
    count_launch();
    return (int)cudaGetLastError();
}

// ---- scale in place ------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* __restrict__ y, const float* __restrict__ scale, size_t n) {
    const float sc = scale[0];
    if (sc == 1.0f) return;  // upstream gradient of exactly 1: the buffer already holds dloss/dxhat
    const size_t nvec = n / VEC;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        float a[VEC];
        load_pack<T, VEC>(y, (long)(v * VEC), a);
#pragma unroll
        for (int k = 0; k < VEC; ++k) a[k] *= sc;
        store_pack<T, VEC>(y, (long)(v * VEC), a);
    }
    // tail (n not a multiple of VEC)
    for (size_t i = nvec * VEC + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = Elem<T>::from_float(Elem<T>::to_float(y[i]) * sc);
}

template <typename T>
static int scale_inplace(T* y, const float* scale, size_t n, cudaStream_t stream) {
    if (!y || !scale) return DDDM_ERR_NULL_POINTER;
    if (n == 0) return DDDM_OK;
    constexpr int V = Elem<T>::kVec;
    const bool vec = aligned16(y);
    const size_t work = vec ? (n + V - 1) / V : n;
    size_t blocks = (work + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (vec)
        scale_inplace_kernel<T, V><<<(unsigned)blocks, 256, 0, stream>>>(y, scale, n);
    else
        scale_inplace_kernel<T, 1><<<(unsigned)blocks, 256, 0, stream>>>(y, scale, n);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace dddm

using namespace dddm;

extern "C" {

int dddm_forward_marginal_expand_f32(const float* x0, const float* t, const float* eps, float* xt, float* xt_rep,
                                     int B, int m, long D, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K2 forward_marginal_expand f32");
    return forward_marginal_expand<float>(x0, t, eps, xt, xt_rep, B, m, D, (cudaStream_t)stream);
}
int dddm_forward_marginal_expand_bf16(const dddm_bf16* x0, const float* t, const dddm_bf16* eps, dddm_bf16* xt,
                                      dddm_bf16* xt_rep, int B, int m, long D, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K2 forward_marginal_expand bf16");
    return forward_marginal_expand<__nv_bfloat16>((const __nv_bfloat16*)x0, t, (const __nv_bfloat16*)eps,
                                                  (__nv_bfloat16*)xt, (__nv_bfloat16*)xt_rep, B, m, D,
                                                  (cudaStream_t)stream);
}

int dddm_forward_marginal_concat_f32(const float* x0, const float* t, const float* eps, const float* xi, void* x6,
                                     int out_dtype, float* x0_tok, int B, int m, int C, int H, int W, int patch,
                                     dddm_stream_t stream) {
    DDDM_NVTX("dddm::K2c forward_marginal_concat f32");
    if (out_dtype == DDDM_DTYPE_F32)
        return forward_marginal_concat<float, float>(x0, t, eps, xi, (float*)x6, x0_tok, B, m, C, H, W, patch,
                                                     (cudaStream_t)stream);
    if (out_dtype == DDDM_DTYPE_BF16)
        return forward_marginal_concat<float, __nv_bfloat16>(x0, t, eps, xi, (__nv_bfloat16*)x6, x0_tok, B, m, C, H, W,
                                                             patch, (cudaStream_t)stream);
    return DDDM_ERR_BAD_ARGUMENT;
}
int dddm_forward_marginal_concat_bf16(const dddm_bf16* x0, const float* t, const dddm_bf16* eps, const dddm_bf16* xi,
                                      dddm_bf16* x6, dddm_bf16* x0_tok, int B, int m, int C, int H, int W, int patch,
                                      dddm_stream_t stream) {
    DDDM_NVTX("dddm::K2c forward_marginal_concat bf16");
    return forward_marginal_concat<__nv_bfloat16, __nv_bfloat16>(
        (const __nv_bfloat16*)x0, t, (const __nv_bfloat16*)eps, (const __nv_bfloat16*)xi, (__nv_bfloat16*)x6,
        (__nv_bfloat16*)x0_tok, B, m, C, H, W, patch, (cudaStream_t)stream);
}

int dddm_sigmoid_weight_sum_f32(const float* t, float bias, float* w, float* w_sum, int B, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K4 sigmoid_weight_sum");
    if (!t || (!w && !w_sum)) return DDDM_ERR_NULL_POINTER;
    if (B < 0) return DDDM_ERR_BAD_SHAPE;
    if (B == 0) {
        if (w_sum) return (int)cudaMemsetAsync(w_sum, 0, sizeof(float), (cudaStream_t)stream);
        return DDDM_OK;
    }
    int threads = B >= 1024 ? 1024 : (B + 31) / 32 * 32;
    sigmoid_weight_sum_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(t, bias, w, w_sum, B);
    count_launch();
    return (int)cudaGetLastError();
}

int dddm_bridge_step_f32(float* x_out, const float* x, const float* xhat0, const float* z, const float* s,
                         const float* t, int st_is_vector, double eps_churn, float* mu_out, float* std_out, long N,
                         long D, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K3 bridge_step f32");
    return bridge_step<float>(x_out, x, xhat0, z, s, t, st_is_vector, eps_churn, mu_out, std_out, N, D,
                              (cudaStream_t)stream);
}
int dddm_bridge_step_bf16(dddm_bf16* x_out, const dddm_bf16* x, const dddm_bf16* xhat0, const dddm_bf16* z,
                          const float* s, const float* t, int st_is_vector, double eps_churn, dddm_bf16* mu_out,
                          float* std_out, long N, long D, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K3 bridge_step bf16");
    return bridge_step<__nv_bfloat16>((__nv_bfloat16*)x_out, (const __nv_bfloat16*)x, (const __nv_bfloat16*)xhat0,
                                      (const __nv_bfloat16*)z, s, t, st_is_vector, eps_churn, (__nv_bfloat16*)mu_out,
                                      std_out, N, D, (cudaStream_t)stream);
}

unsigned long long dddm_philox_increment(long numel) {
    long G, iters;
    torch_philox_plan(numel, &G, &iters);
    return 4ull * (unsigned long long)iters;
}
int dddm_bridge_step_philox_f32(float* x_out, const float* x, const float* xhat0, float* xi_next, const float* s,
                                const float* t, double eps_churn, const unsigned long long* philox_dev,
                                unsigned long long seed, unsigned long long offset_z, unsigned long long offset_xi, long N,
                                long D, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K3 bridge_step + philox f32");
    return bridge_step_philox<float>(x_out, x, xhat0, xi_next, s, t, eps_churn, philox_dev, seed, offset_z, offset_xi, N, D,
                                     (cudaStream_t)stream);
}
int dddm_bridge_step_philox_bf16(dddm_bf16* x_out, const dddm_bf16* x, const dddm_bf16* xhat0, dddm_bf16* xi_next,
                                 const float* s, const float* t, double eps_churn, const unsigned long long* philox_dev,
                                 unsigned long long seed, unsigned long long offset_z, unsigned long long offset_xi, long N,
                                 long D, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K3 bridge_step + philox bf16");
    return bridge_step_philox<__nv_bfloat16>((__nv_bfloat16*)x_out, (const __nv_bfloat16*)x, (const __nv_bfloat16*)xhat0,
                                             (__nv_bfloat16*)xi_next, s, t, eps_churn, philox_dev, seed, offset_z, offset_xi,
                                             N, D, (cudaStream_t)stream);
}

int dddm_scale_inplace_f32(float* y, const float* scale, size_t n, dddm_stream_t stream) {
    return scale_inplace<float>(y, scale, n, (cudaStream_t)stream);
}
int dddm_scale_inplace_bf16(dddm_bf16* y, const float* scale, size_t n, dddm_stream_t stream) {
    return scale_inplace<__nv_bfloat16>((__nv_bfloat16*)y, scale, n, (cudaStream_t)stream);
}

}  // extern "C"
