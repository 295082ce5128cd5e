// common.cuh — shared device helpers for the DDDM sm_100a kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dddm_b200.h"

namespace dddm {

constexpr float kPowEps = 1e-12f;    // dddm/losses.py:14,24
constexpr float kWeightEps = 1e-12f; // dddm/losses.py:33-34
constexpr float kBridgeEps = 1e-8f;  // dddm/schedules.py:47

// ---- launch bookkeeping (api.cu) ---------------------------------------------------------
void count_launch();
struct Tuning {
    int cluster = 0;  // CTAs per row (0 = auto)
    int nv = 0;       // 16-byte vectors per thread (0 = auto)
    int variant = 0;  // 0 auto, 1 register-resident, 2 generic smem/TMA tile, 3 TMA-staged packed-fp32 (m <= 8),
                      // 4 blocked (m = 16, 32), 5 single-wave register-resident (m <= 8, one CTA per SM)
    int threads = 0;  // threads per CTA for variant 3 (0 = auto)
    int pdl = 1;      // programmatic dependent launch (the prologue overlaps the previous kernel's tail; every
                      // kernel waits on cudaGridDependencySynchronize() before touching global memory)
    int cols = 0;     // experiment: columns per thread step of the bf16 variant-3 kernel (0 = default)
    int ksmem = 0;    // single-wave kernel: pass-2 coefficients streamed from shared memory (1) or held in registers (0)
    int loader = 0;   // TMA-staged kernel: 0 auto, 1 TMA bulk copies, 2 cp.async commit groups (thread-private slots)
    int window = 0;   // cp.async loader: column chunks in flight per CTA (0 = default)
    int ldhint = 0;   // single-wave kernel: L2 eviction priority of the input loads (0 normal, 1 first, 2 last, 3 unchanged)
    int sthint = 0;   // ... and of the gradient stores
    int nostore = 0;  // diagnostics: pass 2 without its global stores (fp32/bf16 m = 8, 256 x 3 plan only)
    int ctas = 0;     // experiment: resident-CTA target the bf16 variant-3 kernel is compiled for (0 = default)
    int finish = 0;   // TMA-staged kernel: cross-row sum 0 auto (polled row slots), 1 arrival ticket, 2 polled row slots
    int nvtx = -1;    // NVTX ranges around the C-ABI entry points: -1 = environment DDDM_NVTX, 0 off, 1 on
    void* trace = nullptr;  // device buffer for in-kernel timeline stamps (diagnostics)
};
Tuning& tuning();

// ---- NVTX ranges around every C-ABI entry point (tracing), off unless DDDM_NVTX=1 or
//      dddm_set_tuning("nvtx", 1): nsys / ncu --nvtx then show K1..K6 by name ----------------------------------
bool nvtx_enabled();
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name) : on(nvtx_enabled()) {
        if (on) nvtx_push(name);
    }
    ~NvtxRange() {
        if (on) nvtx_pop();
    }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define DDDM_NVTX(name) ::dddm::NvtxRange dddm_nvtx_range_(name)

// ---- per-device host-side caches (a process may drive several GPUs; function attributes are per device) -------
int device_sm_count();  // multiprocessors of the CURRENT device (api.cu)
constexpr int kMaxDevices = 64;
// Largest dynamic shared-memory size a kernel has been opted into, per device.
struct SmemOptIn {
    size_t per_device[kMaxDevices] = {};
    template <typename K>
    int ensure(K kernel, size_t bytes, size_t threshold) {
        if (bytes <= threshold) return 0;
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = -1;
        if (dev >= 0 && bytes <= per_device[dev]) return 0;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0) per_device[dev] = bytes;  // benign race: a concurrent caller at worst repeats the opt-in
        return 0;
    }
};

// ---- element traits ----------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static constexpr int kVec = 4;  // elements per 16-byte vector
    __device__ static __forceinline__ float to_float(float v) { return v; }
    __device__ static __forceinline__ float from_float(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    __device__ static __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};

// A group of VEC consecutive elements held as fp32 in registers.
template <typename T, int VEC>
struct Pack {
    float v[VEC];
};

// 16-byte streaming load (read once: do not allocate in L1) / store.
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ void stg_stream8(void* p, const uint2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream4(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Load VEC elements starting at element index `e` of row pointer `p` into fp32 registers.
template <typename T, int VEC>
__device__ __forceinline__ void load_pack(const T* __restrict__ p, long e, float (&out)[VEC]) {
    if constexpr (VEC == 1) {
        out[0] = Elem<T>::to_float(p[e]);
    } else if constexpr (sizeof(T) == 4) {
        static_assert(VEC == 4, "fp32 vectors are 4 wide");
        uint4 r = ldg_stream16(p + e);
        out[0] = __uint_as_float(r.x);
        out[1] = __uint_as_float(r.y);
        out[2] = __uint_as_float(r.z);
        out[3] = __uint_as_float(r.w);
    } else {
        static_assert(VEC == 8, "bf16 vectors are 8 wide");
        uint4 r = ldg_stream16(p + e);
        out[0] = bf16lo(r.x); out[1] = bf16hi(r.x);
        out[2] = bf16lo(r.y); out[3] = bf16hi(r.y);
        out[4] = bf16lo(r.z); out[5] = bf16hi(r.z);
        out[6] = bf16lo(r.w); out[7] = bf16hi(r.w);
    }
}

template <typename T, int VEC>
__device__ __forceinline__ void store_pack(T* __restrict__ p, long e, const float (&in)[VEC]) {
    if constexpr (VEC == 1) {
        p[e] = Elem<T>::from_float(in[0]);
    } else if constexpr (sizeof(T) == 4) {
        uint4 r;
        r.x = __float_as_uint(in[0]); r.y = __float_as_uint(in[1]);
        r.z = __float_as_uint(in[2]); r.w = __float_as_uint(in[3]);
        stg_stream16(p + e, r);
    } else {
        uint4 r;
        r.x = pack_bf16x2(in[0], in[1]); r.y = pack_bf16x2(in[2], in[3]);
        r.z = pack_bf16x2(in[4], in[5]); r.w = pack_bf16x2(in[6], in[7]);
        stg_stream16(p + e, r);
    }
}

// ---- the beta-power term -----------------------------------------------------------------
// mode: 0 general pow, 1 beta == 1 (sqrt), 2 beta == 2.0 exactly (identity, no epsilon).
struct PowSpec {
    float half_beta;  // beta / 2
    int mode;
};
__host__ __device__ inline PowSpec make_pow_spec(float beta) {
    PowSpec s;
    s.half_beta = 0.5f * beta;
    s.mode = (beta == 2.0f) ? 2 : (beta == 1.0f ? 1 : 0);
    return s;
}
// f(d2) of dddm/losses.py:11-14, 21-24
__device__ __forceinline__ float pow_value(float d2, const PowSpec& s) {
    if (s.mode == 2) return d2;
    float x = d2 + kPowEps;
    return s.mode == 1 ? sqrtf(x) : powf(x, s.half_beta);
}
// f'(d2)
__device__ __forceinline__ float pow_deriv(float d2, const PowSpec& s) {
    if (s.mode == 2) return 1.0f;
    float x = d2 + kPowEps;
    return s.mode == 1 ? 0.5f * rsqrtf(x) : s.half_beta * powf(x, s.half_beta - 1.0f);
}

// f(d2) and f'(d2) together, off the slow path: powf + an IEEE division cost ~600 ns of dependent
// latency per row on the kernel's critical path (tools/trace_energy.py).  Here x^h = 2^(h*log2 x) with
// the 1-ulp log2f, the rounding error of the product h*log2(x) carried as a first-order correction,
// and f' = h * f / x through the reciprocal unit.  Relative error <= ~1e-6 for beta < 2 over the whole
// range [1e-12, 1e30] (|log2 x| < 128), ~1e-7 for beta = 0.1 — inside the 1e-5 budget of the path.
__device__ __forceinline__ void pow_value_deriv(float d2, const PowSpec& s, float& val, float& der) {
    if (s.mode == 2) {
        val = d2;
        der = 1.0f;
        return;
    }
    const float x = d2 + kPowEps;
    if (s.mode == 1) {
        val = sqrtf(x);
        der = 0.5f * __fdividef(val, x);
        return;
    }
    const float l = log2f(x);
    const float e_hi = s.half_beta * l;
    const float e_lo = fmaf(s.half_beta, l, -e_hi);
    val = exp2f(e_hi) * fmaf(e_lo, 0.693147181f, 1.0f);
    der = s.half_beta * __fdividef(val, x);
}

// ---- warp-level reduction of P per-lane partial sums ----------------------------------------
// Butterfly "reduce-scatter": instead of 5 shuffles per value (5*P), halve the number of live
// values at every butterfly level, which costs P2-1 shuffles for P2 (a power of two) values.
// After reducing a chunk of P2 <= 32 values, lane l holds the warp total of slot (l % P2).
template <int P, int BASE, int N>
__device__ __forceinline__ void fold_step(float (&v)[P], int lane) {
    // N live values at v[BASE .. BASE+N); lanes with (lane & N/2) keep the upper half.
    constexpr int H = N / 2;
    const bool upper = (lane & H) != 0;
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float keep = upper ? v[BASE + k + H] : v[BASE + k];
        float send = upper ? v[BASE + k] : v[BASE + k + H];
        v[BASE + k] = keep + __shfl_xor_sync(0xffffffffu, send, H);
    }
    if constexpr (H > 1) fold_step<P, BASE, H>(v, lane);
}

template <int P, int BASE, int P2>
__device__ __forceinline__ float reduce_scatter_chunk(float (&v)[P], int lane) {
    static_assert(P2 >= 1 && P2 <= 32 && (P2 & (P2 - 1)) == 0, "chunk must be a power of two <= 32");
    // butterfly levels wider than the chunk: plain all-reduce of every value
#pragma unroll
    for (int off = 16; off >= P2; off >>= 1) {
#pragma unroll
        for (int k = 0; k < P2; ++k) v[BASE + k] += __shfl_xor_sync(0xffffffffu, v[BASE + k], off);
    }
    if constexpr (P2 > 1) fold_step<P, BASE, P2>(v, lane);
    return v[BASE];
}

constexpr int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// Reduce acc[0..P) across the warp and write the P totals to dst[0..P) (shared memory).
// acc is clobbered.  PPAD = P rounded so that the tail chunk is a power of two.
template <int P>
struct WarpReduce {
    static constexpr int kFull = P / 32;
    static constexpr int kRem = P % 32;
    static constexpr int kRemP2 = kRem ? next_pow2(kRem) : 0;
    static constexpr int kPadded = kFull * 32 + kRemP2;

    template <int C>
    __device__ static __forceinline__ void full_chunks(float (&v)[kPadded], float* dst, int lane) {
        if constexpr (C < kFull) {
            float t = reduce_scatter_chunk<kPadded, C * 32, 32>(v, lane);
            dst[C * 32 + lane] = t;
            full_chunks<C + 1>(v, dst, lane);
        }
    }
    // v must have kPadded entries with v[P..kPadded) == 0.
    __device__ static __forceinline__ void run(float (&v)[kPadded], float* dst, int lane) {
        full_chunks<0>(v, dst, lane);
        if constexpr (kRem > 0) {
            float t = reduce_scatter_chunk<kPadded, kFull * 32, kRemP2>(v, lane);
            if (lane < kRem) dst[kFull * 32 + lane] = t;
        }
    }
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

}  // namespace dddm
