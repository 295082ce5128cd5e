// backbone_ops.cu — two HBM-bound kernels for the DiT training step around the hot path.
// The backbone stays PyTorch (GEMMs: cuBLAS, attention: SDPA); these replace the two items that dominated its
// profile at 65 536 tokens x 384 channels per step (profiles/r01_dit.md):
//   LayerNorm forward / backward  — ATen's gamma/beta backward kernel took 0.56 ms per LayerNorm at this
//                                   shape (17 per step); here forward = one pass, backward = one pass that also
//                                   produces per-CTA partial sums of dgamma / dbeta, folded by a second tiny launch
//                                   in a fixed order (deterministic, no atomics);
//   column sum                    — bias gradients sum_rows dY[rows, C] (88 reductions per step).
// Arithmetic: fp32 inside, fp32 or bf16 storage; LayerNorm as torch.nn.functional.layer_norm (biased variance,
// eps inside the square root).
#include <type_traits>

#include "common.cuh"

namespace dddm {

constexpr int kLnMaxVec = 8;     // 4-element vectors cached per lane: C <= 32 * 4 * 8 = 1024
constexpr int kLnWarps = 8;      // warps (rows in flight) per CTA

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        const float4 r = *reinterpret_cast<const float4*>(p);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    } else {
        const uint2 r = *reinterpret_cast<const uint2*>(p);
        v[0] = bf16lo(r.x); v[1] = bf16hi(r.x); v[2] = bf16lo(r.y); v[3] = bf16hi(r.y);
    }
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
}

// One warp per row, the row cached in registers (NV vectors of 4 per lane, nvec = C / 4 <= 32 * NV).
template <typename T, int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
layer_norm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ y,
                      float* __restrict__ mean, float* __restrict__ rstd, long N, int C, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = C / 4;
    float g[NV][4], bt[NV][4];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + 32 * k;
        if (v < nvec) {
            ld4<T>(gamma + 4 * v, g[k]);
            ld4<T>(beta + 4 * v, bt[k]);
        }
    }
    const float inv_c = 1.0f / (float)C;
    for (long r = (long)blockIdx.x * kLnWarps + warp; r < N; r += (long)gridDim.x * kLnWarps) {
        const T* xr = x + r * C;
        float xv[NV][4];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
                ld4<T>(xr + 4 * v, xv[k]);
                s += (xv[k][0] + xv[k][1]) + (xv[k][2] + xv[k][3]);
            }
        }
        const float mu = warp_sum(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = xv[k][e] - mu;
                    q = fmaf(d, d, q);
                }
            }
        }
        const float rs = rsqrtf(warp_sum(q) * inv_c + eps);
        if (lane == 0) {
            mean[r] = mu;
            rstd[r] = rs;
        }
        T* yr = y + r * C;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = fmaf((xv[k][e] - mu) * rs, g[k][e], bt[k][e]);
                st4<T>(yr + 4 * v, o);
            }
        }
    }
}

// 8-byte (bf16) / 16-byte (fp32) raw vectors: loaded one row ahead, unpacked when used
template <typename T>
struct Raw4 {
    using type = typename std::conditional<sizeof(T) == 4, float4, uint2>::type;
};
template <typename T>
__device__ __forceinline__ void unpack4(const typename Raw4<T>::type& r, float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    } else {
        v[0] = bf16lo(r.x); v[1] = bf16hi(r.x); v[2] = bf16lo(r.y); v[3] = bf16hi(r.y);
    }
}

// Backward: dx per row; per-CTA partial sums of dgamma = sum dy * xhat and dbeta = sum dy into part[cta][2][C].
// One warp per row with the NEXT row's x / dy already in flight while the current row is reduced (a row is only
// 2 x 768 B: without the prefetch each warp has one row's worth of loads outstanding and the kernel sits at 0.35 of
// the HBM peak).
template <typename T, int NV>
__global__ void __launch_bounds__(kLnWarps * 32, 2)
layer_norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const T* __restrict__ gamma, T* __restrict__ dx,
                      float* __restrict__ part, long N, int C) {
    using R = typename Raw4<T>::type;
    extern __shared__ float s_part[];  // [kLnWarps][2][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = C / 4;
    float g[NV][4], dg[NV][4], db[NV][4];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + 32 * k;
        if (v < nvec) ld4<T>(gamma + 4 * v, g[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) dg[k][e] = db[k][e] = 0.f;
    }
    const float inv_c = 1.0f / (float)C;
    const long stride = (long)gridDim.x * kLnWarps;
    long r = (long)blockIdx.x * kLnWarps + warp;
    R cx[NV], cd[NV];
    float cmu = 0.f, crs = 0.f;
    auto fetch = [&](long row, R (&fx)[NV], R (&fd)[NV], float& mu, float& rs) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
                fx[k] = *reinterpret_cast<const R*>(x + row * C + 4 * v);
                fd[k] = *reinterpret_cast<const R*>(dy + row * C + 4 * v);
            }
        }
        mu = mean[row];
        rs = rstd[row];
    };
    if (r < N) fetch(r, cx, cd, cmu, crs);
    while (r < N) {
        const long rn = r + stride;
        R nx[NV], nd[NV];
        float nmu = 0.f, nrs = 0.f;
        if (rn < N) fetch(rn, nx, nd, nmu, nrs);
        const float mu = cmu, rs = crs;
        float xh[NV][4], gy[NV][4];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
                float xv[4], dv[4];
                unpack4<T>(cx[k], xv);
                unpack4<T>(cd[k], dv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    xh[k][e] = (xv[e] - mu) * rs;
                    gy[k][e] = dv[e] * g[k][e];
                    c1 += gy[k][e];
                    c2 = fmaf(gy[k][e], xh[k][e], c2);
                    dg[k][e] = fmaf(dv[e], xh[k][e], dg[k][e]);
                    db[k][e] += dv[e];
                }
            }
        }
        c1 = warp_sum(c1) * inv_c;
        c2 = warp_sum(c2) * inv_c;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = rs * (gy[k][e] - c1 - xh[k][e] * c2);
                st4<T>(dx + r * C + 4 * v, o);
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            cx[k] = nx[k];
            cd[k] = nd[k];
        }
        cmu = nmu;
        crs = nrs;
        r = rn;
    }
    // fold the warps of this CTA in a fixed order, then publish the CTA's partial
    float* mine = s_part + (size_t)warp * 2 * C;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + 32 * k;
        if (v < nvec) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                mine[4 * v + e] = dg[k][e];
                mine[C + 4 * v + e] = db[k][e];
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < kLnWarps; ++w) t += s_part[(size_t)w * 2 * C + c];
        part[(size_t)blockIdx.x * 2 * C + c] = t;
    }
}

// out[c] = sum_p part[p][c] in a fixed order, c < width; written as T.  32 columns x 8 part-phases per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
fold_partials_kernel(const float* __restrict__ part, int nparts, int width, T* __restrict__ out0, T* __restrict__ out1,
                     int split) {
    __shared__ float s_acc[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a0 = 0.f, a1 = 0.f;
    if (c < width) {
        int p = ty;
        for (; p + 8 < nparts; p += 16) {
            a0 += part[(size_t)p * width + c];
            a1 += part[(size_t)(p + 8) * width + c];
        }
        if (p < nparts) a0 += part[(size_t)p * width + c];
    }
    s_acc[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty == 0 && c < width) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s_acc[k][tx];
        if (c < split) out0[c] = Elem<T>::from_float(t);
        else out1[c - split] = Elem<T>::from_float(t);
    }
}

// Column sums of a [N, C] matrix: CTA (bx, by) sums rows by, by + gridDim.y, ... of 4-column groups.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ a, float* __restrict__ part, long N, int C) {
    const int nvec = C / 4;
    const int v = blockIdx.x * 64 + (threadIdx.x & 63);   // 64 column groups per CTA
    const int sub = threadIdx.x >> 6;                      // 4 row phases per CTA
    __shared__ float s_acc[4][64][4];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (v < nvec) {
#pragma unroll 4
        for (long r = (long)blockIdx.y * 4 + sub; r < N; r += (long)gridDim.y * 4) {
            float t[4];
            ld4<T>(a + r * C + 4 * v, t);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] += t[e];
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) s_acc[sub][threadIdx.x & 63][e] = acc[e];
    __syncthreads();
    if (sub == 0 && v < nvec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int l = threadIdx.x & 63;
            part[(size_t)blockIdx.y * C + 4 * v + e] = (s_acc[0][l][e] + s_acc[1][l][e]) + (s_acc[2][l][e] + s_acc[3][l][e]);
        }
    }
}

static int sms() { return device_sm_count(); }  // per-device cache in api.cu
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T, int NV>
static int ln_fwd_nv(const T* x, const T* gamma, const T* beta, T* y, float* mean, float* rstd, long N, int C, float eps,
                     cudaStream_t stream) {
    const long want = (N + kLnWarps - 1) / kLnWarps;
    const int grid = (int)(want < (long)sms() * 8 ? want : (long)sms() * 8);
    layer_norm_fwd_kernel<T, NV><<<grid, kLnWarps * 32, 0, stream>>>(x, gamma, beta, y, mean, rstd, N, C, eps);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename T>
static int ln_fwd(const T* x, const T* gamma, const T* beta, T* y, float* mean, float* rstd, long N, int C, float eps,
                  cudaStream_t stream) {
    if (!x || !gamma || !beta || !y || !mean || !rstd) return DDDM_ERR_NULL_POINTER;
    if (N < 0 || C < 4 || C % 4 != 0 || C > 32 * 4 * kLnMaxVec) return DDDM_ERR_UNSUPPORTED;
    if (!al16(x) || !al16(y) || !al16(gamma) || !al16(beta) || ((long)C * (long)sizeof(T)) % 16 != 0) return DDDM_ERR_BAD_ALIGNMENT;
    if (N == 0) return DDDM_OK;
    const int nv = (C / 4 + 31) / 32;
    if (nv <= 1) return ln_fwd_nv<T, 1>(x, gamma, beta, y, mean, rstd, N, C, eps, stream);
    if (nv <= 2) return ln_fwd_nv<T, 2>(x, gamma, beta, y, mean, rstd, N, C, eps, stream);
    if (nv <= 3) return ln_fwd_nv<T, 3>(x, gamma, beta, y, mean, rstd, N, C, eps, stream);
    if (nv <= 4) return ln_fwd_nv<T, 4>(x, gamma, beta, y, mean, rstd, N, C, eps, stream);
    return ln_fwd_nv<T, kLnMaxVec>(x, gamma, beta, y, mean, rstd, N, C, eps, stream);
}

template <typename T, int NV>
static int ln_bwd_nv(const T* dy, const T* x, const float* mean, const float* rstd, const T* gamma, T* dx, T* dgamma,
                     T* dbeta, float* part, int grid, long N, int C, cudaStream_t stream) {
    auto kernel = layer_norm_bwd_kernel<T, NV>;
    const size_t smem = (size_t)kLnWarps * 2 * C * sizeof(float);
    static SmemOptIn configured;  // per instantiation and per device
    if (int e = configured.ensure(kernel, smem, 48 * 1024)) return e;
    kernel<<<grid, kLnWarps * 32, smem, stream>>>(dy, x, mean, rstd, gamma, dx, part, N, C);
    count_launch();
    fold_partials_kernel<T><<<(2 * C + 31) / 32, 256, 0, stream>>>(part, grid, 2 * C, dgamma, dbeta, C);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename T>
static int ln_bwd(const T* dy, const T* x, const float* mean, const float* rstd, const T* gamma, T* dx, T* dgamma,
                  T* dbeta, float* part, size_t part_bytes, long N, int C, cudaStream_t stream) {
    if (!dy || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || !part) return DDDM_ERR_NULL_POINTER;
    if (N < 1 || C < 4 || C % 4 != 0 || C > 32 * 4 * kLnMaxVec) return DDDM_ERR_UNSUPPORTED;
    if (!al16(x) || !al16(dy) || !al16(dx) || !al16(gamma) || ((long)C * (long)sizeof(T)) % 16 != 0) return DDDM_ERR_BAD_ALIGNMENT;
    const long want = (N + kLnWarps - 1) / kLnWarps;
    int grid = (int)(want < (long)sms() * 4 ? want : (long)sms() * 4);  // 4 CTAs x 8 warps per SM: enough rows in flight
    const size_t per = (size_t)2 * C * sizeof(float);
    if (part_bytes < per) return DDDM_ERR_BAD_ARGUMENT;
    if ((size_t)grid * per > part_bytes) grid = (int)(part_bytes / per);
    const int nv = (C / 4 + 31) / 32;
    if (nv <= 1) return ln_bwd_nv<T, 1>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, part, grid, N, C, stream);
    if (nv <= 2) return ln_bwd_nv<T, 2>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, part, grid, N, C, stream);
    if (nv <= 3) return ln_bwd_nv<T, 3>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, part, grid, N, C, stream);
    if (nv <= 4) return ln_bwd_nv<T, 4>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, part, grid, N, C, stream);
    return ln_bwd_nv<T, kLnMaxVec>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, part, grid, N, C, stream);
}

template <typename T>
static int colsum(const T* a, T* out, float* part, size_t part_bytes, long N, int C, cudaStream_t stream) {
    if (!a || !out || !part) return DDDM_ERR_NULL_POINTER;
    if (N < 1 || C < 4 || C % 4 != 0) return DDDM_ERR_UNSUPPORTED;
    if (!al16(a) || ((long)C * (long)sizeof(T)) % 16 != 0) return DDDM_ERR_BAD_ALIGNMENT;
    const int gx = (C / 4 + 63) / 64;
    long gy = ((long)sms() * 4 + gx - 1) / gx;
    if (gy > (N + 3) / 4) gy = (N + 3) / 4;
    const size_t per = (size_t)C * sizeof(float);
    if (part_bytes < per) return DDDM_ERR_BAD_ARGUMENT;
    if ((size_t)gy * per > part_bytes) gy = (long)(part_bytes / per);
    colsum_partial_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(a, part, N, C);
    count_launch();
    fold_partials_kernel<T><<<(C + 31) / 32, 256, 0, stream>>>(part, (int)gy, C, out, out, C);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace dddm

using namespace dddm;
using bf16 = __nv_bfloat16;

extern "C" {

size_t dddm_backbone_scratch_bytes(int C) {
    if (C < 1) C = 1;
    return (size_t)148 * 4 * 2 * (size_t)C * sizeof(float);  // enough for every launch plan of the kernels below
}

int dddm_layer_norm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                            long N, int C, float eps, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K6 layer_norm_fwd f32");
    return ln_fwd<float>(x, gamma, beta, y, mean, rstd, N, C, eps, (cudaStream_t)stream);
}
int dddm_layer_norm_fwd_bf16(const dddm_bf16* x, const dddm_bf16* gamma, const dddm_bf16* beta, dddm_bf16* y, float* mean,
                             float* rstd, long N, int C, float eps, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K6 layer_norm_fwd bf16");
    return ln_fwd<bf16>((const bf16*)x, (const bf16*)gamma, (const bf16*)beta, (bf16*)y, mean, rstd, N, C, eps,
                        (cudaStream_t)stream);
}
int dddm_layer_norm_bwd_f32(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                            float* dx, float* dgamma, float* dbeta, float* scratch, size_t scratch_bytes, long N, int C,
                            dddm_stream_t stream) {
    DDDM_NVTX("dddm::K6 layer_norm_bwd f32");
    return ln_bwd<float>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, scratch, scratch_bytes, N, C, (cudaStream_t)stream);
}
int dddm_layer_norm_bwd_bf16(const dddm_bf16* dy, const dddm_bf16* x, const float* mean, const float* rstd,
                             const dddm_bf16* gamma, dddm_bf16* dx, dddm_bf16* dgamma, dddm_bf16* dbeta, float* scratch,
                             size_t scratch_bytes, long N, int C, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K6 layer_norm_bwd bf16");
    return ln_bwd<bf16>((const bf16*)dy, (const bf16*)x, mean, rstd, (const bf16*)gamma, (bf16*)dx, (bf16*)dgamma,
                        (bf16*)dbeta, scratch, scratch_bytes, N, C, (cudaStream_t)stream);
}
int dddm_colsum_f32(const float* a, float* out, float* scratch, size_t scratch_bytes, long N, int C, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K6 colsum f32");
    return colsum<float>(a, out, scratch, scratch_bytes, N, C, (cudaStream_t)stream);
}
int dddm_colsum_bf16(const dddm_bf16* a, dddm_bf16* out, float* scratch, size_t scratch_bytes, long N, int C,
                     dddm_stream_t stream) {
    DDDM_NVTX("dddm::K6 colsum bf16");
    return colsum<bf16>((const bf16*)a, (bf16*)out, scratch, scratch_bytes, N, C, (cudaStream_t)stream);
}

}  // extern "C"
