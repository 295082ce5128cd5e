// metrics.cu — evaluation-side pairwise kernel sums for rbf_mmd2 (dddm/metrics.py:140-163).
// The reference forms three full Gram matrices a @ b.T, turns them into squared distances
// a2 + b2 - 2 a.b, exponentiates, masks the diagonal with a boolean gather and takes means: five n x n
// temporaries per term.  Here the Gram tile comes from the library GEMM (cuBLAS through torch.matmul — a plain
// GEMM, the one place a library call is the right tool) and ONE pass over it produces
//   sum_{i,j} [row_i + shift != col_j] exp(-gamma * (a2_i + b2_j - 2 G_ij))
// without materialising anything else; partial sums are folded in double in a fixed order (deterministic).
#include "common.cuh"

namespace dddm {

// a2[i] = sum_k x[i,k]^2, one warp per row (fp32, like the reference's (a * a).sum(-1))
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ x, float* __restrict__ out, long n, long D) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* xr = x + row * D;
    float s = 0.f;
    for (long k = lane; k < D; k += 32) s = fmaf(xr[k], xr[k], s);
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// grid (col tiles of 1024, row groups); each thread owns 4 columns and walks the rows of its group
__global__ void __launch_bounds__(256)
rbf_sum_partial_kernel(const float* __restrict__ G, long ldg, const float* __restrict__ a2, const float* __restrict__ b2,
                       long rows, long cols, float gamma, long shift, int skip_diag, double* __restrict__ part) {
    __shared__ double s_red[8];
    const long c0 = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
    float bcol[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (c0 + 3 < cols) && (ldg % 4 == 0) && ((reinterpret_cast<uintptr_t>(G) & 15u) == 0);
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (c0 + e < cols) bcol[e] = b2[c0 + e];
    float acc = 0.f;
    double total = 0.0;
    int since = 0;
    for (long r = blockIdx.y; r < rows; r += gridDim.y) {
        const float ar = a2[r];
        const float* g = G + r * ldg + c0;
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec) {
            const float4 t = *reinterpret_cast<const float4*>(g);
            gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (c0 + e < cols) gv[e] = g[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (c0 + e < cols && !(skip_diag && r + shift == c0 + e)) {
                const float d2 = (ar + bcol[e]) - 2.0f * gv[e];  // metrics.py:146: a2 + b2 - 2.0 * (a @ b.T)
                acc += expf(-gamma * d2);
            }
        }
        if (++since == 64) {  // bound the fp32 partial before it loses digits
            total += (double)acc;
            acc = 0.f;
            since = 0;
        }
    }
    total += (double)acc;
    // CTA reduction in double, fixed order
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) fold_double_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
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

int dddm_row_sqnorm_f32(const float* x, float* out, long n, long D, dddm_stream_t stream) {
    if (!x || !out) return DDDM_ERR_NULL_POINTER;
    if (n < 0 || D < 1) return DDDM_ERR_BAD_SHAPE;
    if (n == 0) return DDDM_OK;
    row_sqnorm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, out, n, D);
    count_launch();
    return (int)cudaGetLastError();
}

size_t dddm_rbf_scratch_bytes(long rows, long cols) {
    (void)rows;
    const long gx = (cols + 1023) / 1024;
    return (size_t)(gx < 1 ? 1 : gx) * 1184 * sizeof(double);
}

int dddm_rbf_kernel_sum_f32(const float* G, long ldg, const float* a2, const float* b2, long rows, long cols, float gamma,
                            long diag_shift, int skip_diag, double* scratch, size_t scratch_bytes, double* out,
                            dddm_stream_t stream) {
    DDDM_NVTX("dddm::K5 rbf_kernel_sum");
    if (!G || !a2 || !b2 || !scratch || !out) return DDDM_ERR_NULL_POINTER;
    if (rows < 1 || cols < 1 || ldg < cols) return DDDM_ERR_BAD_SHAPE;
    const long gx = (cols + 1023) / 1024;
    long gy = (148L * 8 + gx - 1) / gx;
    if (gy > rows) gy = rows;
    if (gy > 65535) gy = 65535;
    while ((size_t)(gx * gy) * sizeof(double) > scratch_bytes && gy > 1) --gy;
    if ((size_t)(gx * gy) * sizeof(double) > scratch_bytes || gx > 2147483647L) return DDDM_ERR_BAD_ARGUMENT;
    rbf_sum_partial_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(
        G, ldg, a2, b2, rows, cols, gamma, diag_shift, skip_diag, scratch);
    count_launch();
    fold_double_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scratch, (int)(gx * gy), out);
    count_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
