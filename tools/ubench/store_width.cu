// store_width.cu — is an SM's global-store path bound per BYTE or per store INSTRUCTION?
//
// K1's pass 2 writes one row's gradient (98 KB fp32 / 49 KB bf16 per SM) with 8 store instructions per thread step and
// takes 2.1 us in both dtypes.  One CTA per SM writes 8 "rows" (stride 12288 B, as grad_xhat at D = 3072 fp32) of
// contiguous data from registers with 4-, 8-, 16- or 32-byte stores per lane (STG.32 / .64 / .128 / .256, the last one is
// new in sm_100), either the same BYTES (98 304 B) or the same number of INSTRUCTIONS; the issue time (globaltimer around
// the store loop) and the time until the writes are performed (after a __threadfence) are printed per width.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/store_width tools/ubench/store_width.cu
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int W>
__device__ __forceinline__ void st(unsigned char* p, unsigned v) {
    if constexpr (W == 4) asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if constexpr (W == 8) asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%1};" ::"l"(p), "r"(v) : "memory");
    if constexpr (W == 16) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    if constexpr (W == 32) asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}

// every thread: `iters` steps, 8 rows per step, W bytes per row and step; a warp's 32 lanes are contiguous
template <int W>
__global__ void push(unsigned char* dst, int iters, long row_stride, long cta_stride, long long* ns_issue, long long* ns_done) {
    dst += (long)blockIdx.x * cta_stride;
    unsigned long long t0, t1, t2;
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int it = 0; it < iters; ++it) {
        unsigned char* p = dst + ((long)it * blockDim.x + threadIdx.x) * W;
#pragma unroll
        for (int r = 0; r < 8; ++r) st<W>(p + r * row_stride, (unsigned)it + r);
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    __threadfence();
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t2));
    if (threadIdx.x == 0) {
        ns_issue[blockIdx.x] = (long long)(t1 - t0);
        ns_done[blockIdx.x] = (long long)(t2 - t0);
    }
}

int main() {
    unsigned char* buf;
    const long cta_stride = 8 * 12288 * 4;  // room for 32-byte stores x 12 steps per row
    CK(cudaMalloc(&buf, 148 * cta_stride));
    long long *ni, *nd;
    CK(cudaMalloc(&ni, 148 * 8));
    CK(cudaMalloc(&nd, 148 * 8));
    std::vector<long long> hi(148), hd(148);
    auto report = [&](const char* what, int W, int threads, int iters, int ctas) {
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hi.data(), ni, ctas * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hd.data(), nd, ctas * 8, cudaMemcpyDeviceToHost));
        std::sort(hi.begin(), hi.begin() + ctas);
        std::sort(hd.begin(), hd.begin() + ctas);
        const long bytes = (long)iters * 8 * threads * W, instr = (long)iters * 8 * (threads / 32);
        printf("%-12s STG.%-3d %4d threads %3d CTAs: %6ld B, %4ld warp-stores per SM: issue %5lld ns, performed %5lld ns "
               "(%.1f B/ns, %.1f ns per warp-store)\n", what, W * 8, threads, ctas, bytes, instr, hi[ctas / 2], hd[ctas / 2],
               bytes / (double)hd[ctas / 2], hd[ctas / 2] / (double)instr);
    };
    for (int rep = 0; rep < 2; ++rep)
        for (int ctas : {1, 128})
            for (int threads : {128, 256}) {
                // same bytes (98 304 B per SM)
                const int base = 98304 / (8 * threads);  // iterations x W
                push<4><<<ctas, threads>>>(buf, base / 4, 12288, cta_stride, ni, nd);   report("same bytes", 4, threads, base / 4, ctas);
                push<8><<<ctas, threads>>>(buf, base / 8, 12288, cta_stride, ni, nd);   report("same bytes", 8, threads, base / 8, ctas);
                push<16><<<ctas, threads>>>(buf, base / 16, 12288, cta_stride, ni, nd); report("same bytes", 16, threads, base / 16, ctas);
                push<32><<<ctas, threads>>>(buf, base / 32, 12288, cta_stride, ni, nd); report("same bytes", 32, threads, base / 32, ctas);
                // same instruction count (6 steps x 8 rows per thread, K1's pass 2)
                push<8><<<ctas, threads>>>(buf, 6, 12288, cta_stride, ni, nd);   report("same instr", 8, threads, 6, ctas);
                push<16><<<ctas, threads>>>(buf, 6, 12288, cta_stride, ni, nd);  report("same instr", 16, threads, 6, ctas);
                push<32><<<ctas, threads>>>(buf, 6, 12288, cta_stride, ni, nd);  report("same instr", 32, threads, 6, ctas);
            }
    return 0;
}
