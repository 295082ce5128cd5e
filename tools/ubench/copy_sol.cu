// copy_sol.cu — "speed of light" of ONE dependent launch that moves the bytes of K1 at B=128, m=8, D=3072:
// read 14.2 MB (cold), write 12.6 MB, back-to-back launches on one stream over rotating buffer sets (> L2), with and
// without programmatic dependent launch.  No arithmetic, no row dependency: what the memory system and the launch
// boundary alone cost at this size.  Also: how fast ONE SM can pull its 110 KB tile (bulk copies / LDG.128).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o copy_sol copy_sol.cu && ./copy_sol
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// read nread 16-byte vectors, write nwrite of them (the first nwrite vectors read are written back out)
template <int UNROLL>
__global__ void copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long nread, long nwrite, int pdl) {
    if (pdl) cudaGridDependencySynchronize();
    const long stride = (long)gridDim.x * blockDim.x;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (; i + (UNROLL - 1) * stride < nread; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(in + i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long j = i + u * stride;
            if (j < nwrite) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + j), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
            else { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
        }
    }
    for (; i < nread; i += stride) {
        uint4 v = in[i];
        if (i < nwrite) out[i] = v; else acc.x ^= v.x;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345677u) out[0] = acc;  // keep the reads alive
    if (pdl) cudaTriggerProgrammaticLaunchCompletion();
}

// one CTA pulls `bytes` into shared memory with `ncopies` bulk copies; stamps ns from start to all landed
__global__ void pull_bulk(const unsigned char* __restrict__ src, int bytes, int ncopies, long long* ns_out, long cta_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    src += (long)blockIdx.x * cta_stride;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    __syncthreads();
    const int per = bytes / ncopies;
    for (int c = threadIdx.x; c < ncopies; c += blockDim.x) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(smem + (long)c * per)),
                     "l"(src + (long)c * per), "r"(per), "r"(bar_a) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar_a) : "memory");
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) ns_out[blockIdx.x] = (long long)(t1 - t0);
}

// one CTA pulls `bytes` with 16-byte loads into registers (xor-reduced), `threads` threads, 8 loads in flight per thread
__global__ void pull_ldg(const uint4* __restrict__ src, int nvec, long long* ns_out, long cta_stride_vec, unsigned* sink) {
    src += (long)blockIdx.x * cta_stride_vec;
    unsigned long long t0, t1;
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    unsigned acc = 0;
    int i = threadIdx.x;
    for (; i + 7 * (int)blockDim.x < nvec; i += 8 * blockDim.x) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * blockDim.x));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < nvec; i += blockDim.x) acc ^= src[i].x;
    if (acc == 0x12345677u) sink[0] = acc;
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) ns_out[blockIdx.x] = (long long)(t1 - t0);
}

// one CTA writes `bytes` from shared memory to global: 16-byte stores from `threads` threads, or `ncopies` bulk stores
__global__ void push_stg(unsigned char* __restrict__ dst, int bytes, long long* ns_out, long cta_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    dst += (long)blockIdx.x * cta_stride;
    for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(i, i, i, i);
    __syncthreads();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(smem + i);
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    __threadfence();
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) ns_out[blockIdx.x] = (long long)(t1 - t0);
}
__global__ void push_bulk(unsigned char* __restrict__ dst, int bytes, int ncopies, long long* ns_out, long cta_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    dst += (long)blockIdx.x * cta_stride;
    for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(i, i, i, i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    const int per = bytes / ncopies;
    for (int c = threadIdx.x; c < ncopies; c += blockDim.x) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (long)c * per),
                     "r"((unsigned)__cvta_generic_to_shared(smem + (long)c * per)), "r"(per) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) ns_out[blockIdx.x] = (long long)(t1 - t0);
}

int main(int argc, char** argv) {
    const bool quick = argc > 1;
    const long read_bytes = (128L * 8 * 3072 + 128L * 3072) * 4, write_bytes = 128L * 8 * 3072 * 4;
    const int nsets = 40;
    std::vector<uint4*> in(nsets), out(nsets);
    for (int s = 0; s < nsets; ++s) {
        CK(cudaMalloc(&in[s], read_bytes));
        CK(cudaMalloc(&out[s], write_bytes));
        CK(cudaMemset(in[s], s + 1, read_bytes));
    }
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](int grid, int threads, int unroll, int pdl, bool with_write, int kind) {
        // kind 0: copy kernel, 1: cudaMemcpyAsync D2D of the write bytes
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int rep = 0; rep < 6; ++rep)
            for (int s = 0; s < nsets; ++s) {
                if (kind == 1) { CK(cudaMemcpyAsync(out[s], in[s], write_bytes, cudaMemcpyDeviceToDevice, st)); continue; }
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(threads);
                cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at;
                cfg.numAttrs = pdl ? 1 : 0;
                const long nr = read_bytes / 16, nw = with_write ? write_bytes / 16 : 0;
                if (unroll == 4) CK(cudaLaunchKernelEx(&cfg, copy_kernel<4>, (const uint4*)in[s], out[s], nr, nw, pdl));
                else CK(cudaLaunchKernelEx(&cfg, copy_kernel<8>, (const uint4*)in[s], out[s], nr, nw, pdl));
            }
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
        CK(cudaStreamSynchronize(st));
        std::vector<float> t;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0, st));
            for (int k = 0; k < 5; ++k) CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            t.push_back(ms * 1e3f / (5 * 6 * nsets));
        }
        std::sort(t.begin(), t.end());
        CK(cudaGraphExecDestroy(ge));
        CK(cudaGraphDestroy(g));
        return t[2];
    };
    printf("bytes per launch: read %ld + write %ld = %ld; 6452.5 GB/s -> %.2f us\n", read_bytes, write_bytes, read_bytes + write_bytes,
           (read_bytes + write_bytes) / 6452.5e3);
    if (!quick) printf("cudaMemcpyAsync D2D of %ld bytes (%ld moved), one stream back to back: %.2f us\n", write_bytes, 2 * write_bytes, run(0, 0, 0, 0, true, 1));
    for (int pdl = 0; pdl < (quick ? 0 : 2); ++pdl)
        for (int grid : {148, 296, 592, 1184, 2368, 4736})
            for (int threads : {256, 512})
                for (int unroll : {4, 8}) {
                    if (pdl == 0 && grid > 1184) continue;
                    float us = run(grid, threads, unroll, pdl, true, 0);
                    float us_r = run(grid, threads, unroll, pdl, false, 0);
                    printf("copy kernel grid=%4d threads=%4d unroll=%d pdl=%d: read+write %.2f us (%.0f GB/s, %.2f of 6452)   read-only %.2f us (%.0f GB/s)\n",
                           grid, threads, unroll, pdl, us, (read_bytes + write_bytes) / us / 1e3, (read_bytes + write_bytes) / us / 1e3 / 6452.5,
                           us_r, read_bytes / us_r / 1e3);
                }
    // ---- how fast can one SM pull a 110 KB tile that is cold in L2? ----
    long long* ns;
    CK(cudaMalloc(&ns, 148 * sizeof(long long)));
    unsigned* sink;
    CK(cudaMalloc(&sink, 4));
    const int tile = 110592;
    CK(cudaFuncSetAttribute(pull_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, tile));
    std::vector<long long> h(148);
    auto flush = [&]() { for (int s = 0; s < nsets; ++s) CK(cudaMemsetAsync(out[s], 1, write_bytes, st)); };  // 500 MB of writes: evicts the inputs
    for (int ctas : {1, 4, 32, 128}) {
        for (int ncopies : {9, 54, 216}) {
            flush();
            pull_bulk<<<ctas, 128, tile, st>>>((const unsigned char*)in[3], tile, ncopies, ns, tile);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(h.data(), ns, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
            std::sort(h.begin(), h.begin() + ctas);
            printf("pull 110592 B per CTA, bulk copies x%3d, %3d CTAs: median %lld ns (%.1f B/ns per SM), max %lld ns\n", ncopies, ctas,
                   h[ctas / 2], tile / (double)h[ctas / 2], h[ctas - 1]);
        }
        for (int threads : {128, 256, 512, 1024}) {
            flush();
            pull_ldg<<<ctas, threads, 0, st>>>((const uint4*)in[5], tile / 16, ns, tile / 16, sink);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(h.data(), ns, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
            std::sort(h.begin(), h.begin() + ctas);
            printf("pull 110592 B per CTA, LDG.128 x %4d threads, %3d CTAs: median %lld ns (%.1f B/ns per SM), max %lld ns\n", threads, ctas,
                   h[ctas / 2], tile / (double)h[ctas / 2], h[ctas - 1]);
        }
    }
    // ---- how fast can one SM write out its 98 KB of gradient rows? ----
    const int gtile = 98304;
    CK(cudaFuncSetAttribute(push_stg, cudaFuncAttributeMaxDynamicSharedMemorySize, gtile));
    CK(cudaFuncSetAttribute(push_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, gtile));
    for (int ctas : {1, 4, 32, 128}) {
        for (int threads : {128, 256, 512}) {
            push_stg<<<ctas, threads, gtile, st>>>((unsigned char*)out[7], gtile, ns, gtile);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(h.data(), ns, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
            std::sort(h.begin(), h.begin() + ctas);
            printf("push 98304 B per CTA, STG.128 x %4d threads, %3d CTAs: median %lld ns (%.1f B/ns per SM), max %lld ns\n", threads, ctas,
                   h[ctas / 2], gtile / (double)h[ctas / 2], h[ctas - 1]);
        }
        for (int ncopies : {8, 24, 96, 192}) {
            push_bulk<<<ctas, 128, gtile, st>>>((unsigned char*)out[9], gtile, ncopies, ns, gtile);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(h.data(), ns, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
            std::sort(h.begin(), h.begin() + ctas);
            printf("push 98304 B per CTA, bulk stores x%3d, %3d CTAs: median %lld ns (%.1f B/ns per SM), max %lld ns\n", ncopies, ctas,
                   h[ctas / 2], gtile / (double)h[ctas / 2], h[ctas - 1]);
        }
    }
    return 0;
}
