// Micro-benchmark (diagnostics, not product): issue cost of the inner loops of the energy kernel on one SM.
// Each CTA walks an (M+1) x NQ tile of float4 "quads" in shared memory `reps` times and reports cycles per quad
// per warp for:  pass1 packed (FADD2+FFMA2), pass1 scalar (FADD+FFMA), pass2 packed antisymmetric with and
// without the global stores, pass2 scalar.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
constexpr int M = 8, P = M * (M + 1) / 2, NQ = 768;
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__host__ __device__ constexpr int slot(int i, int j) { return M + i * M - i * (i + 1) / 2 + (j - i - 1); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* in, float* out, float* gout, long long* cyc, int reps) {
    extern __shared__ float4 tile[];  // [M+1][NQ]
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < (M + 1) * NQ; i += nthr) tile[i] = reinterpret_cast<const float4*>(in)[i];
    __syncthreads();
    float k1[P];
    for (int s = 0; s < P; ++s) k1[s] = in[s] * 1e-3f;
    long long t0 = clock64();
    if (MODE == 0) {  // pass 1 packed
        float2 acc[P];
#pragma unroll
        for (int s = 0; s < P; ++s) acc[s] = make_float2(0.f, 0.f);
        for (int r = 0; r < reps; ++r)
            for (int q = tid; q < NQ; q += nthr) {
                float2 x[M + 1][2];
#pragma unroll
                for (int a = 0; a <= M; ++a) { float4 v = tile[a * NQ + q]; x[a][0] = make_float2(v.x, v.y); x[a][1] = make_float2(v.z, v.w); }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < M; ++i) { float2 d = sub2(x[i][h], x[M][h]); acc[i] = __ffma2_rn(d, d, acc[i]); }
#pragma unroll
                    for (int i = 0; i < M; ++i)
#pragma unroll
                        for (int j = i + 1; j < M; ++j) { float2 d = sub2(x[i][h], x[j][h]); acc[slot(i, j)] = __ffma2_rn(d, d, acc[slot(i, j)]); }
                }
            }
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < P; ++s) t += acc[s].x + acc[s].y;
        out[blockIdx.x * nthr + tid] = t;
    } else if (MODE == 1) {  // pass 1 scalar
        float acc[P];
#pragma unroll
        for (int s = 0; s < P; ++s) acc[s] = 0.f;
        for (int r = 0; r < reps; ++r)
            for (int q = tid; q < NQ; q += nthr) {
                float x[M + 1][4];
#pragma unroll
                for (int a = 0; a <= M; ++a) { float4 v = tile[a * NQ + q]; x[a][0] = v.x; x[a][1] = v.y; x[a][2] = v.z; x[a][3] = v.w; }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
#pragma unroll
                    for (int i = 0; i < M; ++i) { float d = x[i][h] - x[M][h]; acc[i] = fmaf(d, d, acc[i]); }
#pragma unroll
                    for (int i = 0; i < M; ++i)
#pragma unroll
                        for (int j = i + 1; j < M; ++j) { float d = x[i][h] - x[j][h]; acc[slot(i, j)] = fmaf(d, d, acc[slot(i, j)]); }
                }
            }
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < P; ++s) t += acc[s];
        out[blockIdx.x * nthr + tid] = t;
    } else if (MODE == 2 || MODE == 3) {  // pass 2 packed antisymmetric; MODE 3 skips the stores (keeps a checksum)
        float2 K2[P];
#pragma unroll
        for (int s = 0; s < P; ++s) K2[s] = make_float2(k1[s], k1[s]);
        float2 chk = make_float2(0.f, 0.f);
        float* grow = gout + (long)blockIdx.x * M * NQ * 4;
        for (int r = 0; r < reps; ++r)
            for (int q = tid; q < NQ; q += nthr) {
                float2 x[M + 1][2], g[M][2];
#pragma unroll
                for (int a = 0; a <= M; ++a) { float4 v = tile[a * NQ + q]; x[a][0] = make_float2(v.x, v.y); x[a][1] = make_float2(v.z, v.w); }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < M; ++i) g[i][h] = __fmul2_rn(K2[i], sub2(x[i][h], x[M][h]));
#pragma unroll
                    for (int i = 0; i < M; ++i)
#pragma unroll
                        for (int j = i + 1; j < M; ++j) {
                            float2 d = sub2(x[i][h], x[j][h]);
                            float2 kk = K2[slot(i, j)];
                            g[i][h] = __ffma2_rn(kk, d, g[i][h]);
                            g[j][h] = __ffma2_rn(make_float2(-kk.x, -kk.y), d, g[j][h]);
                        }
                }
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < M; ++i)
                        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(grow + ((long)i * NQ + q) * 4),
                                     "f"(g[i][0].x), "f"(g[i][0].y), "f"(g[i][1].x), "f"(g[i][1].y) : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < M; ++i) { chk = __fadd2_rn(chk, g[i][0]); chk = __fadd2_rn(chk, g[i][1]); }
                }
            }
        out[blockIdx.x * nthr + tid] = chk.x + chk.y;
    } else if (MODE == 4 || MODE == 5) {  // pass 2 scalar, direct (non-antisymmetric: 8 x (1 + 7) differences), MODE 5 antisymmetric
        float chk = 0.f;
        float* grow = gout + (long)blockIdx.x * M * NQ * 4;
        for (int r = 0; r < reps; ++r)
            for (int q = tid; q < NQ; q += nthr) {
                float x[M + 1][4], g[M][4];
#pragma unroll
                for (int a = 0; a <= M; ++a) { float4 v = tile[a * NQ + q]; x[a][0] = v.x; x[a][1] = v.y; x[a][2] = v.z; x[a][3] = v.w; }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
#pragma unroll
                    for (int i = 0; i < M; ++i) g[i][h] = k1[i] * (x[i][h] - x[M][h]);
                    if (MODE == 4) {
#pragma unroll
                        for (int i = 0; i < M; ++i)
#pragma unroll
                            for (int j = 0; j < M; ++j) {
                                if (j == i) continue;
                                g[i][h] = fmaf(k1[i < j ? slot(i, j) : slot(j, i)], x[i][h] - x[j][h], g[i][h]);
                            }
                    } else {
#pragma unroll
                        for (int i = 0; i < M; ++i)
#pragma unroll
                            for (int j = i + 1; j < M; ++j) {
                                float d = x[i][h] - x[j][h];
                                g[i][h] = fmaf(k1[slot(i, j)], d, g[i][h]);
                                g[j][h] = fmaf(-k1[slot(i, j)], d, g[j][h]);
                            }
                    }
                }
#pragma unroll
                for (int i = 0; i < M; ++i)
                    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(grow + ((long)i * NQ + q) * 4),
                                 "f"(g[i][0]), "f"(g[i][1]), "f"(g[i][2]), "f"(g[i][3]) : "memory");
            }
        out[blockIdx.x * nthr + tid] = chk;
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}


// raw issue rate: 24 independent dependency chains of one opcode
template <int OP>
__global__ void __launch_bounds__(512, 1) raw(const float* in, float* out, long long* cyc, int iters) {
    const int tid = threadIdx.x;
    float2 a[24];
    for (int s = 0; s < 24; ++s) a[s] = make_float2(in[s + tid], in[s + 32 + tid]);
    const float2 b = make_float2(in[tid + 7], in[tid + 9]), c = make_float2(in[tid + 3], in[tid + 5]);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 24; ++s) {
            if (OP == 0) a[s] = __ffma2_rn(a[s], b, c);                       // FFMA2
            if (OP == 1) { a[s].x = fmaf(a[s].x, b.x, c.x); a[s].y = fmaf(a[s].y, b.y, c.y); }  // 2 x FFMA
            if (OP == 2) a[s] = __fadd2_rn(a[s], b);                          // FADD2
            if (OP == 3) { a[s].x = a[s].x + b.x; a[s].y = a[s].y + b.y; }    // 2 x FADD
            if (OP == 4) a[s] = __ffma2_rn(b, c, a[s]);                       // FFMA2, accumulate form
            if (OP == 5) { a[s].x = fmaf(b.x, c.x, a[s].x); a[s].y = fmaf(b.y, c.y, a[s].y); }
        }
    }
    long long t1 = clock64();
    float t = 0.f;
    for (int s = 0; s < 24; ++s) t += a[s].x + a[s].y;
    out[blockIdx.x * blockDim.x + tid] = t;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run_raw(const char* name, int threads, float* in, float* out, long long* cyc) {
    const int iters = 2000;
    raw<OP><<<1, threads>>>(in, out, cyc, iters);
    raw<OP><<<1, threads>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 24.0) / (threads / 128.0);
    printf("raw %-28s threads=%3d  %.2f cyc per (2 flop-lanes x 32) group per SMSP\n", name, threads, per);
}

template <int MODE>
void run(const char* name, int threads, int ctas, float* in, float* out, float* gout, long long* cyc, int reps) {
    size_t smem = (size_t)(M + 1) * NQ * 16;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<ctas, threads, smem>>>(in, out, gout, cyc, reps);
    k<MODE><<<ctas, threads, smem>>>(in, out, gout, cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (auto v : h) mean += v;
    mean /= ctas;
    const double quads_per_warp = (double)reps * NQ / threads;  // each thread (hence each warp) walks this many quads
    const int warps_per_smsp = threads / 128 > 0 ? threads / 128 : 1;
    printf("%-34s threads=%3d ctas=%3d  %8.0f cyc  %7.1f cyc/quad/warp  %6.1f cyc/quad/SMSP   (%s)\n", name, threads, ctas, mean,
           mean / quads_per_warp, mean / quads_per_warp / warps_per_smsp, cudaGetErrorString(e));
}

int main() {
    float *in, *out, *gout;
    long long* cyc;
    const int ctas = 148;
    cudaMalloc(&in, (M + 1) * NQ * 16);
    cudaMalloc(&out, ctas * 512 * 4);
    cudaMalloc(&gout, (size_t)ctas * M * NQ * 16);
    cudaMalloc(&cyc, ctas * 8);
    std::vector<float> h((M + 1) * NQ * 4);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.f;
    cudaMemcpy(in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int threads : {128, 256, 512}) {
        run_raw<0>("FFMA2 a=a*b+c", threads, in, out, cyc);
        run_raw<1>("2xFFMA a=a*b+c", threads, in, out, cyc);
        run_raw<2>("FADD2", threads, in, out, cyc);
        run_raw<3>("2xFADD", threads, in, out, cyc);
        run_raw<4>("FFMA2 a=b*c+a", threads, in, out, cyc);
        run_raw<5>("2xFFMA a=b*c+a", threads, in, out, cyc);
    }
    printf("per quad (4 columns): pass1 = 144 packed | 288 scalar ops; pass2 antisym = 200 packed | 400 scalar, direct = 256 | 512\n");
    for (int threads : {128, 256}) {
        for (int ctas_run : {1, 148}) {
            run<0>("pass1 packed FADD2+FFMA2", threads, ctas_run, in, out, gout, cyc, 20);
            run<1>("pass1 scalar FADD+FFMA", threads, ctas_run, in, out, gout, cyc, 20);
            run<2>("pass2 packed antisym + STG", threads, ctas_run, in, out, gout, cyc, 20);
            run<3>("pass2 packed antisym, no stores", threads, ctas_run, in, out, gout, cyc, 20);
            run<4>("pass2 scalar direct + STG", threads, ctas_run, in, out, gout, cyc, 20);
            run<5>("pass2 scalar antisym + STG", threads, ctas_run, in, out, gout, cyc, 20);
        }
    }
    return 0;
}
