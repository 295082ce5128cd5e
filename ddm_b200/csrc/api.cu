// api.cu — extern "C" entry points of the energy score, argument validation, kernel selection.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "energy.cuh"
#include "energy_smem_plan.h"

#include <nvtx3/nvToolsExt.h>

#include <cstdlib>

namespace dddm {
bool nvtx_enabled() {
    int v = tuning().nvtx;
    if (v < 0) {
        const char* e = getenv("DDDM_NVTX");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
        tuning().nvtx = v;
    }
    return v != 0;
}
void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }
}  // namespace dddm

namespace dddm {

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
Tuning& tuning() {
    static Tuning t;
    return t;
}
static thread_local int g_last_error = 0;
void set_last_error(int e) { g_last_error = e; }

// Multiprocessor count of the CURRENT device (cached per device: a process may drive several).
int device_sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

static bool is_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Register-resident plan: m in [2,8]; a row slab of ceil(nvec/cluster) vectors must fit
// threads * nv with threads <= kRegMaxThreads (256).
RegPlan plan_reg(int m, int D, int elem_size, bool aligned16, bool is_bwd) {
    RegPlan r{};
    r.ok = false;
    if (m < 2 || m > 8 || D < 1) return r;
    const int vecw = 16 / elem_size;
    r.vec = (aligned16 && D % vecw == 0) ? vecw : 1;
    const long nvec = D / r.vec;
    const int max_nv = (r.vec > 1 && elem_size == 4) ? 2 : 1;
    const Tuning& t = tuning();
    int nv = (t.nv >= 1 && t.nv <= max_nv) ? t.nv : 1;
    if (is_bwd) {
        r.cluster = 1;
        r.nv = nv;
        long thr = (nvec + nv - 1) / nv;
        r.threads = (int)(thr >= 256 ? 256 : (thr + 31) / 32 * 32);
        r.ok = true;
        return r;
    }
    int cluster = 0;
    if (t.cluster == 1 || t.cluster == 2 || t.cluster == 4 || t.cluster == 8) cluster = t.cluster;
    if (cluster == 0) {
        // auto: smallest power-of-two cluster whose slab fits 128 threads, capped at 8 CTAs per row
        cluster = 1;
        while (cluster < 8 && (nvec + cluster - 1) / cluster > 128L * nv) cluster *= 2;
    }
    long per_cta = (nvec + cluster - 1) / cluster;
    if (per_cta > 256L * nv && t.nv == 0 && max_nv == 2) nv = 2;
    if (per_cta > 256L * nv) return r;  // slab too wide for the register tile: use the smem-tile kernel
    r.cluster = cluster;
    r.nv = nv;
    long thr = (per_cta + nv - 1) / nv;
    r.threads = (int)((thr + 31) / 32 * 32);
    if (r.threads < 32) r.threads = 32;
    r.ok = true;
    return r;
}

static int validate(const void* xhat, const void* x0, const void* out, const void* ws, int B, int m, int D) {
    if (!xhat || !x0 || !out || !ws) return DDDM_ERR_NULL_POINTER;
    if (m < 2) return DDDM_ERR_BAD_SHAPE;  // training.py:57-58
    if (B < 1 || D < 1 || B > 65535) return DDDM_ERR_BAD_SHAPE;
    if (!is_aligned16(ws)) return DDDM_ERR_BAD_ALIGNMENT;
    return DDDM_OK;
}

template <typename T>
static int run_forward(EnergyParams& p, void* workspace, cudaStream_t stream) {
    auto* ws = static_cast<EnergyWorkspace*>(workspace);
    p.ticket = &ws->ticket;
    p.row_partials = reinterpret_cast<float*>(ws + 1);
    p.trace = static_cast<unsigned long long*>(tuning().trace);
    p.window = tuning().window > 0 ? tuning().window : 3;
    p.ld_hint = tuning().ldhint;
    p.st_hint = tuning().sthint;
    p.finish = tuning().finish;
    const bool al = is_aligned16(p.xhat) && is_aligned16(p.x0) && (!p.grad_xhat || is_aligned16(p.grad_xhat)) &&
                    ((long)p.D * (long)sizeof(T)) % 16 == 0;
    // kernel selection: TMA-staged packed-fp32 kernel (m <= 8, aligned rows) > register-resident kernel
    // (m <= 8, any alignment) > blocked packed-fp32 kernel (m = 16, 32, aligned rows) > chunked smem-tile
    // kernel (any m <= 64)
    const int variant = tuning().variant;
    if (variant == 6 && p.mode != kModeBwd) {  // row-pipelined cluster kernel (single-launch latency path)
        PipePlan pp = plan_pipe(p.B, p.m, p.D, (int)sizeof(T), al, p.x0_f32 ? 2 : 1);
        if (pp.ok) return launch_energy_pipe<T>(p, pp, stream);
        return DDDM_ERR_UNSUPPORTED;
    }
    if constexpr (sizeof(T) == 2) {
        // m = 16 / 32 bf16 draws: Gram + coefficient mixing on the tensor cores (takes bf16 or fp32 x0)
        // auto for both: m = 32 is 2.2x the blocked kernel; m = 16 is ahead on ONE launch (19.7 vs 24.3 us: what a training
        // step issues) and behind with several launches in flight (15.3 vs 12.5 us; energy.variant = 4 selects the blocked kernel)
        if ((variant == 7 || variant == 0) && p.mode != kModeBwd) {
            TcPlan tp = plan_tc(p.B, p.m, p.D, (int)sizeof(T), al);
            if (tp.ok) return launch_energy_tc(p, tp, stream);
            if (variant == 7) return DDDM_ERR_UNSUPPORTED;
        }
    }
    if (p.x0_f32) {  // bf16 draws + fp32 x0: the TMA-staged kernel is the only one with a mixed tile
        SmemPlan sp = plan_smem(p.m, p.D, (int)sizeof(T), al, 2, p.B);
        if (sp.ok) return launch_energy_smem<T>(p, sp, stream);
        return DDDM_ERR_UNSUPPORTED;
    }
    if (variant == 0 || variant == 5) {
        WavePlan wp = plan_wave(p.B, p.m, p.D, (int)sizeof(T), al);
        if (wp.ok) return launch_energy_wave<T>(p, wp, stream);
        if (variant == 5) return DDDM_ERR_UNSUPPORTED;
    }
    if (variant == 0 || variant == 3) {
        SmemPlan sp = plan_smem(p.m, p.D, (int)sizeof(T), al, 1, p.B);
        if (sp.ok) return launch_energy_smem<T>(p, sp, stream);
        if (variant == 3) return DDDM_ERR_UNSUPPORTED;
    }
    if (variant == 0 || variant == 1) {
        RegPlan plan = plan_reg(p.m, p.D, (int)sizeof(T), al, false);
        if (plan.ok) return launch_energy_reg<T>(p, plan, stream);
        if (variant == 1) return DDDM_ERR_UNSUPPORTED;
    }
    if (variant == 0 || variant == 4) {
        SmemPlan bp = plan_blk(p.m, p.D, (int)sizeof(T), al);
        if (bp.ok) return launch_energy_blk<T>(p, bp, stream);
        if (variant == 4) return DDDM_ERR_UNSUPPORTED;
    }
    TilePlan tp = plan_tile(p.m, p.D, (int)sizeof(T), al);
    if (!tp.ok) return DDDM_ERR_UNSUPPORTED;
    return launch_energy_tile<T>(p, tp, stream);
}

template <typename T>
static int energy_fused(const T* xhat, const void* x0, const float* weight_dev, float weight_scale, T* grad, float* out,
                        void* workspace, int B, int m, int D, float beta, float lam, cudaStream_t stream,
                        bool x0_f32 = false) {
    int st = validate(xhat, x0, out, workspace, B, m, D);
    if (st != DDDM_OK) return st;
    if (!weight_dev) return DDDM_ERR_NULL_POINTER;
    EnergyParams p{};
    p.xhat = xhat;
    p.x0 = x0;
    p.grad_xhat = grad;
    p.weight_dev = weight_dev;
    p.weight_scale = weight_scale;
    p.out = out;
    p.B = B;
    p.m = m;
    p.D = D;
    p.lam = lam;
    p.pw = make_pow_spec(beta);
    p.mode = kModeLoss;
    p.x0_f32 = x0_f32 ? 1 : 0;
    return run_forward<T>(p, workspace, stream);
}

template <typename T>
static int energy_terms_fwd(const T* xhat, const T* x0, float* dist, float* out, void* workspace, int B, int m, int D,
                            float beta, cudaStream_t stream) {
    int st = validate(xhat, x0, out, workspace, B, m, D);
    if (st != DDDM_OK) return st;
    EnergyParams p{};
    p.xhat = xhat;
    p.x0 = x0;
    p.dist = dist;
    p.out = out;
    p.B = B;
    p.m = m;
    p.D = D;
    p.lam = 0.f;
    p.pw = make_pow_spec(beta);
    p.mode = kModeTerms;
    return run_forward<T>(p, workspace, stream);
}

template <typename T>
static int energy_terms_bwd(const T* xhat, const T* x0, const float* dist, const float* g_conf, const float* g_inter,
                            T* grad_xhat, T* grad_x0, int B, int m, int D, float beta, cudaStream_t stream) {
    if (!xhat || !x0 || !dist || !g_conf || !g_inter || !grad_xhat) return DDDM_ERR_NULL_POINTER;
    if (m < 2 || B < 1 || D < 1 || B > 65535) return DDDM_ERR_BAD_SHAPE;
    EnergyParams p{};
    p.xhat = xhat;
    p.x0 = x0;
    p.dist = const_cast<float*>(dist);
    p.g_conf = g_conf;
    p.g_inter = g_inter;
    p.grad_xhat = grad_xhat;
    p.grad_x0 = grad_x0;
    p.B = B;
    p.m = m;
    p.D = D;
    p.pw = make_pow_spec(beta);
    p.mode = kModeTerms;
    const bool al = is_aligned16(xhat) && is_aligned16(x0) && is_aligned16(grad_xhat) &&
                    (!grad_x0 || is_aligned16(grad_x0)) && ((long)D * (long)sizeof(T)) % 16 == 0;
    if constexpr (sizeof(T) == 2) {
        // m = 16 / 32 bf16 draws: the tensor-core kernel in backward mode (coefficient mixing from the saved distances)
        if (tuning().variant == 0 || tuning().variant == 7) {
            TcPlan tp = plan_tc(B, m, D, (int)sizeof(T), al);
            if (tp.ok) {
                p.mode = kModeBwd;
                p.trace = static_cast<unsigned long long*>(tuning().trace);
                p.ld_hint = tuning().ldhint;
                return launch_energy_tc(p, tp, stream);
            }
            if (tuning().variant == 7) return DDDM_ERR_UNSUPPORTED;
        }
    }
    if (tuning().variant == 0 || tuning().variant == 3) {
        // TMA-staged packed-fp32 kernel in backward mode (pass 2 only, coefficients from the saved distances)
        SmemPlan sp = plan_smem(m, D, (int)sizeof(T), al, 1, B);
        if (sp.ok) {
            p.mode = kModeBwd;
            p.trace = static_cast<unsigned long long*>(tuning().trace);
            return launch_energy_smem<T>(p, sp, stream);
        }
        if (tuning().variant == 3) return DDDM_ERR_UNSUPPORTED;
    }
    if (tuning().variant == 0 || tuning().variant == 4) {
        // blocked packed-fp32 kernel (m = 16, 24, 32) in backward mode: pass 2 only, coefficients from the saved distances
        SmemPlan bp = plan_blk(m, D, (int)sizeof(T), al);
        if (bp.ok) {
            p.mode = kModeBwd;
            return launch_energy_blk<T>(p, bp, stream);
        }
        if (tuning().variant == 4) return DDDM_ERR_UNSUPPORTED;
    }
    if (tuning().variant != 2) {
        RegPlan plan = plan_reg(m, D, (int)sizeof(T), al, true);
        if (plan.ok) return launch_energy_bwd_reg<T>(p, plan, stream);
    }
    TilePlan tp = plan_tile(m, D, (int)sizeof(T), al);
    if (!tp.ok) return DDDM_ERR_UNSUPPORTED;
    return launch_energy_bwd_tile<T>(p, tp, stream);
}

}  // namespace dddm

using namespace dddm;
using bf16 = __nv_bfloat16;

extern "C" {

int dddm_abi_version(void) { return DDDM_ABI_VERSION; }

const char* dddm_strerror(int status) {
    switch (status) {
        case DDDM_OK: return "ok";
        case DDDM_ERR_NULL_POINTER: return "dddm: required pointer is NULL";
        case DDDM_ERR_BAD_SHAPE: return "dddm: bad shape (need B >= 1, D >= 1, m >= 2 to form interaction pairs)";
        case DDDM_ERR_BAD_ALIGNMENT: return "dddm: workspace must be 16-byte aligned";
        case DDDM_ERR_UNSUPPORTED: return "dddm: no kernel variant supports this shape";
        case DDDM_ERR_BAD_ARGUMENT: return "dddm: bad argument";
        case DDDM_ERR_NO_DEVICE: return "dddm: no CUDA device";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "dddm: unknown status";
}

size_t dddm_energy_workspace_bytes(int B, int m) {
    (void)m;
    if (B < 1) B = 1;
    return sizeof(EnergyWorkspace) + (size_t)B * 2 * sizeof(float);
}
int dddm_energy_workspace_reset(void* workspace, int B, int m, dddm_stream_t stream) {
    if (!workspace) return DDDM_ERR_NULL_POINTER;
    if (B < 1) return DDDM_ERR_BAD_SHAPE;
    return (int)cudaMemsetAsync(workspace, 0, dddm_energy_workspace_bytes(B, m), (cudaStream_t)stream);
}
size_t dddm_energy_dist_per_row(int m) { return m < 2 ? 0 : (size_t)m + (size_t)m * (m - 1) / 2; }

int dddm_energy_fused_f32(const float* xhat, const float* x0, const float* weight_dev, float weight_scale,
                          float* grad_xhat, float* out, void* workspace, int B, int m, int D, float beta, float lam,
                          dddm_stream_t stream) {
    DDDM_NVTX("dddm::K1 energy_fused f32");
    return energy_fused<float>(xhat, x0, weight_dev, weight_scale, grad_xhat, out, workspace, B, m, D, beta, lam,
                               (cudaStream_t)stream);
}
int dddm_energy_fused_bf16(const dddm_bf16* xhat, const dddm_bf16* x0, const float* weight_dev, float weight_scale,
                           dddm_bf16* grad_xhat, float* out, void* workspace, int B, int m, int D, float beta,
                           float lam, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K1 energy_fused bf16");
    return energy_fused<bf16>((const bf16*)xhat, (const bf16*)x0, weight_dev, weight_scale, (bf16*)grad_xhat, out,
                              workspace, B, m, D, beta, lam, (cudaStream_t)stream);
}
int dddm_energy_fused_bf16_x0f32(const dddm_bf16* xhat, const float* x0, const float* weight_dev, float weight_scale,
                                 dddm_bf16* grad_xhat, float* out, void* workspace, int B, int m, int D, float beta,
                                 float lam, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K1 energy_fused bf16 x0=f32");
    return energy_fused<bf16>((const bf16*)xhat, x0, weight_dev, weight_scale, (bf16*)grad_xhat, out, workspace, B, m, D,
                              beta, lam, (cudaStream_t)stream, true);
}
int dddm_energy_fused_bf16_x0f32_supported(int m, int D) {
    const bool al = ((long)D * 2) % 16 == 0;
    const int variant = tuning().variant;
    if ((variant == 0 || variant == 7) && plan_tc(1, m, D, 2, al).ok) return 1;  // tensor-core kernel, m = 16 / 32
    return plan_smem(m, D, 2, al, 2).ok ? 1 : 0;
}
int dddm_energy_terms_fwd_f32(const float* xhat, const float* x0, float* dist, float* out, void* workspace, int B,
                              int m, int D, float beta, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K1b energy_terms_fwd f32");
    return energy_terms_fwd<float>(xhat, x0, dist, out, workspace, B, m, D, beta, (cudaStream_t)stream);
}
int dddm_energy_terms_fwd_bf16(const dddm_bf16* xhat, const dddm_bf16* x0, float* dist, float* out, void* workspace,
                               int B, int m, int D, float beta, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K1b energy_terms_fwd bf16");
    return energy_terms_fwd<bf16>((const bf16*)xhat, (const bf16*)x0, dist, out, workspace, B, m, D, beta,
                                  (cudaStream_t)stream);
}
int dddm_energy_terms_bwd_f32(const float* xhat, const float* x0, const float* dist, const float* g_conf,
                              const float* g_inter, float* grad_xhat, float* grad_x0, int B, int m, int D, float beta,
                              dddm_stream_t stream) {
    DDDM_NVTX("dddm::K1b energy_terms_bwd f32");
    return energy_terms_bwd<float>(xhat, x0, dist, g_conf, g_inter, grad_xhat, grad_x0, B, m, D, beta,
                                   (cudaStream_t)stream);
}
int dddm_energy_terms_bwd_bf16(const dddm_bf16* xhat, const dddm_bf16* x0, const float* dist, const float* g_conf,
                               const float* g_inter, dddm_bf16* grad_xhat, dddm_bf16* grad_x0, int B, int m, int D,
                               float beta, dddm_stream_t stream) {
    DDDM_NVTX("dddm::K1b energy_terms_bwd bf16");
    return energy_terms_bwd<bf16>((const bf16*)xhat, (const bf16*)x0, dist, g_conf, g_inter, (bf16*)grad_xhat,
                                  (bf16*)grad_x0, B, m, D, beta, (cudaStream_t)stream);
}

int dddm_set_tuning(const char* key, int value) {
    if (!key) return DDDM_ERR_NULL_POINTER;
    Tuning& t = tuning();
    if (!strcmp(key, "energy.cluster")) t.cluster = value;
    else if (!strcmp(key, "energy.nv")) t.nv = value;
    else if (!strcmp(key, "energy.variant")) t.variant = value;
    else if (!strcmp(key, "energy.pdl")) t.pdl = value;
    else if (!strcmp(key, "energy.threads")) t.threads = value;
    else if (!strcmp(key, "energy.ctas")) t.ctas = value;
    else if (!strcmp(key, "energy.cols")) t.cols = value;
    else if (!strcmp(key, "energy.ksmem")) t.ksmem = value;
    else if (!strcmp(key, "energy.loader")) t.loader = value;
    else if (!strcmp(key, "energy.window")) t.window = value;
    else if (!strcmp(key, "energy.ldhint")) t.ldhint = value;
    else if (!strcmp(key, "energy.sthint")) t.sthint = value;
    else if (!strcmp(key, "energy.nostore")) t.nostore = value;
    else if (!strcmp(key, "energy.finish")) t.finish = value;
    else if (!strcmp(key, "nvtx")) t.nvtx = value ? 1 : 0;
    else return DDDM_ERR_BAD_ARGUMENT;
    return DDDM_OK;
}
int dddm_get_tuning(const char* key) {
    if (!key) return DDDM_ERR_NULL_POINTER;
    const Tuning& t = tuning();
    if (!strcmp(key, "energy.cluster")) return t.cluster;
    if (!strcmp(key, "energy.nv")) return t.nv;
    if (!strcmp(key, "energy.variant")) return t.variant;
    if (!strcmp(key, "energy.pdl")) return t.pdl;
    if (!strcmp(key, "energy.threads")) return t.threads;
    if (!strcmp(key, "energy.ctas")) return t.ctas;
    if (!strcmp(key, "energy.cols")) return t.cols;
    if (!strcmp(key, "energy.ksmem")) return t.ksmem;
    if (!strcmp(key, "energy.loader")) return t.loader;
    if (!strcmp(key, "energy.window")) return t.window;
    if (!strcmp(key, "energy.ldhint")) return t.ldhint;
    if (!strcmp(key, "energy.sthint")) return t.sthint;
    if (!strcmp(key, "energy.finish")) return t.finish;
    if (!strcmp(key, "energy.nostore")) return t.nostore;
    if (!strcmp(key, "nvtx")) return t.nvtx;
    return DDDM_ERR_BAD_ARGUMENT;
}
int dddm_set_trace_buffer(void* device_buffer) {
    tuning().trace = device_buffer;
    return DDDM_OK;
}
unsigned long long dddm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dddm_energy_describe(int B, int m, int D, int dtype, char* buf, int buflen) {
    if (!buf || buflen < 1) return DDDM_ERR_NULL_POINTER;
    const int es = dtype == 1 ? 2 : 4;
    const bool al = ((long)D * es) % 16 == 0;
    int n;
    const int variant = tuning().variant;
    if ((variant == 7 || variant == 0) && dtype == 1) {
        TcPlan tp = plan_tc(B, m, D, es, al);
        if (tp.ok)
            return snprintf(buf, buflen, "tc<bf16,M=%d> tcgen05 gram + mixing, tma-2d sw128, 1 row per CTA, threads=320 stages=%d smem=%zu", m,
                            tp.stages, tp.smem_bytes);
        if (variant == 7) return snprintf(buf, buflen, "unsupported");
    }
    if (variant == 6) {
        PipePlan pp = plan_pipe(B, m, D, es, al);
        if (pp.ok)
            return snprintf(buf, buflen, "pipe<%s,M=%d> tma-bulk f32x2 rows-per-cluster=%d threads=%d cols=%d slab_vecs=%d window=%d smem=%zu",
                            dtype == 1 ? "bf16" : "f32", m, pp.cluster, pp.threads, pp.cols, pp.slab_vecs, pp.window, pp.smem_bytes);
        return snprintf(buf, buflen, "unsupported");
    }
    if (variant == 0 || variant == 5) {
        WavePlan wp = plan_wave(B, m, D, es, al);
        if (wp.ok)
            return snprintf(buf, buflen, "wave<%s,M=%d,NV=%d> ldg.128 f32x2 register-resident cluster=1 threads=%d coef=%s",
                            dtype == 1 ? "bf16" : "f32", m, wp.nv, wp.threads, wp.ksmem == 2 ? "pair-major" : (wp.ksmem ? "smem" : "regs"));
        if (variant == 5) return snprintf(buf, buflen, "unsupported");
    }
    if (variant == 0 || variant == 3) {
        SmemPlan sp = plan_smem(m, D, es, al, 1, B);
        if (sp.ok) {
            const bool ldgsts = tuning().loader == 2;
            return snprintf(buf, buflen, "smem<%s,M=%d> %s f32x2 cluster=%d threads=%d slab_vecs=%d chunk_vecs=%d smem=%zu",
                            dtype == 1 ? "bf16" : "f32", m, ldgsts ? "cp.async" : "tma-bulk", sp.cluster, sp.threads,
                            sp.slab_vecs, sp.chunk_vecs, sp.smem_bytes);
        }
        if (variant == 3) return snprintf(buf, buflen, "unsupported");
    }
    if (variant == 0 || variant == 1) {
        RegPlan r = plan_reg(m, D, es, al, false);
        if (r.ok) {
            n = snprintf(buf, buflen, "reg<%s,M=%d,VEC=%d,NV=%d> cluster=%d threads=%d", dtype == 1 ? "bf16" : "f32", m,
                         r.vec, r.nv, r.cluster, r.threads);
            return n;
        }
    }
    if (variant == 0 || variant == 4) {
        SmemPlan bp = plan_blk(m, D, es, al);
        if (bp.ok)
            return snprintf(buf, buflen, "blk<%s,M=%d> tma-bulk f32x2 cluster=%d threads=%d slab_vecs=%d chunk_vecs=%d smem=%zu",
                            dtype == 1 ? "bf16" : "f32", m, bp.cluster, bp.threads, bp.slab_vecs, bp.chunk_vecs, bp.smem_bytes);
        if (variant == 4) return snprintf(buf, buflen, "unsupported");
    }
    TilePlan t = plan_tile(m, D, es, al);
    if (t.ok)
        n = snprintf(buf, buflen, "tile<%s,m=%d> cluster=%d threads=%d chunk=%d slab=%d smem=%zu %s",
                     dtype == 1 ? "bf16" : "f32", m, t.cluster, t.threads, t.chunk_cols, t.slab_cols, t.smem_bytes,
                     t.bulk ? "tma-bulk" : "ldg");
    else
        n = snprintf(buf, buflen, "unsupported");
    return n;
}

int dddm_last_error(void) { return g_last_error; }

}  // extern "C"
