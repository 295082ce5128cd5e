// energy.cuh — shared declarations of the energy-score kernels (K1 fused, K1b split pair).
#pragma once

#include "common.cuh"

namespace dddm {

// What one energy launch computes.
enum EnergyMode : int {
    kModeLoss = 0,   // training.py:84-85: out = {loss, conf, inter, W}; grad (optional) = dloss/dxhat
    kModeTerms = 1,  // losses.py:5-25:    out = {conf, inter}; dist saved for the backward
    kModeBwd = 2,    // backward of kModeTerms from the saved distances: grad_xhat (+ grad_x0), no pass 1, no row sums
};

struct EnergyParams {
    const void* xhat;  // [B,m,D]
    const void* x0;    // [B,D]
    void* grad_xhat;   // [B,m,D] or null
    void* grad_x0;     // [B,D] or null (backward kernel only)
    const float* weight_dev;  // kModeLoss: W = weight_dev[0] * weight_scale
    float weight_scale;
    const float* g_conf;   // backward kernel: upstream gradients (device scalars)
    const float* g_inter;
    float* dist;           // [B, m + m(m-1)/2] saved squared distances (written in kModeTerms, read by bwd)
    float* out;            // device scalars
    float* row_partials;   // [B][2] workspace
    unsigned* ticket;      // arrival counter (workspace)
    int B, m, D;
    float lam;
    PowSpec pw;
    int mode;
    int x0_f32;                 // bf16 kernels: x0 is fp32 (mixed entry point; TMA-staged kernel only)
    int window;                 // cp.async loader of the TMA-staged kernel: column chunks in flight per CTA
    int ld_hint, st_hint;       // L2 eviction priority of the streaming loads / stores (single-wave kernel; 0 = normal)
    int finish;                 // TMA-staged kernel: cross-row sum by the arrival ticket (1) or by polled row slots (0 / 2)
    unsigned long long* trace;  // diagnostics: 8 globaltimer stamps per CTA, or null (dddm_set_trace_buffer)
};

struct EnergyWorkspace {  // layout of the caller-provided workspace
    unsigned ticket;
    unsigned pad[3];
    // float row_partials[B][2] follows
};

// Launch plan of the register-resident kernels for a shape (energy_reg.cu).
struct RegPlan {
    bool ok;
    int vec;      // elements per thread vector (1 or 16 bytes worth)
    int nv;       // vectors per thread
    int cluster;  // CTAs per row
    int threads;
};
RegPlan plan_reg(int m, int D, int elem_size, bool aligned16, bool is_bwd);

template <typename T>
int launch_energy_reg(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream);
template <typename T>
int launch_energy_bwd_reg(const EnergyParams& p, const RegPlan& plan, cudaStream_t stream);

// Launch with optional cluster dimension and programmatic dependent launch (PDL).
template <typename K, typename... Args>
inline int launch_with_attrs(K kernel, dim3 grid, dim3 block, size_t smem, int cluster, cudaStream_t stream,
                             Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    int n = 0;
    if (cluster > 1) {
        attrs[n].id = cudaLaunchAttributeClusterDimension;
        attrs[n].val.clusterDim.x = cluster;
        attrs[n].val.clusterDim.y = 1;
        attrs[n].val.clusterDim.z = 1;
        ++n;
    }
    if (tuning().pdl) {
        attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    count_launch();
    return (int)e;
}

// Shared-memory tile kernels for any m (energy_tile.cu).
struct TilePlan {
    bool ok;
    int cluster;
    int threads;
    int chunk_cols;    // columns per staged chunk
    int slab_cols;     // columns per CTA
    size_t smem_bytes;
    bool bulk;         // rows are 16-byte aligned: stage with cp.async.bulk (TMA)
};
TilePlan plan_tile(int m, int D, int elem_size, bool aligned16);
template <typename T>
int launch_energy_tile(const EnergyParams& p, const TilePlan& plan, cudaStream_t stream);
template <typename T>
int launch_energy_bwd_tile(const EnergyParams& p, const TilePlan& plan, cudaStream_t stream);

__device__ __forceinline__ void cluster_arrive_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- deterministic cross-row reduction, executed by the last arriving row ------------------
// Called by ONE warp of the CTA that owns row b after its per-row sums are known.
//
// No __threadfence(): a GPU-scope fence issued while the other warps of the SM stream the gradient
// out costs ~1 us (measured with tools/trace_energy.py) and sat on the kernel's critical path.
// Instead the row's two sums are published with ONE returning 64-bit atomic exchange (performed at
// L2, the point of coherence), and the ticket increment consumes the exchange's return value, so it
// cannot be issued before the exchange has been performed.  The last arriver therefore reads every
// row's sums (ld.global.cg = L2) after they are in L2.  Both sums are non-negative (sign bits are
// cleared, which also canonicalises NaNs), hence bit 63 of whatever the slot held before — zeros
// from the initial memset or an earlier launch's sums — is 0 and the increment is exactly 1.
__device__ __forceinline__ void finish_row(const EnergyParams& p, int b, float conf_row, float inter_row, float W,
                                           int lane) {
    unsigned old = 0;
    if (lane == 0) {
        const unsigned long long packed = (unsigned long long)(__float_as_uint(conf_row) & 0x7fffffffu) |
                                          ((unsigned long long)(__float_as_uint(inter_row) & 0x7fffffffu) << 32);
        const unsigned long long prev = atomicExch(reinterpret_cast<unsigned long long*>(p.row_partials) + b, packed);
        // acq_rel at GPU scope: the exchange above is ordered before the increment (release) and the last arriver's
        // reads of the other rows' sums after it (acquire) by the memory model too, not only by the data dependency
        const unsigned inc = 1u + (unsigned)(prev >> 63);
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p.ticket), "r"(inc) : "memory");
    }
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != (unsigned)(p.B - 1)) return;
    // last row: sum every row's partials in a fixed order
    float c = 0.f, i = 0.f;
    for (int r = lane; r < p.B; r += 32) {
        float2 v = __ldcg(reinterpret_cast<const float2*>(p.row_partials) + r);
        c += v.x;
        i += v.y;
    }
    c = warp_sum(c);
    i = warp_sum(i);
    if (lane == 0) {
        const float conf = c / ((float)p.B * (float)p.m);
        const float inter = i / ((float)p.B * (float)p.m * (float)(p.m - 1));
        if (p.mode == kModeLoss) {
            const float cl = p.lam / (2.0f * (float)(p.m - 1));
            p.out[0] = W * (conf - cl * inter);
            p.out[1] = conf;
            p.out[2] = inter;
            p.out[3] = W;
        } else {
            p.out[0] = conf;
            p.out[1] = inter;
        }
        *p.ticket = 0u;  // leave the workspace reusable
    }
}

// The same cross-row sum without a ticket (kernels with exactly one publishing CTA per row, all rows of the launch
// resident or at least schedulable while one CTA waits: the TMA-staged kernel).  The ticket protocol costs the LAST row
// three dependent round trips to L2 (exchange, increment, reads of the other rows) while L2 is busy with pass 2's
// stores — ~1 us each, and the launch ends with them.  Here a row's sums ARE its flag: one 8-byte store with bit 63
// set (both sums are non-negative, their sign bits are free); the control warp of the launch's last row (b = B - 1)
// polls the B slots, adds them in the same fixed order as finish_row (bit-identical result), and clears them, so the
// workspace stays usable by either protocol: the ticket protocol leaves bit 63 clear (= "not arrived" here), this one
// leaves zeros.  Critical path of the last row: one store + one load round trip.
__device__ __forceinline__ void finish_row_polled(const EnergyParams& p, int b, float conf_row, float inter_row, float W,
                                                  int lane) {
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(p.row_partials);
    if (lane == 0) {
        const unsigned long long packed = (unsigned long long)(__float_as_uint(conf_row) & 0x7fffffffu) |
                                          ((unsigned long long)(__float_as_uint(inter_row) | 0x80000000u) << 32);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slots + b), "l"(packed) : "memory");
    }
    if (b != p.B - 1) return;
    float c = 0.f, i = 0.f;
    for (int r0 = 0; r0 < p.B; r0 += 128) {
        unsigned long long v[4];
        bool all;
        unsigned spins = 0;
        do {
            all = true;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * 32 + lane;
                v[u] = 1ull << 63;
                if (r < p.B) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v[u]) : "l"(slots + r) : "memory");
                all = all && (v[u] >> 63);
            }
            // A row that never arrives means the workspace is shared with a concurrent launch (a contract violation:
            // one workspace per stream) or an earlier launch was aborted: after ~1 s give up with NaN sums, never hang.
            if (!all && ++spins > (1u << 20)) {
                c = __int_as_float(0x7fc00000);
                break;
            }
        } while (!all);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * 32 + lane;
            if (r < p.B) {
                c += __uint_as_float((unsigned)v[u]);
                i += __uint_as_float((unsigned)(v[u] >> 32) & 0x7fffffffu);
                slots[r] = 0ull;
            }
        }
    }
    c = warp_sum(c);
    i = warp_sum(i);
    if (lane == 0) {
        const float conf = c / ((float)p.B * (float)p.m);
        const float inter = i / ((float)p.B * (float)p.m * (float)(p.m - 1));
        if (p.mode == kModeLoss) {
            const float cl = p.lam / (2.0f * (float)(p.m - 1));
            p.out[0] = W * (conf - cl * inter);
            p.out[1] = conf;
            p.out[2] = inter;
            p.out[3] = W;
        } else {
            p.out[0] = conf;
            p.out[1] = inter;
        }
    }
}

// Coefficients multiplying (x_i - x0) and (x_i - x_j) in the gradient, given upstream (gc, gi):
//   d/dxhat_i [gc*conf + gi*inter] = sum_k coef_k * difference_k      (SURVEY.md §8a closed form)
__device__ __forceinline__ float conf_coef(float d2, float gc, const EnergyParams& p) {
    return 2.0f * gc / ((float)p.B * (float)p.m) * pow_deriv(d2, p.pw);
}
__device__ __forceinline__ float pair_coef(float d2, float gi, const EnergyParams& p) {
    return 4.0f * gi / ((float)p.B * (float)p.m * (float)(p.m - 1)) * pow_deriv(d2, p.pw);
}

}  // namespace dddm
