/*
 * dddm_b200.h — C ABI of the B200-native (sm_100a) DDDM hot path.
 *
 * The reference (edluyuan/ddm) is pure Python: it has no FFI / plugin interface, its "operator
 * API" for this path is six Python functions.  This header is the drop-in
 * boundary underneath them: every entry point below is what a ctypes / cffi / pybind binding of
 * the corresponding reference function binds to, and each one cites the reference lines it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - all data pointers are DEVICE pointers owned by the caller unless the name says `_host`;
 *     nothing is allocated or freed inside (except by the explicit dddm_session_* objects);
 *   - every call is asynchronous on `stream` (a cudaStream_t / CUstream passed as void*), acts on the CURRENT
 *     device (host-side caches are keyed by device, so one process may drive several GPUs), may be called from
 *     several threads at once, and never throws: it returns 0 or a negative dddm_status / positive cudaError_t.
 *     The one exception is dddm_set_tuning / dddm_set_trace_buffer below: they change PROCESS-GLOBAL state read by
 *     later launches and are meant for benchmarks and tests on a quiescent library, not for concurrent use;
 *   - `bf16` data is raw uint16 storage of IEEE bfloat16; accumulation is always fp32;
 *   - scalars produced on the device stay on the device (no hidden host synchronisation).
 */
#ifndef DDDM_B200_H_
#define DDDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define DDDM_ABI_VERSION 1

typedef void* dddm_stream_t; /* cudaStream_t */
typedef uint16_t dddm_bf16;  /* raw bfloat16 bits */

enum dddm_status {
    DDDM_OK = 0,
    DDDM_ERR_NULL_POINTER = -1,
    DDDM_ERR_BAD_SHAPE = -2,    /* B, m, D, N out of range (m must be >= 2: training.py:57-58) */
    DDDM_ERR_BAD_ALIGNMENT = -3,
    DDDM_ERR_UNSUPPORTED = -4,  /* shape/arch not supported by any kernel variant */
    DDDM_ERR_BAD_ARGUMENT = -5,
    DDDM_ERR_NO_DEVICE = -6
};

int dddm_abi_version(void);
/* Human-readable message for a status returned by any function here (negative: dddm_status,
 * positive: cudaError_t). Never NULL. */
const char* dddm_strerror(int status);

/* ------------------------------------------------------------------------------------------
 * Energy-score workspace.  `workspace` passed to the energy kernels must hold
 * dddm_energy_workspace_bytes(B, m) bytes, 16-byte aligned, and must be ZERO-INITIALISED once
 * before first use; the kernels leave it reusable (the arrival ticket is reset by the last CTA, the
 * polled row slots of the TMA-staged kernel are cleared by the CTA that sums them).
 * One workspace must not be shared by launches that may run concurrently (one workspace per stream):
 * their row sums would mix; the TMA-staged kernel then reports NaN sums after a bounded wait.
 * ------------------------------------------------------------------------------------------ */
size_t dddm_energy_workspace_bytes(int B, int m);
/* Re-zero a workspace on `stream` (asynchronous).  Only needed after a launch that used it was aborted (a sticky CUDA
 * error, a killed process sharing the allocation): its ticket / row slots may then be left half-written. */
int dddm_energy_workspace_reset(void* workspace, int B, int m, dddm_stream_t stream);
/* number of floats per row in the saved-distance buffer of the split fwd/bwd pair: m + m(m-1)/2 */
size_t dddm_energy_dist_per_row(int m);

/*
 * K1 — fused forward + backward of the loss of distributional_training_step
 * (dddm/training.py:77-85 with dddm/losses.py:5-25):
 *     W     = weight_dev[0] * weight_scale          (the batch-mean logistic weight, training.py:84;
 *                                                    weight_dev may hold an all-reduced SUM and
 *                                                    weight_scale 1/(ranks*B): SURVEY.md §8e)
 *     conf  = mean_{b,i}      f(|x0_b - xhat_bi|^2)
 *     inter = mean_{b,i,j!=i} f(|xhat_bi - xhat_bj|^2),  f(d2) = (d2 + 1e-12)^(beta/2), or d2 if beta == 2.0
 *     loss  = W * (conf - lam/(2(m-1)) * inter)
 *     out[0..3] = {loss, conf, inter, W}           (fp32, device)
 *     grad_xhat = d loss / d xhat                  (same dtype/layout as xhat; nullable -> forward only)
 * xhat [B,m,D] and x0 [B,D] are contiguous row-major.
 */
int dddm_energy_fused_f32(const float* xhat, const float* x0, const float* weight_dev, float weight_scale,
                          float* grad_xhat, float* out, void* workspace, int B, int m, int D, float beta,
                          float lam, dddm_stream_t stream);
int dddm_energy_fused_bf16(const dddm_bf16* xhat, const dddm_bf16* x0, const float* weight_dev,
                           float weight_scale, dddm_bf16* grad_xhat, float* out, void* workspace, int B, int m,
                           int D, float beta, float lam, dddm_stream_t stream);
/*
 * Mixed entry point for a bf16 backbone (dddm/training.py:75-85 under autocast): the draws arrive in bf16 as the
 * backbone wrote them, the data x0 stays fp32 (not rounded), the gradient leaves in bf16 — no up-conversion pass
 * over xhat and no down-conversion pass over the gradient.  m <= 8 and rows of 16-byte multiples (D % 8 == 0);
 * dddm_energy_fused_bf16_x0f32_supported(m, D) tells whether a shape is covered (else DDDM_ERR_UNSUPPORTED).
 */
int dddm_energy_fused_bf16_x0f32(const dddm_bf16* xhat, const float* x0, const float* weight_dev, float weight_scale,
                                 dddm_bf16* grad_xhat, float* out, void* workspace, int B, int m, int D, float beta,
                                 float lam, dddm_stream_t stream);
int dddm_energy_fused_bf16_x0f32_supported(int m, int D);

/*
 * K1b — the API-faithful split pair behind generalized_energy_terms(x0hats, x0, beta, lam)
 * (dddm/losses.py:5-25; `lam` is unused there and therefore absent here).
 *   fwd: out[0..1] = {conf, inter}; dist [B, m + m(m-1)/2] fp32 receives the squared distances
 *        (first the m confinement distances, then pairs (i<j) in row-major order) for the backward.
 *   bwd: grad_xhat = g_conf[0]*dconf/dxhat + g_inter[0]*dinter/dxhat, and, when grad_x0 != NULL,
 *        grad_x0 = g_conf[0]*dconf/dx0.  g_conf / g_inter are device scalars (autograd upstream grads).
 */
int dddm_energy_terms_fwd_f32(const float* xhat, const float* x0, float* dist, float* out, void* workspace,
                              int B, int m, int D, float beta, dddm_stream_t stream);
int dddm_energy_terms_fwd_bf16(const dddm_bf16* xhat, const dddm_bf16* x0, float* dist, float* out,
                               void* workspace, int B, int m, int D, float beta, dddm_stream_t stream);
int dddm_energy_terms_bwd_f32(const float* xhat, const float* x0, const float* dist, const float* g_conf,
                              const float* g_inter, float* grad_xhat, float* grad_x0, int B, int m, int D,
                              float beta, dddm_stream_t stream);
int dddm_energy_terms_bwd_bf16(const dddm_bf16* xhat, const dddm_bf16* x0, const float* dist,
                               const float* g_conf, const float* g_inter, dddm_bf16* grad_xhat,
                               dddm_bf16* grad_x0, int B, int m, int D, float beta, dddm_stream_t stream);

/* y[i] *= scale[0] for i < n unless scale[0] == 1 (device-side test, no host sync): hands the
 * pre-multiplied gradient of K1 to autograd when the upstream gradient is not exactly 1. */
int dddm_scale_inplace_f32(float* y, const float* scale, size_t n, dddm_stream_t stream);
int dddm_scale_inplace_bf16(dddm_bf16* y, const float* scale, size_t n, dddm_stream_t stream);

/*
 * K2 — forward_marginal_sample (dddm/schedules.py:17-25) fused with the m-fold expansion of
 * dddm/training.py:70:  xt[b,:] = (1 - t[b]) * x0[b,:] + t[b] * eps[b,:];  xt_rep[b*m+i,:] = xt[b,:].
 * xt and xt_rep are each nullable (but not both); m is ignored when xt_rep == NULL.
 */
int dddm_forward_marginal_expand_f32(const float* x0, const float* t, const float* eps, float* xt,
                                     float* xt_rep, int B, int m, long D, dddm_stream_t stream);
int dddm_forward_marginal_expand_bf16(const dddm_bf16* x0, const float* t, const dddm_bf16* eps, dddm_bf16* xt,
                                      dddm_bf16* xt_rep, int B, int m, long D, dddm_stream_t stream);

/*
 * K2c — the same marginal written m-fold straight into the backbone's channel-concatenated input
 * (dddm/model.py:236 `torch.cat([xt, xi], dim=1)`, after the expansion of dddm/training.py:70-73):
 *   x6[b*m+i, 0:C] = (1 - t[b]) * x0[b] + t[b] * eps[b];   x6[b*m+i, C:2C] = xi[b, i]      ([B*m, 2C, H, W])
 * x6 is written in out_dtype (DDDM_DTYPE_F32 / DDDM_DTYPE_BF16: the down-conversion of a bf16 backbone's input
 * is fused in).  x0_tok (nullable) receives x0 re-ordered into the patch-token layout of PatchUnembed
 * (dddm/model.py:125-128: [B, (H/p)(W/p), C*p*p]) so the loss can read the backbone's tokens without the
 * unpatchify copy — the energy score is invariant to a common permutation of the D axis.
 * Needs W % 4 == 0 (and p % 4 == 0 when x0_tok is requested) and 16-byte aligned pointers.
 */
#define DDDM_DTYPE_F32 0
#define DDDM_DTYPE_BF16 1
int dddm_forward_marginal_concat_f32(const float* x0, const float* t, const float* eps, const float* xi, void* x6,
                                     int out_dtype, float* x0_tok, int B, int m, int C, int H, int W, int patch,
                                     dddm_stream_t stream);
int dddm_forward_marginal_concat_bf16(const dddm_bf16* x0, const float* t, const dddm_bf16* eps, const dddm_bf16* xi,
                                      dddm_bf16* x6, dddm_bf16* x0_tok, int B, int m, int C, int H, int W, int patch,
                                      dddm_stream_t stream);

/*
 * K4 — sigmoid_weight (dddm/losses.py:28-35): w[b] = sigmoid(log((1-t)^2/(t^2+1e-12) + 1e-12) - bias);
 * w (nullable) receives the per-row weights, w_sum (nullable) their sum over B in a fixed order
 * (the caller all-reduces it across ranks and passes 1/(ranks*B) as weight_scale to K1).
 */
int dddm_sigmoid_weight_sum_f32(const float* t, float bias, float* w, float* w_sum, int B, dddm_stream_t stream);

/*
 * K3 — one Algorithm-2 update (dddm/sampling.py:29-31 with gaussian_bridge_mu_sigma,
 * dddm/schedules.py:28-78):   x_out = c_xt*x + c_x0*xhat0 + std*z,  (c_xt, c_x0, std) from (s, t, eps_churn).
 * s, t: device fp32, one value (st_is_vector == 0) or N values.  z nullable (-> x_out = mu).
 * x_out may alias x.  mu_out / std_out nullable: when given they receive mu [N,D] and std [1 or N]
 * (the return values of gaussian_bridge_mu_sigma itself).
 */
int dddm_bridge_step_f32(float* x_out, const float* x, const float* xhat0, const float* z, const float* s,
                         const float* t, int st_is_vector, double eps_churn, float* mu_out, float* std_out,
                         long N, long D, dddm_stream_t stream);
int dddm_bridge_step_bf16(dddm_bf16* x_out, const dddm_bf16* x, const dddm_bf16* xhat0, const dddm_bf16* z,
                          const float* s, const float* t, int st_is_vector, double eps_churn, dddm_bf16* mu_out,
                          float* std_out, long N, long D, dddm_stream_t stream);

/*
 * K3 with the Gaussian draws of the step fused in (dddm/sampling.py:27,30: xi = randn_like(x), z = randn_like(x)).
 *   x_out   = c_xt*x + c_x0*xhat0 + std*z,   z generated in registers (never stored);
 *   xi_next = the NEXT step's xi (nullable), written by the same launch.
 * Both draws are bit-identical to torch.randn_like on the same device for the Philox (seed, offset) given:
 * offset_z / offset_xi are the generator offsets at which the reference's two randn_like calls would run
 * (multiples of 4; every such draw advances the generator by dddm_philox_increment(N*D)).
 * philox_dev (nullable): device array {seed, offset_z, offset_xi} read instead of the by-value arguments, so that
 * one captured launch can be replayed with new offsets.  s, t: one device fp32 value each.  x_out may alias x.
 */
unsigned long long dddm_philox_increment(long numel);
int dddm_bridge_step_philox_f32(float* x_out, const float* x, const float* xhat0, float* xi_next, const float* s,
                                const float* t, double eps_churn, const unsigned long long* philox_dev,
                                unsigned long long seed, unsigned long long offset_z, unsigned long long offset_xi,
                                long N, long D, dddm_stream_t stream);
int dddm_bridge_step_philox_bf16(dddm_bf16* x_out, const dddm_bf16* x, const dddm_bf16* xhat0, dddm_bf16* xi_next,
                                 const float* s, const float* t, double eps_churn,
                                 const unsigned long long* philox_dev, unsigned long long seed,
                                 unsigned long long offset_z, unsigned long long offset_xi, long N, long D,
                                 dddm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer sessions: the same fused loss called with HOST pointers (what a reference-side
 * binding without device tensors would call; bench.py's `e2e` figure).  A session owns device
 * buffers, pinned staging and two streams sized for (B, m, D); step_host copies the inputs in,
 * runs K4 + K1 and copies {loss, conf, inter, W} and (optionally) grad_xhat back.
 * ------------------------------------------------------------------------------------------ */
typedef struct dddm_session dddm_session;
/* dtype: 0 = fp32, 1 = bf16.  Returns NULL on failure (see dddm_last_error). */
dddm_session* dddm_session_create(int B, int m, int D, int dtype, int device);
void dddm_session_destroy(dddm_session*);
/* Synchronous: returns when out_host[4] (and grad_host, if not NULL) are filled. */
int dddm_session_step_host(dddm_session*, const void* xhat_host, const void* x0_host, const float* t_host,
                           float w_bias, float beta, float lam, void* grad_host, float* out_host);
/* Pipelined: enqueue step k (copy-in / compute / copy-out on rotating buffers) without waiting;
 * dddm_session_wait drains everything enqueued so far.  Host buffers must stay valid until then
 * and should be pinned (dddm_host_alloc) for the copies to overlap. */
int dddm_session_enqueue_host(dddm_session*, const void* xhat_host, const void* x0_host, const float* t_host,
                              float w_bias, float beta, float lam, void* grad_host, float* out_host);
int dddm_session_wait(dddm_session*);
/* Packed host layout: when the caller keeps a step's inputs in ONE host buffer [xhat | x0 | t] with x0 at
 * `x0_offset` and t at `t_offset` (in_bytes total), and its outputs in one buffer [grad | out[4]] with out at
 * `out_offset` (out_bytes total), dddm_session_enqueue_host recognises the pointers and issues exactly one copy per
 * direction.  Any argument may be NULL. */
int dddm_session_packed_layout(const dddm_session*, size_t* in_bytes, size_t* x0_offset, size_t* t_offset,
                               size_t* out_bytes, size_t* out_offset);
void* dddm_host_alloc(size_t bytes); /* pinned host memory */
/* Pinned, WRITE-COMBINED host memory for buffers the CPU only writes and the GPU only reads (a step's inputs):
 * not snooped during the transfer; slow to read back on the CPU, so never use it for outputs.  Free with dddm_host_free. */
void* dddm_host_alloc_input(size_t bytes);
void dddm_host_free(void*);
int dddm_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Backbone helpers for the DiT training step (the backbone itself stays PyTorch).  These replace the
 * two HBM-bound items that dominated its profile at 65 536 tokens x 384 channels (profiles/r01_dit.md):
 *   LayerNorm forward / backward (torch.nn.functional.layer_norm semantics, dddm/model.py:158-163, 203: biased variance,
 *   eps inside the square root; mean/rstd [N] fp32 saved for the backward; dgamma/dbeta reduced in a fixed order), and
 *   column sums out[c] = sum_n a[n, c] (bias gradients of every Linear).
 * x, y, dy, dx: [N, C] contiguous, C % 4 == 0, rows 16-byte aligned, C <= 1024 for LayerNorm.  scratch: at least
 * dddm_backbone_scratch_bytes(C) bytes of device memory (no initialisation needed).
 * ------------------------------------------------------------------------------------------ */
size_t dddm_backbone_scratch_bytes(int C);
int dddm_layer_norm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                            long N, int C, float eps, dddm_stream_t stream);
int dddm_layer_norm_fwd_bf16(const dddm_bf16* x, const dddm_bf16* gamma, const dddm_bf16* beta, dddm_bf16* y, float* mean,
                             float* rstd, long N, int C, float eps, dddm_stream_t stream);
int dddm_layer_norm_bwd_f32(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                            float* dx, float* dgamma, float* dbeta, float* scratch, size_t scratch_bytes, long N, int C,
                            dddm_stream_t stream);
int dddm_layer_norm_bwd_bf16(const dddm_bf16* dy, const dddm_bf16* x, const float* mean, const float* rstd,
                             const dddm_bf16* gamma, dddm_bf16* dx, dddm_bf16* dgamma, dddm_bf16* dbeta, float* scratch,
                             size_t scratch_bytes, long N, int C, dddm_stream_t stream);
int dddm_colsum_f32(const float* a, float* out, float* scratch, size_t scratch_bytes, long N, int C, dddm_stream_t stream);
int dddm_colsum_bf16(const dddm_bf16* a, dddm_bf16* out, float* scratch, size_t scratch_bytes, long N, int C,
                     dddm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation: rbf_mmd2 (dddm/metrics.py:140-163).  The Gram tile G = A B^T comes from the caller's
 * library GEMM; dddm_rbf_kernel_sum_f32 makes ONE pass over it:
 *   out[0] = sum_{r < rows, c < cols, not (skip_diag and r + diag_shift == c)} exp(-gamma * (a2[r] + b2[c] - 2 G[r, c]))
 * (the reference's pdist2 + exp + boolean-mask gather + mean, metrics.py:143-162, without its five n x n temporaries).
 * Partial sums are folded in double in a fixed order.  scratch: dddm_rbf_scratch_bytes(rows, cols) bytes.
 * dddm_row_sqnorm_f32: out[i] = sum_k x[i, k]^2 (metrics.py:144-145).
 * ------------------------------------------------------------------------------------------ */
/*
 * The same sums WITHOUT a Gram matrix in memory (the fused path rbf_mmd2 takes for evaluation-sized sets):
 *   dddm_rbf_split_bf16     x fp32 [n, D] -> hi, lo bf16 [n, Dp], Dp = dddm_rbf_tc_padded_cols(D) (zero padded),
 *                           x ~= hi + lo to 2^-17 relative (one streaming pre-pass per input set);
 *   dddm_rbf_kernel_sum_tc  out[0] = sum_{i,j} w_ij exp(-gamma (a2_i + b2_j - 2 a_i.b_j)), the dot products accumulated
 *                           in fp32 on the tensor cores as hi.hi + hi.lo + lo.hi inside a persistent tcgen05 kernel whose
 *                           epilogue reads the accumulator from tensor memory.  symmetric != 0 (a == b, rows_a == rows_b):
 *                           w_ij = [i != j], computed from the strict upper triangle; else w_ij = 1.
 *                           scratch: dddm_rbf_tc_scratch_bytes() bytes.  Deterministic (static tile schedule, fixed folds).
 */
long dddm_rbf_tc_padded_cols(long D);
int dddm_rbf_split_bf16(const float* x, dddm_bf16* hi, dddm_bf16* lo, long n, long D, dddm_stream_t stream);
size_t dddm_rbf_tc_scratch_bytes(void);
int dddm_rbf_kernel_sum_tc(const dddm_bf16* a_hi, const dddm_bf16* a_lo, const dddm_bf16* b_hi, const dddm_bf16* b_lo,
                           const float* a2, const float* b2, long rows_a, long rows_b, long D, float gamma, int symmetric,
                           double* scratch, size_t scratch_bytes, double* out, dddm_stream_t stream);
int dddm_row_sqnorm_f32(const float* x, float* out, long n, long D, dddm_stream_t stream);
size_t dddm_rbf_scratch_bytes(long rows, long cols);
int dddm_rbf_kernel_sum_f32(const float* G, long ldg, const float* a2, const float* b2, long rows, long cols, float gamma,
                            long diag_shift, int skip_diag, double* scratch, size_t scratch_bytes, double* out,
                            dddm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tuning / introspection (benchmarks and tests; never needed for correctness).
 *   keys: "energy.variant" (0 = auto, 1 = register-resident, 2 = chunked shared-memory tile for any m,
 *         3 = TMA-staged packed-fp32 kernel for m <= 8, 4 = blocked packed-fp32 kernel for m = 16, 32), "energy.cluster" (CTAs per row, 0 = auto: whole rows, split along D only for rows too wide for one tile and — variant 3 — for
 *         minibatches of fewer rows than half the SMs),
 *         "energy.threads" (threads per CTA of variant 3, 0 = auto), "energy.nv" (16-byte vectors per
 *         thread of variant 1, 0 = auto), "energy.pdl" (programmatic dependent launch, default 1), "energy.ctas" (experiment),
 *         "energy.variant" = 5 (single-wave register-resident kernel for m <= 8: "energy.threads" 128/256/384 x "energy.nv"
 *         1..3 vectors per thread), "energy.loader" (TMA-staged kernel: 0/1 TMA bulk copies, 2 cp.async commit groups with
 *         "energy.window" chunks in flight), "energy.ldhint" / "energy.sthint" (L2 eviction priority of the single-wave
 *         kernel's loads / stores: 0 normal, 1 first, 2 last, 3 unchanged), "energy.nostore" (diagnostics).
 * Process-global and NOT thread-safe: set them while no other thread is launching.
 * dddm_launch_count returns the number of kernels this library has launched in this process.
 * ------------------------------------------------------------------------------------------ */
int dddm_set_tuning(const char* key, int value);
int dddm_get_tuning(const char* key);
unsigned long long dddm_launch_count(void);
/* Diagnostics: while a device buffer of >= B * cluster * 16 uint64 is set, the TMA-staged energy kernel writes up
 * to 16 %globaltimer stamps per CTA (entry, inputs ready, first chunk landed, pass 1 done, coefficients ready,
 * pass 2 done, TMA issued, row finished).  NULL (the default) disables it.  tools/trace_energy.py. */
int dddm_set_trace_buffer(void* device_buffer);
/* Describes the kernel variant the current tuning would pick for a shape, e.g.
 * "reg<f32,M=8,VEC=4,NV=1> cluster=8 threads=96". Returns chars written (excluding NUL). */
int dddm_energy_describe(int B, int m, int D, int dtype, char* buf, int buflen);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DDDM_B200_H_ */
