// host_session.cu — the fused loss called with HOST buffers (C-ABI `dddm_session_*`).
//
// What a reference-side binding without device tensors calls, and what bench.py reports as `e2e`:
// host -> device copies of (xhat, x0, t), K4 (logistic weight sum) + K1 (fused energy score),
// device -> host copies of {loss, conf, inter, W} and dloss/dxhat.  Four rotating buffer sets and
// one stream per direction let step k+1's upload and step k-1's download overlap step k's kernels
// (PCIe is full duplex); every dependency is an event, nothing blocks the host until _wait.
// A slot's inputs live in ONE device allocation [xhat | x0 | t] and its outputs in one [grad | out]: when the
// caller's host buffers have the same packed layout (dddm_session_packed_offsets) a step is exactly one
// cudaMemcpyAsync per direction; separate host pointers still work and cost three + two copies.
#include <cstdlib>
#include <new>

#include "common.cuh"

namespace dddm {
void set_last_error(int e);
}

struct dddm_session {
    static constexpr int kSlots = 4;
    int B, m, D, dtype, device;
    size_t esz;
    size_t nx, n0, off_x0, off_t, in_bytes, off_out, out_bytes;  // packed layouts (offsets are 256-byte aligned)
    struct Slot {
        unsigned char* in = nullptr;    // [xhat | x0 | t]
        unsigned char* outb = nullptr;  // [grad | out[4]]
        void* xhat = nullptr;
        void* x0 = nullptr;
        void* grad = nullptr;
        float* t = nullptr;
        float* wsum = nullptr;
        float* out = nullptr;
        void* ws = nullptr;
        cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr;
        bool busy = false;
    } slot[kSlots];
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    unsigned long long next = 0;
};

#define SESSION_TRY(expr)                          \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) {                  \
            dddm::set_last_error((int)e__);        \
            return (int)e__;                       \
        }                                          \
    } while (0)

static int session_alloc(dddm_session* s) {
    SESSION_TRY(cudaSetDevice(s->device));
    SESSION_TRY(cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking));
    SESSION_TRY(cudaStreamCreateWithFlags(&s->s_run, cudaStreamNonBlocking));
    SESSION_TRY(cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking));
    const size_t nws = dddm_energy_workspace_bytes(s->B, s->m);
    for (auto& k : s->slot) {
        SESSION_TRY(cudaMalloc((void**)&k.in, s->in_bytes));
        SESSION_TRY(cudaMalloc((void**)&k.outb, s->out_bytes));
        k.xhat = k.in;
        k.x0 = k.in + s->off_x0;
        k.t = reinterpret_cast<float*>(k.in + s->off_t);
        k.grad = k.outb;
        k.out = reinterpret_cast<float*>(k.outb + s->off_out);
        SESSION_TRY(cudaMalloc((void**)&k.wsum, sizeof(float)));
        SESSION_TRY(cudaMalloc(&k.ws, nws));
        SESSION_TRY(cudaMemsetAsync(k.ws, 0, nws, s->s_run));  // ordered before the first launch on the run stream
        SESSION_TRY(cudaEventCreateWithFlags(&k.uploaded, cudaEventDisableTiming));
        SESSION_TRY(cudaEventCreateWithFlags(&k.computed, cudaEventDisableTiming));
        SESSION_TRY(cudaEventCreateWithFlags(&k.downloaded, cudaEventDisableTiming));
    }
    SESSION_TRY(cudaStreamSynchronize(s->s_run));
    return DDDM_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" {

dddm_session* dddm_session_create(int B, int m, int D, int dtype, int device) {
    if (B < 1 || m < 2 || D < 1 || (dtype != 0 && dtype != 1)) {
        dddm::set_last_error(DDDM_ERR_BAD_SHAPE);
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        dddm::set_last_error(DDDM_ERR_NO_DEVICE);
        return nullptr;
    }
    auto* s = new (std::nothrow) dddm_session();
    if (!s) return nullptr;
    s->B = B;
    s->m = m;
    s->D = D;
    s->dtype = dtype;
    s->device = device;
    s->esz = dtype == 1 ? 2 : 4;
    s->nx = (size_t)B * m * D * s->esz;
    s->n0 = (size_t)B * D * s->esz;
    s->off_x0 = align_up(s->nx, 256);
    s->off_t = s->off_x0 + align_up(s->n0, 256);
    s->in_bytes = s->off_t + align_up((size_t)B * sizeof(float), 256);
    s->off_out = align_up(s->nx, 256);
    s->out_bytes = s->off_out + 256;
    if (session_alloc(s) != DDDM_OK) {
        dddm_session_destroy(s);
        return nullptr;
    }
    return s;
}

void dddm_session_destroy(dddm_session* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (auto& k : s->slot) {
        cudaFree(k.in);
        cudaFree(k.outb);
        cudaFree(k.wsum);
        cudaFree(k.ws);
        if (k.uploaded) cudaEventDestroy(k.uploaded);
        if (k.computed) cudaEventDestroy(k.computed);
        if (k.downloaded) cudaEventDestroy(k.downloaded);
    }
    if (s->s_in) cudaStreamDestroy(s->s_in);
    if (s->s_run) cudaStreamDestroy(s->s_run);
    if (s->s_out) cudaStreamDestroy(s->s_out);
    delete s;
}

int dddm_session_enqueue_host(dddm_session* s, const void* xhat_host, const void* x0_host, const float* t_host,
                              float w_bias, float beta, float lam, void* grad_host, float* out_host) {
    DDDM_NVTX("dddm::session enqueue_host (H2D + K4 + K1 + D2H)");
    if (!s || !xhat_host || !x0_host || !t_host || !out_host) return DDDM_ERR_NULL_POINTER;
    SESSION_TRY(cudaSetDevice(s->device));
    auto& k = s->slot[s->next % dddm_session::kSlots];
    ++s->next;
    if (k.busy) SESSION_TRY(cudaEventSynchronize(k.downloaded));  // slot's previous results have left the device
    const size_t nx = s->nx, n0 = s->n0;
    const unsigned char* base = static_cast<const unsigned char*>(xhat_host);
    if (static_cast<const unsigned char*>(x0_host) == base + s->off_x0 &&
        reinterpret_cast<const unsigned char*>(t_host) == base + s->off_t) {
        // the caller's buffer has the session's packed layout: one copy
        SESSION_TRY(cudaMemcpyAsync(k.in, base, s->off_t + (size_t)s->B * sizeof(float), cudaMemcpyHostToDevice, s->s_in));
    } else {
        SESSION_TRY(cudaMemcpyAsync(k.xhat, xhat_host, nx, cudaMemcpyHostToDevice, s->s_in));
        SESSION_TRY(cudaMemcpyAsync(k.x0, x0_host, n0, cudaMemcpyHostToDevice, s->s_in));
        SESSION_TRY(cudaMemcpyAsync(k.t, t_host, (size_t)s->B * sizeof(float), cudaMemcpyHostToDevice, s->s_in));
    }
    SESSION_TRY(cudaEventRecord(k.uploaded, s->s_in));
    SESSION_TRY(cudaStreamWaitEvent(s->s_run, k.uploaded, 0));
    int st = dddm_sigmoid_weight_sum_f32(k.t, w_bias, nullptr, k.wsum, s->B, s->s_run);
    if (st != DDDM_OK) return st;
    const float scale = 1.0f / (float)s->B;
    void* grad_dev = grad_host ? k.grad : nullptr;
    if (s->dtype == 0)
        st = dddm_energy_fused_f32((const float*)k.xhat, (const float*)k.x0, k.wsum, scale, (float*)grad_dev, k.out,
                                   k.ws, s->B, s->m, s->D, beta, lam, s->s_run);
    else
        st = dddm_energy_fused_bf16((const dddm_bf16*)k.xhat, (const dddm_bf16*)k.x0, k.wsum, scale,
                                    (dddm_bf16*)grad_dev, k.out, k.ws, s->B, s->m, s->D, beta, lam, s->s_run);
    if (st != DDDM_OK) return st;
    SESSION_TRY(cudaEventRecord(k.computed, s->s_run));
    SESSION_TRY(cudaStreamWaitEvent(s->s_out, k.computed, 0));
    if (grad_host && reinterpret_cast<unsigned char*>(out_host) == static_cast<unsigned char*>(grad_host) + s->off_out) {
        SESSION_TRY(cudaMemcpyAsync(grad_host, k.outb, s->off_out + 4 * sizeof(float), cudaMemcpyDeviceToHost, s->s_out));
    } else {
        SESSION_TRY(cudaMemcpyAsync(out_host, k.out, 4 * sizeof(float), cudaMemcpyDeviceToHost, s->s_out));
        if (grad_host) SESSION_TRY(cudaMemcpyAsync(grad_host, k.grad, nx, cudaMemcpyDeviceToHost, s->s_out));
    }
    SESSION_TRY(cudaEventRecord(k.downloaded, s->s_out));
    // No device-side edge from this step's kernels back to the upload stream: the next upload into THIS slot is four
    // steps away and is preceded by the host-side wait on `downloaded` above (download follows compute), while the
    // next step's upload goes to another slot and may overlap these kernels.
    k.busy = true;
    return DDDM_OK;
}

int dddm_session_packed_layout(const dddm_session* s, size_t* in_bytes, size_t* x0_offset, size_t* t_offset,
                               size_t* out_bytes, size_t* out_offset) {
    if (!s) return DDDM_ERR_NULL_POINTER;
    if (in_bytes) *in_bytes = s->in_bytes;
    if (x0_offset) *x0_offset = s->off_x0;
    if (t_offset) *t_offset = s->off_t;
    if (out_bytes) *out_bytes = s->out_bytes;
    if (out_offset) *out_offset = s->off_out;
    return DDDM_OK;
}

int dddm_session_wait(dddm_session* s) {
    if (!s) return DDDM_ERR_NULL_POINTER;
    SESSION_TRY(cudaSetDevice(s->device));
    SESSION_TRY(cudaStreamSynchronize(s->s_in));
    SESSION_TRY(cudaStreamSynchronize(s->s_run));
    SESSION_TRY(cudaStreamSynchronize(s->s_out));
    for (auto& k : s->slot) k.busy = false;
    return DDDM_OK;
}

int dddm_session_step_host(dddm_session* s, const void* xhat_host, const void* x0_host, const float* t_host,
                           float w_bias, float beta, float lam, void* grad_host, float* out_host) {
    int st = dddm_session_enqueue_host(s, xhat_host, x0_host, t_host, w_bias, beta, lam, grad_host, out_host);
    if (st != DDDM_OK) return st;
    return dddm_session_wait(s);
}

void* dddm_host_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        dddm::set_last_error((int)e);
        return nullptr;
    }
    return p;
}
void* dddm_host_alloc_input(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocWriteCombined);
    if (e != cudaSuccess) {
        dddm::set_last_error((int)e);
        return nullptr;
    }
    return p;
}
void dddm_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
