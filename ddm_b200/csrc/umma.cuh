// umma.cuh — the sm_100a tensor-core plumbing shared by the tcgen05 kernels (inline PTX, no library):
// mbarriers, 2-D TMA tensor loads, tensor-memory (TMEM) allocation, shared-memory matrix descriptors, the
// tcgen05.mma instruction descriptor, tcgen05.mma / commit / ld wrappers.
//
// Conventions used by every kernel built on it:
//   * operand tiles sit in shared memory in the canonical 128-byte-swizzled layout: rows of 128 bytes, groups of 8 rows
//     (1024 bytes) contiguous, 16-byte chunk c of row r stored at chunk position c ^ (r & 7); tile bases are 1024-byte
//     aligned.  A TMA tensor map with CU_TENSOR_MAP_SWIZZLE_128B and a 128-byte inner box writes exactly this.
//   * read as a K-major operand, the 128-byte row is 64 bf16 of K; one tcgen05.mma (kind::f16) consumes K = 16 = 32
//     bytes, and the next K step is the same descriptor with the start address advanced by 32 bytes;
//   * read as an MN-major operand, the 128-byte row is 64 consecutive M (or N) indices of ONE k, and 8 rows are 8 k's;
//   * accumulators live in TMEM: lane = row of D (M = 128), column = column of D (fp32, one column per element).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dddm {
namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "UMMA_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra UMMA_WAIT_DONE;\n"
        "bra UMMA_WAIT_LOOP;\n"
        "UMMA_WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// ---- TMA: 2-D tensor load global -> shared (box given by the tensor map), completion on an mbarrier ---------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int crd_inner, int crd_outer) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_addr(dst)),
                 "l"(map), "r"(smem_addr(bar)), "r"(crd_inner), "r"(crd_outer)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_descriptor(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- TMA: 2-D tensor store shared -> global (bulk async-group completion) ----------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int crd_inner, int crd_outer) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_addr(src)), "r"(crd_inner), "r"(crd_outer)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's store groups may still be READING shared memory / still be in flight
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------------
// One full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (PTX "matrix-descriptor", sm_100 version 1): start address, leading- and
// stride-dimension byte offsets (all >> 4), layout type in bits [61,64).
constexpr uint64_t kLayoutNone = 0, kLayoutSw128 = 2;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
// K-major operand in the 128-byte-swizzled tile: 8-row groups 1024 bytes apart; the leading offset is unused (one
// MMA's K extent, 32 bytes, stays inside a swizzled row) and set to 16 bytes as CUTLASS does.
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t addr) { return make_smem_desc(addr, 16, 1024, kLayoutSw128); }
// MN-major operand in the same tile: the 128-byte row holds 64 consecutive M/N indices; the next 64 indices start
// `mn_group_bytes` further (leading offset), the next 8 k's 1024 bytes further (stride offset).
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t addr, uint32_t mn_group_bytes) {
    return make_smem_desc(addr, mn_group_bytes, 1024, kLayoutSw128);
}
// K-major operand WITHOUT swizzle, stored as 8x8 core matrices (8 rows x 16 bytes contiguous = 128 bytes):
// the next core matrix along K is `k_step_bytes` away (leading offset), the next 8 rows `mn_step_bytes` away.
__device__ __forceinline__ uint64_t make_desc_kmajor_core(uint32_t addr, uint32_t k_step_bytes, uint32_t mn_step_bytes) {
    return make_smem_desc(addr, k_step_bytes, mn_step_bytes, kLayoutNone);
}

// Instruction descriptor of tcgen05.mma kind::f16 / kind::tf32 with fp32 accumulation (dense, no negation):
// bits [4,6) D format (1 = f32), [7,10) A format, [10,13) B format (0 = f16, 1 = bf16, 2 = tf32), 15 / 16 A / B major
// (0 = K-major, 1 = MN-major), [17,23) N >> 3, [24,29) M >> 4.
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2;
constexpr uint32_t make_instr_desc(uint32_t fmt_a, uint32_t fmt_b, int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA issue / completion (ONE thread issues on behalf of the CTA) -------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; accumulate == 0 overwrites D.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier when every MMA issued so far by this thread has completed (implies the before-sync fence).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes (the warp's quarter: lanes 32*(warp%4)..) x 32 consecutive fp32 columns -----------
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- host: tensor map of a row-major [rows, cols] bf16 matrix, box = [box_rows, 64 columns], 128-byte swizzle ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
inline int make_tensor_map_bf16_rows(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return -1;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// row-major [rows, cols] bf16 matrix, box = [box_rows, box_cols], no swizzle (dense box in shared memory: TMA stores)
inline int make_tensor_map_bf16_dense(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                                      uint32_t box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return -1;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace umma
}  // namespace dddm
