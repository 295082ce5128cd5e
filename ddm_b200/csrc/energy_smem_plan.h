// energy_smem_plan.h — host-visible plan of the TMA-staged packed-fp32 kernel (no device code).
#pragma once
#include <cuda_runtime.h>

#include "energy.cuh"

namespace dddm {
struct SmemPlan {
    bool ok;
    int cluster, threads, slab_vecs, chunk_vecs;
    size_t smem_bytes;
};
// B > 0: rows of the launch (small minibatches are split along D so that B x cluster CTAs cover the SMs); 0 = not known
SmemPlan plan_smem(int m, int D, int elem_size, bool aligned16, int x0_rows = 1, int B = 0);
template <typename T>
int launch_energy_smem(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream);
// single-wave register-resident variant (energy_wave.cuh)
struct WavePlan {
    bool ok;
    int threads;  // compute threads (a control warp is added at launch)
    int nv;       // 16-byte vectors per thread
    int ksmem;    // pass-2 coefficients streamed from shared memory instead of held in 72 registers
};
WavePlan plan_wave(int B, int m, int D, int elem_size, bool aligned16);
template <typename T>
int launch_energy_wave(const EnergyParams& p, const WavePlan& plan, cudaStream_t stream);
// row-pipelined cluster kernel (energy_pipe.cuh): C CTAs own C rows, one column slab of each per CTA
struct PipePlan {
    bool ok;
    int cluster, threads, cols, slab_vecs, window;
    size_t smem_bytes;
};
PipePlan plan_pipe(int B, int m, int D, int elem_size, bool aligned16, int x0_rows = 1);
template <typename T>
int launch_energy_pipe(const EnergyParams& p, const PipePlan& plan, cudaStream_t stream);
// tensor-core kernel for m = 16, 32 bf16 draws (energy_tc.cu): Gram + coefficient mixing on tcgen05
struct TcPlan {
    bool ok;
    int stages;  // depth of the TMA ring (512-column stages)
    size_t smem_bytes;
};
TcPlan plan_tc(int B, int m, int D, int elem_size, bool aligned16);
int launch_energy_tc(const EnergyParams& p, const TcPlan& plan, cudaStream_t stream);
// blocked variant for m = 16, 32 (energy_blk.cuh)
SmemPlan plan_blk(int m, int D, int elem_size, bool aligned16);
template <typename T>
int launch_energy_blk(const EnergyParams& p, const SmemPlan& plan, cudaStream_t stream);
}  // namespace dddm
