// tc_api.cuh -- entry of the tensor-core path (mlp_tc.cu), called from mlp_fwd_mc.cu.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace pinn {
// TMA tensor map of the input matrix x [n][8] fp32 (row-major) with a [128 rows x 8 features] box: one
// `cp.async.bulk.tensor.2d` stages a tile's 4 KB of inputs into shared memory, rows past n arrive as zeros (mlp_tc.cu).
// false: no map (x not 16-byte aligned, n >= 2^31, or the driver entry point is missing) -- the kernels then load x with LDG.
bool make_x_tensor_map(CUtensorMap* map, const float* x, int64_t n);
struct TcOut {
  float* u; float* s;                                                                   // K1
  float* pred_mean; float* a_u; float* e_u; float* raw_mean; float* raw_m2; float* raw_slv;  // K4
};
// 1: launched on the tcgen05 path; 0: shape not covered (use the FFMA kernels); -1: error in *err.
int launch_tc(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
              cudaStream_t st, int* err, void* workspace = nullptr, size_t workspace_bytes = 0);
// Three-group fp16-pair form of the same kernel (mlp_tc3.cu): tried first by launch_tc; same return convention.
int launch_tc3(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
               cudaStream_t st, int* err, void* workspace, size_t workspace_bytes);
// long sweeps are cut into pass chunks (a function of T alone) whose Welford triples live in the workspace until merged
int mc_pass_chunks(int T);
size_t tc_mc_workspace_bytes(int64_t n);
// fold per-chunk Welford triples [chunk][3][n] (chunk k = passes [k Tc, (k+1) Tc), Tc = ceil(T / C)) in chunk order and finish the sample
void launch_mc_merge(const float* part, int64_t n, int T, int C, const TcOut& out, cudaStream_t st);
// Wide nets (H = 128 / 256) on the tensor cores, one GEMM launch per layer (mlp_wide_tc.cu); same return convention.
// flags < 0: enough for either 256-wide path; otherwise what the call carrying these pinn_net_t.flags needs (the resident-
// activation kernel needs weight images + per-chunk statistics, the per-layer GEMM path 5 KB per sample of operand planes)
size_t wide_tc_workspace_bytes(int H, int L, int64_t n, int flags = -1);
int launch_wide_tc(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
                   void* workspace, size_t workspace_bytes, cudaStream_t st, int* err);
// 256-wide nets, forward / MC sweep with the activations resident on the SM (mlp_wide_res.cu); same return convention.
size_t wide_res_workspace_bytes(int L, int64_t n);
int launch_wide_res(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
                    void* workspace, size_t workspace_bytes, cudaStream_t st, int* err);
bool wide_tc_bwd_covers(const pinn_net_t* net);
size_t wide_tc_bwd_workspace_bytes(int H, int L, int64_t n);
int launch_wide_tc_bwd(const pinn_net_t* net, const float* x, int64_t n, const DropParams& dp, const float* grad_u, const float* grad_s,
                       const float* y, int64_t n_global, float* grad_flat, double* loss_sums, void* workspace, size_t workspace_bytes,
                       cudaStream_t st);
// dependent-launch mode of this call (pinn_net_t.flags): 0 never, 1 small batches, 2 always (mlp_tc_bwd.cu)
int dependent_launch_mode(const pinn_net_t* net);
// Tensor-core backward (mlp_tc_bwd.cu): 64-wide nets with 2..4 hidden layers.
bool tc_bwd_covers(const pinn_net_t* net);
size_t tc_bwd_workspace_bytes(int L, int64_t n, int flags = -1);     // flags < 0: enough for any path
int launch_tc_bwd(const pinn_net_t* net, const float* x, int64_t n, const DropParams& dp, const float* grad_u,
                  const float* grad_s, const float* y, int64_t n_global, float* grad_flat, double* loss_sums, void* workspace,
                  size_t workspace_bytes, cudaStream_t st, const FusedAdam* fused = nullptr);
}  // namespace pinn
