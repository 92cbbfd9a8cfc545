// tc_api.cuh -- entry of the tensor-core path (mlp_tc.cu), called from mlp_fwd_mc.cu.
#pragma once
#include "common.cuh"

namespace pinn {
struct TcOut {
  float* u; float* s;                                                                   // K1
  float* pred_mean; float* a_u; float* e_u; float* raw_mean; float* raw_m2; float* raw_slv;  // K4
};
// 1: launched on the tcgen05 path; 0: shape not covered (use the FFMA kernels); -1: error in *err.
int launch_tc(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
              cudaStream_t st, int* err);
}  // namespace pinn
