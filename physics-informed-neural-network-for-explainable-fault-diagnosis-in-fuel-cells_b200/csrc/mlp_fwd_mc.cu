// mlp_fwd_mc.cu -- K1 (DNN.forward, 01:421-438) and K4 (get_MC_samples, 01:1413-1491).
//
// K4 design: the sweep over T dropout passes happens INSIDE the kernel.  Each thread
// keeps its sample's pass-invariant layer-0 activation tanh(W0 x + b0) (dropout acts
// after tanh, 01:401-404 -- SURVEY H6) in a private shared-memory column, redraws the
// Philox masks for every pass in registers, and folds (u_t, logvar_t) into running
// Welford statistics.  HBM traffic is 32 B in + 12 B out per sample per SWEEP; the
// (T,N,1) host arrays of 01:1475-1477 never exist.
#include "net.cuh"
#include "tc_api.cuh"

namespace pinn {

PINN_D void load_row(const float* __restrict__ x, int64_t s, float (&r)[PINN_N_IN]) {
  const float4* p = reinterpret_cast<const float4*>(x + s * PINN_N_IN);
  float4 a = __ldg(p), b = __ldg(p + 1);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
  r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}

// Column set-up shared by both kernels.  SMALL: columns in shared memory after the
// weight arena; LARGE: columns in the global scratch, one slot per resident thread.
template <int H, bool LARGE>
struct Cols {
  Col a0, bufA, bufB;
  __device__ Cols(float* smem_after_arena, float* scratch, bool need_a0) {
    const int nt = blockDim.x;
    if constexpr (!LARGE) {
      float* base = smem_after_arena + threadIdx.x;
      bufA = Col{base, nt};
      bufB = bufA;  // in place
      a0 = Col{base + static_cast<size_t>(H) * nt, nt};
    } else {
      const size_t slots = static_cast<size_t>(gridDim.x) * nt;
      float* base = scratch + static_cast<size_t>(blockIdx.x) * nt + threadIdx.x;
      bufA = Col{base, static_cast<int>(slots)};
      bufB = Col{base + static_cast<size_t>(H) * slots, static_cast<int>(slots)};
      a0 = Col{base + 2 * static_cast<size_t>(H) * slots, static_cast<int>(slots)};
    }
    (void)need_a0;
  }
};

template <int H, bool LARGE>
__global__ void __launch_bounds__(256)
mlp_fwd_kernel(pinn_net_t net, ParamLayout lay, const float* __restrict__ x, int64_t n, DropParams dp,
               float* __restrict__ out_u, float* __restrict__ out_s, float* scratch) {
  extern __shared__ __align__(16) float smem[];
  const int D = lay.L * H + H / 2;
  Weights<LARGE> w{&net, smem, &lay};
  if constexpr (!LARGE) {
    stage_weights(smem, net, lay);
    __syncthreads();
  }
  Cols<H, LARGE> cols(smem + lay.total, scratch, false);
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < n;
       s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float xr[PINN_N_IN];
    load_row(x, s, xr);
    DropCtx dc = make_ctx(dp, s, 0, D, true);
    TanhDropStore<8> e0{cols.bufA, &dc, 0u, 0u};
    layer0<H, LARGE>(xr, w.W(0), w.b(0), e0);
    float u, v;
    forward_tail<H, LARGE>(w, lay.L, cols.bufA, cols.bufB, dc, u, v);
    out_u[s] = u;
    out_s[s] = logvar_out(v, (net.flags & PINN_NET_NO_LOGVAR) != 0);
  }
}

template <int H, bool LARGE>
__global__ void __launch_bounds__(256)
mc_dropout_kernel(pinn_net_t net, ParamLayout lay, const float* __restrict__ x, int64_t n, int T,
                  DropParams dp, float* __restrict__ pred_mean, float* __restrict__ a_u,
                  float* __restrict__ e_u, float* __restrict__ raw_mean, float* __restrict__ raw_m2,
                  float* __restrict__ raw_slv, float* scratch) {
  extern __shared__ __align__(16) float smem[];
  const int D = lay.L * H + H / 2;
  Weights<LARGE> w{&net, smem, &lay};
  if constexpr (!LARGE) {
    stage_weights(smem, net, lay);
    __syncthreads();
  }
  Cols<H, LARGE> cols(smem + lay.total, scratch, true);
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < n;
       s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float xr[PINN_N_IN];
    load_row(x, s, xr);
    // pass-invariant layer-0 activation
    {
      TanhStore<8> e0{cols.a0};
      layer0<H, LARGE>(xr, w.W(0), w.b(0), e0);
    }
    // eval forward == the reference's mean over T identical eval passes (01:1442-1445,1480)
    DropCtx dc = make_ctx(dp, s, 0, D, false);
    float u, v;
    if (pred_mean != nullptr) {
#pragma unroll 4
      for (int k = 0; k < H; ++k) cols.bufA.set(k, cols.a0.get(k));
      forward_tail<H, LARGE>(w, lay.L, cols.bufA, cols.bufB, dc, u, v);
      pred_mean[s] = u;
    }
    float mean = 0.f, m2 = 0.f, slv = 0.f;
    for (int t = 0; t < T; ++t) {
      dc = make_ctx(dp, s, t, D, true);
#pragma unroll 1
      for (int k = 0; k < H; k += 8) {
        float m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (dc.active) drop8(dc, 0u, k, 0u, m);
#pragma unroll
        for (int q = 0; q < 8; ++q) cols.bufA.set(k + q, cols.a0.get(k + q) * m[q]);
      }
      forward_tail<H, LARGE>(w, lay.L, cols.bufA, cols.bufB, dc, u, v);
      float d = u - mean;
      mean += d / static_cast<float>(t + 1);
      m2 = fmaf(d, u - mean, m2);
      slv += logvar_out(v, (net.flags & PINN_NET_NO_LOGVAR) != 0);
    }
    if (raw_mean) raw_mean[s] = mean;
    if (raw_m2) raw_m2[s] = m2;
    if (raw_slv) raw_slv[s] = slv;
    const float invT = 1.0f / static_cast<float>(T > 0 ? T : 1);
    if (a_u) a_u[s] = sqrtf(expf(slv * invT));
    if (e_u) e_u[s] = sqrtf(fmaxf(m2, 0.f) * invT);
  }
}

// ----------------------------------------------------------------- launch planning
struct Plan {
  bool large;
  int nt;
  int grid;
  size_t smem;
  size_t scratch_floats;
};
constexpr size_t kMaxSmem = 227 * 1024;

// ncols = private columns per thread (each H floats).
static Plan plan_tps(int H, int L, int64_t n, int ncols) {
  ParamLayout lay = make_layout(H, L);
  Plan p{};
  const int sms = sm_count();
  if (H <= 64) {
    // largest block that fits; step down while the grid would leave SMs idle
    for (int nt : {256, 128, 64}) {
      size_t need = (static_cast<size_t>(lay.total) + static_cast<size_t>(ncols) * H * nt) * sizeof(float);
      if (need > kMaxSmem) continue;
      int64_t want = (n + nt - 1) / nt;
      if (want < sms && nt > 64) continue;
      p.large = false; p.nt = nt; p.smem = need;
      int per_sm = static_cast<int>((228 * 1024) / (need + 1024));
      if (per_sm < 1) per_sm = 1;
      if (per_sm * nt > 1024) per_sm = 1024 / nt;
      int64_t cap = static_cast<int64_t>(sms) * per_sm;
      p.grid = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
      p.scratch_floats = 0;
      return p;
    }
  }
  p.large = true; p.nt = 128; p.smem = 0;
  int64_t want = (n + p.nt - 1) / p.nt;
  int64_t cap = static_cast<int64_t>(sms) * 4;
  p.grid = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  p.scratch_floats = static_cast<size_t>(3) * H * p.grid * p.nt;
  return p;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    PINN_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return 0;
}

#define PINN_DISPATCH_H(H_, LARGE_, CALL)                            \
  switch (H_) {                                                      \
    case 32:  if (LARGE_) { CALL(32, true) } else { CALL(32, false) } break;   \
    case 64:  if (LARGE_) { CALL(64, true) } else { CALL(64, false) } break;   \
    case 128: { CALL(128, true) } break;                             \
    case 256: { CALL(256, true) } break;                             \
    default: return PINN_E_SHAPE;                                    \
  }

}  // namespace pinn

using namespace pinn;

// ncols = private activation columns per thread of the FFMA path (1 forward, 2 sweep); flags < 0: enough for any path
static size_t fwd_mc_workspace(int32_t width, int32_t n_hidden, int64_t n, int ncols, int flags) {
  Plan p = plan_tps(width, n_hidden, n, ncols);
  size_t a = p.scratch_floats * sizeof(float);
  // a call known to take a tensor-core path of the wide nets does not need the FFMA path's private columns
  if (flags >= 0 && (width == 128 || width == 256) && !(flags & PINN_NET_NO_WIDE_TC)) a = 0;
  const size_t b = wide_tc_workspace_bytes(width, n_hidden, n, flags);
  const size_t c = width == 64 ? tc_mc_workspace_bytes(n) : 0;      // per-chunk Welford triples of long sweeps (mlp_tc.cu)
  if (b > a) a = b;
  return c > a ? c : a;
}
extern "C" size_t pinn_mlp_fwd_workspace_bytes(int32_t width, int32_t n_hidden, int64_t n) { return fwd_mc_workspace(width, n_hidden, n, 1, -1); }
extern "C" size_t pinn_mc_workspace_bytes(int32_t width, int32_t n_hidden, int64_t n) { return fwd_mc_workspace(width, n_hidden, n, 2, -1); }
extern "C" size_t pinn_mlp_fwd_workspace_bytes_flags(int32_t width, int32_t n_hidden, int64_t n, int32_t flags) {
  return fwd_mc_workspace(width, n_hidden, n, 1, flags < 0 ? 0 : flags);
}
extern "C" size_t pinn_mc_workspace_bytes_flags(int32_t width, int32_t n_hidden, int64_t n, int32_t flags) {
  return fwd_mc_workspace(width, n_hidden, n, 2, flags < 0 ? 0 : flags);
}

extern "C" int pinn_mlp_fwd(const pinn_net_t* net, const float* x, int64_t n, const pinn_dropout_t* drop,
                            float* out_u, float* out_logvar, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (int e = validate_net(net)) return e;
  if (n < 0 || (n > 0 && (!x || !out_u || !out_logvar))) return PINN_E_ARG;
  if (n == 0) return 0;
  if (!aligned16(x)) return PINN_E_ALIGN;
  const int H = net->width, L = net->n_hidden;
  {
    int err = 0;
    TcOut o{out_u, out_logvar, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const int r = launch_tc(false, net, x, n, 1, make_drop_params(drop), o, static_cast<cudaStream_t>(stream), &err);
    if (r == 1) return 0;
    if (r < 0) return err;
    const int rw = launch_wide_tc(false, net, x, n, 1, make_drop_params(drop), o, workspace, workspace_bytes,
                                  static_cast<cudaStream_t>(stream), &err);
    if (rw == 1) return 0;
    if (rw < 0) return err;
  }
  Plan p = plan_tps(H, L, n, 1);
  if (p.large) {
    if (workspace_bytes < p.scratch_floats * sizeof(float) || !workspace) return PINN_E_WORKSPACE;
    for (int l = 0; l < L; ++l) if (!aligned16(net->W[l])) return PINN_E_ALIGN;
    if (!aligned16(net->Wp) || !aligned16(net->Wv0) || !aligned16(net->Wv1) || !aligned16(net->Wv2)) return PINN_E_ALIGN;
  }
  ParamLayout lay = make_layout(H, L);
  DropParams dp = make_drop_params(drop);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(HH, LG)                                                                          \
  {                                                                                           \
    if (int e = set_smem(mlp_fwd_kernel<HH, LG>, p.smem)) return e;                          \
    mlp_fwd_kernel<HH, LG><<<p.grid, p.nt, p.smem, st>>>(*net, lay, x, n, dp, out_u, out_logvar, \
                                                         static_cast<float*>(workspace));    \
  }
  PINN_DISPATCH_H(H, p.large, CALL)
#undef CALL
  return static_cast<int>(cudaGetLastError());
}

extern "C" int pinn_mc_dropout(const pinn_net_t* net, const float* x, int64_t n, int32_t T,
                               const pinn_dropout_t* drop, float* pred_mean, float* a_u, float* e_u,
                               float* raw_mean, float* raw_m2, float* raw_sum_logvar, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (int e = validate_net(net)) return e;
  if (n < 0 || T < 0 || (n > 0 && !x)) return PINN_E_ARG;
  if (n == 0) return 0;
  if (!aligned16(x)) return PINN_E_ALIGN;
  const int H = net->width, L = net->n_hidden;
  {
    int err = 0;
    TcOut o{nullptr, nullptr, pred_mean, a_u, e_u, raw_mean, raw_m2, raw_sum_logvar};
    const int r = launch_tc(true, net, x, n, T, make_drop_params(drop), o, static_cast<cudaStream_t>(stream), &err, workspace,
                            workspace_bytes);
    if (r == 1) return 0;
    if (r < 0) return err;
    const int rw = launch_wide_tc(true, net, x, n, T, make_drop_params(drop), o, workspace, workspace_bytes,
                                  static_cast<cudaStream_t>(stream), &err);
    if (rw == 1) return 0;
    if (rw < 0) return err;
  }
  Plan p = plan_tps(H, L, n, 2);
  if (p.large) {
    if (workspace_bytes < p.scratch_floats * sizeof(float) || !workspace) return PINN_E_WORKSPACE;
    for (int l = 0; l < L; ++l) if (!aligned16(net->W[l])) return PINN_E_ALIGN;
    if (!aligned16(net->Wp) || !aligned16(net->Wv0) || !aligned16(net->Wv1) || !aligned16(net->Wv2)) return PINN_E_ALIGN;
  }
  ParamLayout lay = make_layout(H, L);
  DropParams dp = make_drop_params(drop);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(HH, LG)                                                                              \
  {                                                                                               \
    if (int e = set_smem(mc_dropout_kernel<HH, LG>, p.smem)) return e;                           \
    mc_dropout_kernel<HH, LG><<<p.grid, p.nt, p.smem, st>>>(*net, lay, x, n, T, dp, pred_mean, a_u, e_u, \
                                                            raw_mean, raw_m2, raw_sum_logvar,     \
                                                            static_cast<float*>(workspace));      \
  }
  PINN_DISPATCH_H(H, p.large, CALL)
#undef CALL
  return static_cast<int>(cudaGetLastError());
}

