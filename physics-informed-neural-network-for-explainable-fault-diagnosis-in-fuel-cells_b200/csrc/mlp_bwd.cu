// mlp_bwd.cu -- K2: backward of DNN.forward (what loss.backward() at 01:953 computes
// through autograd), with the aleatoric loss 01:916-927 optionally fused in.
//
// One kernel per training step: each CTA walks tiles of NT samples; for a tile it
//   1. recomputes the forward per sample (thread-per-sample, net.cuh), keeping every
//      masked activation in the CTA's private rows (shared memory for H<=64, a global
//      scratch otherwise) -- activations never round-trip HBM for the headline net;
//   2. forms dL/du, dL/dlogvar in registers (fused loss) or reads them (autograd path);
//   3. back-propagates per sample (dgrad, same FFMA2 machinery, transposed access);
//   4. contracts delta x activation over the tile's samples into weight gradients with
//      a register-tiled 4x8 micro-kernel over the [feature][sample] rows (the batch is
//      the contraction index -- SURVEY H4), accumulating into this CTA's partial.
// A second tiny kernel sums the per-CTA partials in fixed order: deterministic, no
// float atomics.  dL/dx of layer 0 -- computed and thrown away by the reference every
// step because self.x has requires_grad (01:446) -- is never formed.
#include "net.cuh"
#include "tc_api.cuh"

namespace pinn {

// out[c] = sum_i W[i][c] * in[i] (+ wx[c] * sx), W natural [KIN][NOUT]; blocks of 16.
template <int KIN, int NOUT, bool WG, class Epi>
PINN_D void tps_dense_T(Col in, const float* W, const float* wx, float sx, Epi epi) {
  constexpr int CB = NOUT >= 16 ? 16 : NOUT;
  static_assert(NOUT % CB == 0 && CB % 4 == 0, "tile");
  float hin[KIN <= 64 ? KIN : 1];
  if constexpr (KIN <= 64) {
#pragma unroll
    for (int i = 0; i < KIN; ++i) hin[i] = in.get(i);
  }
#pragma unroll 1
  for (int c0 = 0; c0 < NOUT; c0 += CB) {
    float2 acc[CB / 2];
#pragma unroll
    for (int q = 0; q < CB / 2; ++q) acc[q] = make_float2(0.f, 0.f);
    if constexpr (KIN <= 64) {
#pragma unroll
      for (int i = 0; i < KIN; ++i) {
        const float2 d2 = make_float2(hin[i], hin[i]);
#pragma unroll
        for (int q = 0; q < CB / 4; ++q) {
          float4 w4 = ldw4<WG>(W + static_cast<size_t>(i) * NOUT + c0 + 4 * q);
          acc[2 * q] = ffma2(make_float2(w4.x, w4.y), d2, acc[2 * q]);
          acc[2 * q + 1] = ffma2(make_float2(w4.z, w4.w), d2, acc[2 * q + 1]);
        }
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < KIN; ++i) {
        const float v = in.get(i);
        const float2 d2 = make_float2(v, v);
#pragma unroll
        for (int q = 0; q < CB / 4; ++q) {
          float4 w4 = ldw4<WG>(W + static_cast<size_t>(i) * NOUT + c0 + 4 * q);
          acc[2 * q] = ffma2(make_float2(w4.x, w4.y), d2, acc[2 * q]);
          acc[2 * q + 1] = ffma2(make_float2(w4.z, w4.w), d2, acc[2 * q + 1]);
        }
      }
    }
    float z[CB];
#pragma unroll
    for (int q = 0; q < CB / 2; ++q) { z[2 * q] = acc[q].x; z[2 * q + 1] = acc[q].y; }
    if (wx != nullptr) {
#pragma unroll
      for (int q = 0; q < CB; ++q) z[q] = fmaf(ldw1<WG>(wx + c0 + q), sx, z[q]);
    }
    epi(c0, z);
  }
}

// delta_pre[c] = delta_post[c] * keep_mask[c] * (1 - a[c]^2), a recovered from the stored
// masked activation (a = a_masked * (1-p) wherever the unit was kept).
template <int CB>
struct DgradEpi {
  Col out, act;
  const DropCtx* dc;
  uint32_t layer, unit_base;
  PINN_D void operator()(int c0, float (&z)[CB]) const {
#pragma unroll
    for (int g = 0; g < CB; g += 8) {
      float m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (dc->active) drop8(*dc, layer, c0 + g, unit_base, m);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float a = act.get(c0 + g + q) * (dc->active ? dc->keep : 1.0f);
        out.set(c0 + g + q, z[g + q] * m[q] * (1.0f - a * a));
      }
    }
  }
};

// dW[j][k] += sum_m D[j][m] * A[k][m] over the tile's samples (rows are [feature][sample],
// row stride RS); thread tile 4x8, FFMA2 paired along m.  db[j] += sum_m D[j][m].
template <bool GLOBAL_ROWS>
PINN_D void wgrad_mt(const float* D, int J, const float* A, int K, size_t RS, int M, float* dW, float* db) {
  const int tiles_k = K / 8, tiles = (J / 4) * tiles_k;
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
    const int jt = t / tiles_k, kt = t - jt * tiles_k;
    const float* d = D + static_cast<size_t>(4 * jt) * RS;
    const float* a = A + static_cast<size_t>(8 * kt) * RS;
    float2 acc[4][8];
    float2 bs[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      bs[jj] = make_float2(0.f, 0.f);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) acc[jj][kk] = make_float2(0.f, 0.f);
    }
#pragma unroll 1
    for (int m = 0; m < M; m += 4) {
      float4 dv[4], av[8];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) dv[jj] = ldw4<GLOBAL_ROWS>(d + jj * RS + m);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) av[kk] = ldw4<GLOBAL_ROWS>(a + kk * RS + m);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        bs[jj].x += dv[jj].x + dv[jj].y;
        bs[jj].y += dv[jj].z + dv[jj].w;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          acc[jj][kk] = ffma2(make_float2(dv[jj].x, dv[jj].y), make_float2(av[kk].x, av[kk].y), acc[jj][kk]);
          acc[jj][kk] = ffma2(make_float2(dv[jj].z, dv[jj].w), make_float2(av[kk].z, av[kk].w), acc[jj][kk]);
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) dW[static_cast<size_t>(4 * jt + jj) * K + 8 * kt + kk] += acc[jj][kk].x + acc[jj][kk].y;
      if (kt == 0) db[4 * jt + jj] += bs[jj].x + bs[jj].y;
    }
  }
}
// J == 1 (predict head, last variance layer): dW[k] += sum_m d[m] * A[k][m].
template <bool GLOBAL_ROWS>
PINN_D void wgrad_vec(const float* d, const float* A, int K, size_t RS, int M, float* dW, float* db) {
  for (int k = threadIdx.x; k < K + 1; k += blockDim.x) {
    float2 acc = make_float2(0.f, 0.f), acc2 = acc;
    if (k < K) {
      const float* a = A + static_cast<size_t>(k) * RS;
      for (int m = 0; m < M; m += 4) {
        float4 dv = ldw4<GLOBAL_ROWS>(d + m), av = ldw4<GLOBAL_ROWS>(a + m);
        acc = ffma2(make_float2(dv.x, dv.y), make_float2(av.x, av.y), acc);
        acc2 = ffma2(make_float2(dv.z, dv.w), make_float2(av.z, av.w), acc2);
      }
      dW[k] += (acc.x + acc.y) + (acc2.x + acc2.y);
    } else {
      float sum = 0.f;
      for (int m = 0; m < M; m += 4) {
        float4 dv = ldw4<GLOBAL_ROWS>(d + m);
        sum += (dv.x + dv.y) + (dv.z + dv.w);
      }
      db[0] += sum;
    }
  }
}

struct BwdArgs {
  const float* x; int64_t n; int64_t n_tiles;
  const float* grad_u; const float* grad_s; const float* y; float inv_n_global;
  float* partial;        // [grid][lay.total]
  double* loss_partial;  // [grid][4]
  float* scratch;        // LARGE rows
};

template <int H, bool LARGE>
__global__ void __launch_bounds__(128)
mlp_bwd_kernel(pinn_net_t net, ParamLayout lay, DropParams dp, BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int L = lay.L, NT = blockDim.x, tid = threadIdx.x;
  const int Dm = L * H + H / 2;
  Weights<LARGE> w{&net, smem, &lay};
  if constexpr (!LARGE) stage_weights(smem, net, lay);
  // row map
  const int rowV0 = L * H, rowV1 = rowV0 + H / 2, rowX = rowV1 + H / 4, rowD = rowX + PINN_N_IN;
  const int rowD2 = LARGE ? rowD + H : rowD, rowHD = LARGE ? rowD + 2 * H : rowD;
  float* rows;
  size_t RS;
  if constexpr (LARGE) {
    RS = static_cast<size_t>(gridDim.x) * NT;
    rows = a.scratch + static_cast<size_t>(blockIdx.x) * NT;
  } else {
    RS = NT + 4;
    rows = smem + lay.total;
  }
  auto col = [&](int row) { return Col{rows + static_cast<size_t>(row) * RS + tid, static_cast<int>(RS)}; };
  auto rowp = [&](int row) { return rows + static_cast<size_t>(row) * RS; };
  float* part = a.partial + static_cast<size_t>(blockIdx.x) * lay.total;
  for (int64_t i = tid; i < lay.total; i += NT) part[i] = 0.f;
  __syncthreads();
  const int hd1 = rowHD, hd2 = rowHD + H / 2, hdv = rowHD + H / 2 + H / 4, hdu = hdv + 1;

  double l_nll = 0.0, l_abs = 0.0, l_mse = 0.0, l_cnt = 0.0;
  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t s = tile * NT + tid;
    const bool valid = s < a.n;
    // ------------------------------------------------ forward recompute (per sample)
    float xr[PINN_N_IN];
    if (valid) {
      const float4* px = reinterpret_cast<const float4*>(a.x + s * PINN_N_IN);
      float4 q0 = __ldg(px), q1 = __ldg(px + 1);
      xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
    } else {
#pragma unroll
      for (int i = 0; i < PINN_N_IN; ++i) xr[i] = 0.f;
    }
    {
      Col cx = col(rowX);
#pragma unroll
      for (int i = 0; i < PINN_N_IN; ++i) cx.set(i, xr[i]);
    }
    DropCtx dc = make_ctx(dp, s, 0, Dm, true, valid);

    {
      TanhDropStore<8> e0{col(0), &dc, 0u, 0u};
      layer0<H, LARGE>(xr, w.W(0), w.b(0), e0);
    }
    for (int l = 1; l < L; ++l) {
      TanhDropStore<8> e{col(l * H), &dc, static_cast<uint32_t>(l), static_cast<uint32_t>(l * H)};
      tps_dense<H, H, 8, LARGE>(col((l - 1) * H), w.W(l), w.b(l), e);
    }
    const Col cL = col((L - 1) * H);
    const float u = tps_dot1<H, LARGE>(cL, w.Wp(), w.bp());
    {
      TanhDropStore<8> e{col(rowV0), &dc, static_cast<uint32_t>(L), static_cast<uint32_t>(L * H)};
      tps_dense<H, H / 2, 8, LARGE>(cL, w.Wv0(), w.bv0(), e);
    }
    {
      TanhStore<8> e{col(rowV1)};
      tps_dense<H / 2, H / 4, 8, LARGE>(col(rowV0), w.Wv1(), w.bv1(), e);
    }
    const float v = tps_dot1<H / 4, LARGE>(col(rowV1), w.Wv2(), w.bv2());
    const bool no_lv = (net.flags & PINN_NET_NO_LOGVAR) != 0;
    const float slv = logvar_out(v, no_lv);
    // ------------------------------------------------ upstream gradients
    float du = 0.f, ds = 0.f;
    if (valid) {
      if (a.grad_u != nullptr) {
        du = __ldg(a.grad_u + s);
        ds = a.grad_s ? __ldg(a.grad_s + s) : 0.f;
      } else {
        const float yv = __ldg(a.y + s);
        const float e = expf(-slv), diff = yv - u;
        du = -e * diff * a.inv_n_global;
        const float sg = slv > 0.f ? 1.f : (slv < 0.f ? -1.f : 0.f);
        ds = (-0.5f * e * diff * diff + 0.5f + 0.01f * sg) * a.inv_n_global;
        l_nll += static_cast<double>(0.5f * e * diff * diff + 0.5f * slv);
        l_abs += static_cast<double>(fabsf(slv));
        l_mse += static_cast<double>(diff * diff);
        l_cnt += 1.0;
      }
    }
    // ------------------------------------------------ variance-head backward (per sample)
    const float dv = no_lv ? 0.f : ds * dlogvar_dv(v);
    col(hdv).set(0, dv);
    col(hdu).set(0, du);
    {
      Col c1 = col(rowV1), o = col(hd2);
      const float* wv2 = w.Wv2();
#pragma unroll 4
      for (int k = 0; k < H / 4; ++k) {
        const float a1 = c1.get(k);
        o.set(k, dv * ldw1<LARGE>(wv2 + k) * (1.0f - a1 * a1));
      }
    }
    {
      DgradEpi<(H / 2 >= 16 ? 16 : H / 2)> e{col(hd1), col(rowV0), &dc, static_cast<uint32_t>(L),
                                             static_cast<uint32_t>(L * H)};
      tps_dense_T<H / 4, H / 2, LARGE>(col(hd2), w.Wv1(), nullptr, 0.f, e);
    }
    __syncthreads();
    // ------------------------------------------------ head weight gradients (CTA tile)
    wgrad_vec<LARGE>(rowp(hdv), rowp(rowV1), H / 4, RS, NT, part + lay.offWv2, part + lay.offbv2);
    wgrad_mt<LARGE>(rowp(hd2), H / 4, rowp(rowV0), H / 2, RS, NT, part + lay.offWv1, part + lay.offbv1);
    wgrad_mt<LARGE>(rowp(hd1), H / 2, rowp((L - 1) * H), H, RS, NT, part + lay.offWv0, part + lay.offbv0);
    wgrad_vec<LARGE>(rowp(hdu), rowp((L - 1) * H), H, RS, NT, part + lay.offWp, part + lay.offbp);
    __syncthreads();
    // ------------------------------------------------ into the trunk
    int cur = rowD, nxt = rowD2;
    {
      DgradEpi<16> e{col(cur), cL, &dc, static_cast<uint32_t>(L - 1), static_cast<uint32_t>((L - 1) * H)};
      tps_dense_T<H / 2, H, LARGE>(col(hd1), w.Wv0(), w.Wp(), du, e);
    }
    for (int l = L - 1; l >= 1; --l) {
      __syncthreads();
      wgrad_mt<LARGE>(rowp(cur), H, rowp((l - 1) * H), H, RS, NT, part + lay.offW[l], part + lay.offb[l]);
      __syncthreads();
      DgradEpi<16> e{col(nxt), col((l - 1) * H), &dc, static_cast<uint32_t>(l - 1),
                     static_cast<uint32_t>((l - 1) * H)};
      tps_dense_T<H, H, LARGE>(col(cur), w.W(l), nullptr, 0.f, e);
      int t = cur; cur = nxt; nxt = t;
    }
    __syncthreads();
    wgrad_mt<LARGE>(rowp(cur), H, rowp(rowX), PINN_N_IN, RS, NT, part + lay.offW[0], part + lay.offb[0]);
    __syncthreads();
  }
  // ---------------------------------------------------- loss partials (block reduce)
  __shared__ double lred[4][4];
  double vals[4] = {l_nll, l_abs, l_mse, l_cnt};
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double t = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && warp < 4) lred[warp][k] = t;
  }
  __syncthreads();
  if (tid < 4) {
    double t = 0.0;
    for (int wv = 0; wv < (NT + 31) / 32 && wv < 4; ++wv) t += lred[wv][tid];
    a.loss_partial[static_cast<size_t>(blockIdx.x) * 4 + tid] = t;
  }
}

__global__ void grad_reduce_kernel(const float* __restrict__ partial, const double* __restrict__ loss_partial,
                                   int nblk, int64_t total, float* __restrict__ grad, double* __restrict__ loss,
                                   const ParamLayout lay) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) {
    double acc = 0.0;
    if (!layout_is_padding(lay, i))          // alignment padding of the bucket: written as 0
      for (int b = 0; b < nblk; ++b) acc += static_cast<double>(partial[static_cast<size_t>(b) * total + i]);
    grad[i] = static_cast<float>(acc);
  }
  if (loss != nullptr && blockIdx.x == 0 && threadIdx.x < 4) {
    double acc = 0.0;
    for (int b = 0; b < nblk; ++b) acc += loss_partial[static_cast<size_t>(b) * 4 + threadIdx.x];
    loss[threadIdx.x] = acc;
  }
}

struct BwdPlan { bool large; int nt, grid; size_t smem, rows; int64_t n_tiles; size_t off_partial, off_scratch, bytes; };

static BwdPlan plan_bwd(int H, int L, int64_t n) {
  ParamLayout lay = make_layout(H, L);
  BwdPlan p{};
  const int sms = sm_count();
  const size_t rows_small = static_cast<size_t>(L) * H + H / 2 + H / 4 + PINN_N_IN + H;
  p.large = true;
  if (H <= 64) {
    for (int nt : {128, 64, 32}) {
      size_t need = (static_cast<size_t>(lay.total) + rows_small * (nt + 4)) * sizeof(float);
      if (need <= 227 * 1024) { p.large = false; p.nt = nt; p.smem = need; p.rows = rows_small; break; }
    }
  }
  if (p.large) { p.nt = 128; p.smem = 0; p.rows = rows_small + 2 * static_cast<size_t>(H); }
  p.n_tiles = (n + p.nt - 1) / p.nt;
  int64_t cap = static_cast<int64_t>(sms) * (p.large ? 2 : 1);
  if (!p.large) { int per = static_cast<int>((228 * 1024) / (p.smem + 1024)); if (per > 1) cap = static_cast<int64_t>(sms) * per; }
  p.grid = static_cast<int>(p.n_tiles < cap ? (p.n_tiles > 0 ? p.n_tiles : 1) : cap);
  size_t off = static_cast<size_t>(p.grid) * 4 * sizeof(double);
  p.off_partial = off;
  off += static_cast<size_t>(p.grid) * lay.total * sizeof(float);
  off = (off + 255) & ~static_cast<size_t>(255);
  p.off_scratch = off;
  if (p.large) off += p.rows * static_cast<size_t>(p.grid) * p.nt * sizeof(float);
  p.bytes = off;
  return p;
}

}  // namespace pinn

using namespace pinn;

extern "C" int64_t pinn_param_count(int32_t width, int32_t n_hidden) {
  if (n_hidden < 1 || n_hidden > PINN_MAX_HIDDEN) return PINN_E_SHAPE;
  return make_layout(width, n_hidden).total;
}

extern "C" size_t pinn_mlp_bwd_workspace_bytes(int32_t width, int32_t n_hidden, int64_t n) {
  size_t a = plan_bwd(width, n_hidden, n).bytes;
  if (width == 64 && n_hidden >= 2 && n_hidden <= 4) {
    size_t b = tc_bwd_workspace_bytes(n_hidden, n);
    if (b > a) a = b;
  }
  const size_t w = wide_tc_bwd_workspace_bytes(width, n_hidden, n);
  if (w > a) a = w;
  return a;
}

extern "C" size_t pinn_mlp_bwd_workspace_bytes_flags(int32_t width, int32_t n_hidden, int64_t n, int32_t flags) {
  if (width == 64 && n_hidden >= 2 && n_hidden <= 4 && !(flags & PINN_NET_NO_TC_BWD)) return tc_bwd_workspace_bytes(n_hidden, n, flags);
  return pinn_mlp_bwd_workspace_bytes(width, n_hidden, n);
}

extern "C" int pinn_mlp_bwd(const pinn_net_t* net, const float* x, int64_t n, const pinn_dropout_t* drop,
                            const float* grad_u, const float* grad_logvar, const float* y, int64_t n_global,
                            float* grad_flat, double* loss_sums, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (int e = validate_net(net)) return e;
  if (n < 0 || !grad_flat || !workspace) return PINN_E_ARG;
  if (n > 0 && !x) return PINN_E_ARG;
  if (n > 0 && !grad_u && !y) return PINN_E_ARG;
  if (!grad_u && n_global <= 0) return PINN_E_ARG;
  if (!aligned16(x) || !aligned16(workspace)) return PINN_E_ALIGN;
  const int H = net->width, L = net->n_hidden;
  if (n > 0 && wide_tc_bwd_covers(net))
    return launch_wide_tc_bwd(net, x, n, make_drop_params(drop), grad_u, grad_logvar, y, n_global, grad_flat, loss_sums, workspace,
                              workspace_bytes, static_cast<cudaStream_t>(stream));
  if (n > 0 && tc_bwd_covers(net))
    return launch_tc_bwd(net, x, n, make_drop_params(drop), grad_u, grad_logvar, y, n_global, grad_flat, loss_sums, workspace,
                         workspace_bytes, static_cast<cudaStream_t>(stream));
  BwdPlan p = plan_bwd(H, L, n);
  if (workspace_bytes < p.bytes) return PINN_E_WORKSPACE;
  if (p.large) {
    for (int l = 0; l < L; ++l) if (!aligned16(net->W[l])) return PINN_E_ALIGN;
    if (!aligned16(net->Wp) || !aligned16(net->Wv0) || !aligned16(net->Wv1) || !aligned16(net->Wv2)) return PINN_E_ALIGN;
  }
  ParamLayout lay = make_layout(H, L);
  DropParams dp = make_drop_params(drop);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  BwdArgs a{};
  a.x = x; a.n = n; a.n_tiles = p.n_tiles;
  a.grad_u = grad_u; a.grad_s = grad_logvar; a.y = y;
  a.inv_n_global = grad_u ? 0.f : static_cast<float>(1.0 / static_cast<double>(n_global));
  a.loss_partial = reinterpret_cast<double*>(ws);
  a.partial = reinterpret_cast<float*>(ws + p.off_partial);
  a.scratch = reinterpret_cast<float*>(ws + p.off_scratch);
#define CALL(HH, LG)                                                                          \
  {                                                                                           \
    if (p.smem > 48 * 1024)                                                                   \
      PINN_CUDA_TRY(cudaFuncSetAttribute(mlp_bwd_kernel<HH, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         static_cast<int>(p.smem)));                          \
    mlp_bwd_kernel<HH, LG><<<p.grid, p.nt, p.smem, st>>>(*net, lay, dp, a);                   \
  }
  switch (H) {
    case 32: if (p.large) CALL(32, true) else CALL(32, false) break;
    case 64: if (p.large) CALL(64, true) else CALL(64, false) break;
    case 128: CALL(128, true) break;
    case 256: CALL(256, true) break;
    default: return PINN_E_SHAPE;
  }
#undef CALL
  PINN_CUDA_TRY(cudaGetLastError());
  const int rb = 256;
  const int rg = static_cast<int>((lay.total + rb - 1) / rb);
  grad_reduce_kernel<<<rg, rb, 0, st>>>(a.partial, a.loss_partial, p.grid, lay.total, grad_flat, loss_sums, lay);
  return static_cast<int>(cudaGetLastError());
}

// One whole train_dnn step (01:948-955: forward in train mode, aleatoric loss, backward, Adam.step, StepLR.step) as
// ONE call.  On the tensor-core backward path the gradient reduce and the optimiser are the same launch; on the other
// paths it is pinn_mlp_bwd followed by pinn_adam_step.  `net`'s tensors must be views into `params_flat` in the
// pinn_param_count layout (that is what makes one flat optimiser launch possible).
static int train_dnn_step_impl(const pinn_net_t* net, const float* x, int64_t n, const pinn_dropout_t* drop, const float* y,
                               int64_t n_global, float* params_flat, float* exp_avg, float* exp_avg_sq, int64_t* step_counter,
                               double lr0, double gamma, int64_t step_size, float* grad_flat, double* loss_sums, void* workspace,
                               size_t workspace_bytes, void* stream, int images_valid) {
  if (int e = validate_net(net)) return e;
  if (n <= 0 || !x || !y || n_global <= 0 || !params_flat || !exp_avg || !exp_avg_sq || !step_counter || step_size <= 0 ||
      !grad_flat || !workspace)
    return PINN_E_ARG;
  ParamLayout lay = make_layout(net->width, net->n_hidden);
  if (net->W[0] != params_flat + lay.offW[0] || net->bv2 != params_flat + lay.offbv2) return PINN_E_ARG;
  if (!aligned16(x) || !aligned16(workspace)) return PINN_E_ALIGN;
  if (!wide_tc_bwd_covers(net) && tc_bwd_covers(net)) {
    FusedAdam fa{params_flat, exp_avg, exp_avg_sq, step_counter, AdamHyper{lr0, gamma, 1.0, step_size}, nullptr, 1.0f, images_valid};
    return launch_tc_bwd(net, x, n, make_drop_params(drop), nullptr, nullptr, y, n_global, grad_flat, loss_sums, workspace,
                         workspace_bytes, static_cast<cudaStream_t>(stream), &fa);
  }
  if (int e = pinn_mlp_bwd(net, x, n, drop, nullptr, nullptr, y, n_global, grad_flat, loss_sums, workspace, workspace_bytes, stream))
    return e;
  return pinn_adam_step(params_flat, grad_flat, exp_avg, exp_avg_sq, lay.total, step_counter, lr0, gamma, step_size, 1.0, nullptr,
                        nullptr, nullptr, 1, stream);
}

extern "C" int pinn_train_dnn_step(const pinn_net_t* net, const float* x, int64_t n, const pinn_dropout_t* drop, const float* y,
                                   int64_t n_global, float* params_flat, float* exp_avg, float* exp_avg_sq, int64_t* step_counter,
                                   double lr0, double gamma, int64_t step_size, float* grad_flat, double* loss_sums, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  return train_dnn_step_impl(net, x, n, drop, y, n_global, params_flat, exp_avg, exp_avg_sq, step_counter, lr0, gamma, step_size,
                             grad_flat, loss_sums, workspace, workspace_bytes, stream, 0);
}

// `n_steps` consecutive train_dnn steps enqueued by one call (the Python loop around pinn_train_dnn_step costs about as
// much host time per step as the three launches take on the device at the reference's batch sizes).  Step i draws its
// dropout masks with pass_offset = drop->pass_offset + i -- what the per-step loop passes; injected masks cannot be
// advanced here, so `drop->masks` must be NULL.  loss_sums holds the sums of the LAST step.
extern "C" int pinn_train_dnn_steps(const pinn_net_t* net, const float* x, int64_t n, const pinn_dropout_t* drop, const float* y,
                                    int64_t n_global, float* params_flat, float* exp_avg, float* exp_avg_sq, int64_t* step_counter,
                                    double lr0, double gamma, int64_t step_size, int64_t n_steps, float* grad_flat,
                                    double* loss_sums, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_steps < 0) return PINN_E_ARG;
  if (drop != nullptr && drop->masks != nullptr && n_steps > 1) return PINN_E_ARG;
  pinn_dropout_t d{};
  if (drop != nullptr) d = *drop;
  for (int64_t i = 0; i < n_steps; ++i) {
    if (drop != nullptr) d.pass_offset = drop->pass_offset + i;
    // from the second step on the one-kernel backward finds its weight images written by the previous optimiser launch
    const int r = train_dnn_step_impl(net, x, n, drop != nullptr ? &d : nullptr, y, n_global, params_flat, exp_avg, exp_avg_sq,
                                      step_counter, lr0, gamma, step_size, grad_flat, loss_sums, workspace, workspace_bytes, stream,
                                      i > 0 ? 1 : 0);
    if (r != 0) return r;
  }
  return 0;
}

// Data-parallel form of pinn_train_dnn_steps (one process per GPU): the sum of the gradient bucket over the ranks happens
// INSIDE the gradient-reduce launch over NVLink peer memory (mlp_tc_bwd.cu, grad_reduce2_kernel), Adam + StepLR follow in
// the same launch, and all `n_steps` steps are enqueued by this one call -- no NCCL call, no extra launch and no host
// round trip per step.  `peer_buffers`: device array of `world` addresses, entry r = rank r's symmetric buffer of
// pinn_dp_bucket_words() 32-bit words (zeroed once before first use); step i carries tag `first_tag + i` (tags must grow
// by one per step over the life of the buffer, starting above 0).  Covers the 64-wide net with 2..4 hidden layers
// (PINN_E_SHAPE otherwise: use pinn_mlp_bwd + pinn_adam_step_p2p).
extern "C" int64_t pinn_dp_bucket_words(int32_t width, int32_t n_hidden, int32_t world) {
  if (n_hidden < 1 || n_hidden > PINN_MAX_HIDDEN || world < 1) return PINN_E_SHAPE;
  return dp_bucket_words(make_layout(width, n_hidden).total, world);
}

extern "C" int pinn_train_dnn_steps_dp(const pinn_net_t* net, const float* x, int64_t n, const pinn_dropout_t* drop, const float* y,
                                       int64_t n_global, float* params_flat, float* exp_avg, float* exp_avg_sq, int64_t* step_counter,
                                       double lr0, double gamma, int64_t step_size, int64_t n_steps, const uint64_t* peer_buffers,
                                       int32_t rank, int32_t world, uint32_t first_tag, float* grad_flat, double* loss_sums,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = validate_net(net)) return e;
  if (n <= 0 || !x || !y || n_global <= 0 || !params_flat || !exp_avg || !exp_avg_sq || !step_counter || step_size <= 0 || !workspace ||
      n_steps < 0 || !peer_buffers || world < 1 || world > 32 || rank < 0 || rank >= world || first_tag == 0)
    return PINN_E_ARG;
  if (drop != nullptr && drop->masks != nullptr && n_steps > 1) return PINN_E_ARG;
  if (wide_tc_bwd_covers(net) || !tc_bwd_covers(net)) return PINN_E_SHAPE;
  ParamLayout lay = make_layout(net->width, net->n_hidden);
  if (net->W[0] != params_flat + lay.offW[0] || net->bv2 != params_flat + lay.offbv2) return PINN_E_ARG;
  if (!aligned16(x) || !aligned16(workspace)) return PINN_E_ALIGN;
  pinn_dropout_t d{};
  if (drop != nullptr) d = *drop;
  for (int64_t i = 0; i < n_steps; ++i) {
    if (drop != nullptr) d.pass_offset = drop->pass_offset + i;
    FusedAdam fa{params_flat, exp_avg, exp_avg_sq, step_counter, AdamHyper{lr0, gamma, 1.0, step_size}, nullptr, 1.0f, i > 0 ? 1 : 0,
                 reinterpret_cast<const unsigned long long*>(peer_buffers), rank, world, first_tag + static_cast<uint32_t>(i)};
    const int r = launch_tc_bwd(net, x, n, make_drop_params(drop != nullptr ? &d : nullptr), nullptr, nullptr, y, n_global, grad_flat,
                                loss_sums, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), &fa);
    if (r != 0) return r;
  }
  return 0;
}
