// mlp_wide_tc.cu -- tensor-core forward / MC-dropout path for the WIDE nets (H = 256: the reference's own
// `Layers = [8,256,256,256,1]` (01:2139) and config 4's 6x256; H = 128 rides along).
//
// A 256-wide layer does not fit the resident-operand design of mlp_tc.cu (hi/lo activation planes alone would
// need 512 + 256 TMEM columns, one layer's split weights 512 KB), so each layer is ONE GEMM launch
//     out[128-row tile][N] = epilogue( A[tile][K] * W[N][K]^T ),      3xTF32 on tcgen05, fp32 accumulation in TMEM,
// whose operands are stored in global memory as pre-split tf32 hi / lo planes ALREADY IN THE SHARED-MEMORY IMAGE
// LAYOUT of the K-major no-swizzle UMMA operand: per 16-wide K chunk a contiguous [hi plane | lo plane] block,
//     byte(row, k) = (k / 4) * 16 * ROWS + row * 16 + (k % 4) * 4          (k = 0..15 inside the chunk).
// Both operands therefore arrive by plain `cp.async.bulk` (TMA 1-D bulk copies, mbarrier transaction bytes): no
// thread touches them.  The epilogue of layer l (bias, tanh, Philox keep-select, re-split) writes the planes of
// layer l+1; weights are split once per call (dropout scale folded in).
//
// CTA = one 128-row tile: warp 8 = producer (bulk copies, 2-3 stage ring), warp 9 = MMA issuer, warps 0..7 =
// epilogue (thread = (row, column half)).  96 KB of shared memory and <= 256 TMEM columns per CTA: two CTAs per SM,
// so one CTA's epilogue overlaps the other's MMAs.  Per tile-layer: 96 N=256 MMAs = 12.3 k clk of tensor work
// against 768 KB of operand traffic -- the L2 (42 B/clk/SM) paces it, about half the tensor peak.
//
// Launches per pass: layer 0 (K = 8, CUDA cores, writes the first planes) + (L-1) hidden GEMMs + heads GEMM
// ([Wv0; Wp], N = H/2 + 16) + variance-head GEMM (H/2 -> H/4) whose epilogue finishes the sample (last dot,
// log-variance, Welford update of the per-sample statistics in global memory).
// The mask stream (Philox counters per (sample, pass, layer, unit / 8)) is the one every other path uses.
#include "net.cuh"
#include "tc.cuh"
#include "tc_api.cuh"

namespace pinn {

constexpr int kWT = 128;          // rows per tile
constexpr int kWKc = 16;          // K per plane chunk = per pipeline stage
PINN_HD constexpr int w_plane(int rows) { return 4 * rows * 16; }            // bytes of one plane of one chunk
PINN_HD constexpr int w_chunk(int rows) { return 2 * w_plane(rows); }        // [hi | lo]
PINN_HD constexpr size_t w_tile_bytes(int K) { return static_cast<size_t>(K / kWKc) * w_chunk(kWT); }   // activation planes of one tile
PINN_HD constexpr size_t w_mat_bytes(int N, int K) { return static_cast<size_t>(K / kWKc) * w_chunk(N); }   // weight planes of one matrix

enum { EPI_HIDDEN = 0, EPI_HEADS = 1, EPI_V1 = 2 };

struct WideArgs {
  const unsigned char* A;      // activation planes in, [tile][K/16][hi|lo]
  const unsigned char* W;      // weight planes, [K/16][hi|lo]
  unsigned char* out;          // activation planes out (EPI_HIDDEN: K' = N; EPI_HEADS: K' = H/2)
  const float* bias;           // bias of this layer (b_l / bv0 / bv1)
  const float* bias2;          // EPI_HEADS: bp;  EPI_V1: bv2
  const float* w2;             // EPI_V1: Wv2
  float* u_io;                 // EPI_HEADS writes u[s]; EPI_V1 reads it
  int64_t n;
  uint32_t layer;              // dropout layer id of this epilogue's mask (EPI_HIDDEN: l, EPI_HEADS: L)
  uint32_t unit_base;          // byte offset of that layer inside an injected mask row
  int mask_row_bytes;          // D
  int active;                  // dropout active on this pass
  int pass;                    // local pass index t (injected-mask row, Welford count)
  float inact;                 // multiplier of an un-masked activation (undoes the folded scale)
  // EPI_V1 outputs
  int mode;                    // 0: forward (u, logvar); 1: eval pass of a sweep (pred_mean); 2: dropout pass t of T
  int T;
  float* out_u; float* out_s; float* pred_mean; float* a_u; float* e_u; float* st_mean; float* st_m2; float* st_slv;
};

// keep decisions of units [j0, j0 + 8) for sample s (global index sg) -- Philox or injected bytes
PINN_D void wide_keep8(const DropParams& dp, const WideArgs& a, uint64_t sg, int64_t s_local, uint32_t j0, bool (&k)[8]) {
  if (dp.masks != nullptr) {
    const uint8_t* mrow = dp.masks + (static_cast<size_t>(a.pass) * dp.mask_n + s_local) * a.mask_row_bytes + a.unit_base + j0;
#pragma unroll
    for (int q = 0; q < 8; ++q) k[q] = mrow[q] != 0;
  } else {
    const uint4 r = Philox::gen_rk(dp.rk, static_cast<uint32_t>(sg), static_cast<uint32_t>(sg >> 32),
                                   static_cast<uint32_t>(dp.pass_offset + a.pass), (a.layer << 16) | (j0 >> 3));
    keep8_from(r, dp.thresh_hi, k);
  }
}
// 8 activations of columns c0..c0+7 of row r -> hi / lo planes of a tile with 128 rows
PINN_D void wide_store8(unsigned char* tile_planes, int c0, int r, const float (&v)[8]) {
  float h[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) h[q] = tc::tf32_hi_fast(v[q]);
  unsigned char* p = tile_planes + static_cast<size_t>(c0 >> 4) * w_chunk(kWT) + ((c0 & 15) >> 2) * (kWT * 16) + r * 16;
  *reinterpret_cast<float4*>(p) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(p + kWT * 16) = make_float4(h[4], h[5], h[6], h[7]);
  *reinterpret_cast<float4*>(p + w_plane(kWT)) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
  *reinterpret_cast<float4*>(p + w_plane(kWT) + kWT * 16) = make_float4(v[4] - h[4], v[5] - h[5], v[6] - h[6], v[7] - h[7]);
}

// ------------------------------------------------------------------ weight planes (once per call)
// dst planes of an [N x K] matrix whose first `rows_a` rows come from `src_a` ([rows_a][K]), row `rows_a` from
// `src_b` (or zero) and the rest are zero; every element multiplied by `c`.
__global__ void wide_split_weights_kernel(const float* __restrict__ src_a, int rows_a, const float* __restrict__ src_b, int N, int K,
                                          float c, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // (row, k4)
  if (idx >= N * (K / 4)) return;
  const int nrow = idx % N, k4 = idx / N;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nrow < rows_a) v = __ldg(reinterpret_cast<const float4*>(src_a + static_cast<size_t>(nrow) * K) + k4);
  else if (nrow == rows_a && src_b != nullptr) v = __ldg(reinterpret_cast<const float4*>(src_b) + k4);
  v = make_float4(v.x * c, v.y * c, v.z * c, v.w * c);
  const float4 h = make_float4(tc::tf32_hi(v.x), tc::tf32_hi(v.y), tc::tf32_hi(v.z), tc::tf32_hi(v.w));
  unsigned char* p = dst + static_cast<size_t>(k4 >> 2) * w_chunk(N) + (k4 & 3) * (N * 16) + nrow * 16;
  *reinterpret_cast<float4*>(p) = h;
  *reinterpret_cast<float4*>(p + w_plane(N)) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// ------------------------------------------------------------------ layer 0 (K = 8): x -> masked a0 -> planes
template <int H>
__global__ void __launch_bounds__(256)
wide_layer0_kernel(const float* __restrict__ x, const float* __restrict__ W0, const float* __restrict__ b0,
                   const __grid_constant__ DropParams dp, WideArgs a) {
  __shared__ float sW[H * PINN_N_IN];
  __shared__ float sb[H];
  for (int i = threadIdx.x; i < H * PINN_N_IN; i += blockDim.x) sW[i] = __ldg(W0 + i) * kTanhArg;
  for (int i = threadIdx.x; i < H; i += blockDim.x) sb[i] = __ldg(b0 + i) * kTanhArg;
  __syncthreads();
  const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int64_t tile = blockIdx.x, s = tile * kWT + r;
  const bool valid = s < a.n;
  float xr[PINN_N_IN];
  if (valid) {
    const float4* px = reinterpret_cast<const float4*>(x + s * PINN_N_IN);
    const float4 q0 = __ldg(px), q1 = __ldg(px + 1);
    xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
  } else {
#pragma unroll
    for (int i = 0; i < PINN_N_IN; ++i) xr[i] = 0.f;
  }
  const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
  unsigned char* tp = a.out + static_cast<size_t>(tile) * w_tile_bytes(H);
  const bool act = a.active && valid;
#pragma unroll 1
  for (int c0 = half * (H / 2); c0 < (half + 1) * (H / 2); c0 += 8) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float* w = sW + (c0 + q) * PINN_N_IN;
      float z = sb[c0 + q];
#pragma unroll
      for (int i = 0; i < PINN_N_IN; ++i) z = fmaf(w[i], xr[i], z);
      v[q] = tanh_pre(z);
    }
    if (act) {
      bool k[8];
      wide_keep8(dp, a, sg, s, static_cast<uint32_t>(c0), k);
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = k[q] ? v[q] : 0.f;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = valid ? v[q] * a.inact : 0.f;
    }
    wide_store8(tp, c0, r, v);
  }
}

// ------------------------------------------------------------------ one layer = one GEMM launch
template <int N, int K, int EPI>
__global__ void __launch_bounds__(320, 2)
wide_gemm_kernel(const __grid_constant__ DropParams dp, WideArgs a) {
  constexpr int S = N >= 256 ? 2 : 3;                            // pipeline stages
  constexpr int A_CH = w_chunk(kWT), W_CH = w_chunk(N), STAGE = A_CH + W_CH;
  constexpr int NCH = K / kWKc;
  constexpr uint32_t TCOLS = N > 128 ? 256u : (N > 64 ? 128u : 64u);
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[S], empty[S], accum;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tc::uniform_warp_idx();
  const int64_t tile = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&accum, 1);
    tc::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, TCOLS); tc::tmem_relinquish(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s;

  if (warp == 8) {
    // ================================================================== producer: bulk copies into the ring
    if (tc::elect_one()) {
      const unsigned char* At = a.A + static_cast<size_t>(tile) * w_tile_bytes(K);
      for (int c = 0; c < NCH; ++c) {
        const int s = c % S;
        if (c >= S) tc::mbar_wait(&empty[s], static_cast<uint32_t>((c / S - 1) & 1));
        tc::mbar_expect_tx(&full[s], STAGE);
        tc::bulk_g2s(smem + s * STAGE, At + static_cast<size_t>(c) * A_CH, A_CH, &full[s]);
        tc::bulk_g2s(smem + s * STAGE + A_CH, a.W + static_cast<size_t>(c) * W_CH, W_CH, &full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ================================================================== MMA issuer
    const uint32_t idesc = tc::make_idesc_tf32(kWT, N);
    constexpr uint32_t LBO_A = kWT * 16, LBO_B = N * 16;
    for (int c = 0; c < NCH; ++c) {
      const int s = c % S;
      tc::mbar_wait(&full[s], static_cast<uint32_t>((c / S) & 1));
      __syncwarp();
      if (tc::elect_one()) {
        tc::fence_after_sync();
        const uint32_t sa = tc::smem_u32(smem + s * STAGE), sw = sa + A_CH;
        const uint64_t a_hi = tc::make_desc(sa, LBO_A, 128), a_lo = tc::make_desc(sa + w_plane(kWT), LBO_A, 128);
        const uint64_t b_hi = tc::make_desc(sw, LBO_B, 128), b_lo = tc::make_desc(sw + w_plane(N), LBO_B, 128);
        const uint64_t as = (2u * LBO_A) >> 4, bs = (2u * LBO_B) >> 4;
#pragma unroll
        for (int term = 0; term < 3; ++term) {      // lo*hi, hi*lo, hi*hi
          const uint64_t aa = term == 0 ? a_lo : a_hi, bb = term == 1 ? b_lo : b_hi;
#pragma unroll
          for (int k8 = 0; k8 < kWKc / 8; ++k8)
            tc::umma_tf32(tb, aa + k8 * as, bb + k8 * bs, idesc, (c | term | k8) != 0 ? 1u : 0u);
        }
        tc::umma_commit(&empty[s]);
        if (c == NCH - 1) tc::umma_commit(&accum);
      }
      __syncwarp();
    }
  } else {
    // ================================================================== epilogue warps 0..7
    const int r = (warp & 3) * 32 + (tid & 31), half = warp >> 2;
    const int64_t s = tile * kWT + r;
    const bool valid = s < a.n;
    const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
    const uint32_t tl = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    tc::mbar_wait(&accum, 0u);
    __syncwarp();
    tc::fence_after_sync();
    const bool act = a.active && valid;
    if constexpr (EPI == EPI_HIDDEN || EPI == EPI_HEADS) {
      constexpr int ND = EPI == EPI_HIDDEN ? N : N - 16;          // activation columns produced (heads: H/2 + the mean column)
      unsigned char* tp = a.out + static_cast<size_t>(tile) * w_tile_bytes(ND);
#pragma unroll 1
      for (int c0 = half * (ND / 2); c0 < (half + 1) * (ND / 2); c0 += 16) {
        float z[16];
        tc::tmem_ld16(tl + static_cast<uint32_t>(c0), z);
        tc::tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 16; g += 8) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = tanh_pre((z[g + q] + __ldg(a.bias + c0 + g + q)) * kTanhArg);
          if (act) {
            bool k[8];
            wide_keep8(dp, a, sg, s, static_cast<uint32_t>(c0 + g), k);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = k[q] ? v[q] : 0.f;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = valid ? v[q] * a.inact : 0.f;
          }
          wide_store8(tp, c0 + g, r, v);
        }
      }
      if constexpr (EPI == EPI_HEADS) {
        if (half == 0) {          // column ND = the mean head
          float zz[8];
          tc::tmem_ld8(tl + static_cast<uint32_t>(ND), zz);
          tc::tmem_wait_ld();
          if (valid) a.u_io[s] = zz[0] + __ldg(a.bias2);
        }
      }
    } else {
      // EPI_V1: a1 = tanh(z1 + bv1) (N = H/4 columns), v = Wv2 . a1 + bv2, log-variance, statistics
      if (half == 0) {
        float vraw = __ldg(a.bias2);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 16) {
          float z[16];
          tc::tmem_ld16(tl + static_cast<uint32_t>(c0), z);
          tc::tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; ++q)
            vraw = fmaf(__ldg(a.w2 + c0 + q), tanh_pre((z[q] + __ldg(a.bias + c0 + q)) * kTanhArg), vraw);
        }
        if (valid) {
          const float lv = logvar_from_v(vraw), u = a.u_io[s];
          if (a.mode == 0) { a.out_u[s] = u; a.out_s[s] = lv; }
          else if (a.mode == 1) { a.pred_mean[s] = u; }
          else {
            float mean = 0.f, m2 = 0.f, slv = 0.f;
            if (a.pass > 0) { mean = a.st_mean[s]; m2 = a.st_m2[s]; slv = a.st_slv[s]; }
            const float d = u - mean;
            mean += d / static_cast<float>(a.pass + 1);
            m2 = fmaf(d, u - mean, m2);
            slv += lv;
            a.st_mean[s] = mean; a.st_m2[s] = m2; a.st_slv[s] = slv;
            if (a.pass == a.T - 1) {
              const float invT = 1.0f / static_cast<float>(a.T);
              if (a.a_u) a.a_u[s] = sqrtf(expf(slv * invT));
              if (a.e_u) a.e_u[s] = sqrtf(fmaxf(m2, 0.f) * invT);
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, TCOLS);
}

// ------------------------------------------------------------------ host side
static int g_wide_tc_enabled = 1;

template <int H>
struct WidePlan {
  static constexpr int NH = H / 2 + 16;          // heads GEMM N: H/2 variance-head rows + mean row + zero rows
  size_t off_w[PINN_MAX_HIDDEN], off_wh, off_wv1, off_p0, off_p1, off_pv, off_u, off_st, bytes;
};
template <int H>
static WidePlan<H> wide_plan(int L, int64_t n) {
  WidePlan<H> p{};
  const size_t tiles = static_cast<size_t>((n + kWT - 1) / kWT);
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += (b + 255) & ~static_cast<size_t>(255); return r; };
  for (int l = 1; l < L; ++l) p.off_w[l] = take(w_mat_bytes(H, H));
  p.off_wh = take(w_mat_bytes(WidePlan<H>::NH, H));
  p.off_wv1 = take(w_mat_bytes(H / 4, H / 2));
  p.off_p0 = take(tiles * w_tile_bytes(H));
  p.off_p1 = take(tiles * w_tile_bytes(H));
  p.off_pv = take(tiles * w_tile_bytes(H / 2));
  p.off_u = take(static_cast<size_t>(n) * sizeof(float));
  p.off_st = take(static_cast<size_t>(3) * n * sizeof(float));
  p.bytes = o;
  return p;
}
size_t wide_tc_workspace_bytes(int H, int L, int64_t n) {
  if (!g_wide_tc_enabled || n <= 0) return 0;
  if (H == 256) return wide_plan<256>(L, n).bytes;
  if (H == 128) return wide_plan<128>(L, n).bytes;
  return 0;
}

template <int H>
static int run_wide(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
  constexpr int NH = WidePlan<H>::NH;
  const int L = net->n_hidden;
  const WidePlan<H> p = wide_plan<H>(L, n);
  if (!workspace || workspace_bytes < p.bytes) return PINN_E_WORKSPACE;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const int tiles = static_cast<int>((n + kWT - 1) / kWT);
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;
  // ---- weight planes
  auto split = [&](const float* sa, int rows_a, const float* sb, int N, int K, float c, unsigned char* dst) {
    const int items = N * (K / 4);
    wide_split_weights_kernel<<<(items + 255) / 256, 256, 0, st>>>(sa, rows_a, sb, N, K, c, dst);
  };
  for (int l = 1; l < L; ++l) split(net->W[l], H, nullptr, H, H, wscale, ws + p.off_w[l]);
  split(net->Wv0, H / 2, net->Wp, NH, H, wscale, ws + p.off_wh);
  split(net->Wv1, H / 4, nullptr, H / 4, H / 2, wscale, ws + p.off_wv1);
  PINN_CUDA_TRY(cudaGetLastError());
  // ---- kernels and their shared-memory sizes
  auto k_hidden = wide_gemm_kernel<H, H, EPI_HIDDEN>;
  auto k_heads = wide_gemm_kernel<NH, H, EPI_HEADS>;
  auto k_v1 = wide_gemm_kernel<H / 4, H / 2, EPI_V1>;
  const int sm_hidden = (H >= 256 ? 2 : 3) * (w_chunk(kWT) + w_chunk(H));
  const int sm_heads = 3 * (w_chunk(kWT) + w_chunk(NH));
  const int sm_v1 = 3 * (w_chunk(kWT) + w_chunk(H / 4));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_hidden, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_hidden));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_heads));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_v1));

  const bool do_eval = mc && out.pred_mean != nullptr;
  const int n_pass = mc ? T + (do_eval ? 1 : 0) : 1;
  float* st_mean = out.raw_mean ? out.raw_mean : reinterpret_cast<float*>(ws + p.off_st);
  float* st_m2 = out.raw_m2 ? out.raw_m2 : reinterpret_cast<float*>(ws + p.off_st) + n;
  float* st_slv = out.raw_slv ? out.raw_slv : reinterpret_cast<float*>(ws + p.off_st) + 2 * n;
  for (int pi = 0; pi < n_pass; ++pi) {
    const bool eval_pass = mc && do_eval && pi == 0;
    WideArgs a{};
    a.n = n;
    a.mask_row_bytes = L * H + H / 2;
    a.active = drop_on && !eval_pass ? 1 : 0;
    a.pass = mc ? (do_eval ? pi - 1 : pi) : 0;
    if (a.pass < 0) a.pass = 0;
    a.inact = drop_on ? dp.keep : 1.0f;
    a.u_io = reinterpret_cast<float*>(ws + p.off_u);
    a.mode = !mc ? 0 : (eval_pass ? 1 : 2);
    a.T = T;
    a.out_u = out.u; a.out_s = out.s; a.pred_mean = out.pred_mean; a.a_u = out.a_u; a.e_u = out.e_u;
    a.st_mean = st_mean; a.st_m2 = st_m2; a.st_slv = st_slv;
    unsigned char* cur = ws + p.off_p0;
    unsigned char* nxt = ws + p.off_p1;
    // layer 0
    a.layer = 0; a.unit_base = 0; a.out = cur;
    wide_layer0_kernel<H><<<tiles, 256, 0, st>>>(x, net->W[0], net->b[0], dp, a);
    for (int l = 1; l < L; ++l) {
      a.A = cur; a.W = ws + p.off_w[l]; a.out = nxt; a.bias = net->b[l];
      a.layer = static_cast<uint32_t>(l); a.unit_base = static_cast<uint32_t>(l * H);
      k_hidden<<<tiles, 320, sm_hidden, st>>>(dp, a);
      unsigned char* t = cur; cur = nxt; nxt = t;
    }
    a.A = cur; a.W = ws + p.off_wh; a.out = ws + p.off_pv; a.bias = net->bv0; a.bias2 = net->bp;
    a.layer = static_cast<uint32_t>(L); a.unit_base = static_cast<uint32_t>(L * H);
    k_heads<<<tiles, 320, sm_heads, st>>>(dp, a);
    a.A = ws + p.off_pv; a.W = ws + p.off_wv1; a.out = nullptr; a.bias = net->bv1; a.bias2 = net->bv2; a.w2 = net->Wv2;
    a.active = 0;
    k_v1<<<tiles, 320, sm_v1, st>>>(dp, a);
  }
  return static_cast<int>(cudaGetLastError());
}

// 1: handled on the wide tensor-core path; 0: shape not covered; -1: error in *err.
int launch_wide_tc(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
                   void* workspace, size_t workspace_bytes, cudaStream_t st, int* err) {
  *err = 0;
  if (!g_wide_tc_enabled || (net->width != 256 && net->width != 128) || net->n_hidden < 1) return 0;
  if (mc && T <= 0) return 0;
  for (int l = 1; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return 0;
  if (!aligned16(net->Wv0) || !aligned16(net->Wp) || !aligned16(net->Wv1)) return 0;
  const int rc = net->width == 256 ? run_wide<256>(mc, net, x, n, T, dp, out, workspace, workspace_bytes, st)
                                   : run_wide<128>(mc, net, x, n, T, dp, out, workspace, workspace_bytes, st);
  if (rc != 0) { *err = rc; return -1; }
  return 1;
}

}  // namespace pinn

// Test / ablation switch: 0 routes the wide nets through the FFMA kernels.
extern "C" int pinn_set_wide_tensor_core_path(int enable) {
  int prev = pinn::g_wide_tc_enabled;
  pinn::g_wide_tc_enabled = enable ? 1 : 0;
  return prev;
}
