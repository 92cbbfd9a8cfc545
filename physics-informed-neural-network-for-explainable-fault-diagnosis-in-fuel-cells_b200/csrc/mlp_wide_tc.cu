// mlp_wide_tc.cu -- tensor-core forward / MC-dropout / training-step path for the WIDE nets (H = 256: the reference's
// own `Layers = [8,256,256,256,1]` (01:2139) and config 4's 6x256; H = 128 rides along).
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
//
// Training step (run_wide_bwd): the same kernel serves dgrad (A = delta planes, W = transposed weight planes, epilogue
// delta' = D * keep * (1 - a^2)) and the weight gradients (split-K over samples: A = transposed delta blocks, W =
// transposed activations + a row of ones for the biases; per-split partials -> deterministic reduce).
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

PINN_D void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

enum { EPI_HIDDEN = 0, EPI_HEADS = 1, EPI_V1 = 2, EPI_V1_TRAIN = 3, EPI_DV0 = 4, EPI_DZ = 5, EPI_WGRAD = 6 };

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
  // ---- training (backward) extras
  unsigned char* outT;         // transposed (sample-contiguous) copy of what this epilogue produces, or nullptr
  int T_rows;                  // B-side copy: rows per chunk of the transposed buffer (features + 16); 0: A-side copy in 128-row blocks
  int64_t T_chunks;            // total 16-sample chunks (= 8 * tiles): block stride of an A-side copy
  const unsigned char* act;    // EPI_DV0 / EPI_DZ: saved K-major planes of the activation whose (1 - a^2) multiplies the delta
  const float* y; const float* grad_u; const float* grad_s; float inv_n_global;     // EPI_V1_TRAIN: loss gradient source
  int no_logvar;               // DNN(logvar=False): log-variance output identically 0 (PINN_NET_NO_LOGVAR)
  float* du;                   // EPI_V1_TRAIN writes d loss / d u per sample; EPI_DV0 reads it
  float* tail_partial;         // EPI_V1_TRAIN: [tile][N + 1] sums of dv * a1[k] and of dv
  double* loss_partial;        // EPI_V1_TRAIN: [tile][4]
  float* wg_partial;           // EPI_WGRAD: [split][m-block][128][N]
  int nch;                     // K chunks of this launch (per CTA for EPI_WGRAD)
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

// The same 8 values into the TRANSPOSED (sample-contiguous) copy the weight-gradient GEMM contracts over:
// element (feature f, sample) of 16-sample chunk `chunk`, position sl = sample % 16 inside it:
//     byte = (sl / 4) * 16 * ROWS + f * 16 + (sl % 4) * 4      (hi plane; lo plane follows)
// B-side copies (activations) keep all their rows in one block per chunk (ROWS = features + 16: a row of ones and
// zero padding follow the data); A-side copies (deltas) are cut into 128-row blocks, block stride = T_chunks chunks.
// The whole warp calls this together (32 consecutive samples x 8 features): the values are transposed through a
// per-warp shared-memory patch (2 planes x 8 rows x 36 floats, conflict-free both ways) so that every lane ends up
// with two (feature, 4 consecutive samples) units per plane and the global stores are 16 bytes wide, eight features
// (128 contiguous bytes) per quarter-warp -- the scattered 4-byte version halved the forward GEMM's speed.
constexpr int kTPatch = 2 * 8 * 36;          // floats per warp
PINN_D void wide_storeT8(const WideArgs& a, int64_t tile, int r, int c0, const float (&v)[8], float* patch) {
  const int lane = r & 31;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float h = tc::tf32_hi_fast(v[q]);
    patch[q * 36 + lane] = h;
    patch[8 * 36 + q * 36 + lane] = v[q] - h;
  }
  __syncwarp();
#pragma unroll
  for (int u2 = 0; u2 < 2; ++u2) {
    const int u = lane + 32 * u2, q = u & 7, quad = u >> 3;            // unit: feature c0 + q, samples 4 quad .. 4 quad + 3 of the warp
    const float4 h4 = *reinterpret_cast<const float4*>(patch + q * 36 + 4 * quad);
    const float4 l4 = *reinterpret_cast<const float4*>(patch + 8 * 36 + q * 36 + 4 * quad);
    const int rs = (r & ~31) + 4 * quad;                               // first sample (row of the tile) of the unit
    const int64_t chunk = tile * (kWT / kWKc) + (rs >> 4);
    const int k4 = (rs & 15) >> 2, f = c0 + q;
    unsigned char* p;
    int plane;
    if (a.T_rows > 0) {
      plane = 4 * a.T_rows * 16;
      p = a.outT + static_cast<size_t>(chunk) * (2 * plane) + k4 * (a.T_rows * 16) + f * 16;
    } else {
      plane = w_plane(kWT);
      p = a.outT + (static_cast<size_t>(f >> 7) * a.T_chunks + chunk) * w_chunk(kWT) + k4 * (kWT * 16) + (f & 127) * 16;
    }
    *reinterpret_cast<float4*>(p) = h4;
    *reinterpret_cast<float4*>(p + plane) = l4;
  }
  __syncwarp();
}
// a = hi + lo of 8 saved activations (K-major planes of a tile with 128 rows)
PINN_D void wide_load8(const unsigned char* tile_planes, int c0, int r, float (&v)[8]) {
  const unsigned char* p = tile_planes + static_cast<size_t>(c0 >> 4) * w_chunk(kWT) + ((c0 & 15) >> 2) * (kWT * 16) + r * 16;
  const float4 h0 = *reinterpret_cast<const float4*>(p), h1 = *reinterpret_cast<const float4*>(p + kWT * 16);
  const float4 l0 = *reinterpret_cast<const float4*>(p + w_plane(kWT)), l1 = *reinterpret_cast<const float4*>(p + w_plane(kWT) + kWT * 16);
  v[0] = h0.x + l0.x; v[1] = h0.y + l0.y; v[2] = h0.z + l0.z; v[3] = h0.w + l0.w;
  v[4] = h1.x + l1.x; v[5] = h1.y + l1.y; v[6] = h1.z + l1.z; v[7] = h1.w + l1.w;
}
// constant rows of a B-side transposed buffer: row `data_rows` = 1 (the bias column of the wgrad product), the rest 0
__global__ void wide_fill_ones_kernel(unsigned char* buf, int rows, int data_rows, int64_t chunks) {
  griddep_launch();
  griddep_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per = 4 * (rows - data_rows);
  if (i >= chunks * per) return;
  const int64_t chunk = i / per;
  const int k4 = static_cast<int>(i % per) / (rows - data_rows), rr = data_rows + static_cast<int>(i % per) % (rows - data_rows);
  const int plane = 4 * rows * 16;
  unsigned char* p = buf + static_cast<size_t>(chunk) * (2 * plane) + k4 * (rows * 16) + rr * 16;
  const float v = rr == data_rows ? 1.0f : 0.0f;
  *reinterpret_cast<float4*>(p) = make_float4(v, v, v, v);
  *reinterpret_cast<float4*>(p + plane) = make_float4(0.f, 0.f, 0.f, 0.f);
}
// transposed weight planes for dgrad: B[n = k][K index = j] = W[j][k] (j < rows_a), = w_b[k] (j == rows_a), 0 beyond; times c
__global__ void wide_split_weights_T_kernel(const float* __restrict__ src_a, int rows_a, const float* __restrict__ src_b, int N, int K,
                                            float c, unsigned char* __restrict__ dst) {
  griddep_launch();
  griddep_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // (n, j4): output row n, 4 consecutive K indices
  if (idx >= N * (K / 4)) return;
  const int nrow = idx % N, j4 = idx / N;
  float vv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = 4 * j4 + q;
    float w = 0.f;
    if (j < rows_a) w = __ldg(src_a + static_cast<size_t>(j) * N + nrow);
    else if (j == rows_a && src_b != nullptr) w = __ldg(src_b + nrow);
    vv[q] = w * c;
  }
  const float4 h = make_float4(tc::tf32_hi(vv[0]), tc::tf32_hi(vv[1]), tc::tf32_hi(vv[2]), tc::tf32_hi(vv[3]));
  unsigned char* p = dst + static_cast<size_t>(j4 >> 2) * w_chunk(N) + (j4 & 3) * (N * 16) + nrow * 16;
  *reinterpret_cast<float4*>(p) = h;
  *reinterpret_cast<float4*>(p + w_plane(N)) = make_float4(vv[0] - h.x, vv[1] - h.y, vv[2] - h.z, vv[3] - h.w);
}

// ------------------------------------------------------------------ weight planes (once per call)
// dst planes of an [N x K] matrix whose first `rows_a` rows come from `src_a` ([rows_a][K]), row `rows_a` from
// `src_b` (or zero) and the rest are zero; every element multiplied by `c`.
__global__ void wide_split_weights_kernel(const float* __restrict__ src_a, int rows_a, const float* __restrict__ src_b, int N, int K,
                                          float c, unsigned char* __restrict__ dst) {
  griddep_launch();
  griddep_wait();
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
  griddep_launch();
  griddep_wait();
  __shared__ float sW[H * PINN_N_IN];
  __shared__ float sb[H];
  __shared__ __align__(16) float tpatch[8][kTPatch];
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
    if (a.outT != nullptr) wide_storeT8(a, tile, r, c0, v, tpatch[threadIdx.x >> 5]);
  }
  if (a.act != nullptr && half == 0) {      // training: x^T (8 rows) as the B operand of dW0; `act` carries the buffer here
    unsigned char* xt = const_cast<unsigned char*>(a.act);
    const int64_t chunk = tile * (kWT / kWKc) + (r >> 4);
    const int sl = r & 15, plane = 4 * 16 * 16;
#pragma unroll
    for (int f = 0; f < PINN_N_IN; ++f) {
      const float h = tc::tf32_hi_fast(xr[f]);
      unsigned char* p = xt + static_cast<size_t>(chunk) * (2 * plane) + (sl >> 2) * (16 * 16) + f * 16 + (sl & 3) * 4;
      *reinterpret_cast<float*>(p) = h;
      *reinterpret_cast<float*>(p + plane) = xr[f] - h;
    }
  }
}

// ------------------------------------------------------------------ one layer = one GEMM launch
// Forward / dgrad: grid.x = 128-row tiles, A = the tile's planes (a.nch chunks), W = weight planes.
// Weight gradient (EPI_WGRAD): grid = (sample splits, 128-row feature blocks of the delta); both operands stream along
// the sample axis: A = transposed deltas (block blockIdx.y), W = transposed activations, chunks [x * nch, (x+1) * nch).
template <int N, int EPI>
__global__ void __launch_bounds__(320, N > 256 ? 1 : 2)
wide_gemm_kernel(const __grid_constant__ DropParams dp, WideArgs a) {
  griddep_launch();
  griddep_wait();
  // A pipeline stage is HALF a plane chunk (8 of its 16 k: the first or second pair of 4-wide sub-chunks of the hi and
  // of the lo plane): four bulk copies, three MMAs.  Twice as many, half as large stages as chunk-sized ones keep more
  // loads in flight per CTA in the same shared memory (TMA latency ~2000 clk vs 384 clk of MMA work per stage).
  constexpr int S = N >= 256 ? 4 : 6;                            // pipeline stages
  constexpr int A_CH = w_chunk(kWT), W_CH = w_chunk(N);          // bytes of a whole [hi | lo] chunk in global memory
  constexpr int A_H = w_plane(kWT) / 2, W_H = w_plane(N) / 2;    // bytes of half a plane
  constexpr int STAGE = 2 * A_H + 2 * W_H;                       // [A hi half | A lo half | W hi half | W lo half]
  constexpr uint32_t TCOLS = N > 256 ? 512u : (N > 128 ? 256u : (N > 64 ? 128u : 64u));
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[S], empty[S], accum;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red_s[4][N <= 64 ? N + 1 : 1];               // EPI_V1_TRAIN: per-warp column sums
  __shared__ double red_d[4][4];
  const int tid = threadIdx.x, warp = tc::uniform_warp_idx();
  const int64_t tile = blockIdx.x;
  int64_t c_begin = 0, nch = a.nch;
  if constexpr (EPI == EPI_WGRAD) {
    c_begin = static_cast<int64_t>(blockIdx.x) * a.nch;
    nch = a.T_chunks - c_begin < a.nch ? a.T_chunks - c_begin : a.nch;
    if (nch < 0) nch = 0;
  }

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
      const unsigned char* At;
      const unsigned char* Wt;
      if constexpr (EPI == EPI_WGRAD) {
        At = a.A + (static_cast<size_t>(blockIdx.y) * a.T_chunks + c_begin) * A_CH;
        Wt = a.W + static_cast<size_t>(c_begin) * W_CH;
      } else {
        At = a.A + static_cast<size_t>(tile) * a.nch * A_CH;
        Wt = a.W;
      }
      for (int64_t c = 0; c < 2 * nch; ++c) {            // c = 2 * chunk + half
        const int s = static_cast<int>(c % S);
        if (c >= S) tc::mbar_wait(&empty[s], static_cast<uint32_t>((c / S - 1) & 1));
        tc::mbar_expect_tx(&full[s], STAGE);
        unsigned char* d = smem + s * STAGE;
        const unsigned char* ga = At + static_cast<size_t>(c >> 1) * A_CH + (c & 1) * A_H;
        const unsigned char* gw = Wt + static_cast<size_t>(c >> 1) * W_CH + (c & 1) * W_H;
        tc::bulk_g2s(d, ga, A_H, &full[s]);
        tc::bulk_g2s(d + A_H, ga + w_plane(kWT), A_H, &full[s]);
        tc::bulk_g2s(d + 2 * A_H, gw, W_H, &full[s]);
        tc::bulk_g2s(d + 2 * A_H + W_H, gw + w_plane(N), W_H, &full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ================================================================== MMA issuer
    constexpr int N1 = N > 256 ? 256 : N, N2 = N - N1;            // one tcgen05.mma covers at most 256 columns
    const uint32_t idesc = tc::make_idesc_tf32(kWT, N1), idesc2 = tc::make_idesc_tf32(kWT, N2 > 0 ? N2 : 16);
    constexpr uint32_t LBO_A = kWT * 16, LBO_B = N * 16;
    for (int64_t c = 0; c < 2 * nch; ++c) {
      const int s = static_cast<int>(c % S);
      tc::mbar_wait(&full[s], static_cast<uint32_t>((c / S) & 1));
      __syncwarp();
      if (tc::elect_one()) {
        tc::fence_after_sync();
        const uint32_t sa = tc::smem_u32(smem + s * STAGE), sw = sa + 2 * A_H;
        const uint64_t a_hi = tc::make_desc(sa, LBO_A, 128), a_lo = tc::make_desc(sa + A_H, LBO_A, 128);
        const uint64_t b_hi = tc::make_desc(sw, LBO_B, 128), b_lo = tc::make_desc(sw + W_H, LBO_B, 128);
#pragma unroll
        for (int term = 0; term < 3; ++term) {      // lo*hi, hi*lo, hi*hi
          const uint64_t aa = term == 0 ? a_lo : a_hi, bb = term == 1 ? b_lo : b_hi;
          const uint32_t acc = (c != 0 || term != 0) ? 1u : 0u;
          tc::umma_tf32(tb, aa, bb, idesc, acc);
          if constexpr (N2 > 0) tc::umma_tf32(tb + N1, aa, bb + ((N1 * 16u) >> 4), idesc2, acc);
        }
        tc::umma_commit(&empty[s]);
        if (c == 2 * nch - 1) tc::umma_commit(&accum);
      }
      __syncwarp();
    }
  } else {
    // ================================================================== epilogue warps 0..7
    const int r = (warp & 3) * 32 + (tid & 31), half = warp >> 2;
    const int64_t s = tile * kWT + r;
    const bool valid = EPI == EPI_WGRAD ? true : s < a.n;
    const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
    const uint32_t tl = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    if (nch > 0) tc::mbar_wait(&accum, 0u);
    __syncwarp();
    tc::fence_after_sync();
    const bool act = a.active && valid;
    float* patch = reinterpret_cast<float*>(smem) + warp * kTPatch;      // the operand ring is idle once `accum` has fired
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
          if (a.outT != nullptr) wide_storeT8(a, tile, r, c0 + g, v, patch);
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
    } else if constexpr (EPI == EPI_V1) {
      // a1 = tanh(z1 + bv1) (N = H/4 columns), v = Wv2 . a1 + bv2, log-variance, statistics
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
          const float lv = logvar_out(vraw, a.no_logvar != 0), u = a.u_io[s];
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
    } else if constexpr (EPI == EPI_V1_TRAIN) {
      // the variance head's last layer forward AND backward: v, log-variance, loss gradient (fused aleatoric loss or the
      // caller's grad_u / grad_logvar), d z1 = dv * Wv2 * (1 - a1^2) as K-major planes (next dgrad GEMM) and transposed
      // (dWv1), per-tile sums of dv * a1 (dWv2), dv (dbv2) and of the loss terms
      if (half == 0) {
        const int lane = tid & 31;
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
        const float lv = logvar_out(vraw, a.no_logvar != 0);
        float du = 0.f, ds = 0.f;
        double l4[4] = {0.0, 0.0, 0.0, 0.0};
        if (valid) {
          const float u = a.u_io[s];
          if (a.grad_u != nullptr) { du = __ldg(a.grad_u + s); ds = a.grad_s ? __ldg(a.grad_s + s) : 0.f; }
          else {
            const float yv = __ldg(a.y + s), e = expf(-lv), diff = yv - u;
            du = -e * diff * a.inv_n_global;
            const float sgn = lv > 0.f ? 1.f : (lv < 0.f ? -1.f : 0.f);
            ds = (-0.5f * e * diff * diff + 0.5f + 0.01f * sgn) * a.inv_n_global;
            l4[0] = static_cast<double>(0.5f * e * diff * diff + 0.5f * lv);
            l4[1] = static_cast<double>(fabsf(lv));
            l4[2] = static_cast<double>(diff * diff);
            l4[3] = 1.0;
          }
        }
        const float dv = a.no_logvar ? 0.f : ds * dlogvar_dv(vraw);
        a.du[s < a.n ? s : a.n] = du;               // slot n = scratch for the tail rows of the last tile
        unsigned char* tp = a.out + static_cast<size_t>(tile) * w_tile_bytes(N);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 8) {
          float z[8], dz[8], col[8];
          tc::tmem_ld8(tl + static_cast<uint32_t>(c0), z);
          tc::tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float a1 = tanh_pre((z[q] + __ldg(a.bias + c0 + q)) * kTanhArg);
            dz[q] = dv * __ldg(a.w2 + c0 + q) * fmaf(-a1, a1, 1.0f);
            col[q] = dv * a1;
          }
          wide_store8(tp, c0, r, dz);
          wide_storeT8(a, tile, r, c0, dz, patch);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float t = col[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) red_s[warp][c0 + q] = t;
          }
        }
        {
          float t = dv;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) red_s[warp][N] = t;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            double d = l4[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            if (lane == 0) red_d[warp][k] = d;
          }
        }
        bar_sync_named(1, 128);
        if (tid <= N) a.tail_partial[static_cast<size_t>(tile) * (N + 1) + tid] = (red_s[0][tid] + red_s[1][tid]) + (red_s[2][tid] + red_s[3][tid]);
        if (tid < 4) a.loss_partial[static_cast<size_t>(tile) * 4 + tid] = (red_d[0][tid] + red_d[1][tid]) + (red_d[2][tid] + red_d[3][tid]);
      }
    } else if constexpr (EPI == EPI_DV0 || EPI == EPI_DZ) {
      // delta = D * keep * (1 - a^2): K-major planes for the next dgrad GEMM + transposed 128-row blocks for the wgrad GEMM.
      // EPI_DV0 (variance-head layer 0) also appends d loss / d u as column N of its K-major output / row 0 of block 1.
      constexpr int KOUT = EPI == EPI_DV0 ? N + 16 : N;
      const unsigned char* ap = a.act + static_cast<size_t>(tile) * w_tile_bytes(N);
      unsigned char* tp = a.out + static_cast<size_t>(tile) * w_tile_bytes(KOUT);
#pragma unroll 1
      for (int c0 = half * (N / 2); c0 < (half + 1) * (N / 2); c0 += 16) {
        float z[16];
        tc::tmem_ld16(tl + static_cast<uint32_t>(c0), z);
        tc::tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 16; g += 8) {
          float av[8], v[8];
          wide_load8(ap, c0 + g, r, av);
          bool k[8] = {true, true, true, true, true, true, true, true};
          if (act) wide_keep8(dp, a, sg, s, static_cast<uint32_t>(c0 + g), k);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = (k[q] && valid) ? z[g + q] * fmaf(-av[q], av[q], 1.0f) : 0.f;
          wide_store8(tp, c0 + g, r, v);
          wide_storeT8(a, tile, r, c0 + g, v, patch);
        }
      }
      if constexpr (EPI == EPI_DV0) {
        if (half == 0) {
          const float du = valid ? a.du[s] : 0.f;
          const float d8[8] = {du, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          wide_store8(tp, N, r, d8);
          wide_store8(tp, N + 8, r, z8);
          wide_storeT8(a, tile, r, N, d8, patch);        // delta rows N .. N+7: row N = du
        }
      }
    } else {
      // EPI_WGRAD: accumulator rows (features of the delta block) x N columns -> this CTA's partial
      float* dst = a.wg_partial + ((static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * kWT + r) * N;
#pragma unroll 1
      for (int c0 = half * 16; c0 < N; c0 += 32) {
        float z[16];
        if (nch > 0) { tc::tmem_ld16(tl + static_cast<uint32_t>(c0), z); tc::tmem_wait_ld(); }
        else {
#pragma unroll
          for (int q = 0; q < 16; ++q) z[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4*>(dst + c0 + q) = make_float4(z[q], z[q + 1], z[q + 2], z[q + 3]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, TCOLS);
}

// dst[row0 + j][k] (leading dimension ld) = scale * sum over splits of partial[split][mb][j][k], j < nrows, k < ncols;
// optionally dstB[j] = sum of column bias_col.  Fixed summation order: deterministic.
__global__ void wide_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int nmb, int N, int mb, int row0, int nrows,
                                         int ncols, int bias_col, float* __restrict__ dstW, int ld, float* __restrict__ dstB, float scale) {
  griddep_launch();
  griddep_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = ncols + (dstB != nullptr ? 1 : 0);
  if (idx >= nrows * per) return;
  const int j = idx / per, kk = idx % per;
  const int col = kk < ncols ? kk : bias_col;
  const float* src = partial + (static_cast<size_t>(mb) * kWT + row0 + j) * N + col;
  const size_t stride = static_cast<size_t>(nmb) * kWT * N;
  double acc = 0.0;
  int sp = 0;
  for (; sp + 8 <= splits; sp += 8) {          // eight loads in flight; the summation order stays fixed
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(src + (sp + q) * stride);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += static_cast<double>(v[q]);
  }
  for (; sp < splits; ++sp) acc += static_cast<double>(__ldg(src + sp * stride));
  if (kk < ncols) dstW[static_cast<size_t>(j) * ld + kk] = static_cast<float>(acc) * scale;
  else dstB[j] = static_cast<float>(acc);
}
// dWv2 / dbv2 and the loss sums from the per-tile partials of EPI_V1_TRAIN: block k < n1 -> dWv2[k], block n1 -> dbv2,
// blocks n1+1 .. n1+4 -> the four loss sums.  Strided per-thread sums, then a fixed-order shared-memory tree: deterministic.
__global__ void __launch_bounds__(256)
wide_tail_reduce_kernel(const float* __restrict__ tail_partial, const double* __restrict__ loss_partial, int tiles, int n1,
                        float* __restrict__ dWv2, float* __restrict__ dbv2, double* __restrict__ loss) {
  griddep_launch();
  griddep_wait();
  __shared__ double sh[256];
  const int k = blockIdx.x, t0 = threadIdx.x;
  double acc = 0.0;
  if (k <= n1) {
    for (int t = t0; t < tiles; t += 256) acc += static_cast<double>(tail_partial[static_cast<size_t>(t) * (n1 + 1) + k]);
  } else {
    if (loss == nullptr) return;
    for (int t = t0; t < tiles; t += 256) acc += loss_partial[static_cast<size_t>(t) * 4 + (k - n1 - 1)];
  }
  sh[t0] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t0 < o) sh[t0] += sh[t0 + o];
    __syncthreads();
  }
  if (t0 == 0) {
    if (k < n1) dWv2[k] = static_cast<float>(sh[0]);
    else if (k == n1) dbv2[0] = static_cast<float>(sh[0]);
    else loss[k - n1 - 1] = sh[0];
  }
}

// ------------------------------------------------------------------ host side
// Programmatic dependent launch for the per-layer launches of a small batch (every kernel of this file starts with
// griddep_launch + griddep_wait: stream order is kept, only the launch latency and CTA scheduling of the successor overlap
// the predecessor's tail).  Same size rule and per-call flags as the 64-wide training step (pinn_net_t.flags).
static bool wide_pdl(const pinn_net_t* net, int64_t n) {
  const int mode = dependent_launch_mode(net);
  return mode == 2 || (mode == 1 && (n + kWT - 1) / kWT <= static_cast<int64_t>(2) * sm_count());
}

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
size_t wide_tc_workspace_bytes(int H, int L, int64_t n, int flags) {
  if (n <= 0) return 0;
  if (flags >= 0 && (flags & PINN_NET_NO_WIDE_TC)) return 0;          // FFMA path: its own scratch only
  if (H == 256) {           // resident-activation kernel (mlp_wide_res.cu) unless the call opts out; flags < 0: either path
    const size_t a = wide_plan<256>(L, n).bytes, b = wide_res_workspace_bytes(L, n);
    if (flags >= 0) return (L >= 2 && !(flags & PINN_NET_NO_WIDE_RESIDENT)) ? b : a;
    return a > b ? a : b;
  }
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
  const bool pdl = wide_pdl(net, n);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const int tiles = static_cast<int>((n + kWT - 1) / kWT);
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;
  // ---- weight planes
  auto split = [&](const float* sa, int rows_a, const float* sb, int N, int K, float c, unsigned char* dst) {
    const int items = N * (K / 4);
    launch_pdl(wide_split_weights_kernel, dim3((items + 255) / 256), dim3(256), 0, st, pdl, sa, rows_a, sb, N, K, c, dst);
  };
  for (int l = 1; l < L; ++l) split(net->W[l], H, nullptr, H, H, wscale, ws + p.off_w[l]);
  split(net->Wv0, H / 2, net->Wp, NH, H, wscale, ws + p.off_wh);
  split(net->Wv1, H / 4, nullptr, H / 4, H / 2, wscale, ws + p.off_wv1);
  PINN_CUDA_TRY(cudaGetLastError());
  // ---- kernels and their shared-memory sizes
  auto k_hidden = wide_gemm_kernel<H, EPI_HIDDEN>;
  auto k_heads = wide_gemm_kernel<NH, EPI_HEADS>;
  auto k_v1 = wide_gemm_kernel<H / 4, EPI_V1>;
  auto smem_of = [](int N) { return (N >= 256 ? 4 : 6) * (w_plane(kWT) + w_plane(N)); };
  const int sm_hidden = smem_of(H), sm_heads = smem_of(NH), sm_v1 = smem_of(H / 4);
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
    a.no_logvar = (net->flags & PINN_NET_NO_LOGVAR) ? 1 : 0;
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
    launch_pdl(wide_layer0_kernel<H>, dim3(tiles), dim3(256), 0, st, pdl, x, net->W[0], net->b[0], dp, a);
    for (int l = 1; l < L; ++l) {
      a.A = cur; a.W = ws + p.off_w[l]; a.out = nxt; a.bias = net->b[l];
      a.layer = static_cast<uint32_t>(l); a.unit_base = static_cast<uint32_t>(l * H); a.nch = H / kWKc;
      launch_pdl(k_hidden, dim3(tiles), dim3(320), sm_hidden, st, pdl, dp, a);
      unsigned char* t = cur; cur = nxt; nxt = t;
    }
    a.A = cur; a.W = ws + p.off_wh; a.out = ws + p.off_pv; a.bias = net->bv0; a.bias2 = net->bp;
    a.layer = static_cast<uint32_t>(L); a.unit_base = static_cast<uint32_t>(L * H); a.nch = H / kWKc;
    launch_pdl(k_heads, dim3(tiles), dim3(320), sm_heads, st, pdl, dp, a);
    a.A = ws + p.off_pv; a.W = ws + p.off_wv1; a.out = nullptr; a.bias = net->bv1; a.bias2 = net->bv2; a.w2 = net->Wv2;
    a.active = 0; a.nch = (H / 2) / kWKc;
    launch_pdl(k_v1, dim3(tiles), dim3(320), sm_v1, st, pdl, dp, a);
  }
  return static_cast<int>(cudaGetLastError());
}

// 1: handled on the wide tensor-core path; 0: shape not covered; -1: error in *err.
int launch_wide_tc(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
                   void* workspace, size_t workspace_bytes, cudaStream_t st, int* err) {
  *err = 0;
  if ((net->flags & PINN_NET_NO_WIDE_TC) || (net->width != 256 && net->width != 128) || net->n_hidden < 1) return 0;
  if (mc && T <= 0) return 0;
  // 256-wide nets: the resident-activation kernel (mlp_wide_res.cu) unless the call opts out or the shape is not covered
  if (const int rr = launch_wide_res(mc, net, x, n, T, dp, out, workspace, workspace_bytes, st, err); rr != 0) return rr;
  for (int l = 1; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return 0;
  if (!aligned16(net->Wv0) || !aligned16(net->Wp) || !aligned16(net->Wv1)) return 0;
  const int rc = net->width == 256 ? run_wide<256>(mc, net, x, n, T, dp, out, workspace, workspace_bytes, st)
                                   : run_wide<128>(mc, net, x, n, T, dp, out, workspace, workspace_bytes, st);
  if (rc != 0) { *err = rc; return -1; }
  return 1;
}

// ------------------------------------------------------------------ training step (K2 for the wide nets)
// forward with every layer's planes kept (K-major for the dgrad epilogues, transposed for the weight gradients) ->
// variance-head tail + loss gradient -> dgrad GEMMs back through the heads and the trunk, each followed by the
// weight-gradient GEMM of the layer it just produced the deltas of.

template <int H>
struct WideBwdPlan {
  static constexpr int NH = H / 2 + 16, RB = H + 16, RV = H / 2 + 16;
  size_t off_wf[PINN_MAX_HIDDEN], off_wfh, off_wfv1, off_wt[PINN_MAX_HIDDEN], off_wth, off_wtv1;
  size_t off_pa[PINN_MAX_HIDDEN], off_pv0, off_at[PINN_MAX_HIDDEN], off_v0t, off_xt;
  size_t off_pd[2], off_pdv0h, off_pdz1, off_dt, off_u, off_du, off_tail, off_loss, off_wg, bytes;
  int splits;
};
template <int H>
static WideBwdPlan<H> wide_bwd_plan(int L, int64_t n) {
  using P = WideBwdPlan<H>;
  P p{};
  const size_t tiles = static_cast<size_t>((n + kWT - 1) / kWT), chunks = tiles * (kWT / kWKc);
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += (b + 255) & ~static_cast<size_t>(255); return r; };
  for (int l = 1; l < L; ++l) { p.off_wf[l] = take(w_mat_bytes(H, H)); p.off_wt[l] = take(w_mat_bytes(H, H)); }
  p.off_wfh = take(w_mat_bytes(P::NH, H));       p.off_wth = take(w_mat_bytes(H, P::NH));
  p.off_wfv1 = take(w_mat_bytes(H / 4, H / 2));  p.off_wtv1 = take(w_mat_bytes(H / 2, H / 4));
  for (int l = 0; l < L; ++l) { p.off_pa[l] = take(tiles * w_tile_bytes(H)); p.off_at[l] = take(chunks * w_chunk(P::RB)); }
  p.off_pv0 = take(tiles * w_tile_bytes(H / 2));
  p.off_v0t = take(chunks * w_chunk(P::RV));
  p.off_xt = take(chunks * w_chunk(16));
  p.off_pd[0] = take(tiles * w_tile_bytes(H));
  p.off_pd[1] = take(tiles * w_tile_bytes(H));
  p.off_pdv0h = take(tiles * w_tile_bytes(P::NH));
  p.off_pdz1 = take(tiles * w_tile_bytes(H / 4));
  p.off_dt = take(static_cast<size_t>(2) * chunks * w_chunk(kWT));
  p.off_u = take((static_cast<size_t>(n) + 1) * sizeof(float));
  p.off_du = take((static_cast<size_t>(n) + 1) * sizeof(float));
  p.off_tail = take(tiles * (H / 4 + 1) * sizeof(float));
  p.off_loss = take(tiles * 4 * sizeof(double));
  const int sms = sm_count();
  // the big products have two 128-row delta blocks and one CTA per SM (512 TMEM columns): splits x 2 = one wave
  const int64_t want = sms / 2 > 0 ? sms / 2 : 1;
  int64_t sp = static_cast<int64_t>(chunks) < want ? static_cast<int64_t>(chunks) : want;
  p.splits = static_cast<int>(sp > 0 ? sp : 1);
  p.off_wg = take(static_cast<size_t>(p.splits) * 2 * kWT * P::RB * sizeof(float));
  p.bytes = o;
  return p;
}
size_t wide_tc_bwd_workspace_bytes(int H, int L, int64_t n) {
  if (n <= 0) return 0;
  if (H == 256) return wide_bwd_plan<256>(L, n).bytes;
  if (H == 128) return wide_bwd_plan<128>(L, n).bytes;
  return 0;
}
bool wide_tc_bwd_covers(const pinn_net_t* net) {
  if ((net->flags & PINN_NET_NO_WIDE_TC) || (net->width != 256 && net->width != 128) || net->n_hidden < 1) return false;
  for (int l = 0; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return false;
  return aligned16(net->Wv0) && aligned16(net->Wp) && aligned16(net->Wv1);
}

template <int H>
static int run_wide_bwd(const pinn_net_t* net, const float* x, int64_t n, const DropParams& dp, const float* grad_u, const float* grad_s,
                        const float* y, int64_t n_global, float* grad_flat, double* loss_sums, void* workspace, size_t workspace_bytes,
                        cudaStream_t st) {
  using P = WideBwdPlan<H>;
  constexpr int NH = P::NH, RB = P::RB, RV = P::RV;
  const int L = net->n_hidden;
  const P p = wide_bwd_plan<H>(L, n);
  if (!workspace || workspace_bytes < p.bytes) return PINN_E_WORKSPACE;
  const bool pdl = wide_pdl(net, n);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const ParamLayout lay = make_layout(H, L);
  const int tiles = static_cast<int>((n + kWT - 1) / kWT);
  const int64_t chunks = static_cast<int64_t>(tiles) * (kWT / kWKc);
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;
  // ---- weight planes: forward (rows) and dgrad (transposed), dropout scale folded into both
  auto split = [&](const float* sa, int rows_a, const float* sb, int N, int K, unsigned char* dst) {
    const int items = N * (K / 4);
    launch_pdl(wide_split_weights_kernel, dim3((items + 255) / 256), dim3(256), 0, st, pdl, sa, rows_a, sb, N, K, wscale, dst);
  };
  auto splitT = [&](const float* sa, int rows_a, const float* sb, int N, int K, unsigned char* dst) {
    const int items = N * (K / 4);
    launch_pdl(wide_split_weights_T_kernel, dim3((items + 255) / 256), dim3(256), 0, st, pdl, sa, rows_a, sb, N, K, wscale, dst);
  };
  for (int l = 1; l < L; ++l) { split(net->W[l], H, nullptr, H, H, ws + p.off_wf[l]); splitT(net->W[l], H, nullptr, H, H, ws + p.off_wt[l]); }
  split(net->Wv0, H / 2, net->Wp, NH, H, ws + p.off_wfh);
  splitT(net->Wv0, H / 2, net->Wp, H, NH, ws + p.off_wth);           // B[n = k][j]: j < H/2 -> Wv0[j][k], j = H/2 -> Wp[k]
  split(net->Wv1, H / 4, nullptr, H / 4, H / 2, ws + p.off_wfv1);
  splitT(net->Wv1, H / 4, nullptr, H / 2, H / 4, ws + p.off_wtv1);   // B[n = i][kk] = Wv1[kk][i]
  // ---- constant rows (ones + zero padding) of the B-side transposed buffers
  auto ones = [&](unsigned char* buf, int rows, int data_rows) {
    const int64_t items = chunks * 4 * (rows - data_rows);
    launch_pdl(wide_fill_ones_kernel, dim3(static_cast<unsigned>((items + 255) / 256)), dim3(256), 0, st, pdl, buf, rows, data_rows, chunks);
  };
  for (int l = 0; l < L; ++l) ones(ws + p.off_at[l], RB, H);
  ones(ws + p.off_v0t, RV, H / 2);
  ones(ws + p.off_xt, 16, PINN_N_IN);
  PINN_CUDA_TRY(cudaGetLastError());

  auto k_hidden = wide_gemm_kernel<H, EPI_HIDDEN>;
  auto k_heads = wide_gemm_kernel<NH, EPI_HEADS>;
  auto k_v1t = wide_gemm_kernel<H / 4, EPI_V1_TRAIN>;
  auto k_dv0 = wide_gemm_kernel<H / 2, EPI_DV0>;
  auto k_dz = wide_gemm_kernel<H, EPI_DZ>;
  auto k_wg_b = wide_gemm_kernel<RB, EPI_WGRAD>;
  auto k_wg_v = wide_gemm_kernel<RV, EPI_WGRAD>;
  auto k_wg_x = wide_gemm_kernel<16, EPI_WGRAD>;
  auto smem_of = [](int N) { return (N >= 256 ? 4 : 6) * (w_plane(kWT) + w_plane(N)); };
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_hidden, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(H)));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(NH)));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_v1t, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(H / 4)));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_dv0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(H / 2)));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_dz, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(H)));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_wg_b, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(RB)));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_wg_v, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(RV)));
  PINN_CUDA_TRY(cudaFuncSetAttribute(k_wg_x, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(16)));

  WideArgs a{};
  a.n = n;
  a.mask_row_bytes = L * H + H / 2;
  a.no_logvar = (net->flags & PINN_NET_NO_LOGVAR) ? 1 : 0;
  a.active = drop_on ? 1 : 0;
  a.pass = 0;
  a.inact = 1.0f;                      // training forward with p = 0: nothing folded, nothing to undo
  a.u_io = reinterpret_cast<float*>(ws + p.off_u);
  a.du = reinterpret_cast<float*>(ws + p.off_du);
  a.T_chunks = chunks;
  a.y = y; a.grad_u = grad_u; a.grad_s = grad_s;
  a.inv_n_global = grad_u ? 0.f : static_cast<float>(1.0 / static_cast<double>(n_global));
  a.tail_partial = reinterpret_cast<float*>(ws + p.off_tail);
  a.loss_partial = reinterpret_cast<double*>(ws + p.off_loss);
  a.wg_partial = reinterpret_cast<float*>(ws + p.off_wg);
  // ============================ forward ============================
  a.layer = 0; a.unit_base = 0; a.out = ws + p.off_pa[0];
  a.outT = ws + p.off_at[0]; a.T_rows = RB; a.act = ws + p.off_xt;
  launch_pdl(wide_layer0_kernel<H>, dim3(tiles), dim3(256), 0, st, pdl, x, net->W[0], net->b[0], dp, a);
  a.act = nullptr;
  for (int l = 1; l < L; ++l) {
    a.A = ws + p.off_pa[l - 1]; a.W = ws + p.off_wf[l]; a.out = ws + p.off_pa[l]; a.bias = net->b[l];
    a.outT = ws + p.off_at[l]; a.T_rows = RB;
    a.layer = static_cast<uint32_t>(l); a.unit_base = static_cast<uint32_t>(l * H); a.nch = H / kWKc;
    launch_pdl(k_hidden, dim3(tiles), dim3(320), smem_of(H), st, pdl, dp, a);
  }
  a.A = ws + p.off_pa[L - 1]; a.W = ws + p.off_wfh; a.out = ws + p.off_pv0; a.bias = net->bv0; a.bias2 = net->bp;
  a.outT = ws + p.off_v0t; a.T_rows = RV;
  a.layer = static_cast<uint32_t>(L); a.unit_base = static_cast<uint32_t>(L * H); a.nch = H / kWKc;
  launch_pdl(k_heads, dim3(tiles), dim3(320), smem_of(NH), st, pdl, dp, a);
  // ============================ tail: last variance layer forward + backward, loss gradient ============================
  a.A = ws + p.off_pv0; a.W = ws + p.off_wfv1; a.out = ws + p.off_pdz1; a.bias = net->bv1; a.bias2 = net->bv2; a.w2 = net->Wv2;
  a.outT = ws + p.off_dt; a.T_rows = 0; a.nch = (H / 2) / kWKc;
  launch_pdl(k_v1t, dim3(tiles), dim3(320), smem_of(H / 4), st, pdl, dp, a);
  launch_pdl(wide_tail_reduce_kernel, dim3(H / 4 + 5), dim3(256), 0, st, pdl, a.tail_partial, a.loss_partial, tiles, H / 4, grad_flat + lay.offWv2, grad_flat + lay.offbv2,
                                             grad_u ? nullptr : loss_sums);
  // weight-gradient GEMM + reduce helpers
  const int per_split = static_cast<int>((chunks + p.splits - 1) / p.splits);
  auto wgrad = [&](auto kern, int N, int nmb, const unsigned char* Bt) {
    a.A = ws + p.off_dt; a.W = Bt; a.nch = per_split;
    launch_pdl(kern, dim3(p.splits, nmb), dim3(320), smem_of(N), st, pdl, dp, a);
  };
  auto reduce = [&](int nmb, int N, int mb, int row0, int nrows, int ncols, int bias_col, float* dW, int ld, float* dB, float scale) {
    const int items = nrows * (ncols + (dB ? 1 : 0));
    launch_pdl(wide_wgrad_reduce_kernel, dim3((items + 255) / 256), dim3(256), 0, st, pdl, a.wg_partial, p.splits, nmb, N, mb, row0, nrows, ncols, bias_col, dW, ld,
                                                                   dB, scale);
  };
  // dWv1 / dbv1 = dz1^T [v0 | 1]
  wgrad(k_wg_v, RV, 1, ws + p.off_v0t);
  reduce(1, RV, 0, 0, H / 4, H / 2, H / 2, grad_flat + lay.offWv1, H / 2, grad_flat + lay.offbv1, wscale);
  // ============================ dgrad through the variance head's first layer ============================
  a.A = ws + p.off_pdz1; a.W = ws + p.off_wtv1; a.out = ws + p.off_pdv0h; a.act = ws + p.off_pv0;
  a.outT = ws + p.off_dt; a.T_rows = 0; a.nch = (H / 4) / kWKc;
  a.layer = static_cast<uint32_t>(L); a.unit_base = static_cast<uint32_t>(L * H);
  launch_pdl(k_dv0, dim3(tiles), dim3(320), smem_of(H / 2), st, pdl, dp, a);
  // dWv0 / dbv0 (block 0) and dWp / dbp (block 1, row 0) = [dz_v0 ; du]^T [a_{L-1} | 1]
  constexpr int NMB_V0 = (H / 2 + 16 + kWT - 1) / kWT;         // delta rows: H/2 of dz_v0, then du at row H/2
  wgrad(k_wg_b, RB, NMB_V0, ws + p.off_at[L - 1]);
  reduce(NMB_V0, RB, 0, 0, H / 2 < kWT ? H / 2 : kWT, H, H, grad_flat + lay.offWv0, H, grad_flat + lay.offbv0, wscale);
  reduce(NMB_V0, RB, (H / 2) / kWT, (H / 2) % kWT, 1, H, H, grad_flat + lay.offWp, H, grad_flat + lay.offbp, wscale);
  // ============================ dgrad through the heads into the trunk ============================
  int cur = 0;
  a.A = ws + p.off_pdv0h; a.W = ws + p.off_wth; a.out = ws + p.off_pd[cur]; a.act = ws + p.off_pa[L - 1];
  a.outT = ws + p.off_dt; a.T_rows = 0; a.nch = NH / kWKc;
  a.layer = static_cast<uint32_t>(L - 1); a.unit_base = static_cast<uint32_t>((L - 1) * H);
  launch_pdl(k_dz, dim3(tiles), dim3(320), smem_of(H), st, pdl, dp, a);
  for (int l = L - 1; l >= 1; --l) {
    // dW_l / db_l = dz_l^T [a_{l-1} | 1]   (two 128-row blocks)
    wgrad(k_wg_b, RB, H / kWT, ws + p.off_at[l - 1]);
    for (int mb = 0; mb < H / kWT; ++mb)
      reduce(H / kWT, RB, mb, 0, kWT, H, H, grad_flat + lay.offW[l] + static_cast<size_t>(mb) * kWT * H, H, grad_flat + lay.offb[l] + mb * kWT, wscale);
    // dz_{l-1} = (dz_l W_l) * keep_{l-1} * (1 - a_{l-1}^2)
    a.A = ws + p.off_pd[cur]; a.W = ws + p.off_wt[l]; a.out = ws + p.off_pd[cur ^ 1]; a.act = ws + p.off_pa[l - 1];
    a.outT = ws + p.off_dt; a.T_rows = 0; a.nch = H / kWKc;
    a.layer = static_cast<uint32_t>(l - 1); a.unit_base = static_cast<uint32_t>((l - 1) * H);
    launch_pdl(k_dz, dim3(tiles), dim3(320), smem_of(H), st, pdl, dp, a);
    cur ^= 1;
  }
  // dW0 / db0 = dz_0^T [x | 1]
  wgrad(k_wg_x, 16, H / kWT, ws + p.off_xt);
  for (int mb = 0; mb < H / kWT; ++mb)
    reduce(H / kWT, 16, mb, 0, kWT, PINN_N_IN, PINN_N_IN, grad_flat + lay.offW[0] + static_cast<size_t>(mb) * kWT * PINN_N_IN, PINN_N_IN,
           grad_flat + lay.offb[0] + mb * kWT, 1.0f);
  return static_cast<int>(cudaGetLastError());
}

int launch_wide_tc_bwd(const pinn_net_t* net, const float* x, int64_t n, const DropParams& dp, const float* grad_u, const float* grad_s,
                       const float* y, int64_t n_global, float* grad_flat, double* loss_sums, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
  return net->width == 256
             ? run_wide_bwd<256>(net, x, n, dp, grad_u, grad_s, y, n_global, grad_flat, loss_sums, workspace, workspace_bytes, st)
             : run_wide_bwd<128>(net, x, n, dp, grad_u, grad_s, y, n_global, grad_flat, loss_sums, workspace, workspace_bytes, st);
}

}  // namespace pinn

