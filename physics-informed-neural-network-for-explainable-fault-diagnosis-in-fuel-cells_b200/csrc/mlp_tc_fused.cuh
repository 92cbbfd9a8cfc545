// mlp_tc_fused.cuh -- K2 for the 64-wide net (2 or 3 hidden layers) as ONE kernel per step: forward recompute,
// dgrad AND every weight / bias gradient of a 128-sample tile inside the CTA that owns the tile.  Nothing per-sample
// ever reaches HBM (the two-kernel form, K2a + K2b in mlp_tc_bwd.cu, hands 2 KB per sample over through a row table:
// 4 GB of traffic per step at N = 1M for a 40 MB problem).  Included by mlp_tc_bwd.cu.
//
// CTA = 16 compute warps + 1 MMA / TMA warp, one tile in flight.  Compute thread (q, c): TMEM lane quadrant q = warp & 3,
// sample row = 32 q + lane, column slice c = warp >> 2 (16 of the 64 units of a layer).
//
//   chain (serial per tile, 2L tensor-core products, A operand in tensor memory as in K2a):
//     a0 -> [W1] -> a1 -> [W2] -> a2 -> [Wv0;Wp] -> tail (CUDA cores) -> [(Wv0;Wp)^T] -> d2 -> [W2^T] -> d1 -> [W1^T] -> d0
//   weight gradients (L + 1 batches per tile, accumulators RESIDENT in tensor memory for the CTA's whole tile range):
//     the contraction index is the SAMPLE, so both operands must be sample-contiguous (K-major; MN-major TF32 operands
//     need the 32-byte-base swizzle, tests/cuda/tc_mn_test.cu).  Thread = sample row means a thread holds one K index of
//     16 features: it scatters them with 4-byte stores into a K-major plane whose chunk stride (LBO) is padded to
//     4 banks mod 32, so a warp's 32 stores hit 32 banks.  Two staging planes:
//       DEL (A operand, M = 128): rows [0,64) = tf32 hi part of up to 64 features, rows [64,128) = their lo part;
//       ACT (B operand):          rows [0,80) = hi part of up to 80 features, rows [80,160) = lo part.
//     One product per K = 8 slab and per B part:  D[128 x N] += DEL * ACT_hi^T,  D += DEL * ACT_lo^T  -- with hi and lo
//     of the A side stacked along M, the two MMAs produce all four cross terms (hi*hi, lo*hi, hi*lo, lo*lo); the read-out
//     adds lanes j and 64 + j.
//       batch H : DEL = [dz_v0 (32) | du]            ACT = [a_{L-1} (64) | 1]    -> dWv0, dWp, dbv0, dbp
//       batch l : DEL = delta_l (64)                 ACT = [a_{l-1} (64) | 1]    -> dW_l, db_l          (l = L-1 .. 1)
//       batch 0T: DEL = [x (8) | av0 (32) | av1 (16) | 1]   ACT = [delta_0 (64) | dz_v1 (16)]   (roles swapped: the result is
//                 transposed)  -> dW0, db0, dWv1, dbv1;  dWv2 / dbv2 (17 numbers) are summed on the CUDA cores.
//     The ones row of ACT (row 64) turns a bias sum into one more accumulator column.
//   Weight planes do not fit next to the staging planes (2L x 32 KB): they are pre-split once per step by
//   weight_image_kernel into the exact shared-memory image of each UMMA operand and streamed through a two-slot ring
//   with one 32 KB bulk copy (TMA) per product, issued one product ahead.
//   Tensor memory (512 columns): D 64 | A hi 64 | A lo 64 | (L + 1) x 80 accumulator columns.
//
// Order of the tensor-core work of a backward phase: the chain product first (the compute warps wait for it), the
// weight-gradient batch behind it -- it runs under the next epilogue.
#pragma once

namespace pinn {

constexpr int kFzThreads = 512;
constexpr uint32_t kFzImgBytes = 32768, kFzPlaneBytes = 16384, kFzImgLbo = 1024;
constexpr uint32_t kDelLbo = 2064, kActLbo = 2576;          // bytes per 4-sample chunk: 128 / 160 rows x 16 B + 16 B of padding
constexpr int kActLo = 80;                                  // first lo row of ACT
constexpr uint32_t kDelBytes = 32 * kDelLbo, kActBytes = 32 * kActLbo;
constexpr uint32_t kColD = 0, kColAhi = 64, kColAlo = 128, kColAcc = 192, kAccW = 80;
// rows of the 0T batch's DEL plane
constexpr int kTX = 0, kTV0 = 8, kTV1 = 40, kTOne = 56;

struct FzSmall {      // float offsets into the small-tensor area of shared memory
  int W0, b[3], bv0, bv1, Wv1f, Wv1t, Wv2, bp, bv2, total;
};
PINN_HD FzSmall make_fz_small(int L) {
  FzSmall t{};
  int o = 0;
  t.W0 = o; o += 2 * 64 * PINN_N_IN;      // layer 0 as UMMA planes: hi [2 chunks][64 rows][4], then lo
  for (int l = 0; l < 3; ++l) { t.b[l] = o; if (l < L) o += 64; }
  t.bv0 = o; o += 32; t.bv1 = o; o += 16;
  t.Wv1f = o; o += 2 * 16 * 32;      // Wv1 [16 x 32] as UMMA planes (B of the 32 -> 16 product): hi [8 chunks][16 rows][4], then lo
  t.Wv1t = o; o += 2 * 16 * 32;      // Wv1^T (B of its dgrad, K = 16): hi [4 chunks][32 rows][4], then lo
  t.Wv2 = o; o += 16;
  t.bp = o; o += 4; t.bv2 = o; o += 4;
  t.total = o;
  return t;
}
PINN_HD size_t fz_smem_bytes(int L) { return 2 * kFzImgBytes + kDelBytes + kActBytes + static_cast<size_t>(make_fz_small(L).total) * sizeof(float); }

// UMMA operand images of the step's weights, 2L x 32 KB in product order:
//   [0, L-1) forward W_1..W_{L-1} | L-1 forward heads [Wv0; Wp; 0] | L heads transposed | L+1.. transposed W_{L-1}..W_1
// every image = tf32 hi plane (16 KB) + lo plane, K-major, 64 rows x 16 B per 4-element K chunk (LBO = 1024 B), dropout
// scale folded in (see mlp_tc.cu).  One thread splits one float4 of one matrix into both orientations.
__global__ void __launch_bounds__(256) weight_image_kernel(pinn_net_t net, float wscale, unsigned char* __restrict__ images) {
  griddep_launch();
  griddep_wait();            // the weights come from the previous step's optimiser launch
  const int L = net.n_hidden;
  const int m = blockIdx.x >> 2;                                 // 0 .. L-2: W_{m+1};  L-1: heads
  const int idx = (blockIdx.x & 3) * 256 + threadIdx.x, j = idx & 63, kc = idx >> 6;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m < L - 1) v = __ldg(reinterpret_cast<const float4*>(net.W[m + 1] + j * 64) + kc);
  else if (j < 32) v = __ldg(reinterpret_cast<const float4*>(net.Wv0 + j * 64) + kc);
  else if (j == 32) v = __ldg(reinterpret_cast<const float4*>(net.Wp) + kc);
  const float vv[4] = {v.x * wscale, v.y * wscale, v.z * wscale, v.w * wscale};
  float h[4], lo[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { h[r] = tc::tf32_hi(vv[r]); lo[r] = vv[r] - h[r]; }
  unsigned char* fw = images + static_cast<size_t>(m) * kFzImgBytes;
  unsigned char* bw = images + static_cast<size_t>(2 * L - 1 - m) * kFzImgBytes;
  *reinterpret_cast<float4*>(fw + kc * kFzImgLbo + j * 16) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(fw + kFzPlaneBytes + kc * kFzImgLbo + j * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
  for (int r = 0; r < 4; ++r) {        // transposed: row n = 4 kc + r, K index j
    const size_t off = static_cast<size_t>(j >> 2) * kFzImgLbo + static_cast<size_t>(4 * kc + r) * 16 + (j & 3) * 4;
    *reinterpret_cast<float*>(bw + off) = h[r];
    *reinterpret_cast<float*>(bw + kFzPlaneBytes + off) = lo[r];
  }
}

// The same image entries for ONE parameter of the flat bucket (entry i, new value p): used by the optimiser launch.
PINN_D void image_write_entry(const ParamLayout& lay, int64_t i, float p, float wscale, unsigned char* __restrict__ images) {
  const int L = lay.L;
  int m = -1, j = 0, k = 0;
  for (int l = 1; l < L; ++l) {
    const int64_t r = i - lay.offW[l];
    if (r >= 0 && r < 64 * 64) { m = l - 1; j = static_cast<int>(r >> 6); k = static_cast<int>(r & 63); }
  }
  {
    const int64_t r = i - lay.offWv0, rp = i - lay.offWp;
    if (r >= 0 && r < 32 * 64) { m = L - 1; j = static_cast<int>(r >> 6); k = static_cast<int>(r & 63); }
    if (rp >= 0 && rp < 64) { m = L - 1; j = 32; k = static_cast<int>(rp); }
  }
  if (m < 0) return;
  const float v = p * wscale, h = tc::tf32_hi(v), lo = v - h;
  unsigned char* fw = images + static_cast<size_t>(m) * kFzImgBytes + static_cast<size_t>(k >> 2) * kFzImgLbo + j * 16 + (k & 3) * 4;
  unsigned char* bw = images + static_cast<size_t>(2 * L - 1 - m) * kFzImgBytes + static_cast<size_t>(j >> 2) * kFzImgLbo + k * 16 + (j & 3) * 4;
  *reinterpret_cast<float*>(fw) = h; *reinterpret_cast<float*>(fw + kFzPlaneBytes) = lo;
  *reinterpret_cast<float*>(bw) = h; *reinterpret_cast<float*>(bw + kFzPlaneBytes) = lo;
}

struct FzArgs {
  const float* x; int64_t n;
  const float* grad_u; const float* grad_s; const float* y; float inv_n_global;
  const unsigned char* images;
  float* partial;            // [grid][lay.total]
  double* loss_partial;      // [grid][4]
  float* park;               // [grid][L + 1][4][512] float4 (see the kernel)
};

#ifdef PINN_TIMELINE
#ifndef PINN_TL_TID
#define PINN_TL_TID 0      // the compute thread that stamps (0: warp 0 = quadrant 0, slice 0, on the MMA warp's scheduler)
#endif
__device__ long long g_tlf[2][256];      // CTA 0: [0] compute thread 0, [1] the MMA warp's elected lane; clock64 at the stamps below, 32 per tile
#define TLF(i) do { if (blockIdx.x == 0 && tid == PINN_TL_TID && (i) < 256) g_tlf[0][i] = clock64(); } while (0)
#define TLM(i) do { if (blockIdx.x == 0 && (i) < 256) g_tlf[1][i] = clock64(); } while (0)
#else
#define TLF(i) do { } while (0)
#define TLM(i) do { } while (0)
#endif

template <int L, bool INJ>
__global__ void __launch_bounds__(kFzThreads + 32, 1)
mlp_tc_fused_kernel(pinn_net_t net, FzSmall sl, const __grid_constant__ DropParams dp, FzArgs a, ParamLayout pl) {
  static_assert(L == 2 || L == 3, "tensor memory holds L + 1 <= 4 accumulator blocks");
  constexpr int H = 64;
  extern __shared__ __align__(1024) unsigned char fsm[];
  __shared__ __align__(8) uint64_t bar_ready, bar_chain, bar_full[2];
  __shared__ __align__(8) uint64_t bar_stage[4], bar_wg[4];      // per lane quadrant = per 32-sample quarter of the staging planes
  __shared__ uint32_t tmem_base_s;
  __shared__ double lred[4][4];
  __shared__ double lacc[3][128];          // loss terms per sample row, summed over the CTA's tiles by the row's slice-0 thread
  __shared__ float wred[16][8];
  unsigned char* const ring = fsm;
  unsigned char* const DEL = fsm + 2 * kFzImgBytes;
  unsigned char* const ACT = DEL + kDelBytes;
  float* const sm = reinterpret_cast<float*>(ACT + kActBytes);
  const int tid = threadIdx.x, warp = tc::uniform_warp_idx();
  const int Dm = L * H + H / 2;
  griddep_launch();

  if (tid == 0) {
    tc::mbar_init(&bar_ready, kFzThreads); tc::mbar_init(&bar_chain, 1);
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&bar_stage[i], kFzThreads / 4); tc::mbar_init(&bar_wg[i], 1); }
    tc::mbar_init(&bar_full[0], 1); tc::mbar_init(&bar_full[1], 1);
    tc::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  // the ones rows of ACT (hi rows 64..79: row 64 = 1, the others 0; rewritten after every 0T batch, which parks dz_v1 there)
  for (int i = tid; i < 16 * 128; i += blockDim.x) {
    const int r = 64 + (i >> 7), s = i & 127;
    *reinterpret_cast<float*>(ACT + (s >> 2) * kActLbo + r * 16 + (s & 3) * 4) = r == 64 ? 1.0f : 0.0f;
  }
  griddep_wait();            // weights (and their images) come from launches earlier in the stream
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;
  if (tid < kFzThreads) {
    // small tensors: one element per thread (W0 and Wv1 are 512 floats each); biases / Wv2 / bp / bv2 laid over the thread index
    {   // W0 [64 x 8] -> K-major hi / lo planes (one K = 8 slab: two 4-column chunks of 64 rows x 16 B)
      const int j = tid >> 3, k = tid & 7;
      const float w = __ldg(net.W[0] + tid), h = tc::tf32_hi(w);
      const int off = (k >> 2) * 256 + j * 4 + (k & 3);
      sm[sl.W0 + off] = h;
      sm[sl.W0 + 512 + off] = w - h;
    }
    {   // Wv1 [16 x 32], element (k, i) = tid: both orientations as K-major hi / lo planes (no dropout scale: av0 carries it)
      const int k = tid >> 5, i = tid & 31;
      const float w = __ldg(net.Wv1 + tid), h = tc::tf32_hi(w);
      const int of = (i >> 2) * 64 + k * 4 + (i & 3), ot = (k >> 2) * 128 + i * 4 + (k & 3);
      sm[sl.Wv1f + of] = h; sm[sl.Wv1f + 512 + of] = w - h;
      sm[sl.Wv1t + ot] = h; sm[sl.Wv1t + 512 + ot] = w - h;
    }
    if (tid < 64 * L) sm[sl.b[tid >> 6] + (tid & 63)] = __ldg(net.b[tid >> 6] + (tid & 63)) * kTanhArg;
    else if (tid >= 256 && tid < 288) sm[sl.bv0 + tid - 256] = __ldg(net.bv0 + tid - 256) * kTanhArg;
    else if (tid >= 288 && tid < 304) sm[sl.bv1 + tid - 288] = __ldg(net.bv1 + tid - 288) * kTanhArg;
    else if (tid >= 304 && tid < 320) sm[sl.Wv2 + tid - 304] = __ldg(net.Wv2 + tid - 304);
    else if (tid == 320) sm[sl.bp] = __ldg(net.bp);
    else if (tid == 321) sm[sl.bv2] = __ldg(net.bv2);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const int64_t n_tiles = (a.n + 127) / 128;

  if (warp == kFzThreads / 32) {
    // ======================================================================================== MMA / TMA warp
    const uint32_t ring_u = tc::smem_u32(ring), del_u = tc::smem_u32(DEL), act_u = tc::smem_u32(ACT);
    const uint32_t idesc64 = tc::make_idesc_tf32(128, 64), idesc48 = tc::make_idesc_tf32(128, 48), idesc80 = tc::make_idesc_tf32(128, 80);
    uint32_t rp = 0, sp = 0, fpar = 0, g = 0;
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&bar_full[0], kFzImgBytes);
      tc::bulk_g2s(ring, a.images, kFzImgBytes, &bar_full[0]);
      tc::mbar_expect_tx(&bar_full[1], kFzImgBytes);
      tc::bulk_g2s(ring + kFzImgBytes, a.images + kFzImgBytes, kFzImgBytes, &bar_full[1]);
    }
    __syncwarp();
    bool first = true;
    int tm = 0;
    (void)tm;
    // Weight-gradient batch of the operands the compute warps have just staged.  The planes are handed over and handed back
    // by QUARTERS: lane quadrant q owns samples 32q..32q+31 = K slabs 4q..4q+3; it arrives on bar_stage[q] when its quarter
    // is written, and bar_wg[q] tells it when the tensor core has read that quarter -- so staging of the next batch and the
    // products of this one overlap slab by slab (both are bound by the same shared-memory port) instead of taking turns.
    auto wgrad_batch = [&](int acc, bool wide_lo) {
      const uint32_t d = tmem + kColAcc + kAccW * static_cast<uint32_t>(acc);
      const uint64_t a0 = tc::make_desc(del_u, kDelLbo, 128);
      const uint64_t bh0 = tc::make_desc(act_u, kActLbo, 128), bl0 = tc::make_desc(act_u + kActLo * 16, kActLbo, 128);
      const uint32_t id_lo = wide_lo ? idesc80 : idesc64;
      constexpr uint64_t as = (2u * kDelLbo) >> 4, bs = (2u * kActLbo) >> 4;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        tc::mbar_wait(&bar_stage[qq], sp);
        __syncwarp();
        if (tc::elect_one()) {
          tc::fence_after_sync();
#pragma unroll
          for (int ks = 4 * qq; ks < 4 * qq + 4; ++ks) {
            tc::umma_tf32(d, a0 + ks * as, bh0 + ks * bs, idesc80, (ks != 0 || !first) ? 1u : 0u);
            tc::umma_tf32(d, a0 + ks * as, bl0 + ks * bs, id_lo, 1u);
          }
          tc::umma_commit(&bar_wg[qq]);
        }
        __syncwarp();
      }
      sp ^= 1u;
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const bool has_next = tile + gridDim.x < n_tiles;
      {   // layer 0: x (K = 8, hi / lo in tensor memory) against the resident W0 planes
        tc::mbar_wait(&bar_ready, rp);
        rp ^= 1u;
        __syncwarp();
        if (tc::elect_one()) {
          tc::fence_after_sync();
          const uint32_t w0_u = tc::smem_u32(sm + sl.W0);
          tc::issue_3xtf32_ts<8>(tmem + kColD, tmem + kColAhi, tmem + kColAlo, tc::make_desc(w0_u, 1024, 128),
                                 tc::make_desc(w0_u + 2048, 1024, 128), 1024, idesc64);
          tc::umma_commit(&bar_chain);
        }
        __syncwarp();
      }
#pragma unroll
      for (int p = 0; p < 2 * L; ++p) {
        tc::mbar_wait(&bar_ready, rp);
        rp ^= 1u;
        __syncwarp();
        if (tc::elect_one()) {
          tc::fence_after_sync();
          TLM(tm + 4 * p);
          // (the image was requested two products ago by compute thread 0, see wait_chain_ring)
          const uint32_t s1 = g & 1u;
          tc::mbar_wait(&bar_full[s1], (fpar >> s1) & 1u);      // still ~450 clk late in the timelines (a spin instead of the
                                                                // suspending wait changes nothing): L2 carries ~4.7 TB/s of images + parking lot
          fpar ^= 1u << s1;
          TLM(tm + 4 * p + 1);
          const uint64_t bh = tc::make_desc(ring_u + s1 * kFzImgBytes, kFzImgLbo, 128);
          const uint64_t bl = tc::make_desc(ring_u + s1 * kFzImgBytes + kFzPlaneBytes, kFzImgLbo, 128);
          if (p == L - 1) tc::issue_3xtf32_ts<64>(tmem + kColD, tmem + kColAhi, tmem + kColAlo, bh, bl, kFzImgLbo, idesc48);
          else if (p == L) tc::issue_3xtf32_ts<48>(tmem + kColD, tmem + kColAhi, tmem + kColAlo, bh, bl, kFzImgLbo, idesc64);
          else tc::issue_3xtf32_ts<64>(tmem + kColD, tmem + kColAhi, tmem + kColAlo, bh, bl, kFzImgLbo, idesc64);
          tc::umma_commit(&bar_chain);
          TLM(tm + 4 * p + 2);
        }
        __syncwarp();
        ++g;
        if (p == L - 1) {
          // the variance head's two small layers: av0 (K = 32) x Wv1 -> 16 columns, then dz_v1 (K = 16) x Wv1^T -> 32 columns
#pragma unroll
          for (int tp = 0; tp < 2; ++tp) {
            tc::mbar_wait(&bar_ready, rp);
            rp ^= 1u;
            __syncwarp();
            if (tc::elect_one()) {
              tc::fence_after_sync();
              const uint32_t w_u = tc::smem_u32(sm + (tp == 0 ? sl.Wv1f : sl.Wv1t));
              if (tp == 0) tc::issue_3xtf32_ts<32>(tmem + kColD, tmem + kColAhi, tmem + kColAlo, tc::make_desc(w_u, 256, 128),
                                                   tc::make_desc(w_u + 2048, 256, 128), 256, tc::make_idesc_tf32(128, 16));
              else tc::issue_3xtf32_ts<16>(tmem + kColD, tmem + kColAhi, tmem + kColAlo, tc::make_desc(w_u, 512, 128),
                                           tc::make_desc(w_u + 2048, 512, 128), 512, tc::make_idesc_tf32(128, 32));
              tc::umma_commit(&bar_chain);
            }
            __syncwarp();
          }
        }
        // (Tried: issuing a batch in two halves around the next chain product so that the product does not queue behind a whole
        // batch -- 0.78 -> 0.83 ms: the later release of the second half's quarters costs more than the chain product gains.)
        if (p >= L) { wgrad_batch(p == L ? 0 : 2 * L - p, false); TLM(tm + 4 * p + 3); }      // staged while the chain product ran
      }
      wgrad_batch(L, true);          // batch 0T
      TLM(tm + 8 * L + 3);
      first = false;
      tm += 32;
    }
  } else {
    // ======================================================================================== compute warps
    const int q = warp & 3, c = warp >> 2, lane = tid & 31, row = q * 32 + lane, cb = 16 * c;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tD = tmem + kColD + lane_sel, tAh = tmem + kColAhi + lane_sel, tAl = tmem + kColAlo + lane_sel;
    unsigned char* const del_s = DEL + (row >> 2) * kDelLbo + (row & 3) * 4;
    unsigned char* const act_s = ACT + (row >> 2) * kActLbo + (row & 3) * 4;
    auto del_st = [&](int f, float h, float l) {
      *reinterpret_cast<float*>(del_s + f * 16) = h;
      *reinterpret_cast<float*>(del_s + (64 + f) * 16) = l;
    };
    auto act_st = [&](int r, float h, float l) {
      *reinterpret_cast<float*>(act_s + r * 16) = h;
      *reinterpret_cast<float*>(act_s + (kActLo + r) * 16) = l;
    };
    uint32_t chain_par = 0, wg_par = 0;
    bool wg_pending = false;
    auto arrive_ready = [&] {          // the chain operand lives in tensor memory: no shared-memory (async-proxy) fence needed here
      tc::tmem_wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(&bar_ready);
    };
    auto arrive_stage = [&] {          // staging planes written: the weight-gradient batch may run
      tc::fence_proxy_async();
      tc::mbar_arrive(&bar_stage[q]);
      wg_pending = true;
    };
    auto wait_chain = [&] {
      tc::mbar_wait(&bar_chain, chain_par);
      chain_par ^= 1u;
      __syncwarp();
      tc::fence_after_sync();
    };
    // A chain product that read ring slot `img & 1` has completed: compute thread 0 -- the first to know -- requests the image
    // after next into that slot (one bulk copy, 32 KB).  Requested by the MMA warp one product ahead, the copy left that warp
    // waiting ~600 clk at every chain product: all 148 SMs pull the same image out of L2 at about the same time.
    auto wait_chain_ring = [&](int img, bool has_next) {
      wait_chain();
      if (tid == 0) {
        const int nxt = img + 2 < 2 * L ? img + 2 : (has_next ? img + 2 - 2 * L : -1);
        if (nxt >= 0) {
          tc::mbar_expect_tx(&bar_full[img & 1], kFzImgBytes);
          tc::bulk_g2s(ring + (img & 1) * kFzImgBytes, a.images + static_cast<size_t>(nxt) * kFzImgBytes, kFzImgBytes, &bar_full[img & 1]);
        }
      }
      __syncwarp();
    };
    auto wait_wg = [&] {          // the batch that reads the staging planes has drained them
      if (wg_pending) { tc::mbar_wait(&bar_wg[q], wg_par); wg_par ^= 1u; wg_pending = false; }
      __syncwarp();
    };
    auto qbar = [&] {             // the four warps of a lane quadrant (the four column slices of 32 sample rows)
      tc::tmem_wait_st();
      tc::fence_before_sync();
      asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory");
      tc::fence_after_sync();
    };
    auto split16 = [&](const float (&v)[16], float (&h)[16], float (&lo)[16]) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        h[i] = tc::tf32_hi_fast(v[i]); h[i + 1] = tc::tf32_hi_fast(v[i + 1]);
        const float2 l2 = __ffma2_rn(make_float2(h[i], h[i + 1]), make_float2(-1.0f, -1.0f), make_float2(v[i], v[i + 1]));   // v - h, exact
        lo[i] = l2.x; lo[i + 1] = l2.y;
      }
    };
    auto keep16 = [&](const KeepSrc<INJ>& ks, bool active, uint32_t layer) {      // bit i = unit cb + i of dropout layer `layer`
      uint32_t bits = 0xffffu;
      if (active) {
        bits = 0u;
#pragma unroll
        for (int g8 = 0; g8 < 16; g8 += 8) {
          bool k[8];
          ks.get8(dp, layer, static_cast<uint32_t>(cb + g8), layer * H, k);
#pragma unroll
          for (int i = 0; i < 8; ++i) bits |= (k[i] ? 1u : 0u) << (g8 + i);
        }
      }
      return bits;
    };
    const bool no_lv = (net.flags & PINN_NET_NO_LOGVAR) != 0;
    if (c == 0) { lacc[0][row] = 0.0; lacc[1][row] = 0.0; lacc[2][row] = 0.0; }
    float g_wv2[4] = {0.f, 0.f, 0.f, 0.f}, g_bv2 = 0.f;
    // Parking lot (global, 32 KB per slot and CTA, L2-resident: rewritten every tile): the activations of layers 0..L-2
    // and the tail's (av0, v1, dz_v1) wait here between the forward and the backward phase that needs them -- 64 live
    // registers per thread otherwise (the 17th warp caps the kernel at 96).  Slot layout [4][512] float4: coalesced.
    float4* const park = reinterpret_cast<float4*>(a.park) + static_cast<size_t>(blockIdx.x) * (L + 1) * 4 * kFzThreads + tid;
    auto park_st = [&](int slot, const float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) park[(slot * 4 + i) * kFzThreads] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    };
    auto park_ld = [&](int slot, float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = __ldcg(park + (slot * 4 + i) * kFzThreads);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    int tl = 0;
    (void)tl;

    // x of the first tile; every later tile's row is fetched while the previous tile's last phase runs
    float4 xq0 = make_float4(0.f, 0.f, 0.f, 0.f), xq1 = xq0;
    auto load_x = [&](int64_t tile) {
      const int64_t s = tile * 128 + row;
      if (tile < n_tiles && s < a.n) {
        const float4* px = reinterpret_cast<const float4*>(a.x + s * PINN_N_IN);
        xq0 = __ldg(px); xq1 = __ldg(px + 1);
      } else {
        xq0 = xq1 = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    load_x(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t s = tile * 128 + row;
      const bool valid = s < a.n, has_next = tile + gridDim.x < n_tiles;
      const bool active = drop_on && (!INJ || valid);       // injected masks: tail rows of the last tile have no mask row
      const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
      KeepSrc<INJ> ks;
      ks.s_lo = static_cast<uint32_t>(sg); ks.s_hi = static_cast<uint32_t>(sg >> 32);
      ks.pass = static_cast<uint32_t>(dp.pass_offset);
      ks.mrow = INJ ? dp.masks + static_cast<size_t>(valid ? s : 0) * Dm : nullptr;
      float acur[16];              // this thread's 16 masked activations of the current trunk layer (without the dropout scale)
      uint32_t kb[L];
      TLF(tl);
      // ============================ forward ============================
      // layer 0 runs on the tensor core too (K = 8): the slice-0 thread of a row parks x as hi / lo columns 0..7
      if (c == 0) {
        const float xr[PINN_N_IN] = {xq0.x, xq0.y, xq0.z, xq0.w, xq1.x, xq1.y, xq1.z, xq1.w};
        float h8[8], l8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h8[i] = tc::tf32_hi_fast(xr[i]); l8[i] = xr[i] - h8[i]; }
        tc::tmem_st8(tAh, h8);
        tc::tmem_st8(tAl, l8);
      }
      arrive_ready();
      TLF(tl + 1);
#pragma unroll
      for (int l = 0; l < L; ++l) {
        kb[l] = keep16(ks, active, static_cast<uint32_t>(l));        // drawn while the tensor core works
        TLF(tl + 2 + 2 * l);
        if (l == 0) wait_chain(); else wait_chain_ring(l - 1, has_next);
        TLF(tl + 3 + 2 * l);
        const float* bl = sm + sl.b[l] + cb;
        // an opaque copy of the keep bits: otherwise the compiler extracts all 16 bit tests ahead of the wait, keeps them for the
        // backward phase that tests the same bits, and spills every one of them (16 STL + 16 LDL per layer and tile)
        uint32_t kbf = kb[l];
        asm volatile("" : "+r"(kbf));
#pragma unroll
        for (int g8 = 0; g8 < 16; g8 += 8) {          // eight columns at a time: half the live registers of a 16-wide pass
          float z[8];
          tc::tmem_ld8(tD + cb + g8, z);
          tc::tmem_wait_ld();
          const float4 bA = *reinterpret_cast<const float4*>(bl + g8), bB = *reinterpret_cast<const float4*>(bl + g8 + 4);
          const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
          float t8[8], h8[8], l8[8];
          tanh8_prescaled(z, bb, t8);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acur[g8 + i] = ((kbf >> (g8 + i)) & 1u) ? t8[i] : 0.f;
            h8[i] = tc::tf32_hi_fast(acur[g8 + i]); l8[i] = acur[g8 + i] - h8[i];
          }
          tc::tmem_st8(tAh + cb + g8, h8);
          tc::tmem_st8(tAl + cb + g8, l8);
        }
        arrive_ready();
        park_st(l < L - 1 ? l : L, acur);          // the last layer's activations sit out the tail in slot L
      }
      // ---- heads (rows 0..31 = Wv0, row 32 = Wp) and the variance head's tail, split over the row's four threads:
      //      thread c owns units 8c..8c+7 of the 32-wide layer and units 4c..4c+3 of the 16-wide layer; three hand-overs
      //      through free tensor-memory columns (A hi 0..31: av0, A hi 48..63: v1, A lo 48..63: dz_v1, D 56: d loss / d raw variance).
      //      The scalar part (log-variance, loss, output gradients, all 16 dz_v1) runs on the row's slice-0 thread only:
      //      the four threads of a row share one scheduler, so doing it four times would cost four times the issue slots.
      uint32_t kbv = 0xffu;
      if (active) {
        bool k[8];
        ks.get8(dp, static_cast<uint32_t>(L), static_cast<uint32_t>(8 * c), static_cast<uint32_t>(L * H), k);
        kbv = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) kbv |= (k[i] ? 1u : 0u) << i;
      }
      // target (or the caller's output gradients) of this row: in flight while the heads product runs
      float y_pre = 0.f, gs_pre = 0.f;
      if (c == 0 && valid) {
        if (a.grad_u != nullptr) { y_pre = __ldg(a.grad_u + s); gs_pre = a.grad_s ? __ldg(a.grad_s + s) : 0.f; }
        else y_pre = __ldg(a.y + s);
      }
      TLF(tl + 8);
      wait_chain_ring(L - 1, has_next);
      TLF(tl + 9);
      float dzv0[8], du = 0.f;
      {
        float tailv[16];           // [0,8) av0 (scaled), [8,12) v1, [12,16) dz_v1: parked for the 0T batch
        float zv[8], zu[8];
        tc::tmem_ld8(tD + 8 * c, zv);
        tc::tmem_ld8(tD + 32, zu);
        tc::tmem_wait_ld();
        const float* bv0 = sm + sl.bv0 + 8 * c;
        uint32_t kbv1 = kbv;
        asm volatile("" : "+r"(kbv1));
#pragma unroll
        for (int i = 0; i < 8; ++i) {          // activation first, select second: a conditional around the MUFU pair compiles to a divergent branch per unit
          const float t = tanh_pre(fmaf(zv[i], kTanhArg, bv0[i])) * wscale;
          tailv[i] = ((kbv1 >> i) & 1u) ? t : 0.f;
        }
        {   // av0 = the A operand (K = 32) of the 32 -> 16 product
          float h8[8], l8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { h8[i] = tc::tf32_hi_fast(tailv[i]); l8[i] = tailv[i] - h8[i]; }
#ifdef PINN_TIMELINE
          asm volatile("" : "+f"(h8[0]), "+f"(h8[7]), "+f"(l8[0]), "+f"(l8[7]));      // values complete before the stamp
#endif
          TLF(tl + 30);
          tc::tmem_st8(tAh + 8 * c, h8);
          tc::tmem_st8(tAl + 8 * c, l8);
          TLF(tl + 31);
        }
        arrive_ready();
        TLF(tl + 10);
        wait_chain();
        TLF(tl + 11);
        {
          float z1[4];
          tc::tmem_ld4(tD + 4 * c, z1);
          tc::tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 4; ++k) tailv[8 + k] = tanh_pre(fmaf(z1[k], kTanhArg, sm[sl.bv1 + 4 * c + k]));
        }
        tc::tmem_st4(tAh + 48 + 4 * c, tailv + 8);
        qbar();                    // the slice-0 thread needs the row's 16 v1
        TLF(tl + 12);
        if (c == 0) {
          float v1a[16], dz[16];
          tc::tmem_ld16(tAh + 48, v1a);
          tc::tmem_wait_ld();
          float vraw = sm[sl.bv2];
#pragma unroll
          for (int k = 0; k < 16; ++k) vraw = fmaf(sm[sl.Wv2 + k], v1a[k], vraw);
          const float u = zu[0] + sm[sl.bp];
          // log-variance head 01:432-434 with one accurate log1p and MUFU for the rest: with t = softplus(v) + 1e-6,
          // logvar = log t, exp(-logvar) = 1 / t, d logvar / d v = sigmoid(v) / t, sigmoid(v) = e^v / (1 + e^v) (v <= 20)
          const bool big = vraw > 20.0f;
          const float ev = __expf(big ? 0.f : vraw);
          const float t = (big ? vraw : log1pf(ev)) + 1e-6f;
          const float rt = __frcp_rn(t);
          const float lv = no_lv ? 0.f : logf(t);
          float ds = 0.f;
          if (valid) {
            if (a.grad_u != nullptr) {
              du = y_pre;
              ds = gs_pre;
            } else {
              const float yv = y_pre;
              const float e = no_lv ? 1.0f : rt, diff = yv - u;
              du = -e * diff * a.inv_n_global;
              const float sgn = lv > 0.f ? 1.f : (lv < 0.f ? -1.f : 0.f);
              ds = (-0.5f * e * diff * diff + 0.5f + 0.01f * sgn) * a.inv_n_global;
              lacc[0][row] += static_cast<double>(0.5f * e * diff * diff + 0.5f * lv);
              lacc[1][row] += static_cast<double>(fabsf(lv));
              lacc[2][row] += static_cast<double>(diff * diff);
            }
          }
          const float sig = big ? 1.0f : __fdividef(ev, 1.0f + ev);
          const float dvv = no_lv ? 0.f : ds * sig * rt;
#pragma unroll
          for (int k = 0; k < 16; ++k) dz[k] = dvv * sm[sl.Wv2 + k] * (1.0f - v1a[k] * v1a[k]);
          float h[16], lo[16];
          split16(dz, h, lo);
          tc::tmem_st16(tAh, h);                   // dz_v1 = the A operand (K = 16) of the product with Wv1^T
          tc::tmem_st16(tAl, lo);
          tc::tmem_st16(tAl + 48, dz);             // and as is, for the row's other threads (they park their four for the 0T batch)
          tc::tmem_st1(tD + 56, dvv);              // D columns 32.. are not touched by the N = 32 product
        }
        arrive_ready();
        TLF(tl + 13);
        wait_chain();
        TLF(tl + 14);
        float dv8[8];
        tc::tmem_ld8(tD + 8 * c, dzv0);            // sum_k Wv1[k][i] dz_v1[k] for this thread's eight units
        tc::tmem_ld4(tAl + 48 + 4 * c, tailv + 12);
        tc::tmem_ld8(tD + 56, dv8);
        tc::tmem_wait_ld();
        const float dv = dv8[0];
        // dz_v0 = d v0 * keep-mask * scale * (1 - a^2)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = tailv[i] * (drop_on ? dp.keep : 1.0f);
          const float mk = ((kbv >> i) & 1u) ? wscale : 0.f;
          dzv0[i] = dzv0[i] * mk * (1.0f - av * av);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) g_wv2[k] = fmaf(dv, tailv[8 + k], g_wv2[k]);
        if (c == 0) g_bv2 += dv;
        // every thread of the row has passed the third barrier, so the hand-over columns below 48 are dead and the
        // heads^T operand (columns 0..47 of both planes) may overwrite them
        // ============================ backward ============================
        // chain operand [dz_v0 (32 cols) | du | 0 ...] (K = 48)
        float h8[8], l8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h8[i] = tc::tf32_hi_fast(dzv0[i]); l8[i] = dzv0[i] - h8[i]; }
        tc::tmem_st8(tAh + 8 * c, h8);
        tc::tmem_st8(tAl + 8 * c, l8);
        const float duh = tc::tf32_hi_fast(du), dul = du - duh;
        if (c == 0) {
          const float e8[8] = {duh, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, f8[8] = {dul, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          tc::tmem_st8(tAh + 32, e8);
          tc::tmem_st8(tAl + 32, f8);
        } else if (c == 1) {
          const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          tc::tmem_st8(tAh + 40, z8);
          tc::tmem_st8(tAl + 40, z8);
        }
        arrive_ready();            // the heads^T product runs while batch H is staged
        TLF(tl + 15);
        park_st(L - 1, tailv);
        park_ld(L, acur);          // lands while the previous tile's 0T batch drains the planes
        wait_wg();
        TLF(tl + 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float*>(act_s + (64 + 4 * c + i) * 16) = (c == 0 && i == 0) ? 1.0f : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) del_st(8 * c + i, h8[i], l8[i]);
        if (c == 0) del_st(32, duh, dul);
        float h[16], lo[16];
        split16(acur, h, lo);
#pragma unroll
        for (int i = 0; i < 16; ++i) act_st(cb + i, h[i], lo[i]);
        arrive_stage();
        TLF(tl + 17);
      }
#pragma unroll
      for (int l = L - 1; l >= 0; --l) {
        float x2[2] = {0.f, 0.f};                       // this thread's two input features for the 0T batch
        if (l == 0) {
          if (valid) { const float2 t2 = __ldg(reinterpret_cast<const float2*>(a.x + s * PINN_N_IN) + c); x2[0] = t2.x; x2[1] = t2.y; }
          load_x(tile + gridDim.x);                     // the next tile's input row
        }
        wait_chain_ring(2 * L - 1 - l, has_next);
        TLF(tl + 18 + 4 * (L - 1 - l));
        float dz[16];
        tc::tmem_ld16(tD + cb, dz);
        tc::tmem_wait_ld();
        uint32_t kbb = kb[l];
        asm volatile("" : "+r"(kbb));      // see the forward pass
#pragma unroll
        for (int i = 0; i < 16; ++i)     // z already carries the dropout scale (folded into W^T); dropped units have a = 0
          dz[i] = ((kbb >> i) & 1u) ? dz[i] * fmaf(-acur[i], acur[i], 1.0f) : 0.f;
        float nxt[16];             // l > 0: activations of layer l - 1;  l == 0: the parked tail values
        if (l > 0) {
          {
            float h[16], lo[16];
            split16(dz, h, lo);
            tc::tmem_st16(tAh + cb, h);
            tc::tmem_st16(tAl + cb, lo);
          }
          arrive_ready();          // the next chain product runs while this layer's batch is staged
          TLF(tl + 19 + 4 * (L - 1 - l));
          park_ld(l - 1, nxt);     // lands while the previous batch drains the planes
          wait_wg();
          TLF(tl + 20 + 4 * (L - 1 - l));
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float hh = tc::tf32_hi_fast(dz[i]); del_st(cb + i, hh, dz[i] - hh); }
#pragma unroll
          for (int i = 0; i < 16; ++i) { acur[i] = nxt[i]; const float hh = tc::tf32_hi_fast(acur[i]); act_st(cb + i, hh, acur[i] - hh); }
        } else {
          // batch 0T (roles swapped): DEL = [x | av0 | av1 | 1], ACT = [delta_0 | dz_v1]
          TLF(tl + 19 + 4 * (L - 1 - l));
          park_ld(L - 1, nxt);
          wait_wg();
          TLF(tl + 20 + 4 * (L - 1 - l));
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float hh = tc::tf32_hi_fast(dz[i]); act_st(cb + i, hh, dz[i] - hh); }
#pragma unroll
          for (int i = 0; i < 2; ++i) { const float hh = tc::tf32_hi_fast(x2[i]); del_st(kTX + 2 * c + i, hh, x2[i] - hh); }
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float hh = tc::tf32_hi_fast(nxt[i]); del_st(kTV0 + 8 * c + i, hh, nxt[i] - hh); }
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float hh = tc::tf32_hi_fast(nxt[8 + i]); del_st(kTV1 + 4 * c + i, hh, nxt[8 + i] - hh); }
          if (c == 0) del_st(kTOne, 1.0f, 0.0f);
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float hh = tc::tf32_hi_fast(nxt[12 + i]); act_st(64 + 4 * c + i, hh, nxt[12 + i] - hh); }
        }
        arrive_stage();
        TLF(tl + 21 + 4 * (L - 1 - l));
      }
      tl += 32;
    }
    TLF(tl);
    // ---------------------------------------------------------------- accumulators -> this CTA's partial vector
    // Lanes j and 64 + j hold the hi and the lo part of the same sum: the lo half of the CTA (quadrants 2, 3) parks its
    // values in the idle staging plane, the hi half adds them and writes the vector out.
    if (wg_pending) tc::mbar_wait(&bar_wg[3], wg_par);      // the LAST quarter of the last batch: every product has completed
    __syncwarp();
    tc::fence_after_sync();
    {
      float* const part = a.partial + static_cast<size_t>(blockIdx.x) * pl.total;
      float* const lo_sm = reinterpret_cast<float*>(DEL);
      const int lr = row & 63;
      const uint32_t tacc = tmem + kColAcc + lane_sel;
      auto readout = [&](float* dst, const float* add) {
        float v[16];
        auto put = [&](int64_t idx, float val) { dst[idx] = add != nullptr ? val + add[idx] : val; };
        auto put16_scaled = [&](int64_t idx) {          // idx % 4 == 0
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o = make_float4(v[i] * wscale, v[i + 1] * wscale, v[i + 2] * wscale, v[i + 3] * wscale);
            if (add != nullptr) { const float4 t = *reinterpret_cast<const float4*>(add + idx + i); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            *reinterpret_cast<float4*>(dst + idx + i) = o;
          }
        };
#pragma unroll
        for (int l = 1; l < L; ++l) {
          tc::tmem_ld16(tacc + kAccW * l + cb, v);
          tc::tmem_wait_ld();
          put16_scaled(pl.offW[l] + lr * 64 + cb);
          if (c == 0) {
            float b8[8];
            tc::tmem_ld8(tacc + kAccW * l + 64, b8);
            tc::tmem_wait_ld();
            put(pl.offb[l] + lr, b8[0]);
          }
        }
        // batch H: lanes 0..31 = dWv0 rows, lane 32 = dWp; column 64 = their biases
        tc::tmem_ld16(tacc + cb, v);
        tc::tmem_wait_ld();
        if (lr < 32) put16_scaled(pl.offWv0 + lr * 64 + cb);
        else if (lr == 32) put16_scaled(pl.offWp + cb);
        if (c == 0) {
          float b8[8];
          tc::tmem_ld8(tacc + 64, b8);
          tc::tmem_wait_ld();
          if (lr < 32) put(pl.offbv0 + lr, b8[0]);
          else if (lr == 32) put(pl.offbp, b8[0]);
        }
        // batch 0T (transposed): lanes 0..7 = input feature i, columns 0..63 = unit j -> dW0[j][i]; lane 56 -> db0, dbv1;
        // lanes 8..39 = av0 unit, columns 64..79 = dz_v1 unit k -> dWv1[k][i]
        tc::tmem_ld16(tacc + kAccW * L + cb, v);
        tc::tmem_wait_ld();
        if (lr < 8) {
#pragma unroll
          for (int i = 0; i < 16; ++i) put(pl.offW[0] + (cb + i) * PINN_N_IN + lr, v[i]);
        } else if (lr == kTOne) {
#pragma unroll
          for (int i = 0; i < 16; ++i) put(pl.offb[0] + cb + i, v[i]);
        }
        if (c == 1) {
          tc::tmem_ld16(tacc + kAccW * L + 64, v);
          tc::tmem_wait_ld();
          if (lr >= kTV0 && lr < kTV0 + 32) {
#pragma unroll
            for (int k = 0; k < 16; ++k) put(pl.offWv1 + k * 32 + lr - kTV0, v[k]);
          } else if (lr == kTOne) {
#pragma unroll
            for (int k = 0; k < 16; ++k) put(pl.offbv1 + k, v[k]);
          }
        }
      };
      if (q >= 2) readout(lo_sm, nullptr);
      asm volatile("bar.sync 5, 512;" ::: "memory");
      if (q < 2) readout(part, lo_sm);
    }
    // dWv2 / dbv2 and the loss sums: xor tree over the warp's 32 rows, then the four quadrants in order
    {
      float w5[5] = {g_wv2[0], g_wv2[1], g_wv2[2], g_wv2[3], g_bv2};
#pragma unroll
      for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w5[k] += __shfl_xor_sync(0xffffffffu, w5[k], o);
        if (lane == 0) wred[warp][k] = w5[k];
      }
      if (c == 0) {
        // the fourth loss sum is the number of valid rows this thread has seen: its row of every tile of the CTA
        double cnt = 0.0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) cnt += (tile * 128 + row < a.n) ? 1.0 : 0.0;
        if (a.grad_u != nullptr) cnt = 0.0;
        double vals[4] = {lacc[0][row], lacc[1][row], lacc[2][row], cnt};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double t = vals[k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) lred[q][k] = t;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 17) {
    float* const part_hi = a.partial + static_cast<size_t>(blockIdx.x) * pl.total;
    const int k = tid;                  // 0..15: dWv2[k] (slice c = k / 4, entry k % 4); 16: dbv2 (slice 0, entry 4)
    const int cc = k < 16 ? k >> 2 : 0, e = k < 16 ? k & 3 : 4;
    const float t = (wred[4 * cc][e] + wred[4 * cc + 1][e]) + (wred[4 * cc + 2][e] + wred[4 * cc + 3][e]);
    if (k < 16) part_hi[pl.offWv2 + k] = t;
    else part_hi[pl.offbv2] = t;
  } else if (tid >= 32 && tid < 36) {
    const int k = tid - 32;
    a.loss_partial[static_cast<size_t>(blockIdx.x) * 4 + k] = (lred[0][k] + lred[1][k]) + (lred[2][k] + lred[3][k]);
  }
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

struct FzPlan { int grid; size_t smem, off_partial, off_images, off_park, bytes; };
static FzPlan plan_fused(int L, int64_t n) {
  FzPlan p{};
  const ParamLayout lay = make_layout(64, L);
  const int sms = sm_count();
  const int64_t tiles = (n + 127) / 128;
  p.grid = static_cast<int>(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
  p.smem = fz_smem_bytes(L);
  size_t off = static_cast<size_t>(p.grid) * 4 * sizeof(double);
  off = (off + 255) & ~static_cast<size_t>(255);
  p.off_partial = off;
  off += static_cast<size_t>(p.grid) * lay.total * sizeof(float);
  off = (off + 255) & ~static_cast<size_t>(255);
  p.off_images = off;
  off += static_cast<size_t>(2 * L) * kFzImgBytes;
  p.off_park = off;
  off += static_cast<size_t>(p.grid) * (L + 1) * 4 * kFzThreads * sizeof(float4);
  p.bytes = off;
  return p;
}

}  // namespace pinn
