// mlp_wide_res.cu -- resident-activation forward / MC-dropout sweep for the 256-wide nets: the reference's own
// `Layers = [8,256,256,256,1]` (01:2139) and config 4's 6x256.  DNN.forward 01:421-438, get_MC_samples 01:1413-1491.
//
// The per-layer GEMM path (mlp_wide_tc.cu) writes every layer's activations to HBM as 8 B/element tf32 hi/lo planes and
// reads them back in the next launch: ~3 000x the algorithmic bytes of a sweep.  Here ONE persistent CTA per SM owns a
// 128-sample tile for ALL passes of the sweep and its activations never leave the SM:
//
//   split    : fp16 pairs instead of tf32 pairs.  a = a_h + a_l with a_h = fp16(a), a_l = fp16(a - a_h) carries 22
//              significant bits for |a| <= 1/(1-p) (tanh outputs), w likewise; the three products
//              a_l*w_h + a_h*w_l + a_h*w_h run as `tcgen05.mma.kind::f16` (twice the tf32 rate) with fp32 accumulation.
//              4 B per element instead of 8: a tile's [128 x 256] activations are 128 KB of shared memory -- resident.
//   A planes : shared memory, K-major no-swizzle UMMA layout, byte(row, k) = (k/8) * 2048 + row * 16 + (k%8) * 2;
//              a K = 16 slab of both planes is 2 x 4 KB.  Written by the epilogue, read only by the tensor core.
//   weights  : pre-split fp16 images in global memory (L2-resident, 256 KB per hidden layer), one contiguous
//              [hi | lo] block per K = 16 slab, streamed by a producer warp with `cp.async.bulk` (1-D TMA, mbarrier
//              transaction bytes) through a 5 x 16 KB ring.  Dropout scale 1/(1-p) folded in.
//   accum    : two 256-column fp32 accumulators in tensor memory (all 512 columns), alternating by MMA phase: the
//              epilogue of phase p reads one while the products of phase p+1 fill the other.
//   epilogue : 16 warps, thread = (row, 64-column quarter).  Per 16 columns: tcgen05.ld, bias + tanh, Philox keep-select,
//              fp16 split, four 16-byte shared-memory stores, fence.proxy.async, arrive on the slab's mbarrier -- the MMA
//              warp issues the next layer's three products of that K slab as soon as the slab and its weights are in, so
//              the tensor pipe works under the epilogue that feeds it.
//   layer 0  : K = 8, pass-invariant: computed once per tile on the CUDA cores and parked in a per-CTA
//              128 KB global scratch (each thread re-reads only what it wrote itself; stays in L2), re-masked per pass.
//   heads    : [Wv0; Wp; 0] as one N = 144 product, Wv1 (128 -> 64) as an N = 64 product, the last 64-wide dot, the
//              log-variance and the Welford update on the CUDA cores; the statistics live in registers across passes.
//
// MMA phases of a pass: hidden layers 1..L-1, heads, variance layer 1.  HBM traffic of a sweep = x in, three result
// vectors out.  The mask stream (Philox counters per (sample, pass, layer, unit / 8)) is the one every other path uses.
#include <cuda_fp16.h>
#include "net.cuh"
#include "tc.cuh"
#include "tc_api.cuh"

namespace pinn {

constexpr int kRH = 256;                    // width
constexpr int kRT = 128;                    // rows per tile
constexpr int kRComputeWarps = 16;
constexpr int kRThreads = (kRComputeWarps + 2) * 32;   // + producer warp + MMA warp
#ifndef RES_STAGES
#define RES_STAGES 5
#endif
constexpr int kRStages = RES_STAGES;
constexpr int kRStageBytes = 64 * kRH;      // one K = 16 slab of a 256-row matrix: [hi 8 KB | lo 8 KB]
constexpr int kRPlane = kRT * kRH * 2;      // one fp16 plane of a tile's activations (64 KB)
constexpr int kRSlabA = kRT * 32;           // bytes of one K = 16 slab of one A plane (4 KB)
constexpr int kRNH = kRH / 2 + 16;          // heads product: 128 variance-head rows + mean row + 15 zero rows
constexpr int kRNV = kRH / 4;               // variance layer 1: 64 rows
// A/B switches of the epilogue (profiles/build_variant.py -DRES_...=0/1; measured numbers in DESIGN.md)
#ifndef RES_PREDRAW
#define RES_PREDRAW 0      // draw a phase's Philox blocks ahead of the wait that precedes its epilogue
#endif
#ifndef RES_UNROLLJ
#define RES_UNROLLJ 1      // unroll the four 16-column groups of a hidden epilogue
#endif
#ifndef RES_WARP_ARRIVE
#define RES_WARP_ARRIVE 1  // one mbarrier arrival per warp and K slab (after __syncwarp) instead of one per thread
#endif
#ifndef RES_LDPF
#define RES_LDPF 1         // tcgen05.ld of the next 16 columns in flight while the current 16 are processed
#endif
constexpr bool kPredraw = RES_PREDRAW != 0, kLdPrefetch = RES_LDPF != 0 && RES_UNROLLJ != 0;
#if RES_UNROLLJ
#define RES_J_UNROLL _Pragma("unroll")
#else
#define RES_J_UNROLL _Pragma("unroll 1")
#endif
static_assert(!kPredraw || RES_UNROLLJ, "pre-drawn blocks are indexed by the group: needs the unrolled form");

PINN_HD constexpr int res_slab_bytes(int N) { return 64 * N; }                              // [hi | lo] of one slab
PINN_HD constexpr size_t res_img_bytes(int N, int K) { return static_cast<size_t>(K / 16) * res_slab_bytes(N); }

struct ResPlan {
  size_t off_w[PINN_MAX_HIDDEN], off_wh, off_wv1, off_a0, bytes;
  int grid;
};
static ResPlan res_plan(int L, int64_t n) {
  ResPlan p{};
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += (b + 255) & ~static_cast<size_t>(255); return r; };
  for (int l = 1; l < L; ++l) p.off_w[l] = take(res_img_bytes(kRH, kRH));
  p.off_wh = take(res_img_bytes(kRNH, kRH));
  p.off_wv1 = take(res_img_bytes(kRNV, kRH / 2));
  const int64_t tiles = (n + kRT - 1) / kRT;
  const int64_t items = tiles * 8;                 // up to 8 pass chunks per tile
  p.grid = static_cast<int>(items < sm_count() ? (items > 0 ? items : 1) : sm_count());
  p.off_a0 = take(static_cast<size_t>(p.grid) * kRT * kRH * sizeof(float));
  p.bytes = o;
  return p;
}
// Pass chunks of a sweep on this path: runs of ~12 passes (a pass of a tile is ~20 us; the per-item overhead is one layer-0
// evaluation), at most 8 -- a function of T ALONE (bitwise shard invariance, as in mlp_tc.cu).
int res_pass_chunks(int T) {
  const int c = (T + 6) / 12;
  return c < 1 ? 1 : (c > 8 ? 8 : c);
}
size_t wide_res_workspace_bytes(int L, int64_t n) {
  return n > 0 ? res_plan(L, n).bytes + static_cast<size_t>(8) * 3 * static_cast<size_t>(n) * sizeof(float) : 0;
}

// two fp32 -> packed fp16 pair (hi) and the packed fp16 pair of the remainders (lo)
PINN_D void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ------------------------------------------------------------------ weight images (once per call)
// [N x K] matrix whose first `rows_a` rows come from `src_a` ([rows_a][K]), row `rows_a` from `src_b` (or zero), the rest
// zero; times c; as per-slab [hi | lo] fp16 blocks: byte(n, k) = (k/16) * 64 N + (k%16 / 8) * 16 N + n * 16 + (k%8) * 2.
__global__ void wide_res_split_kernel(const float* __restrict__ src_a, int rows_a, const float* __restrict__ src_b, int N, int K, float c,
                                      unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // (row, k8)
  if (idx >= N * (K / 8)) return;
  const int nrow = idx % N, k8 = idx / N;
  float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
  if (nrow < rows_a) {
    const float4* p = reinterpret_cast<const float4*>(src_a + static_cast<size_t>(nrow) * K) + 2 * k8;
    v0 = __ldg(p); v1 = __ldg(p + 1);
  } else if (nrow == rows_a && src_b != nullptr) {
    const float4* p = reinterpret_cast<const float4*>(src_b) + 2 * k8;
    v0 = __ldg(p); v1 = __ldg(p + 1);
  }
  uint4 h, l;
  split_h2(v0.x * c, v0.y * c, h.x, l.x); split_h2(v0.z * c, v0.w * c, h.y, l.y);
  split_h2(v1.x * c, v1.y * c, h.z, l.z); split_h2(v1.z * c, v1.w * c, h.w, l.w);
  unsigned char* p = dst + static_cast<size_t>(k8 >> 1) * res_slab_bytes(N) + (k8 & 1) * (N * 16) + nrow * 16;
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + N * 32) = l;
}

PINN_HD constexpr uint32_t make_idesc_f16(int M, int N) {      // D = F32, A = B = F16, both K-major
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
PINN_D void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
PINN_D float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
PINN_D void st_cg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

struct ResArgs {
  const unsigned char* img_w[PINN_MAX_HIDDEN];   // hidden layers 1..L-1
  const unsigned char* img_h;                    // heads
  const unsigned char* img_v1;                   // variance layer 1
  float* a0;                                     // [grid][64 column quads][128 rows][4]
  int L, T, mc, do_eval;
  int chunks;                                    // pass chunks per tile (a function of T alone), work item = (tile, chunk)
  float* part;                                   // chunks > 1: per-chunk Welford triples [chunk][3][n], folded by mc_merge
  float inact;                                   // multiplier of an un-masked activation (undoes the folded scale)
  int no_logvar;
};

// slab issued i-th in a phase whose K slabs were produced by the four column quarters, `spq` slabs each: the quarters
// work in parallel, so their j-th slabs become ready together
PINN_D int res_slab_order(int i, int spq) { return (i & 3) * spq + (i >> 2); }

__global__ void __launch_bounds__(kRThreads, 1)
wide_res_kernel(const __grid_constant__ pinn_net_t net, const float* __restrict__ x, int64_t n, const __grid_constant__ DropParams dp,
                const __grid_constant__ ResArgs a, TcOut out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kRStages], empty[kRStages], ready[16], done;
  __shared__ uint32_t tmem_base_s;
  unsigned char* const ring = smem + 2 * kRPlane;
  float* const fsm = reinterpret_cast<float*>(ring + kRStages * kRStageBytes);
  // float area: b[l] (l = 1..L-1, 256 each, pre-scaled by kTanhArg) | bv0 (128, pre-scaled) | bv1 (64, pre-scaled) | Wv2 (64) | part (4 x 128)
  const int L = a.L;
  float* const s_b = fsm;
  float* const s_bv0 = fsm + (PINN_MAX_HIDDEN - 1) * kRH;
  float* const s_bv1 = s_bv0 + kRH / 2;
  float* const s_wv2 = s_bv1 + kRNV;
  float* const s_part = s_wv2 + kRNV;

  const int tid = threadIdx.x, warp = tc::uniform_warp_idx(), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kRStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int c = 0; c < 16; ++c) tc::mbar_init(&ready[c], RES_WARP_ARRIVE ? 4 : 128);
    tc::mbar_init(&done, 1);
    tc::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  for (int l = 1; l < L; ++l)
    for (int i = tid; i < kRH; i += blockDim.x) s_b[(l - 1) * kRH + i] = __ldg(net.b[l] + i) * kTanhArg;
  for (int i = tid; i < kRH / 2; i += blockDim.x) s_bv0[i] = __ldg(net.bv0 + i) * kTanhArg;
  for (int i = tid; i < kRNV; i += blockDim.x) { s_bv1[i] = __ldg(net.bv1 + i) * kTanhArg; s_wv2[i] = __ldg(net.Wv2 + i); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s;

  const int64_t n_tiles = (n + kRT - 1) / kRT;
  // Work item = (tile, pass chunk): a sweep is cut into `chunks` runs of Tc consecutive passes (chosen from T alone, so a
  // sample's arithmetic does not depend on the batch or its sharding); every run keeps its own Welford triple and the
  // merge launch folds them in chunk order.  At the reference's own size (N = 20 000: 157 tiles on 148 SMs) this is what
  // fills the machine.
  const int C = a.chunks, Tc = (a.T + C - 1) / C;
  const int64_t n_items = n_tiles * C;
  auto item_passes = [&](int chunk) {            // dropout passes of the chunk (+ the eval pass, which rides with chunk 0)
    if (!a.mc) return 1;
    const int t0 = chunk * Tc, cnt = a.T - t0 < Tc ? a.T - t0 : Tc;
    return (cnt > 0 ? cnt : 0) + ((a.do_eval && chunk == 0) ? 1 : 0);
  };
  const int n_phase = L + 1;                                   // MMA phases per pass: L-1 hidden, heads, variance layer 1
  // phase ph of a pass: image, rows N, K slabs, slabs per producing quarter
  auto phase_img = [&](int ph) { return ph < L - 1 ? a.img_w[ph + 1] : (ph == L - 1 ? a.img_h : a.img_v1); };
  auto phase_N = [&](int ph) { return ph < L - 1 ? kRH : (ph == L - 1 ? kRNH : kRNV); };
  auto phase_slabs = [&](int ph) { return ph <= L - 1 ? kRH / 16 : (kRH / 2) / 16; };

  if (warp == kRComputeWarps) {
    // ================================================================== producer: weight slabs into the ring
    if (tc::elect_one()) {
      uint32_t cnt = 0;
      for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x)
        for (int pi = 0, n_pass = item_passes(static_cast<int>(item % C)); pi < n_pass; ++pi)
          for (int ph = 0; ph < n_phase; ++ph) {
            const unsigned char* img = phase_img(ph);
            const int N = phase_N(ph), ns = phase_slabs(ph), spq = ns / 4;
            const uint32_t bytes = static_cast<uint32_t>(res_slab_bytes(N));
            for (int i = 0; i < ns; ++i, ++cnt) {
              const uint32_t s = cnt % kRStages;
              if (cnt >= kRStages) tc::mbar_wait(&empty[s], ((cnt / kRStages) - 1u) & 1u);
              tc::mbar_expect_tx(&full[s], bytes);
              tc::bulk_g2s(ring + s * kRStageBytes, img + static_cast<size_t>(res_slab_order(i, spq)) * bytes, bytes, &full[s]);
            }
          }
    }
    __syncwarp();
  } else if (warp == kRComputeWarps + 1) {
    // ================================================================== MMA issuer
    uint32_t cnt = 0, rpar = 0u, acc_sel = 0u;
    const uint32_t a_hi_s = tc::smem_u32(smem), a_lo_s = a_hi_s + kRPlane, ring_s = tc::smem_u32(ring);
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x)
      for (int pi = 0, n_pass = item_passes(static_cast<int>(item % C)); pi < n_pass; ++pi)
        for (int ph = 0; ph < n_phase; ++ph, acc_sel ^= 1u) {
          const int N = phase_N(ph), ns = phase_slabs(ph), spq = ns / 4;
          const uint32_t idesc = make_idesc_f16(kRT, N), d_t = tb + acc_sel * 256u;
          const uint32_t lbo_b = static_cast<uint32_t>(N) * 16u;
#pragma unroll 1
          for (int i = 0; i < ns; ++i, ++cnt) {
            const uint32_t s = cnt % kRStages;
            const int slab = res_slab_order(i, spq);
            tc::mbar_wait(&ready[slab], (rpar >> slab) & 1u);
            rpar ^= 1u << slab;
            tc::mbar_wait(&full[s], (cnt / kRStages) & 1u);
            __syncwarp();
            if (tc::elect_one()) {
              tc::fence_after_sync();
              const uint64_t ah = tc::make_desc(a_hi_s + slab * kRSlabA, kRT * 16, 128), al = tc::make_desc(a_lo_s + slab * kRSlabA, kRT * 16, 128);
              const uint32_t sw = ring_s + s * kRStageBytes;
              const uint64_t bh = tc::make_desc(sw, lbo_b, 128), bl = tc::make_desc(sw + static_cast<uint32_t>(N) * 32u, lbo_b, 128);
              umma_f16(d_t, al, bh, idesc, i != 0 ? 1u : 0u);      // small terms first: lo*hi, hi*lo, then hi*hi
              umma_f16(d_t, ah, bl, idesc, 1u);
              umma_f16(d_t, ah, bh, idesc, 1u);
              tc::umma_commit(&empty[s]);
              if (i == ns - 1) tc::umma_commit(&done);
            }
            __syncwarp();
          }
        }
  } else {
    // ================================================================== compute warps: thread = (row, column quarter)
    const int q = warp >> 2, r = (warp & 3) * 32 + lane;
    const uint32_t tlane = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    unsigned char* const my_a = smem + r * 16;
    float* const my_a0 = a.a0 + static_cast<size_t>(blockIdx.x) * kRT * kRH + static_cast<size_t>(r) * 4;
    const bool drop_on = dp.p > 0.f, inj = dp.masks != nullptr;
    const int Dm = L * kRH + kRH / 2;
    uint32_t dpar = 0u, acc_sel = 0u;

    // 16 masked activations -> both planes of A slab `slab`, then hand the slab to the MMA warp
    auto emit_slab = [&](const float (&v)[16], int slab) {
      uint4 h0, h1, l0, l1;
      split_h2(v[0], v[1], h0.x, l0.x);   split_h2(v[2], v[3], h0.y, l0.y);
      split_h2(v[4], v[5], h0.z, l0.z);   split_h2(v[6], v[7], h0.w, l0.w);
      split_h2(v[8], v[9], h1.x, l1.x);   split_h2(v[10], v[11], h1.y, l1.y);
      split_h2(v[12], v[13], h1.z, l1.z); split_h2(v[14], v[15], h1.w, l1.w);
      unsigned char* p = my_a + slab * kRSlabA;
      *reinterpret_cast<uint4*>(p) = h0;
      *reinterpret_cast<uint4*>(p + kRT * 16) = h1;
      *reinterpret_cast<uint4*>(p + kRPlane) = l0;
      *reinterpret_cast<uint4*>(p + kRPlane + kRT * 16) = l1;
      tc::fence_proxy_async();
#if RES_WARP_ARRIVE
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&ready[slab]);
#else
      tc::mbar_arrive(&ready[slab]);
#endif
    };
    auto wait_done = [&]() {
      tc::mbar_wait(&done, dpar);
      dpar ^= 1u;
      __syncwarp();
      tc::fence_after_sync();
    };

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int64_t tile = item / C;
      const int chunk = static_cast<int>(item % C), t0 = chunk * Tc, n_pass = item_passes(chunk);
      const bool eval_item = a.do_eval && chunk == 0;
      if (n_pass == 0) continue;
      const int64_t s = tile * kRT + r;
      const bool valid = s < n;
      const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
      const uint32_t s_lo = static_cast<uint32_t>(sg), s_hi = static_cast<uint32_t>(sg >> 32);
      // keep-select 16 activations of units [j0, j0 + 16) of dropout layer `layer`
      auto select16 = [&](float (&v)[16], bool active, const uint4& r0, const uint4& r1, uint32_t pass, int tloc, uint32_t layer, uint32_t j0) {
        if (active) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            bool k[8];
            if (inj) {
              const uint8_t* mrow = dp.masks + (static_cast<size_t>(tloc) * dp.mask_n + s) * Dm + layer * kRH + j0 + 8 * g;
              const uint2 mb = *reinterpret_cast<const uint2*>(mrow);
#pragma unroll
              for (int e = 0; e < 4; ++e) { k[e] = ((mb.x >> (8 * e)) & 0xffu) != 0; k[4 + e] = ((mb.y >> (8 * e)) & 0xffu) != 0; }
            } else {
              if constexpr (kPredraw) keep8_from(g == 0 ? r0 : r1, dp.thresh_hi, k);
              else keep8_from(Philox::gen_rk(dp.rk, s_lo, s_hi, pass, (layer << 16) | ((j0 >> 3) + g)), dp.thresh_hi, k);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[8 * g + e] = k[e] ? v[8 * g + e] : 0.f;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] *= a.inact;
        }
      };

      // ---- layer 0 (pass-invariant): this thread's 64 columns -> the CTA's scratch
      {
        float xr[PINN_N_IN];
        if (valid) {
          const float4* px = reinterpret_cast<const float4*>(x + s * PINN_N_IN);
          const float4 q0 = __ldg(px), q1 = __ldg(px + 1);
          xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
        } else {
#pragma unroll
          for (int i = 0; i < PINN_N_IN; ++i) xr[i] = 0.f;
        }
#pragma unroll 1
        for (int c4 = 16 * q; c4 < 16 * q + 16; ++c4) {
          float o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * c4 + e;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(net.W[0] + j * PINN_N_IN));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(net.W[0] + j * PINN_N_IN) + 1);
            float z = __ldg(net.b[0] + j);
            z = fmaf(w0.x, xr[0], z); z = fmaf(w0.y, xr[1], z); z = fmaf(w0.z, xr[2], z); z = fmaf(w0.w, xr[3], z);
            z = fmaf(w1.x, xr[4], z); z = fmaf(w1.y, xr[5], z); z = fmaf(w1.z, xr[6], z); z = fmaf(w1.w, xr[7], z);
            o4[e] = tanh_pre(z * kTanhArg);
          }
          st_cg4(my_a0 + static_cast<size_t>(c4) * (kRT * 4), make_float4(o4[0], o4[1], o4[2], o4[3]));
        }
      }

      // Philox blocks of the coming epilogue are drawn AHEAD of the wait that precedes it (they depend on nothing the tensor
      // core produces): the generator is a third of an epilogue's instructions and the wait is otherwise idle -- the last
      // four K slabs' products (~1 500 clk) cannot start before the previous epilogue's last stores.
      uint4 rk[8] = {};
      auto draw = [&](int nblk, uint32_t pass, uint32_t layer, uint32_t j0) {
#pragma unroll
        for (int b = 0; b < 8; ++b)
          if (b < nblk) rk[b] = Philox::gen_rk(dp.rk, s_lo, s_hi, pass, (layer << 16) | ((j0 >> 3) + b));
#pragma unroll
        for (int b = 0; b < 8; ++b)
          if (b < nblk) asm volatile("" : "+r"(rk[b].x), "+r"(rk[b].y), "+r"(rk[b].z), "+r"(rk[b].w));   // pin before the wait
      };
      const bool drawn = kPredraw && drop_on && !inj;           // masks come from Philox, drawn ahead
      if (drawn && !eval_item) draw(8, static_cast<uint32_t>(dp.pass_offset + t0), 0u, static_cast<uint32_t>(64 * q));

      float mean = 0.f, m2 = 0.f, slv = 0.f;
#pragma unroll 1
      for (int pi = 0; pi < n_pass; ++pi) {
        const bool eval_pass = a.mc && eval_item && pi == 0;
        const int tl = a.mc ? (eval_item ? pi - 1 : pi) : 0;                 // pass index inside the chunk (Welford count)
        const int t = t0 + tl;                                               // pass index of the sweep (mask stream)
        const bool active = drop_on && !eval_pass && (!inj || valid);
        const uint32_t pass = static_cast<uint32_t>(dp.pass_offset + t);
        // ---- stage the masked layer-0 activations as the first A operand
        {
          float4 cur[4], nxt[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) cur[e] = ld_cg4(my_a0 + static_cast<size_t>(16 * q + e) * (kRT * 4));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < 3) {
#pragma unroll
              for (int e = 0; e < 4; ++e) nxt[e] = ld_cg4(my_a0 + static_cast<size_t>(16 * q + 4 * (j + 1) + e) * (kRT * 4));
            }
            float v[16] = {cur[0].x, cur[0].y, cur[0].z, cur[0].w, cur[1].x, cur[1].y, cur[1].z, cur[1].w,
                           cur[2].x, cur[2].y, cur[2].z, cur[2].w, cur[3].x, cur[3].y, cur[3].z, cur[3].w};
            select16(v, active, rk[(2 * j) & 7], rk[(2 * j + 1) & 7], pass, t, 0u, static_cast<uint32_t>(64 * q + 16 * j));
            emit_slab(v, 4 * q + j);
#pragma unroll
            for (int e = 0; e < 4; ++e) cur[e] = nxt[e];
          }
        }
        // ---- hidden layers
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
          if (drawn && active) draw(8, pass, static_cast<uint32_t>(l), static_cast<uint32_t>(64 * q));
          wait_done();
          const uint32_t acc = tlane + acc_sel * 256u + static_cast<uint32_t>(64 * q);
          acc_sel ^= 1u;
          const float* bl = s_b + (l - 1) * kRH + 64 * q;
          float zb[2][16];
          if constexpr (kLdPrefetch) tc::tmem_ld16(acc, zb[0]);
          RES_J_UNROLL
          for (int j = 0; j < 4; ++j) {
            float v[16];
            float* z = zb[kLdPrefetch ? (j & 1) : 0];
            if constexpr (kLdPrefetch) {
              tc::tmem_wait_ld();
              if (j < 3) tc::tmem_ld16(acc + 16u * (j + 1), zb[(j + 1) & 1]);
            } else {
              tc::tmem_ld16(acc + 16u * j, z);
              tc::tmem_wait_ld();
            }
#pragma unroll
            for (int g = 0; g < 16; g += 8) {
              const float4 bA = *reinterpret_cast<const float4*>(bl + 16 * j + g), bB = *reinterpret_cast<const float4*>(bl + 16 * j + g + 4);
              const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
              float t8[8];
              tanh8_prescaled(z + g, bb, t8);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[g + e] = t8[e];
            }
            select16(v, active, rk[(2 * j) & 7], rk[(2 * j + 1) & 7], pass, t, static_cast<uint32_t>(l), static_cast<uint32_t>(64 * q + 16 * j));
            emit_slab(v, 4 * q + j);
          }
        }
        // ---- heads: 128 variance-head units (32 per quarter) + the mean head (column 128)
        float u = 0.f;
        {
          if (drawn && active) draw(4, pass, static_cast<uint32_t>(L), static_cast<uint32_t>(32 * q));
          wait_done();
          const uint32_t acc = tlane + acc_sel * 256u;
          acc_sel ^= 1u;
          if (q == 0) {
            float zz[4];
            tc::tmem_ld4(acc + static_cast<uint32_t>(kRH / 2), zz);
            tc::tmem_wait_ld();
            u = zz[0] + __ldg(net.bp);
          }
          RES_J_UNROLL
          for (int j = 0; j < 2; ++j) {
            const int c0 = 32 * q + 16 * j;
            float z[16], v[16];
            tc::tmem_ld16(acc + static_cast<uint32_t>(c0), z);
            tc::tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < 16; g += 8) {
              const float4 bA = *reinterpret_cast<const float4*>(s_bv0 + c0 + g), bB = *reinterpret_cast<const float4*>(s_bv0 + c0 + g + 4);
              const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
              float t8[8];
              tanh8_prescaled(z + g, bb, t8);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[g + e] = t8[e];
            }
            select16(v, active, rk[(2 * j) & 7], rk[(2 * j + 1) & 7], pass, t, static_cast<uint32_t>(L), static_cast<uint32_t>(c0));
            emit_slab(v, 2 * q + j);
          }
        }
        // ---- variance layer 1 (64 units, 16 per quarter) + the last dot; thread (row, 0) finishes the sample
        {
          if (drawn && pi + 1 < n_pass) draw(8, static_cast<uint32_t>(dp.pass_offset + t + 1), 0u, static_cast<uint32_t>(64 * q));   // next pass's layer-0 masks
          wait_done();
          const uint32_t acc = tlane + acc_sel * 256u + static_cast<uint32_t>(16 * q);
          acc_sel ^= 1u;
          float z[16];
          tc::tmem_ld16(acc, z);
          tc::tmem_wait_ld();
          float part = 0.f;
#pragma unroll
          for (int g = 0; g < 16; g += 8) {
            const float4 bA = *reinterpret_cast<const float4*>(s_bv1 + 16 * q + g), bB = *reinterpret_cast<const float4*>(s_bv1 + 16 * q + g + 4);
            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
            float t8[8];
            tanh8_prescaled(z + g, bb, t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) part = fmaf(s_wv2[16 * q + g + e], t8[e], part);
          }
          if (q != 0) {
            s_part[q * kRT + r] = part;
            __threadfence_block();
            asm volatile("bar.arrive %0, %1;" ::"r"(1 + (warp & 3)), "r"(128) : "memory");
          } else {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp & 3)), "r"(128) : "memory");
            const float vraw = __ldg(net.bv2) + ((part + s_part[kRT + r]) + (s_part[2 * kRT + r] + s_part[3 * kRT + r]));
            const float lv = logvar_out(vraw, a.no_logvar != 0);
            if (!a.mc) {
              if (valid) { out.u[s] = u; out.s[s] = lv; }
            } else if (eval_pass) {
              if (valid) out.pred_mean[s] = u;
            } else {
              const float d = u - mean;
              mean += d / static_cast<float>(tl + 1);
              m2 = fmaf(d, u - mean, m2);
              slv += lv;
            }
          }
        }
      }
      if (a.mc && valid && q == 0 && C > 1) {
        float* pp = a.part + static_cast<size_t>(chunk) * 3 * n + s;
        pp[0] = mean; pp[n] = m2; pp[2 * n] = slv;
      } else if (a.mc && valid && q == 0) {
        if (out.raw_mean) out.raw_mean[s] = mean;
        if (out.raw_m2) out.raw_m2[s] = m2;
        if (out.raw_slv) out.raw_slv[s] = slv;
        const float invT = 1.0f / static_cast<float>(a.T > 0 ? a.T : 1);
        if (out.a_u) out.a_u[s] = sqrtf(expf(slv * invT));
        if (out.e_u) out.e_u[s] = sqrtf(fmaxf(m2, 0.f) * invT);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

constexpr size_t kRSmemBytes = static_cast<size_t>(2) * kRPlane + static_cast<size_t>(kRStages) * kRStageBytes +
                               (static_cast<size_t>(PINN_MAX_HIDDEN - 1) * kRH + kRH / 2 + 2 * kRNV + 4 * kRT) * sizeof(float);

// 1: handled; 0: shape not covered (the per-layer GEMM path takes it); -1: error in *err.
int launch_wide_res(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
                    void* workspace, size_t workspace_bytes, cudaStream_t st, int* err) {
  *err = 0;
  if ((net->flags & (PINN_NET_NO_WIDE_TC | PINN_NET_NO_WIDE_RESIDENT)) || net->width != kRH || net->n_hidden < 1 || n <= 0) return 0;
  if (mc && T <= 0) return 0;
  for (int l = 0; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return 0;
  if (!aligned16(net->Wv0) || !aligned16(net->Wp) || !aligned16(net->Wv1) || !aligned16(x)) return 0;
  if (dp.masks != nullptr && (((net->n_hidden * kRH + kRH / 2) & 7) != 0 || (reinterpret_cast<uintptr_t>(dp.masks) & 7u) != 0)) return 0;
  const int L = net->n_hidden;
  const ResPlan p = res_plan(L, n);
  if (!workspace || workspace_bytes < p.bytes) { *err = PINN_E_WORKSPACE; return -1; }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const float wscale = dp.p > 0.f ? dp.scale : 1.0f;
  auto split = [&](const float* sa, int rows_a, const float* sb, int N, int K, unsigned char* dst) {
    const int items = N * (K / 8);
    wide_res_split_kernel<<<(items + 255) / 256, 256, 0, st>>>(sa, rows_a, sb, N, K, wscale, dst);
  };
  ResArgs a{};
  for (int l = 1; l < L; ++l) { split(net->W[l], kRH, nullptr, kRH, kRH, ws + p.off_w[l]); a.img_w[l] = ws + p.off_w[l]; }
  split(net->Wv0, kRH / 2, net->Wp, kRNH, kRH, ws + p.off_wh);
  split(net->Wv1, kRNV, nullptr, kRNV, kRH / 2, ws + p.off_wv1);
  a.img_h = ws + p.off_wh; a.img_v1 = ws + p.off_wv1;
  a.a0 = reinterpret_cast<float*>(ws + p.off_a0);
  const int C = mc ? res_pass_chunks(T) : 1;
  const size_t part_bytes = C > 1 ? static_cast<size_t>(C) * 3 * n * sizeof(float) : 0;
  if (workspace_bytes < p.bytes + part_bytes) { *err = PINN_E_WORKSPACE; return -1; }
  a.chunks = C;
  a.part = reinterpret_cast<float*>(ws + p.bytes);
  a.L = L; a.T = T; a.mc = mc ? 1 : 0; a.do_eval = (mc && out.pred_mean != nullptr) ? 1 : 0;
  a.inact = dp.p > 0.f ? dp.keep : 1.0f;
  a.no_logvar = (net->flags & PINN_NET_NO_LOGVAR) ? 1 : 0;
  cudaError_t e = cudaFuncSetAttribute(wide_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kRSmemBytes));
  if (e != cudaSuccess) { *err = static_cast<int>(e); return -1; }
  const int64_t items = ((n + kRT - 1) / kRT) * C;
  const int grid = static_cast<int>(items < sm_count() ? items : sm_count());
  wide_res_kernel<<<grid, kRThreads, kRSmemBytes, st>>>(*net, x, n, dp, a, out);
  if (C > 1) launch_mc_merge(a.part, n, T, C, out, st);
  *err = static_cast<int>(cudaGetLastError());
  return *err == 0 ? 1 : -1;
}

}  // namespace pinn
