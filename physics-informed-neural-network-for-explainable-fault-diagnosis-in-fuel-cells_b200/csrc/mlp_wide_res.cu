// mlp_wide_res.cu -- resident-activation forward / MC-dropout sweep for the 256-wide nets: the reference's own
// `Layers = [8,256,256,256,1]` (01:2139) and config 4's 6x256.  DNN.forward 01:421-438, get_MC_samples 01:1413-1491.
//
// The per-layer GEMM path (mlp_wide_tc.cu) writes every layer's activations to HBM as 8 B/element tf32 hi/lo planes and
// reads them back in the next launch: ~3 000x the algorithmic bytes of a sweep.  Here ONE persistent CTA per SM owns a
// 128-sample tile for ALL passes of a work item and its activations never leave the SM:
//
//   split    : fp16 pairs instead of tf32 pairs.  a = a_h + a_l with a_h = fp16(a), a_l = fp16(a - a_h) carries 22
//              significant bits for |a| <= 1/(1-p) (tanh outputs), w likewise; the three products
//              a_l*w_h + a_h*w_l + a_h*w_h run as `tcgen05.mma.kind::f16` (twice the tf32 rate) with fp32 accumulation.
//   A planes : TENSOR MEMORY, packed fp16 pairs written by the epilogue with tcgen05.st (lane = sample row, column c of a
//              plane = features 2c, 2c+1): the tensor core fetches only B from shared memory.
//   weights  : pre-split fp16 images in global memory (L2-resident, 256 KB per hidden layer), one contiguous
//              [hi | lo] block per K = 16 slab (K-major no-swizzle UMMA layout), streamed by a producer warp with
//              `cp.async.bulk` (1-D TMA, mbarrier transaction bytes) through a 5 x 16 KB ring.  Dropout scale folded in.
//   tensor memory : [0, 256) the fp32 accumulator | [256, 384) A hi plane | [384, 512) A lo plane.
//   epilogue : 16 warps, thread = (row, 64-column quarter).  It first pulls all its accumulator columns into registers
//              and releases the accumulator (`accfree`); then per 16 columns: bias + tanh, Philox keep-select, fp16
//              split, two tcgen05.st, arrive on the K slab's mbarrier -- the MMA warp issues the next layer's three
//              products of that slab as soon as the slab and its weights are in, so the tensor pipe works under the
//              epilogue that feeds it.
//   layer 0  : K = 8, pass-invariant: computed once per work item on the CUDA cores and parked in shared
//              memory (128 KB fp32), re-masked per pass.
//   heads    : [Wv0; Wp; 0] as one N = 144 product, Wv1 (128 -> 64) as an N = 64 product, the last 64-wide dot, the
//              log-variance and the Welford update on the CUDA cores; the statistics live in registers across passes.
//
// MMA phases of a pass: hidden layers 1..L-1, heads, variance layer 1.  HBM traffic of a sweep = x in, three result
// vectors out.  The mask stream (Philox counters per (sample, pass, layer, unit / 8)) is the one every other path uses.
//
// What bounds it (profiles/r2_wide_res_ab*.log): the tensor pipe.  The three products execute ~950 TFLOP/s
// (58 % of the measured dense bf16 peak) at N = 262 144; a first form with the A planes in shared memory was 2 % slower,
// and pre-drawn Philox blocks, a prefetched tcgen05.ld, per-warp arrivals or a deeper ring moved nothing -- what is left
// is the tensor pipe idling at the start of each epilogue until the first K slab of the next layer is ready.
#include <string.h>
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
constexpr int kRNH = kRH / 2 + 16;          // heads product: 128 variance-head rows + mean row + 15 zero rows
constexpr int kRNV = kRH / 4;               // variance layer 1: 64 rows

PINN_HD constexpr int res_slab_bytes(int N) { return 64 * N; }                              // [hi | lo] of one slab
PINN_HD constexpr size_t res_img_bytes(int N, int K) { return static_cast<size_t>(K / 16) * res_slab_bytes(N); }

struct ResPlan {
  size_t off_w[PINN_MAX_HIDDEN], off_wh, off_wv1, bytes;
};
static ResPlan res_plan(int L) {
  ResPlan p{};
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += (b + 255) & ~static_cast<size_t>(255); return r; };
  for (int l = 1; l < L; ++l) p.off_w[l] = take(res_img_bytes(kRH, kRH));
  p.off_wh = take(res_img_bytes(kRNH, kRH));
  p.off_wv1 = take(res_img_bytes(kRNV, kRH / 2));
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
  return n > 0 ? res_plan(L).bytes + static_cast<size_t>(8) * 3 * static_cast<size_t>(n) * sizeof(float) : 0;
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
  tc::split_h2(v0.x * c, v0.y * c, h.x, l.x); tc::split_h2(v0.z * c, v0.w * c, h.y, l.y);
  tc::split_h2(v1.x * c, v1.y * c, h.z, l.z); tc::split_h2(v1.z * c, v1.w * c, h.w, l.w);
  unsigned char* p = dst + static_cast<size_t>(k8 >> 1) * res_slab_bytes(N) + (k8 & 1) * (N * 16) + nrow * 16;
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + N * 32) = l;
}

struct ResArgs {
  const unsigned char* img_w[PINN_MAX_HIDDEN];   // hidden layers 1..L-1
  const unsigned char* img_h;                    // heads
  const unsigned char* img_v1;                   // variance layer 1
  int L, T, mc, do_eval;
  int chunks;                                    // pass chunks per tile (a function of T alone), work item = (tile, chunk)
  float* part;                                   // chunks > 1: per-chunk Welford triples [chunk][3][n], folded by mc_merge
  float inact;                                   // multiplier of an un-masked activation (undoes the folded scale)
  int no_logvar;
};

// slab issued i-th in a phase whose K slabs were produced by the four column quarters, `spq` slabs each: the quarters
// work in parallel, so their j-th slabs become ready together
// K slab issued i-th in a phase: the column quarters work as two pairs, thread (row, quarter q) owning the 8-column half
// (q & 1) of the K slabs of pair (q >> 1); a pair finishes one slab per 8-column step, so the pairs' slabs alternate.  The
// first slabs of the next layer are ready after an eighth of the epilogue and only two slabs' products trail its end
// (whole 16-column steps per quarter: 3.21 vs 3.08 ms).
PINN_D int res_slab_order(int i, int ns) { return (i & 1) * (ns >> 1) + (i >> 1); }

#ifdef PINN_TIMELINE
// Debug build (profiles/timeline_wide.py): clock stamps of CTA 0 -- compute warps 0 (quarter 0) and 12 (quarter 3), lane 0:
// one record per epilogue [kind, before wait, after wait, after release, after slab 0..3]; MMA warp: one record per phase
// [after accfree, issue time of every slab, after the last commit].
__device__ long long g_rtl[2][512][8];
__device__ long long g_mtl[512][20];
__device__ long long g_stl[4][8][8];      // fine stamps inside the 8 steps of one hidden epilogue (4th recorded epilogue of each role)
#define STL(k, j) do { if (tl_on && tl_i == 17) g_stl[tl_w * 2 + (l == 1 ? 0 : 1)][k][j] = clock64(); } while (0)
#define RTL(k) do { if (tl_on && tl_i < 512) g_rtl[tl_w][tl_i][k] = clock64(); } while (0)
#define RTL_KIND(v) do { if (tl_on && tl_i < 512) g_rtl[tl_w][tl_i][0] = (v); } while (0)
#define RTL_NEXT() do { ++tl_i; } while (0)
#else
#define RTL(k) do {} while (0)
#define RTL_KIND(v) do {} while (0)
#define RTL_NEXT() do {} while (0)
#define STL(k, j) do {} while (0)
#endif

__global__ void __launch_bounds__(kRThreads, 1)
wide_res_ts_kernel(const __grid_constant__ pinn_net_t net, const float* __restrict__ x, int64_t n, const __grid_constant__ DropParams dp,
                   const __grid_constant__ ResArgs a, TcOut out, const __grid_constant__ CUtensorMap xmap, int use_tma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kRStages], empty[kRStages], ready[16], done, accfree, xbar;
  __shared__ __align__(128) float xs[kRT * PINN_N_IN];      // the work item's input tile, staged by TMA (tensor map of x)
  __shared__ uint32_t tmem_base_s;
  unsigned char* const ring = smem;
  float* const fsm = reinterpret_cast<float*>(ring + kRStages * kRStageBytes);
  const int L = a.L;
  float* const s_b = fsm;
  float* const s_bv0 = fsm + (PINN_MAX_HIDDEN - 1) * kRH;
  float* const s_bv1 = s_bv0 + kRH / 2;
  float* const s_wv2 = s_bv1 + kRNV;
  float* const s_part = s_wv2 + kRNV;
  float4* const a0s = reinterpret_cast<float4*>(s_part + 4 * kRT);        // [64 column quads][128 rows]

  const int tid = threadIdx.x, warp = tc::uniform_warp_idx(), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kRStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int c = 0; c < 16; ++c) tc::mbar_init(&ready[c], 256);
    tc::mbar_init(&done, 1);
    tc::mbar_init(&accfree, kRComputeWarps * 32);
    tc::mbar_init(&xbar, 1);
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
  auto item_passes = [&](int chunk) {
    if (!a.mc) return 1;
    const int t0 = chunk * Tc, cnt = a.T - t0 < Tc ? a.T - t0 : Tc;
    return (cnt > 0 ? cnt : 0) + ((a.do_eval && chunk == 0) ? 1 : 0);
  };
  const int n_phase = L + 1;
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
            const int N = phase_N(ph), ns = phase_slabs(ph);
            const uint32_t bytes = static_cast<uint32_t>(res_slab_bytes(N));
            for (int i = 0; i < ns; ++i, ++cnt) {
              const uint32_t s = cnt % kRStages;
              if (cnt >= kRStages) tc::mbar_wait(&empty[s], ((cnt / kRStages) - 1u) & 1u);
              tc::mbar_expect_tx(&full[s], bytes);
              tc::bulk_g2s(ring + s * kRStageBytes, img + static_cast<size_t>(res_slab_order(i, ns)) * bytes, bytes, &full[s]);
            }
          }
    }
    __syncwarp();
  } else if (warp == kRComputeWarps + 1) {
    // ================================================================== MMA issuer
    uint32_t cnt = 0, rpar = 0u, fpar = 0u;
    bool first_phase = true;
#ifdef PINN_TIMELINE
    int mtl_i = 0;
#endif
    const uint32_t ring_s = tc::smem_u32(ring), a_hi_t = tb + 256u, a_lo_t = tb + 384u;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x)
      for (int pi = 0, n_pass = item_passes(static_cast<int>(item % C)); pi < n_pass; ++pi)
        for (int ph = 0; ph < n_phase; ++ph) {
          const int N = phase_N(ph), ns = phase_slabs(ph);
          const uint32_t idesc = tc::make_idesc_f16(kRT, N);
          const uint32_t lbo_b = static_cast<uint32_t>(N) * 16u;
          // the accumulator is free once every compute thread holds its columns of the previous phase in registers
          if (!first_phase) { tc::mbar_wait(&accfree, fpar); fpar ^= 1u; }
          first_phase = false;
#ifdef PINN_TIMELINE
          const bool mtl_on = blockIdx.x == 0 && lane == 0 && mtl_i < 512;
          if (mtl_on) { g_mtl[mtl_i][0] = ph; g_mtl[mtl_i][1] = clock64(); }
#endif
#pragma unroll 1
          for (int i = 0; i < ns; ++i, ++cnt) {
            const uint32_t s = cnt % kRStages;
            const int slab = res_slab_order(i, ns);
            tc::mbar_wait(&ready[slab], (rpar >> slab) & 1u);
            rpar ^= 1u << slab;
            tc::mbar_wait(&full[s], (cnt / kRStages) & 1u);
            __syncwarp();
#ifdef PINN_TIMELINE
            if (mtl_on) g_mtl[mtl_i][2 + i] = clock64();
#endif
            if (tc::elect_one()) {
              tc::fence_after_sync();
              const uint32_t sw = ring_s + s * kRStageBytes, ac = 8u * static_cast<uint32_t>(slab);
              const uint64_t bh = tc::make_desc(sw, lbo_b, 128), bl = tc::make_desc(sw + static_cast<uint32_t>(N) * 32u, lbo_b, 128);
              tc::umma_f16_ts(tb, a_lo_t + ac, bh, idesc, i != 0 ? 1u : 0u);      // small terms first: lo*hi, hi*lo, then hi*hi
              tc::umma_f16_ts(tb, a_hi_t + ac, bl, idesc, 1u);
              tc::umma_f16_ts(tb, a_hi_t + ac, bh, idesc, 1u);
              tc::umma_commit(&empty[s]);
              if (i == ns - 1) tc::umma_commit(&done);
            }
            __syncwarp();
          }
#ifdef PINN_TIMELINE
          if (mtl_on) { g_mtl[mtl_i][18] = clock64(); }
          ++mtl_i;
#endif
        }
  } else {
    // ================================================================== compute warps: thread = (row, column quarter)
    const int q = warp >> 2, r = (warp & 3) * 32 + lane;
#ifdef PINN_TIMELINE
    const bool tl_on = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 12);
    const int tl_w = warp == 0 ? 0 : 1;
    int tl_i = 0;
#endif
    const uint32_t tlane = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t a_hi_l = tlane + 256u, a_lo_l = tlane + 384u;
    float4* const my_a0 = a0s + r;
    const bool drop_on = dp.p > 0.f, inj = dp.masks != nullptr;
    const int Dm = L * kRH + kRH / 2;
    uint32_t dpar = 0u, xpar = 0u;

    // 8 masked activations = columns [c0, c0 + 8) (one half of K slab c0 / 16) -> packed fp16 pairs in both A planes, then this
    // thread's arrival on the slab's barrier
    auto emit_half = [&](const float (&v)[8], int c0) {
      uint32_t h[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) tc::split_h2(v[2 * e], v[2 * e + 1], h[e], lo[e]);
      tc::tmem_st4(a_hi_l + static_cast<uint32_t>(c0 >> 1), reinterpret_cast<const float*>(h));
      tc::tmem_st4(a_lo_l + static_cast<uint32_t>(c0 >> 1), reinterpret_cast<const float*>(lo));
      tc::tmem_wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(&ready[c0 >> 4]);
    };
    const int pr = q >> 1, hf = q & 1;
    auto wait_done = [&]() {
      tc::mbar_wait(&done, dpar);
      dpar ^= 1u;
      __syncwarp();
      tc::fence_after_sync();
    };
    auto release_acc = [&]() {          // this thread's accumulator columns are in registers
      tc::tmem_wait_ld();
      tc::fence_before_sync();
      tc::mbar_arrive(&accfree);
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
      // keep-select 8 activations of units [j0, j0 + 8) of dropout layer `layer` (one Philox block / 8 injected bytes)
      auto select8 = [&](float (&v)[8], bool active, uint32_t pass, int tloc, uint32_t layer, uint32_t j0) {
        if (active) {
          bool k[8];
          if (inj) {
            const uint8_t* mrow = dp.masks + (static_cast<size_t>(tloc) * dp.mask_n + s) * Dm + layer * kRH + j0;
            const uint2 mb = *reinterpret_cast<const uint2*>(mrow);
#pragma unroll
            for (int e = 0; e < 4; ++e) { k[e] = ((mb.x >> (8 * e)) & 0xffu) != 0; k[4 + e] = ((mb.y >> (8 * e)) & 0xffu) != 0; }
          } else {
            keep8_from(Philox::gen_rk(dp.rk, s_lo, s_hi, pass, (layer << 16) | (j0 >> 3)), dp.thresh_hi, k);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = k[e] ? v[e] : 0.f;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= a.inact;
        }
      };
      auto tanh8b = [&](const float* z, const float* bias, float (&v)[8]) {
        const float4 bA = *reinterpret_cast<const float4*>(bias), bB = *reinterpret_cast<const float4*>(bias + 4);
        const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
        tanh8_prescaled(z, bb, v);
      };
      // ---- layer 0 (pass-invariant): this thread's 64 columns -> the tile's park in shared memory
      {
        float xr[PINN_N_IN];
        if (use_tma) {
          // [128 rows x 8 features] box of the tensor map, rows past n zero-filled; every earlier reader of `xs` has long
          // passed (it is read once, at the start of an item)
          if (tid == 0) {
            tc::mbar_expect_tx(&xbar, kRT * PINN_N_IN * sizeof(float));
            tc::tma_load_2d(xs, &xmap, 0, static_cast<int>(tile * kRT), &xbar);
          }
          tc::mbar_wait(&xbar, xpar);
          xpar ^= 1u;
          const float4* px = reinterpret_cast<const float4*>(xs + r * PINN_N_IN);
          const float4 q0 = px[0], q1 = px[1];
          xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
        } else if (valid) {
          const float4* px = reinterpret_cast<const float4*>(x + s * PINN_N_IN);
          const float4 q0 = __ldg(px), q1 = __ldg(px + 1);
          xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
        } else {
#pragma unroll
          for (int i = 0; i < PINN_N_IN; ++i) xr[i] = 0.f;
        }
#pragma unroll 1
        for (int i4 = 0; i4 < 16; ++i4) {
          const int c4 = 4 * (8 * pr + (i4 >> 1)) + 2 * hf + (i4 & 1);      // this thread's i4-th column quad
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
          my_a0[c4 * kRT] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
      }

      float mean = 0.f, m2 = 0.f, slv = 0.f;
#pragma unroll 1
      for (int pi = 0; pi < n_pass; ++pi) {
        const bool eval_pass = a.mc && eval_item && pi == 0;
        const int tl = a.mc ? (eval_item ? pi - 1 : pi) : 0;                 // pass index inside the chunk (Welford count)
        const int t = t0 + tl;                                               // pass index of the sweep (mask stream)
        const bool active = drop_on && !eval_pass && (!inj || valid);
        const uint32_t pass = static_cast<uint32_t>(dp.pass_offset + t);
        // ---- stage the masked layer-0 activations as the first A operand
        RTL_KIND(100); RTL(1); RTL(2); RTL(3);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c0 = 16 * (8 * pr + k) + 8 * hf;
          const float4 f0 = my_a0[(c0 >> 2) * kRT], f1 = my_a0[((c0 >> 2) + 1) * kRT];
          float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          select8(v, active, pass, t, 0u, static_cast<uint32_t>(c0));
          emit_half(v, c0);
          if (k & 1) RTL(4 + (k >> 1));
        }
        RTL_NEXT();
        // ---- hidden layers
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
          RTL_KIND(l); RTL(1);
          wait_done();
          RTL(2);
          float z[64];
#pragma unroll
          for (int k = 0; k < 8; ++k) tc::tmem_ld8(tlane + static_cast<uint32_t>(16 * (8 * pr + k) + 8 * hf), z + 8 * k);
          release_acc();
          RTL(3);
          {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c0 = 16 * (8 * pr + k) + 8 * hf;
              float v[8];
              STL(k, 0);
              tanh8b(z + 8 * k, s_b + (l - 1) * kRH + c0, v);
#ifdef PINN_TIMELINE
              asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
#endif
              STL(k, 1);
              select8(v, active, pass, t, static_cast<uint32_t>(l), static_cast<uint32_t>(c0));
#ifdef PINN_TIMELINE
              asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
              STL(k, 2);
              {
                uint32_t h[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) tc::split_h2(v[2 * e], v[2 * e + 1], h[e], lo[e]);
                asm volatile("" : "+r"(h[0]), "+r"(h[1]), "+r"(h[2]), "+r"(h[3]), "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]));
                STL(k, 3);
                tc::tmem_st4(a_hi_l + static_cast<uint32_t>(c0 >> 1), reinterpret_cast<const float*>(h));
                tc::tmem_st4(a_lo_l + static_cast<uint32_t>(c0 >> 1), reinterpret_cast<const float*>(lo));
                STL(k, 4);
                tc::tmem_wait_st();
                STL(k, 5);
                tc::fence_before_sync();
                tc::mbar_arrive(&ready[c0 >> 4]);
                STL(k, 6);
              }
#else
              emit_half(v, c0);
#endif
              if (k & 1) RTL(4 + (k >> 1));
            }
          }
          RTL_NEXT();
        }
        // ---- heads: 128 variance-head units (32 per quarter) + the mean head (column 128)
        float u = 0.f;
        {
          RTL_KIND(200); RTL(1);
          wait_done();
          RTL(2);
          float z[32], zz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::tmem_ld8(tlane + static_cast<uint32_t>(16 * (4 * pr + k) + 8 * hf), z + 8 * k);
          if (q == 0) tc::tmem_ld4(tlane + static_cast<uint32_t>(kRH / 2), zz);
          release_acc();
          RTL(3);
          u = zz[0] + __ldg(net.bp);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c0 = 16 * (4 * pr + k) + 8 * hf;
            float v[8];
            tanh8b(z + 8 * k, s_bv0 + c0, v);
            select8(v, active, pass, t, static_cast<uint32_t>(L), static_cast<uint32_t>(c0));
            emit_half(v, c0);
            if (k & 1) RTL(4 + (k >> 1));
          }
          RTL_NEXT();
        }
        // ---- variance layer 1 (64 units, 16 per quarter) + the last dot; thread (row, 0) finishes the sample
        {
          RTL_KIND(300); RTL(1);
          wait_done();
          RTL(2);
          float z[16], v[8];
          tc::tmem_ld16(tlane + static_cast<uint32_t>(16 * q), z);
          release_acc();
          RTL(3);
          float part = 0.f;
#pragma unroll
          for (int g = 0; g < 16; g += 8) {
            tanh8b(z + g, s_bv1 + 16 * q + g, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) part = fmaf(s_wv2[16 * q + g + e], v[e], part);
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
          RTL(4);
          RTL_NEXT();
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

constexpr size_t kRSmemBytesTs = static_cast<size_t>(kRStages) * kRStageBytes +
                                 (static_cast<size_t>(PINN_MAX_HIDDEN - 1) * kRH + kRH / 2 + 2 * kRNV + 4 * kRT) * sizeof(float) +
                                 static_cast<size_t>(kRT) * kRH * sizeof(float);
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
  const ResPlan p = res_plan(L);
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
  const int C = mc ? res_pass_chunks(T) : 1;
  const size_t part_bytes = C > 1 ? static_cast<size_t>(C) * 3 * n * sizeof(float) : 0;
  if (workspace_bytes < p.bytes + part_bytes) { *err = PINN_E_WORKSPACE; return -1; }
  a.chunks = C;
  a.part = reinterpret_cast<float*>(ws + p.bytes);
  a.L = L; a.T = T; a.mc = mc ? 1 : 0; a.do_eval = (mc && out.pred_mean != nullptr) ? 1 : 0;
  a.inact = dp.p > 0.f ? dp.keep : 1.0f;
  a.no_logvar = (net->flags & PINN_NET_NO_LOGVAR) ? 1 : 0;
  auto kern = wide_res_ts_kernel;
  const size_t smem_bytes = kRSmemBytesTs;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
  if (e != cudaSuccess) { *err = static_cast<int>(e); return -1; }
  const int64_t items = ((n + kRT - 1) / kRT) * C;
  const int grid = static_cast<int>(items < sm_count() ? items : sm_count());
  alignas(64) CUtensorMap xmap;
  memset(&xmap, 0, sizeof(xmap));
  const int use_tma = (net->flags & PINN_NET_NO_TMA_INPUT) ? 0 : (make_x_tensor_map(&xmap, x, n) ? 1 : 0);
  kern<<<grid, kRThreads, smem_bytes, st>>>(*net, x, n, dp, a, out, xmap, use_tma);
  if (C > 1) launch_mc_merge(a.part, n, T, C, out, st);
  *err = static_cast<int>(cudaGetLastError());
  return *err == 0 ? 1 : -1;
}

}  // namespace pinn

#ifdef PINN_TIMELINE
extern "C" int pinn_debug_wide_timeline(long long* compute_out, long long* mma_out) {
  cudaError_t e = cudaMemcpyFromSymbol(compute_out, pinn::g_rtl, sizeof(pinn::g_rtl));
  if (e != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(cudaMemcpyFromSymbol(mma_out, pinn::g_mtl, sizeof(pinn::g_mtl)));
}
extern "C" int pinn_debug_wide_steps(long long* out) { return static_cast<int>(cudaMemcpyFromSymbol(out, pinn::g_stl, sizeof(pinn::g_stl))); }
#endif
