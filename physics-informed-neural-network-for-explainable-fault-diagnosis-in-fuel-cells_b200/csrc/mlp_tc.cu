// mlp_tc.cu -- tensor-core (tcgen05 + TMEM) path of K1 / K4 for the headline 64-wide net.
//
// CTA = 18 warps: two compute groups of 256 threads + one MMA warp per group.
//
//   compute group g : owns one 128-sample tile at a time.  Threads (r, 0) and (r, 1) own the two
//              32-column halves of sample r (= TMEM lane r).  Philox counters are per (sample,
//              pass, layer, unit/8): the mask stream is identical to the thread-per-sample
//              kernels, only the 64x64 contractions moved to the tensor cores.
//   MMA warps: one per group; an elected lane waits for the group's "operands ready" mbarrier, issues the
//              3 x 8 tcgen05.mma (lo*hi + hi*lo + hi*hi: 3xTF32,
//              fp32 accumulation in TMEM) and commits to that group's "done" mbarrier.  Issuing
//              24 MMAs occupies the issuing thread for as long as the tensor pipe needs to
//              execute them (~1000 clk, profiles/README.md), so it cannot be a compute thread.
//
//   weights  : every hidden layer W_l (l >= 1) and the stacked head matrix [Wv0; Wp; 0] are
//              split once per CTA into tf32 hi/lo planes (UMMA K-major layout) and stay resident
//              in shared memory, pre-multiplied by the dropout scale 1/(1-p);
//   operands : activations never touch shared memory: the A operand of every MMA is read from
//              TENSOR MEMORY (hi and lo planes, lane = sample row), written by the epilogue with
//              tcgen05.st.  With A in shared memory an M128 N64 K8 tf32 MMA pulls 6 KB through
//              the 128 B/clk shared-memory port (48 clk vs a 32 clk math floor);
//   per layer: a compute thread signals "ready", draws the Philox blocks of the coming epilogue
//              while the tensor core works, waits for "done", pulls its 32 accumulator columns
//              (tcgen05.ld), applies bias + tanh + keep-select, re-splits and stores the next
//              A planes;
//   heads    : one N = 48 MMA gives the 32 variance-head pre-activations and the mean head; the
//              32 -> 16 -> 1 tail is split between the row's two threads (hand-over through
//              spare accumulator columns) and runs on CUDA cores.
//
// Layer 0 (K = 8) is pass-invariant: computed once per tile and parked in tensor
// memory, re-read (not recomputed) by every pass.
//
// Tensor-memory map (512 columns, one CTA per SM):
//   [64 g, +64)        accumulators of group g (columns 48..63 double as the tail hand-over)
//   [128 + 128 g, +64) activation hi plane, [192 + 128 g, +64) lo plane
//   [384 + 64 g, +64)  layer-0 activations of the group's current tile
#include "net.cuh"
#include "tc.cuh"
#include "tc_api.cuh"
#include <string.h>

namespace pinn {

constexpr int kTcH = 64;
constexpr int kTcTile = 128;
constexpr int kHeadN = 48;  // 32 variance-head rows + 1 mean-head row + 15 zero rows (N % 16 == 0)
constexpr int kTcThreads = 512 + 64;   // 16 compute warps + one MMA warp per group

struct TcLayout {  // offsets in floats from the dynamic shared-memory base
  int L;
  int b_hi[PINN_MAX_HIDDEN], b_lo[PINN_MAX_HIDDEN];  // hidden layer l >= 1
  int h_hi, h_lo;                                    // stacked heads
  int W0, b0, b[PINN_MAX_HIDDEN], bv0, bp, Wv1, bv1, Wv2, bv2;
  int total;
};
PINN_HD TcLayout make_tc_layout(int L) {
  TcLayout t;
  t.L = L;
  int o = 0;
  for (int l = 0; l < PINN_MAX_HIDDEN; ++l) { t.b_hi[l] = t.b_lo[l] = t.b[l] = 0; }
  for (int l = 1; l < L; ++l) { t.b_hi[l] = o; o += kTcH * kTcH; t.b_lo[l] = o; o += kTcH * kTcH; }
  t.h_hi = o; o += kHeadN * kTcH;
  t.h_lo = o; o += kHeadN * kTcH;
  t.W0 = o; o += kTcH * PINN_N_IN;
  t.b0 = o; o += kTcH;
  for (int l = 1; l < L; ++l) { t.b[l] = o; o += kTcH; }
  t.bv0 = o; o += kTcH / 2;
  t.bp = o; o += 4;
  t.Wv1 = o; o += (kTcH / 4) * (kTcH / 2);
  t.bv1 = o; o += kTcH / 4;
  t.Wv2 = o; o += kTcH / 4;
  t.bv2 = o; o += 4;
  t.total = o;
  return t;
}

#ifdef PINN_TIMELINE
// Debug build (profiles/timeline_mc.py): clock stamps of CTA 0 / group 0, warps 0 (half 0) and 4 (half 1).
__device__ long long g_tl[3][64][8];
#define TL(k) do { if (tl_on && tl_i < 64) g_tl[half][tl_i][k] = clock64(); } while (0)
#define TL_NEXT() do { ++tl_i; } while (0)
#else
#define TL(k) do {} while (0)
#define TL_NEXT() do {} while (0)
#endif

PINN_D void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
PINN_D void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// MC = true : eval pass (if pred_mean) + T dropout passes with Welford.   MC = false: one pass.
// INJ = true: keep decisions come from an injected mask tensor (parity runs) instead of Philox.
//
// Instruction diet of the epilogue (it, not the tensor pipe, bounds this kernel):
//  * the dropout scale 1/(1-p) is folded into the resident weight planes (and into Wv1), so a
//    kept activation is stored as is: one FSEL per unit instead of FSEL + FMUL; passes without
//    dropout (the eval pass) multiply by (1-p) instead;
//  * biases (and layer 0's weights) are pre-scaled by 2 log2(e): tanh = FFMA, EX2, FADD, RCP, FFMA;
//  * Philox round keys are constant-bank operands, 16-bit draws are compared in place;
//  * the tf32 split of an activation is IADD + LOP3 + FADD.
template <bool MC, bool INJ, bool CH>      // CH: the sweep is cut into pass chunks (long sweeps only; C = 1 folds away otherwise)
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_tc_kernel(pinn_net_t net, TcLayout lay, const float* __restrict__ x, int64_t n, int T, const __grid_constant__ DropParams dp,
              TcOut out, int chunks, float* __restrict__ part, const __grid_constant__ CUtensorMap xmap, int use_tma) {
  constexpr int H = kTcH, HH = kTcH / 2;
  constexpr uint32_t LBO_B = H * 16, LBO_H = kHeadN * 16;
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t ready[2][4], done[2];     // ready[group][chunk]: one barrier per hand-over of a layer, so no
                                                              // thread can arrive twice on a barrier within one of its phases
  __shared__ uint32_t tmem_base_s;
  // input tiles staged by TMA (tensor map of x, [128 rows x 8 features] box, rows past n zero-filled): per group two 4 KB
  // buffers behind the resident weights (use_tma = their byte offset; < 0 when the seven-layer net leaves no room), the
  // NEXT work item's tile is requested while the current one runs its passes
  __shared__ __align__(8) uint64_t xbar[2][2];
  float (*const xs)[2][kTcTile * PINN_N_IN] =
      reinterpret_cast<float (*)[2][kTcTile * PINN_N_IN]>(reinterpret_cast<unsigned char*>(smem) + (use_tma >= 0 ? use_tma : 0));
  const int L = lay.L, tid = threadIdx.x, row = tid & 127;
  const int warp = tc::uniform_warp_idx(), grp = (warp >> 3) & 1, half = (warp >> 2) & 1;  // warp-uniform roles
  const bool mma_warp = warp >= 16;
  const int Dm = L * H + H / 2;
  const int cb = half * HH;  // first column owned by this thread
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;    // folded into every matrix that consumes masked activations
  const float inact = drop_on ? dp.keep : 1.0f;      // multiplier of an un-masked activation (undoes the fold)

  // ---------------------------------------------------------------- one-time CTA set-up
  if (tid == 0) {
    for (int c = 0; c < 4; ++c) { tc::mbar_init(&ready[0][c], 256); tc::mbar_init(&ready[1][c], 256); }
    tc::mbar_init(&done[0], 1);
    tc::mbar_init(&done[1], 1);
    for (int g = 0; g < 2; ++g) { tc::mbar_init(&xbar[g][0], 1); tc::mbar_init(&xbar[g][1], 1); }
    tc::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  auto scaled4 = [](float4 v, float c) { return make_float4(v.x * c, v.y * c, v.z * c, v.w * c); };
  for (int l = 1; l < L; ++l)
    for (int idx = tid; idx < H * (H / 4); idx += blockDim.x) {
      const int nrow = idx % H, kc = idx / H;
      tc::store_split4(smem + lay.b_hi[l], smem + lay.b_lo[l], LBO_B, nrow, kc,
                       scaled4(__ldg(reinterpret_cast<const float4*>(net.W[l] + nrow * H) + kc), wscale));
    }
  for (int idx = tid; idx < kHeadN * (H / 4); idx += blockDim.x) {
    const int nrow = idx % kHeadN, kc = idx / kHeadN;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nrow < H / 2) v = __ldg(reinterpret_cast<const float4*>(net.Wv0 + nrow * H) + kc);
    else if (nrow == H / 2) v = __ldg(reinterpret_cast<const float4*>(net.Wp) + kc);
    tc::store_split4(smem + lay.h_hi, smem + lay.h_lo, LBO_H, nrow, kc, scaled4(v, wscale));
  }
  stage_tensor_scaled(smem + lay.W0, net.W[0], H * PINN_N_IN, kTanhArg);
  stage_tensor_scaled(smem + lay.b0, net.b[0], H, kTanhArg);
  for (int l = 1; l < L; ++l) stage_tensor_scaled(smem + lay.b[l], net.b[l], H, kTanhArg);
  stage_tensor_scaled(smem + lay.bv0, net.bv0, H / 2, kTanhArg);
  stage_tensor(smem + lay.bp, net.bp, 1);
  stage_tensor_scaled(smem + lay.Wv1, net.Wv1, (H / 4) * (H / 2), kTanhArg * wscale);
  stage_tensor_scaled(smem + lay.bv1, net.bv1, H / 4, kTanhArg);
  stage_tensor(smem + lay.Wv2, net.Wv2, H / 4);
  stage_tensor(smem + lay.bv2, net.bv2, 1);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const int64_t n_tiles = (n + kTcTile - 1) / kTcTile;
  const bool do_eval = MC && out.pred_mean != nullptr;
  // Work item = (tile, pass chunk): a long sweep is cut into `chunks` runs of Tc consecutive passes per tile (chosen from T
  // alone, so the arithmetic of a sample does not depend on the batch or its sharding); every run keeps its own Welford
  // triple, mc_merge_kernel folds them in chunk order (Chan).  With few tiles per SM this is what fills the machine:
  // 977 tiles x T = 1000 are 3.3 waves of whole-tile items but 13.2 waves of quarter-sweeps.
  const int C = (MC && CH) ? chunks : 1, Tc = (T + C - 1) / C;
  auto item_passes = [&](int chunk) {            // dropout passes of the chunk (+ the eval pass, which rides with chunk 0)
    const int t0 = chunk * Tc, cnt = (T - t0 < Tc ? T - t0 : Tc);
    return MC ? (cnt > 0 ? cnt : 0) + ((do_eval && chunk == 0) ? 1 : 0) : 1;
  };
  const uint32_t idesc64 = tc::make_idesc_tf32(kTcTile, H), idesc48 = tc::make_idesc_tf32(kTcTile, kHeadN);

  if (mma_warp) {
    // ================================================================== MMA warp
    // One issuing warp per group; it walks the group's deterministic schedule (tile -> pass -> layers
    // 1..L-1 -> heads) and sleeps on the group's "ready" barrier in between.
    const int g = warp - 16;
    const int64_t first = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(g) * gridDim.x;     // same map as the compute warps below
    const int64_t n_items = n_tiles * C;
    const uint32_t d_t = tmem_base_s + static_cast<uint32_t>(g * 64);
    const uint32_t a_hi = tmem_base_s + static_cast<uint32_t>(128 + g * 128);
    uint32_t par = 0u;
#ifdef PINN_TIMELINE
    int tlm = 0;
#endif
    for (int64_t item = first; item < n_items; item += 2 * static_cast<int64_t>(gridDim.x))
    for (int it = 0, np = item_passes(static_cast<int>(item % C)); it < np; ++it) {
#pragma unroll 1
      for (int l = 1; l <= L; ++l) {
        const bool heads = l == L;
        const uint64_t b_hi_d = tc::make_desc(tc::smem_u32(smem + (heads ? lay.h_hi : lay.b_hi[l])), heads ? LBO_H : LBO_B, 128);
        const uint64_t b_lo_d = tc::make_desc(tc::smem_u32(smem + (heads ? lay.h_lo : lay.b_lo[l])), heads ? LBO_H : LBO_B, 128);
        // The compute threads hand their new A columns over in four chunks (8 columns of each half = K slabs c and 4 + c):
        // the products of a slab are issued as soon as it is complete, so the tensor pipe works under the epilogue that
        // feeds it and only the last slab pair's MMAs (+ the fixed latency) are exposed after the epilogue ends.
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tc::mbar_wait(&ready[g][c], par);
          __syncwarp();
#ifdef PINN_TIMELINE
          if (blockIdx.x == 0 && g == 0 && !heads && c == 0 && (tid & 31) == 0 && tlm < 64) g_tl[2][tlm][0] = clock64();
#endif
          if (tc::elect_one()) {
            tc::fence_after_sync();
            if (heads) tc::issue_3xtf32_ts_slabs(d_t, a_hi, a_hi + 64u, b_hi_d, b_lo_d, LBO_H, idesc48, c, 4 + c, c == 0);
            else tc::issue_3xtf32_ts_slabs(d_t, a_hi, a_hi + 64u, b_hi_d, b_lo_d, LBO_B, idesc64, c, 4 + c, c == 0);
            if (c == 3) tc::umma_commit(&done[g]);
          }
          __syncwarp();
        }
        par ^= 1u;
#ifdef PINN_TIMELINE
        if (blockIdx.x == 0 && g == 0 && !heads && (tid & 31) == 0 && tlm < 64) { g_tl[2][tlm][1] = clock64(); ++tlm; }
#endif
      }
    }
  } else {
    // ================================================================== compute groups
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;          // this warp's 32 TMEM lanes
    const uint32_t d_lane = tmem_base_s + static_cast<uint32_t>(grp * 64) + lane_sel;
    const uint32_t x_lane = d_lane + 48u;                                            // tail hand-over: accumulator columns the heads MMA (N = 48) leaves alone
    const uint32_t a_hi_l = tmem_base_s + static_cast<uint32_t>(128 + grp * 128) + lane_sel, a_lo_l = a_hi_l + 64u;
    const uint32_t a0_l = tmem_base_s + static_cast<uint32_t>(384 + grp * 64) + lane_sel;
    uint32_t phase = 0;
#ifdef PINN_TIMELINE
    const bool tl_on = blockIdx.x == 0 && grp == 0 && (warp & 3) == 0 && (tid & 31) == 0;
    int tl_i = 0;
#endif

    // One 8-column chunk of this row's next A operand: split into tf32 hi / lo planes in tensor memory.  The hand-over of
    // the PREVIOUS chunk (its stores have long landed) goes out between the split and the stores of this one, so the MMA
    // warp can start on a K slab pair while the rest of the epilogue is still running; the last chunk is handed over by
    // signal_ready() below.
    auto store8 = [&](int c0, const float (&v)[8], int chunk) {
      float h[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; q += 2) {
        h[q] = tc::tf32_hi_fast(v[q]); h[q + 1] = tc::tf32_hi_fast(v[q + 1]);
        const float2 l2 = __ffma2_rn(make_float2(h[q], h[q + 1]), make_float2(-1.0f, -1.0f), make_float2(v[q], v[q + 1]));   // v - h, exact
        lo[q] = l2.x; lo[q + 1] = l2.y;
      }
      if (chunk > 0) {
        tc::tmem_wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&ready[grp][chunk - 1]);
      }
      tc::tmem_st8(a_hi_l + static_cast<uint32_t>(c0), h);
      tc::tmem_st8(a_lo_l + static_cast<uint32_t>(c0), lo);
    };
    // Philox blocks are drawn AHEAD of the MMA wait that precedes their use (they depend on nothing the tensor core
    // produces) and pinned there.  Tried and measured slower (profiles/README.md, round 2): drawing them one phase ahead
    // inside the previous epilogue, as raw blocks (11.6 ms) or as packed keep bits (12.4 ms), against 10.3 ms here -- the
    // epilogue is bound by the MUFU pipe and dependent-issue latency and does not absorb the extra integer work.
    auto draw = [&](uint4* r, int nblk, const KeepSrc<INJ>& ks, uint32_t pass, uint32_t layer, int c0) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nblk) r[c] = Philox::gen_rk(dp.rk, ks.s_lo, ks.s_hi, pass, (layer << 16) | static_cast<uint32_t>((c0 >> 3) + c));
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nblk) asm volatile("" : "+r"(r[c].x), "+r"(r[c].y), "+r"(r[c].z), "+r"(r[c].w));   // pin before the wait
    };
    // keep-select 8 activations: pre-drawn block `r` (Philox), injected mask bytes, or no dropout
    auto select8 = [&](const uint4& r, const KeepSrc<INJ>& ks, bool active, uint32_t layer, int c0, const float (&t)[8], float (&v)[8]) {
      if (active) {
        bool k[8];
        if (INJ) ks.get8(dp, layer, static_cast<uint32_t>(c0), layer * H, k);
        else keep8_from(r, dp.thresh_hi, k);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = k[q] ? t[q] : 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = t[q] * inact;
      }
    };
    // hand the last chunk of the A planes to the MMA warp / wait for its accumulators
    auto signal_ready = [&]() {
      tc::tmem_wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(&ready[grp][3]);
    };
    auto wait_done = [&]() {
      tc::mbar_wait(&done[grp], phase);
      phase ^= 1u;
      __syncwarp();
      tc::fence_after_sync();
    };

    // tile -> (CTA, group): group 0 of every CTA first, then group 1: up to one tile per SM every tile runs alone (a lone
    // tile finishes ~1.4x sooner than a tile of an interleaved pair; pairing only buys throughput once all SMs are busy)
    const int64_t item0 = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(grp) * gridDim.x, item_step = static_cast<int64_t>(gridDim.x) * 2;
    const bool x_leader = use_tma >= 0 && (tid & 255) == 0;          // one thread per group requests the tiles
    auto request_x = [&](int64_t it, int buf) {
      tc::mbar_expect_tx(&xbar[grp][buf], kTcTile * PINN_N_IN * sizeof(float));
      tc::tma_load_2d(xs[grp][buf], &xmap, 0, static_cast<int>((it / C) * kTcTile), &xbar[grp][buf]);
    };
    if (x_leader && item0 < n_tiles * C) request_x(item0, 0);
    uint32_t xcount = 0;
    for (int64_t item = item0; item < n_tiles * C; item += item_step, ++xcount) {
      const int64_t tile = item / C;
      const int chunk = static_cast<int>(item % C), t0 = chunk * Tc;
      const bool eval_item = do_eval && chunk == 0;
      const int n_pass = item_passes(chunk);
      const int64_t s = tile * kTcTile + row;
      const bool valid = s < n;
      // layer 0 (pass-invariant): this thread's 32 columns -> tensor memory
      {
        float xr[PINN_N_IN];
        if (use_tma >= 0) {
          const int buf = static_cast<int>(xcount & 1u);
          if (x_leader && item + item_step < n_tiles * C) request_x(item + item_step, buf ^ 1);    // that buffer was last read one item ago
          tc::mbar_wait(&xbar[grp][buf], (xcount >> 1) & 1u);
          const float4* px = reinterpret_cast<const float4*>(xs[grp][buf] + row * PINN_N_IN);
          const float4 q0 = px[0], q1 = px[1];
          xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
        } else if (valid) {
          const float4* px = reinterpret_cast<const float4*>(x + s * PINN_N_IN);
          float4 q0 = __ldg(px), q1 = __ldg(px + 1);
          xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
        } else {
#pragma unroll
          for (int i = 0; i < PINN_N_IN; ++i) xr[i] = 0.f;
        }
        const float* W0 = smem + lay.W0 + cb * PINN_N_IN;
        const float* b0 = smem + lay.b0 + cb;
#pragma unroll
        for (int g = 0; g < HH; g += 8) {
          float a8[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int j = g + q;
            const float4 w0 = *reinterpret_cast<const float4*>(W0 + j * PINN_N_IN);
            const float4 w1 = *reinterpret_cast<const float4*>(W0 + j * PINN_N_IN + 4);
            float z = b0[j];
            z = fmaf(w0.x, xr[0], z); z = fmaf(w0.y, xr[1], z); z = fmaf(w0.z, xr[2], z); z = fmaf(w0.w, xr[3], z);
            z = fmaf(w1.x, xr[4], z); z = fmaf(w1.y, xr[5], z); z = fmaf(w1.z, xr[6], z); z = fmaf(w1.w, xr[7], z);
            a8[q] = tanh_pre(z);
          }
          tc::tmem_st8(a0_l + static_cast<uint32_t>(cb + g), a8);
        }
        tc::tmem_wait_st();
      }

      float mean = 0.f, m2 = 0.f, slv = 0.f;
      const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
      KeepSrc<INJ> ks;
      ks.s_lo = static_cast<uint32_t>(sg); ks.s_hi = static_cast<uint32_t>(sg >> 32);
      uint4 r0[4] = {};        // draws of the coming pass's layer-0 staging
      if (!INJ && drop_on && !eval_item) draw(r0, 4, ks, static_cast<uint32_t>(dp.pass_offset + t0), 0u, cb);
#pragma unroll 1
      for (int pi = 0; pi < n_pass; ++pi) {
        const bool eval_pass = MC && eval_item && pi == 0;
        const int tl = MC ? (eval_item ? pi - 1 : pi) : 0;      // pass index inside the chunk (Welford count)
        const int t = t0 + tl;                                  // pass index of the sweep (mask stream)
        // dropout is active on this pass?  (injected masks: tail rows of the last tile have no mask row)
        const bool active = drop_on && !eval_pass && (!INJ || valid);
        ks.pass = static_cast<uint32_t>(dp.pass_offset + t);
        ks.mrow = INJ ? dp.masks + (static_cast<size_t>(t) * dp.mask_n + (valid ? s : 0)) * Dm : nullptr;
        // ---- stage layer-0 activations (masked) as the first A operand
        {
          float a0[HH];
          tc::tmem_ld16(a0_l + static_cast<uint32_t>(cb), a0);
          tc::tmem_ld16(a0_l + static_cast<uint32_t>(cb + 16), a0 + 16);
          tc::tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < HH; g += 8) {
            const float t8[8] = {a0[g], a0[g + 1], a0[g + 2], a0[g + 3], a0[g + 4], a0[g + 5], a0[g + 6], a0[g + 7]};
            float v[8];
            select8(r0[g / 8], ks, active, 0u, cb + g, t8, v);
            store8(cb + g, v, g / 8);
          }
        }
        // ---- hidden layers on the tensor cores
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
          TL(0);
          signal_ready();
          TL(1);
          uint4 rl[4] = {};
          if (!INJ && active) draw(rl, 4, ks, ks.pass, static_cast<uint32_t>(l), cb);
          TL(2);
          wait_done();
          TL(3);
          const float* bl = smem + lay.b[l] + cb;
          float z[HH];
          tc::tmem_ld16(d_lane + cb, z);
          tc::tmem_ld16(d_lane + cb + 16, z + 16);
          tc::tmem_wait_ld();
          TL(4);
#pragma unroll
          for (int g = 0; g < HH; g += 8) {
            const float4 bA = *reinterpret_cast<const float4*>(bl + g), bB = *reinterpret_cast<const float4*>(bl + g + 4);
            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
            float t8[8], v[8];
            tanh8_prescaled<true>(z + g, bb, t8);
            select8(rl[g / 8], ks, active, static_cast<uint32_t>(l), cb + g, t8, v);
            store8(cb + g, v, g / 8);
          }

          TL(5);
          TL_NEXT();
        }
        // ---- heads: [Wv0; Wp] in one N = 48 MMA
        TL(0);
        signal_ready();
        TL(1);
        uint4 rv[2] = {};
        if (!INJ && active) draw(rv, 2, ks, ks.pass, static_cast<uint32_t>(L), 16 * half);
        if (!INJ && drop_on && pi + 1 < n_pass) draw(r0, 4, ks, static_cast<uint32_t>(dp.pass_offset + t + 1), 0u, cb);
        TL(2);
        wait_done();
        TL(3);
        {
          // Variance head, split between the row's two threads: each activates 16 of the 32 head units
          // and forms its share of the sixteen 32 -> 16 sums; thread (row, 1) parks its share in spare
          // accumulator columns of the row's lane and moves on to the next pass, thread (row, 0) adds the
          // two shares and finishes the sample.  Hand-over = named barrier (producers arrive, consumers sync).
          float v0[16], part[16];
          float u = 0.f;
          tc::tmem_ld16(d_lane + 16 * half, v0);
          if (half == 0) { float zz[8]; tc::tmem_ld8(d_lane + 32, zz); tc::tmem_wait_ld(); u = zz[0] + smem[lay.bp]; }
          else tc::tmem_wait_ld();
          const float* bv0 = smem + lay.bv0 + 16 * half;
#pragma unroll
          for (int g = 0; g < 16; g += 8) {
            const float4 bA = *reinterpret_cast<const float4*>(bv0 + g), bB = *reinterpret_cast<const float4*>(bv0 + g + 4);
            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
            float t8[8], v[8];
            tanh8_prescaled<true>(v0 + g, bb, t8);
            select8(rv[g / 8], ks, active, static_cast<uint32_t>(L), 16 * half + g, t8, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) v0[g + q] = v[q];
          }
          TL(4);
          const float* Wv1 = smem + lay.Wv1 + 16 * half;     // pre-scaled by 2 log2(e) / (1-p)
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float2 acc = make_float2(0.f, 0.f), acc2 = acc;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 w = *reinterpret_cast<const float4*>(Wv1 + k * HH + 4 * i4);
              acc = ffma2(make_float2(w.x, w.y), make_float2(v0[4 * i4], v0[4 * i4 + 1]), acc);
              acc2 = ffma2(make_float2(w.z, w.w), make_float2(v0[4 * i4 + 2], v0[4 * i4 + 3]), acc2);
            }
            part[k] = (acc.x + acc.y) + (acc2.x + acc2.y);
          }
          TL(5);
          if (half == 1) {
            tc::tmem_st16(x_lane, part);
            tc::tmem_wait_st();
            tc::fence_before_sync();
            bar_arrive_n(1 + grp, 256);
            TL(6); TL(7);
          } else {
            bar_sync_n(1 + grp, 256);
            tc::fence_after_sync();
            float p1[16];
            tc::tmem_ld16(x_lane, p1);
            tc::tmem_wait_ld();
            TL(6);
            float vraw = smem[lay.bv2];
            const float* bv1 = smem + lay.bv1;
            const float* Wv2 = smem + lay.Wv2;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
              const float2 a1 = tanh_pre2<true>(make_float2((part[k] + p1[k]) + bv1[k], (part[k + 1] + p1[k + 1]) + bv1[k + 1]));
              vraw = fmaf(Wv2[k], a1.x, vraw);
              vraw = fmaf(Wv2[k + 1], a1.y, vraw);
            }
            const float lv = logvar_out(vraw, (net.flags & PINN_NET_NO_LOGVAR) != 0);
            if (!MC) {
              if (valid) { out.u[s] = u; out.s[s] = lv; }
            } else if (eval_pass) {
              if (valid) out.pred_mean[s] = u;
            } else {
              const float d = u - mean;
              mean += d / static_cast<float>(tl + 1);
              m2 = fmaf(d, u - mean, m2);
              slv += lv;
            }
            TL(7);
          }
        }
        TL_NEXT();
      }
      if (MC && valid && half == 0 && C > 1) {
        float* pp = part + static_cast<size_t>(chunk) * 3 * n + s;
        pp[0] = mean; pp[n] = m2; pp[2 * n] = slv;
      } else if (MC && valid && half == 0) {
        if (out.raw_mean) out.raw_mean[s] = mean;
        if (out.raw_m2) out.raw_m2[s] = m2;
        if (out.raw_slv) out.raw_slv[s] = slv;
        const float invT = 1.0f / static_cast<float>(T > 0 ? T : 1);
        if (out.a_u) out.a_u[s] = sqrtf(expf(slv * invT));
        if (out.e_u) out.e_u[s] = sqrtf(fmaxf(m2, 0.f) * invT);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}


// Launch helper used by pinn_mlp_fwd / pinn_mc_dropout.  Returns 1 if the TC path took the
// call, 0 if the shape is not covered (caller falls through to the FFMA kernels), <0 / >1 on error.
// Fold the per-chunk Welford triples of a sweep in chunk order (Chan's update, fp32 like the kernel's own running form)
// and finish the sample: a_u = sqrt(exp(mean_t logvar)), e_u = sqrt(var_t u) (01:1483-1486).
__global__ void __launch_bounds__(256) mc_merge_kernel(const float* __restrict__ part, int64_t n, int T, int C, TcOut out) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int Tc = (T + C - 1) / C;
  float cnt = 0.f, mean = 0.f, m2 = 0.f, slv = 0.f;
  for (int k = 0; k < C; ++k) {
    const int t0 = k * Tc, ck = T - t0 < Tc ? T - t0 : Tc;
    if (ck <= 0) break;
    const float* pp = part + static_cast<size_t>(k) * 3 * n + s;
    const float mb = pp[0], m2b = pp[n], sb = pp[2 * n], cb = static_cast<float>(ck);
    if (k == 0) { cnt = cb; mean = mb; m2 = m2b; slv = sb; }
    else {
      const float tot = cnt + cb, d = mb - mean;
      mean = mean + d * (cb / tot);
      m2 = m2 + m2b + d * d * (cnt * cb / tot);
      slv += sb;
      cnt = tot;
    }
  }
  if (out.raw_mean) out.raw_mean[s] = mean;
  if (out.raw_m2) out.raw_m2[s] = m2;
  if (out.raw_slv) out.raw_slv[s] = slv;
  const float invT = 1.0f / static_cast<float>(T > 0 ? T : 1);
  if (out.a_u) out.a_u[s] = sqrtf(expf(slv * invT));
  if (out.e_u) out.e_u[s] = sqrtf(fmaxf(m2, 0.f) * invT);
}

// Pass chunks of a sweep: runs of ~250 passes, at most 8 -- a function of T ALONE, so that a sample's numbers do not depend
// on the batch size or on how the batch is sharded (tests assert bitwise shard invariance).
int mc_pass_chunks(int T) {
  const int c = (T + 125) / 250;
  return c < 1 ? 1 : (c > 8 ? 8 : c);
}
size_t tc_mc_workspace_bytes(int64_t n) { return static_cast<size_t>(8) * 3 * static_cast<size_t>(n > 0 ? n : 0) * sizeof(float); }

void launch_mc_merge(const float* part, int64_t n, int T, int C, const TcOut& out, cudaStream_t st) {
  mc_merge_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(part, n, T, C, out);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
bool make_x_tensor_map(CUtensorMap* map, const float* x, int64_t n) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (encode == nullptr || !aligned16(x) || n <= 0 || n >= (static_cast<int64_t>(1) << 31) - kTcTile) return false;
  const cuuint64_t dims[2] = {PINN_N_IN, static_cast<cuuint64_t>(n)};
  const cuuint64_t strides[1] = {PINN_N_IN * sizeof(float)};
  const cuuint32_t box[2] = {PINN_N_IN, kTcTile}, estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_tc(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
              cudaStream_t st, int* err, void* workspace, size_t workspace_bytes) {
  *err = 0;
  if (const int r3 = launch_tc3(mc, net, x, n, T, dp, out, st, err, workspace, workspace_bytes); r3 != 0) return r3;
  if ((net->flags & PINN_NET_NO_TC_FWD) || net->width != kTcH || net->n_hidden < 2 || net->n_hidden > PINN_MAX_HIDDEN) return 0;
  for (int l = 1; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return 0;
  if (!aligned16(net->Wv0) || !aligned16(net->Wp)) return 0;
  const TcLayout lay = make_tc_layout(net->n_hidden);
  const size_t smem = static_cast<size_t>(lay.total) * sizeof(float);
  if (smem > 226 * 1024) return 0;          // resident weight planes: 32 KB per hidden layer (7 hidden layers fit)
  const int64_t tiles = (n + kTcTile - 1) / kTcTile;
  const int C = mc ? mc_pass_chunks(T) : 1;
  const int64_t want = tiles * C;              // one CTA per work item until the SMs run out, then two items in flight per CTA
  const int grid = static_cast<int>(want < sm_count() ? (want > 0 ? want : 1) : sm_count());
  const bool inj = dp.p > 0.f && dp.masks != nullptr;
  alignas(64) CUtensorMap xmap;
  memset(&xmap, 0, sizeof(xmap));
  // the TMA-staged input tiles (2 groups x 2 buffers x 4 KB) sit behind the resident weights when they fit
  const size_t xs_off = (smem + 127) & ~static_cast<size_t>(127), xs_bytes = static_cast<size_t>(4) * kTcTile * PINN_N_IN * sizeof(float);
  int use_tma = -1;
  if (!(net->flags & PINN_NET_NO_TMA_INPUT) && xs_off + xs_bytes <= 226 * 1024 && make_x_tensor_map(&xmap, x, n)) use_tma = static_cast<int>(xs_off);
  const size_t smem_launch = use_tma >= 0 ? xs_off + xs_bytes : smem;
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_launch));
    if (e != cudaSuccess) return e;
    kern<<<grid, kTcThreads, smem_launch, st>>>(*net, lay, x, n, T, dp, out, C, static_cast<float*>(workspace), xmap, use_tma);
    return cudaSuccess;
  };
  if (C > 1 && (workspace == nullptr || workspace_bytes < static_cast<size_t>(C) * 3 * n * sizeof(float))) { *err = PINN_E_WORKSPACE; return -1; }
  cudaError_t e;
  if (mc && C > 1) e = inj ? go(mlp_tc_kernel<true, true, true>) : go(mlp_tc_kernel<true, false, true>);
  else if (mc) e = inj ? go(mlp_tc_kernel<true, true, false>) : go(mlp_tc_kernel<true, false, false>);
  else e = inj ? go(mlp_tc_kernel<false, true, false>) : go(mlp_tc_kernel<false, false, false>);
  if (e != cudaSuccess) { *err = static_cast<int>(e); return -1; }
  if (C > 1) mc_merge_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(static_cast<const float*>(workspace), n, T, C, out);
  *err = static_cast<int>(cudaGetLastError());
  return *err == 0 ? 1 : -1;
}

}  // namespace pinn

#ifdef PINN_TIMELINE
extern "C" int pinn_debug_timeline(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, pinn::g_tl, sizeof(pinn::g_tl)));
}
#endif
