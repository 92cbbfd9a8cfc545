// mlp_tc.cu -- tensor-core (tcgen05 + TMEM) path of K1 / K4 for the headline 64-wide net.
//
// CTA = 2 groups x 256 threads; each group owns one 128-sample tile at a time and runs
// independently (named barrier + its own mbarrier), so one group's MMAs overlap the other
// group's CUDA-core epilogue.  Threads (r, 0) and (r, 1) of a group own the two 32-column
// halves of sample r of the tile (= TMEM lane r); Philox counters are per (sample, pass,
// layer, unit/8), so the mask stream is identical to the thread-per-sample kernels -- only
// the 64x64 contractions moved:
//
//   weights  : every hidden layer W_l (l>=1) and the stacked head matrix [Wv0; Wp; 0] are
//              split once per CTA into tf32 hi/lo planes (UMMA K-major layout) and stay
//              resident in shared memory for all tiles and passes;
//   per layer: the group's activations sit as hi/lo A planes; one elected thread issues
//              3 x 8 tcgen05.mma (lo*hi + hi*lo + hi*hi, fp32 accumulate in TMEM) and commits
//              to the group's mbarrier; all 128 threads then pull their row out of TMEM
//              (tcgen05.ld 32x32b), add bias, tanh, draw the Philox mask, re-split and store
//              the next layer's A planes (conflict-free 128-bit stores);
//   heads    : one N=48 MMA gives the 32 variance-head pre-activations and the mean head;
//              the 32->16->1 tail is 528 FMAs per sample on CUDA cores.
//
// Layer 0 (K = 8) is pass-invariant: computed once per sample into registers.
#include "net.cuh"
#include "tc.cuh"
#include "tc_api.cuh"

namespace pinn {

constexpr int kTcH = 64;
constexpr int kTcTile = 128;
constexpr int kHeadN = 48;  // 32 variance-head rows + 1 mean-head row + 15 zero rows (N % 16 == 0)

struct TcLayout {  // offsets in floats from the dynamic shared-memory base
  int L, nwg;
  int b_hi[PINN_MAX_HIDDEN], b_lo[PINN_MAX_HIDDEN];  // hidden layer l >= 1
  int h_hi, h_lo;                                    // stacked heads
  int a_hi[2], a_lo[2];                              // per warpgroup
  int W0, b0, b[PINN_MAX_HIDDEN], bv0, bp, Wv1, bv1, Wv2, bv2;
  int total;
};
PINN_HD TcLayout make_tc_layout(int L, int nwg) {
  TcLayout t;
  t.L = L; t.nwg = nwg;
  int o = 0;
  for (int l = 0; l < PINN_MAX_HIDDEN; ++l) { t.b_hi[l] = t.b_lo[l] = t.b[l] = 0; }
  for (int l = 1; l < L; ++l) { t.b_hi[l] = o; o += kTcH * kTcH; t.b_lo[l] = o; o += kTcH * kTcH; }
  t.h_hi = o; o += kHeadN * kTcH;
  t.h_lo = o; o += kHeadN * kTcH;
  for (int g = 0; g < 2; ++g) t.a_hi[g] = t.a_lo[g] = 0;      // activations live in tensor memory
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

// named barrier for one 256-thread group (ids 1, 2; id 0 is __syncthreads)
PINN_D void grp_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

#ifdef PINN_TIMELINE
// Debug build (profiles/timeline_mc.py): clock stamps of CTA 0 / group 0, warps 0 (half 0) and 4 (half 1).
__device__ long long g_tl[2][64][8];
#define TL(slot, k) do { if (tl_on && tl_i < 64) g_tl[half][tl_i][k] = clock64(); } while (0)
#else
#define TL(slot, k) do {} while (0)
#endif

PINN_D void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
PINN_D void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// MC = true : eval pass (if pred_mean) + T dropout passes with Welford.   MC = false: one pass.
// INJ = true: keep decisions come from an injected mask tensor (parity runs) instead of Philox.
// A "group" is 256 threads working on one 128-sample tile: thread (row, half) owns columns
// [32*half, 32*half+32) of sample `row`'s activations (TMEM lane `row`), so 16 warps per SM
// keep the schedulers fed while each thread's working set stays at 32 values.
//
// Instruction diet of the epilogue (it, not the tensor pipe, bounds this kernel):
//  * the dropout scale 1/(1-p) is folded into the resident weight planes (and into Wv1), so a
//    kept activation is stored as is: one FSEL per unit instead of FSEL + FMUL; passes without
//    dropout (the eval pass) multiply by (1-p) instead;
//  * biases (and layer 0's weights) are pre-scaled by 2 log2(e): tanh = FFMA, EX2, FADD, RCP, FFMA;
//  * Philox round keys are constant-bank operands, 16-bit draws are compared in place;
//  * the tf32 split of an activation is IADD + LOP3 + FADD.
template <bool MC, bool INJ>
__global__ void __launch_bounds__(512, 1)
mlp_tc_kernel(pinn_net_t net, TcLayout lay, const float* __restrict__ x, int64_t n, int T, const __grid_constant__ DropParams dp,
              TcOut out) {
  constexpr int H = kTcH, HH = kTcH / 2;
  constexpr uint32_t LBO_A = kTcTile * 16, LBO_B = H * 16, LBO_H = kHeadN * 16;
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  const int L = lay.L, tid = threadIdx.x, row = tid & 127;
  const int warp = tc::uniform_warp_idx(), grp = warp >> 3, half = (warp >> 2) & 1;  // warp-uniform roles
  const int ngrp = blockDim.x >> 8;
  const int Dm = L * H + H / 2;
  const int cb = half * HH;  // first column owned by this thread
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;    // folded into every matrix that consumes masked activations
  const float inact = drop_on ? dp.keep : 1.0f;      // multiplier of an un-masked activation (undoes the fold)

  // ---------------------------------------------------------------- one-time CTA set-up
  if (tid == 0) {
    tc::mbar_init(&mbar[0], 1);
    tc::mbar_init(&mbar[1], 1);
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

  const uint32_t d_tmem = tmem_base_s + static_cast<uint32_t>(grp * 64);           // this group's 64 columns
  const uint32_t d_lane = d_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16); // this warp's 32 lanes
  // Tensor-memory map (512 columns, one CTA per SM): accumulators [64 g, +64); activation planes of
  // group g: hi [128 + 128 g, +64), lo [192 + 128 g, +64); tail hand-over [384 + 16 g, +16).
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t x_lane = tmem_base_s + static_cast<uint32_t>(384 + grp * 16) + lane_sel;
  const uint32_t a_hi_t = tmem_base_s + static_cast<uint32_t>(128 + grp * 128), a_lo_t = a_hi_t + 64u;
  const uint64_t h_hi_d = tc::make_desc(tc::smem_u32(smem + lay.h_hi), LBO_H, 128);
  const uint64_t h_lo_d = tc::make_desc(tc::smem_u32(smem + lay.h_lo), LBO_H, 128);
  const uint32_t idesc64 = tc::make_idesc_tf32(kTcTile, H), idesc48 = tc::make_idesc_tf32(kTcTile, kHeadN);
  const bool issuer_warp = (warp & 7) == 0;
  uint32_t phase = 0;
#ifdef PINN_TIMELINE
  const bool tl_on = blockIdx.x == 0 && grp == 0 && (warp & 3) == 0 && (tid & 31) == 0;
  int tl_i = 0;
#endif

  // activations t[0..8) of columns c0.. -> masked (or rescaled) -> split -> A planes
  auto store8 = [&](int c0, const float (&v)[8]) {        // split -> this row's hi / lo columns in tensor memory
    float h[8], lo[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { h[q] = tc::tf32_hi_fast(v[q]); lo[q] = v[q] - h[q]; }
    tc::tmem_st8(a_hi_t + lane_sel + static_cast<uint32_t>(c0), h);
    tc::tmem_st8(a_lo_t + lane_sel + static_cast<uint32_t>(c0), lo);
  };
  // Philox blocks are drawn AHEAD of the MMA wait that precedes their use (they depend on nothing the
  // tensor core produces), so the integer work fills the group's otherwise idle wait window.
  auto draw4 = [&](uint4 (&r)[4], const KeepSrc<INJ>& ks, uint32_t pass, uint32_t layer, int c0) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      r[c] = Philox::gen_rk(dp.rk, ks.s_lo, ks.s_hi, pass, (layer << 16) | static_cast<uint32_t>((c0 >> 3) + c));
#pragma unroll
    for (int c = 0; c < 4; ++c) asm volatile("" : "+r"(r[c].x), "+r"(r[c].y), "+r"(r[c].z), "+r"(r[c].w));   // pin before the wait
  };
  auto stage8r = [&](const uint4& r, const KeepSrc<INJ>& ks, bool active, uint32_t layer, int c0, const float (&t)[8]) {
    float v[8];
    if (INJ || !active) {
      if (active) {
        bool k[8];
        ks.get8(dp, layer, static_cast<uint32_t>(c0), layer * H, k);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = k[q] ? t[q] : 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = t[q] * inact;
      }
    } else {
      bool k[8];
      keep8_from(r, dp.thresh_hi, k);
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = k[q] ? t[q] : 0.f;
    }
    store8(c0, v);
  };
  auto stage8 = [&](const KeepSrc<INJ>& ks, bool active, uint32_t layer, int c0, const float (&t)[8]) {
    float v[8];
    if (active) {
      bool k[8];
      ks.get8(dp, layer, static_cast<uint32_t>(c0), layer * H, k);
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = k[q] ? t[q] : 0.f;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = t[q] * inact;
    }
    store8(c0, v);
  };

  // ---------------------------------------------------------------- tiles of this group
  const int64_t n_tiles = (n + kTcTile - 1) / kTcTile;
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * ngrp + grp; tile < n_tiles;
       tile += static_cast<int64_t>(gridDim.x) * ngrp) {
    const int64_t s = tile * kTcTile + row;
    const bool valid = s < n;
    // layer 0 into registers (pass-invariant): this thread's 32 columns
    float a0[HH];
    {
      float xr[PINN_N_IN];
      if (valid) {
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
      for (int j = 0; j < HH; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(W0 + j * PINN_N_IN);
        const float4 w1 = *reinterpret_cast<const float4*>(W0 + j * PINN_N_IN + 4);
        float z = b0[j];
        z = fmaf(w0.x, xr[0], z); z = fmaf(w0.y, xr[1], z); z = fmaf(w0.z, xr[2], z); z = fmaf(w0.w, xr[3], z);
        z = fmaf(w1.x, xr[4], z); z = fmaf(w1.y, xr[5], z); z = fmaf(w1.z, xr[6], z); z = fmaf(w1.w, xr[7], z);
        a0[j] = tanh_pre(z);
      }
    }

    float mean = 0.f, m2 = 0.f, slv = 0.f;
    const bool do_eval = MC && out.pred_mean != nullptr;
    const int n_pass = MC ? T + (do_eval ? 1 : 0) : 1;
    const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
    uint4 r0[4] = {};        // draws of the coming pass's layer-0 staging
    if (!INJ && drop_on && !do_eval) {
      KeepSrc<INJ> k0;
      k0.s_lo = static_cast<uint32_t>(sg); k0.s_hi = static_cast<uint32_t>(sg >> 32); k0.pass = 0; k0.mrow = nullptr;
      draw4(r0, k0, static_cast<uint32_t>(dp.pass_offset), 0u, cb);
    }
    for (int pi = 0; pi < n_pass; ++pi) {
      const bool eval_pass = MC && do_eval && pi == 0;
      const int t = MC ? (do_eval ? pi - 1 : pi) : 0;
      // dropout is active on this pass?  (injected masks: tail rows of the last tile have no mask row)
      const bool active = drop_on && !eval_pass && (!INJ || valid);
      KeepSrc<INJ> ks;
      ks.s_lo = static_cast<uint32_t>(sg); ks.s_hi = static_cast<uint32_t>(sg >> 32);
      ks.pass = static_cast<uint32_t>(dp.pass_offset + t);
      ks.mrow = INJ ? dp.masks + (static_cast<size_t>(t) * dp.mask_n + (valid ? s : 0)) * Dm : nullptr;
      // ---- stage layer-0 activations (masked) as the first A operand
#pragma unroll
      for (int g = 0; g < HH; g += 8) {
        const float t8[8] = {a0[g], a0[g + 1], a0[g + 2], a0[g + 3], a0[g + 4], a0[g + 5], a0[g + 6], a0[g + 7]};
        stage8r(r0[g / 8], ks, active, 0u, cb + g, t8);
      }
      // ---- hidden layers on the tensor cores
      for (int l = 1; l < L; ++l) {
        TL(tl_i, 0);
        tc::tmem_wait_st();
        tc::fence_before_sync();
        grp_sync(grp);
        TL(tl_i, 1);
        if (issuer_warp) {
          const uint64_t b_hi_d = tc::make_desc(tc::smem_u32(smem + lay.b_hi[l]), LBO_B, 128);
          const uint64_t b_lo_d = tc::make_desc(tc::smem_u32(smem + lay.b_lo[l]), LBO_B, 128);
          if (tc::elect_one()) {
            tc::fence_after_sync();
            tc::issue_3xtf32_ts<H>(d_tmem, a_hi_t, a_lo_t, b_hi_d, b_lo_d, LBO_B, idesc64);
            tc::umma_commit(&mbar[grp]);
          }
          __syncwarp();
        }
        TL(tl_i, 2);
        uint4 rl[4] = {};
        if (!INJ && active) draw4(rl, ks, ks.pass, static_cast<uint32_t>(l), cb);
        tc::mbar_wait(&mbar[grp], phase);
        phase ^= 1u;
        __syncwarp();
        tc::fence_after_sync();
        TL(tl_i, 3);
        const float* bl = smem + lay.b[l] + cb;
        float z[HH];
        tc::tmem_ld16(d_lane + cb, z);
        tc::tmem_ld16(d_lane + cb + 16, z + 16);
        tc::tmem_wait_ld();
        TL(tl_i, 4);
#pragma unroll
        for (int g = 0; g < HH; g += 8) {
          const float4 bA = *reinterpret_cast<const float4*>(bl + g), bB = *reinterpret_cast<const float4*>(bl + g + 4);
          const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
          float t8[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) t8[q] = tanh_pre(fmaf(z[g + q], kTanhArg, bb[q]));
          stage8r(rl[g / 8], ks, active, static_cast<uint32_t>(l), cb + g, t8);
        }
        TL(tl_i, 5);
#ifdef PINN_TIMELINE
        ++tl_i;
#endif
      }
      // ---- heads: [Wv0; Wp] in one N = 48 MMA
      tc::tmem_wait_st();
      tc::fence_before_sync();
      grp_sync(grp);
      if (issuer_warp) {
        if (tc::elect_one()) {
          tc::fence_after_sync();
          tc::issue_3xtf32_ts<H>(d_tmem, a_hi_t, a_lo_t, h_hi_d, h_lo_d, LBO_H, idesc48);
          tc::umma_commit(&mbar[grp]);
        }
        __syncwarp();
      }
      uint4 rv[2] = {};
      if (!INJ && active) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          rv[c] = Philox::gen_rk(dp.rk, ks.s_lo, ks.s_hi, ks.pass, (static_cast<uint32_t>(L) << 16) | static_cast<uint32_t>(2 * half + c));
#pragma unroll
        for (int c = 0; c < 2; ++c) asm volatile("" : "+r"(rv[c].x), "+r"(rv[c].y), "+r"(rv[c].z), "+r"(rv[c].w));
      }
      if (!INJ && drop_on && pi + 1 < n_pass) draw4(r0, ks, static_cast<uint32_t>(dp.pass_offset + t + 1), 0u, cb);
      tc::mbar_wait(&mbar[grp], phase);
      phase ^= 1u;
      __syncwarp();
      tc::fence_after_sync();
      {
        // Variance head, split between the row's two threads: each activates 16 of the 32 head units
        // and forms its share of the sixteen 32 -> 16 sums; thread (row, 1) parks its share in spare
        // TMEM columns of the row's lane and moves on to the next pass, thread (row, 0) adds the two
        // shares and finishes the sample.  Hand-over = named barrier (producers arrive, consumers sync).
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
          float t8[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) t8[q] = tanh_pre(fmaf(v0[g + q], kTanhArg, bb[q]));
          if (active) {
            bool k[8];
            if (INJ) ks.get8(dp, static_cast<uint32_t>(L), static_cast<uint32_t>(16 * half + g), static_cast<uint32_t>(L * H), k);
            else keep8_from(rv[g / 8], dp.thresh_hi, k);
#pragma unroll
            for (int q = 0; q < 8; ++q) v0[g + q] = k[q] ? t8[q] : 0.f;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v0[g + q] = t8[q] * inact;
          }
        }
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
        if (half == 1) {
          tc::tmem_st16(x_lane, part);
          tc::tmem_wait_st();
          tc::fence_before_sync();
          bar_arrive_n(3 + grp, 256);
        } else {
          bar_sync_n(3 + grp, 256);
          tc::fence_after_sync();
          float p1[16];
          tc::tmem_ld16(x_lane, p1);
          tc::tmem_wait_ld();
          float vraw = smem[lay.bv2];
          const float* bv1 = smem + lay.bv1;
          const float* Wv2 = smem + lay.Wv2;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float a1 = tanh_pre((part[k] + p1[k]) + bv1[k]);
            vraw = fmaf(Wv2[k], a1, vraw);
          }
          const float lv = logvar_from_v(vraw);
          if (!MC) {
            if (valid) { out.u[s] = u; out.s[s] = lv; }
          } else if (eval_pass) {
            if (valid) out.pred_mean[s] = u;
          } else {
            const float d = u - mean;
            mean += d / static_cast<float>(t + 1);
            m2 = fmaf(d, u - mean, m2);
            slv += lv;
          }
        }
      }
    }
    if (MC && valid && half == 0) {
      if (out.raw_mean) out.raw_mean[s] = mean;
      if (out.raw_m2) out.raw_m2[s] = m2;
      if (out.raw_slv) out.raw_slv[s] = slv;
      const float invT = 1.0f / static_cast<float>(T > 0 ? T : 1);
      if (out.a_u) out.a_u[s] = sqrtf(expf(slv * invT));
      if (out.e_u) out.e_u[s] = sqrtf(fmaxf(m2, 0.f) * invT);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

static int g_tc_enabled = 1;


// Launch helper used by pinn_mlp_fwd / pinn_mc_dropout.  Returns 1 if the TC path took the
// call, 0 if the shape is not covered (caller falls through to the FFMA kernels), <0 / >1 on error.
int launch_tc(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
              cudaStream_t st, int* err) {
  *err = 0;
  if (!g_tc_enabled || net->width != kTcH || net->n_hidden < 2 || net->n_hidden > 5) return 0;
  for (int l = 1; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return 0;
  if (!aligned16(net->Wv0) || !aligned16(net->Wp)) return 0;
  int nwg = 2;  // 256-thread groups (one 128-sample tile each) per CTA
  TcLayout lay = make_tc_layout(net->n_hidden, nwg);
  if (static_cast<size_t>(lay.total) * sizeof(float) > 226 * 1024) {
    nwg = 1;
    lay = make_tc_layout(net->n_hidden, nwg);
    if (static_cast<size_t>(lay.total) * sizeof(float) > 226 * 1024) return 0;
  }
  const size_t smem = static_cast<size_t>(lay.total) * sizeof(float);
  const int64_t tiles = (n + kTcTile - 1) / kTcTile;
  int64_t want = (tiles + nwg - 1) / nwg;
  const int grid = static_cast<int>(want < sm_count() ? (want > 0 ? want : 1) : sm_count());
  const bool inj = dp.p > 0.f && dp.masks != nullptr;
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, 256 * nwg, smem, st>>>(*net, lay, x, n, T, dp, out);
    return cudaSuccess;
  };
  cudaError_t e;
  if (mc) e = inj ? go(mlp_tc_kernel<true, true>) : go(mlp_tc_kernel<true, false>);
  else e = inj ? go(mlp_tc_kernel<false, true>) : go(mlp_tc_kernel<false, false>);
  if (e != cudaSuccess) { *err = static_cast<int>(e); return -1; }
  *err = static_cast<int>(cudaGetLastError());
  return *err == 0 ? 1 : -1;
}

}  // namespace pinn

#ifdef PINN_TIMELINE
extern "C" int pinn_debug_timeline(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, pinn::g_tl, sizeof(pinn::g_tl)));
}
#endif
// Test / ablation switch: 0 routes the 64-wide net through the FFMA kernels as well.
extern "C" int pinn_set_tensor_core_path(int enable) {
  int prev = pinn::g_tc_enabled;
  pinn::g_tc_enabled = enable ? 1 : 0;
  return prev;
}
