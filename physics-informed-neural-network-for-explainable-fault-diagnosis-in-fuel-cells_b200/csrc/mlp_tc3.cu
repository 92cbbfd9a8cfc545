// mlp_tc3.cu -- K1 / K4 for the 64-wide net with THREE tile groups per CTA (fp16-pair operands).
//
// mlp_tc.cu keeps two 128-sample tiles in flight per SM and its warp schedulers issue on 64 % of their cycles: with 18
// warps per SM the epilogue's dependent chains (tanh, Philox, split) are not covered.  A third group does not fit its
// tensor-memory map (3xTF32: 64 accumulator + 128 operand + 64 layer-0 columns per group).  Here the operands are
// fp16 pairs (a = a_h + a_l, a_h = fp16(a), a_l = fp16(a - a_h): 22 significant bits, the split of mlp_wide_res.cu):
//   tensor memory per group : 64 accumulator columns + 32 (A hi, packed pairs) + 32 (A lo)  -> three groups = 384 of 512
//   layer-0 park            : shared memory (32 KB fp32 per group) instead of tensor memory
//   weights                 : resident fp16 hi / lo images, 16 KB per hidden layer (K-major no-swizzle, K-PERMUTED: K slab c
//                             = the c-th 8-column chunk of both column halves, so a slab is complete when both threads of
//                             a row have handed over their c-th chunk)
//   products                : 3 kind::f16 MMAs (K = 16) per chunk hand-over, 12 per layer instead of 24 tf32 ones
// CTA = 3 x 256 compute threads + 3 MMA warps = 27 warps; 72 registers per thread.  Everything else -- thread = (row,
// 32-column half), Philox counters, the variance-head tail split between the row's two threads, Welford in registers,
// pass chunks, TMA-staged input tiles -- is mlp_tc.cu's, and the mask stream is identical.
//
// DNN.forward 01:421-438, get_MC_samples 01:1413-1491.
#include <string.h>
#include "net.cuh"
#include "tc.cuh"
#include "tc_api.cuh"

namespace pinn {

#ifndef TC3_NG
#define TC3_NG 3           // tile groups per CTA (2: the fp16-pair form of the two-group kernel, for attribution runs)
#endif
constexpr int k3H = 64, k3HH = 32, k3Tile = 128, k3HeadN = 48, k3NG = TC3_NG;
constexpr int k3Threads = k3NG * 256 + k3NG * 32;
#ifndef TC3_PREDRAW
#define TC3_PREDRAW 1      // Philox blocks of a hidden epilogue drawn ahead of the wait that precedes it
#endif
#ifndef TC3_TANH_PAIR
#define TC3_TANH_PAIR 0    // one reciprocal per pair of tanh: 8.69 vs 8.60 ms without -- with 27 warps the issue slots bind, not the XU pipe
#endif

struct Tc3Layout {  // byte offsets from the dynamic shared-memory base
  int L;
  int w[PINN_MAX_HIDDEN];    // hidden layer l >= 1: [hi plane 8 KB | lo plane 8 KB]
  int wh;                    // stacked heads [Wv0; Wp; 0] (N = 48): [hi 6 KB | lo 6 KB]
  int W0, b0, b[PINN_MAX_HIDDEN], bv0, bp, Wv1, bv1, Wv2, bv2;
  int park;                  // k3NG x [16 column quads][128 rows] float4
  int xs;                    // k3NG x 2 x 4 KB input tiles (TMA)
  int total;
};
static Tc3Layout make_tc3_layout(int L) {
  Tc3Layout t{};
  t.L = L;
  int o = 0;
  auto take = [&](int bytes, int align) { o = (o + align - 1) / align * align; const int r = o; o += bytes; return r; };
  for (int l = 1; l < L; ++l) t.w[l] = take(2 * k3H * k3H * 2, 128);
  t.wh = take(2 * k3HeadN * k3H * 2, 128);
  t.W0 = take(k3H * PINN_N_IN * 4, 16);
  t.b0 = take(k3H * 4, 16);
  for (int l = 1; l < L; ++l) t.b[l] = take(k3H * 4, 16);
  t.bv0 = take(k3HH * 4, 16);
  t.bp = take(16, 16);
  t.Wv1 = take(16 * k3HH * 4, 16);
  t.bv1 = take(16 * 4, 16);
  t.Wv2 = take(16 * 4, 16);
  t.bv2 = take(16, 16);
  t.park = take(k3NG * k3Tile * k3H * 4, 128);
  t.xs = take(k3NG * 2 * k3Tile * PINN_N_IN * 4, 128);
  t.total = o;
  return t;
}

#ifndef TC3_TRUNC_SPLIT
#define TC3_TRUNC_SPLIT 0
#endif
// fp16-pair split of two activations.  TC3_TRUNC_SPLIT: hi = the fp32 value with its low 13 mantissa bits cleared (exactly
// a normal fp16 number for |v| >= 2^-14), lo = v - hi in fp32, both packed with one F2FP each -- five instructions per pair
// instead of six, |lo| < 2^-10 |v| instead of 2^-11.
PINN_D void split3(float a, float b, uint32_t& hi, uint32_t& lo) {
#if TC3_TRUNC_SPLIT
  const float ha = __uint_as_float(__float_as_uint(a) & 0xffffe000u), hb = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
  const float2 l2 = __fadd2_rn(make_float2(a, b), make_float2(-ha, -hb));
  const __half2 h = __floats2half2_rn(ha, hb), l = __floats2half2_rn(l2.x, l2.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
#else
  tc::split_h2(a, b, hi, lo);
#endif
}

PINN_D void bar3_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
PINN_D void bar3_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// fp16 image of an [N x 64] matrix (rows beyond `rows_a` / the `src_b` row are zero), times c, K-permuted:
// kk = 16 c + 8 h + e  <->  column 32 h + 8 c + e;  byte(n, kk) = (kk / 8) * 16 N + n * 16 + (kk % 8) * 2, lo plane after hi.
PINN_D void stage_image(unsigned char* dst, const float* src_a, int rows_a, const float* src_b, int N, float c) {
  for (int idx = threadIdx.x; idx < N * 8; idx += blockDim.x) {
    const int nrow = idx % N, k8 = idx / N, cc = k8 >> 1, h = k8 & 1;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    const float* p = nullptr;
    if (nrow < rows_a) p = src_a + nrow * k3H + 32 * h + 8 * cc;
    else if (nrow == rows_a && src_b != nullptr) p = src_b + 32 * h + 8 * cc;
    if (p != nullptr) { v0 = __ldg(reinterpret_cast<const float4*>(p)); v1 = __ldg(reinterpret_cast<const float4*>(p) + 1); }
    uint4 hi, lo;
    tc::split_h2(v0.x * c, v0.y * c, hi.x, lo.x); tc::split_h2(v0.z * c, v0.w * c, hi.y, lo.y);
    tc::split_h2(v1.x * c, v1.y * c, hi.z, lo.z); tc::split_h2(v1.z * c, v1.w * c, hi.w, lo.w);
    unsigned char* q = dst + k8 * (N * 16) + nrow * 16;
    *reinterpret_cast<uint4*>(q) = hi;
    *reinterpret_cast<uint4*>(q + N * k3H * 2) = lo;
  }
}

template <bool MC, bool INJ, bool CH>
__global__ void __launch_bounds__(k3Threads, 1)
mlp_tc3_kernel(const __grid_constant__ pinn_net_t net, const __grid_constant__ Tc3Layout lay, const float* __restrict__ x, int64_t n, int T,
               const __grid_constant__ DropParams dp, TcOut out, int chunks, float* __restrict__ part,
               const __grid_constant__ CUtensorMap xmap, int use_tma) {
  constexpr int H = k3H, HH = k3HH;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t ready[k3NG][4], done[k3NG], xbar[k3NG][2];
  __shared__ uint32_t tmem_base_s;
  const int L = lay.L, tid = threadIdx.x, row = tid & 127;
  const int warp = tc::uniform_warp_idx();
  const bool mma_warp = warp >= k3NG * 8;
  const int grp = mma_warp ? warp - k3NG * 8 : warp >> 3, half = (warp >> 2) & 1;
  const int Dm = L * H + H / 2;
  const int cb = half * HH;
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;
  const float inact = drop_on ? dp.keep : 1.0f;
  auto fl = [&](int off) { return reinterpret_cast<float*>(smem + off); };

  if (tid == 0) {
    for (int g = 0; g < k3NG; ++g) {
      for (int c = 0; c < 4; ++c) tc::mbar_init(&ready[g][c], 256);
      tc::mbar_init(&done[g], 1);
      tc::mbar_init(&xbar[g][0], 1);
      tc::mbar_init(&xbar[g][1], 1);
    }
    tc::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  for (int l = 1; l < L; ++l) stage_image(smem + lay.w[l], net.W[l], H, nullptr, H, wscale);
  stage_image(smem + lay.wh, net.Wv0, HH, net.Wp, k3HeadN, wscale);
  stage_tensor_scaled(fl(lay.W0), net.W[0], H * PINN_N_IN, kTanhArg);
  stage_tensor_scaled(fl(lay.b0), net.b[0], H, kTanhArg);
  for (int l = 1; l < L; ++l) stage_tensor_scaled(fl(lay.b[l]), net.b[l], H, kTanhArg);
  stage_tensor_scaled(fl(lay.bv0), net.bv0, HH, kTanhArg);
  stage_tensor(fl(lay.bp), net.bp, 1);
  stage_tensor_scaled(fl(lay.Wv1), net.Wv1, 16 * HH, kTanhArg * wscale);
  stage_tensor_scaled(fl(lay.bv1), net.bv1, 16, kTanhArg);
  stage_tensor(fl(lay.Wv2), net.Wv2, 16);
  stage_tensor(fl(lay.bv2), net.bv2, 1);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const int64_t n_tiles = (n + k3Tile - 1) / k3Tile;
  const bool do_eval = MC && out.pred_mean != nullptr;
  const int C = (MC && CH) ? chunks : 1, Tc = (T + C - 1) / C;
  const uint32_t n_items = static_cast<uint32_t>(n_tiles * C);        // < 2^31: checked by the launcher (32-bit item arithmetic)
  auto item_passes = [&](int chunk) {
    const int t0 = chunk * Tc, cnt = (T - t0 < Tc ? T - t0 : Tc);
    return MC ? (cnt > 0 ? cnt : 0) + ((do_eval && chunk == 0) ? 1 : 0) : 1;
  };
  const uint32_t item0 = blockIdx.x + static_cast<uint32_t>(grp) * gridDim.x, item_step = gridDim.x * k3NG, uC = static_cast<uint32_t>(C);
  // tensor memory: accumulators [64 g, +64) | A hi planes [192 + 64 g, +32) | A lo planes [224 + 64 g, +32)
  const uint32_t acc_t = tmem_base_s + static_cast<uint32_t>(64 * grp);
  const uint32_t ahi_t = tmem_base_s + static_cast<uint32_t>(192 + 64 * grp), alo_t = ahi_t + 32u;

  if (mma_warp) {
    // ================================================================== MMA warp of group `grp`
    const uint32_t idesc64 = tc::make_idesc_f16(k3Tile, H), idesc48 = tc::make_idesc_f16(k3Tile, k3HeadN);
    uint32_t par = 0u;
    for (uint32_t item = item0; item < n_items; item += item_step)
      for (int it = 0, np = item_passes(static_cast<int>(item % uC)); it < np; ++it) {
#pragma unroll 1
        for (int l = 1; l <= L; ++l) {
          const bool heads = l == L;
          const int N = heads ? k3HeadN : H;
          const uint32_t wb = tc::smem_u32(smem + (heads ? lay.wh : lay.w[l]));
          const uint64_t b_hi = tc::make_desc(wb, static_cast<uint32_t>(N) * 16u, 128);
          const uint64_t b_lo = tc::make_desc(wb + static_cast<uint32_t>(N) * H * 2u, static_cast<uint32_t>(N) * 16u, 128);
          const uint32_t idesc = heads ? idesc48 : idesc64;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {          // K slab c = the c-th chunk of both column halves
            tc::mbar_wait(&ready[grp][c], par);
            __syncwarp();
            if (tc::elect_one()) {
              tc::fence_after_sync();
              const uint64_t bs = static_cast<uint64_t>(c) * static_cast<uint64_t>(2 * N);       // 2 K-chunks of 16 N bytes, in 16-byte units
              tc::umma_f16_ts(acc_t, alo_t + 8u * c, b_hi + bs, idesc, c != 0 ? 1u : 0u);      // small terms first
              tc::umma_f16_ts(acc_t, ahi_t + 8u * c, b_lo + bs, idesc, 1u);
              tc::umma_f16_ts(acc_t, ahi_t + 8u * c, b_hi + bs, idesc, 1u);
              if (c == 3) tc::umma_commit(&done[grp]);
            }
            __syncwarp();
          }
          par ^= 1u;
        }
      }
  } else {
    // ================================================================== compute groups
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t d_lane = acc_t + lane_sel, x_lane = d_lane + 48u;
    const uint32_t ahi_l = ahi_t + lane_sel + static_cast<uint32_t>(4 * half), alo_l = ahi_l + 32u;
    float4* const park = reinterpret_cast<float4*>(smem + lay.park) + static_cast<size_t>(grp) * (16 * k3Tile) + row;   // [c4][row]
    float* const xs = fl(lay.xs) + static_cast<size_t>(grp) * 2 * (k3Tile * PINN_N_IN);
    uint32_t phase = 0;

    // chunk `chunk` (8 columns) of this thread's half -> packed fp16 pairs in both A planes; the hand-over of the PREVIOUS
    // chunk goes out between the split and the stores of this one (its stores have long landed)
    auto store8 = [&](const float (&v)[8], int chunk) {
      uint32_t h[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split3(v[2 * e], v[2 * e + 1], h[e], lo[e]);
      if (chunk > 0) {
        tc::tmem_wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&ready[grp][chunk - 1]);
      }
      tc::tmem_st4(ahi_l + 8u * chunk, reinterpret_cast<const float*>(h));
      tc::tmem_st4(alo_l + 8u * chunk, reinterpret_cast<const float*>(lo));
    };
    auto draw = [&](uint4* r, int nblk, const KeepSrc<INJ>& ks, uint32_t pass, uint32_t layer, int c0) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nblk) r[c] = Philox::gen_rk(dp.rk, ks.s_lo, ks.s_hi, pass, (layer << 16) | static_cast<uint32_t>((c0 >> 3) + c));
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nblk) asm volatile("" : "+r"(r[c].x), "+r"(r[c].y), "+r"(r[c].z), "+r"(r[c].w));
    };
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
    const bool x_leader = use_tma && (tid & 255) == 0;
    auto request_x = [&](uint32_t it, int buf) {
      tc::mbar_expect_tx(&xbar[grp][buf], k3Tile * PINN_N_IN * sizeof(float));
      tc::tma_load_2d(xs + buf * (k3Tile * PINN_N_IN), &xmap, 0, static_cast<int>((it / uC) * k3Tile), &xbar[grp][buf]);
    };
    if (x_leader && item0 < n_items) request_x(item0, 0);
    uint32_t xcount = 0;

    for (uint32_t item = item0; item < n_items; item += item_step, ++xcount) {
      const int64_t tile = item / uC;
      const int chunk = static_cast<int>(item % uC), t0 = chunk * Tc;
      const bool eval_item = do_eval && chunk == 0;
      const int n_pass = item_passes(chunk);
      const int64_t s = tile * k3Tile + row;
      const bool valid = s < n;
      // layer 0 (pass-invariant): this thread's 32 columns -> the group's park in shared memory
      {
        float xr[PINN_N_IN];
        if (use_tma) {
          const int buf = static_cast<int>(xcount & 1u);
          if (x_leader && item + item_step < n_items) request_x(item + item_step, buf ^ 1);
          tc::mbar_wait(&xbar[grp][buf], (xcount >> 1) & 1u);
          const float4* px = reinterpret_cast<const float4*>(xs + buf * (k3Tile * PINN_N_IN) + row * PINN_N_IN);
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
        const float* W0 = fl(lay.W0) + cb * PINN_N_IN;
        const float* b0 = fl(lay.b0) + cb;
#pragma unroll 2
        for (int j4 = 0; j4 < 8; ++j4) {
          float o4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = 4 * j4 + q;
            const float4 w0 = *reinterpret_cast<const float4*>(W0 + j * PINN_N_IN);
            const float4 w1 = *reinterpret_cast<const float4*>(W0 + j * PINN_N_IN + 4);
            float z = b0[j];
            z = fmaf(w0.x, xr[0], z); z = fmaf(w0.y, xr[1], z); z = fmaf(w0.z, xr[2], z); z = fmaf(w0.w, xr[3], z);
            z = fmaf(w1.x, xr[4], z); z = fmaf(w1.y, xr[5], z); z = fmaf(w1.z, xr[6], z); z = fmaf(w1.w, xr[7], z);
            o4[q] = tanh_pre(z);
          }
          park[(8 * half + j4) * k3Tile] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
      }

      float mean = 0.f, m2 = 0.f, slv = 0.f;
      const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
      KeepSrc<INJ> ks;
      ks.s_lo = static_cast<uint32_t>(sg); ks.s_hi = static_cast<uint32_t>(sg >> 32);
      uint4 r0[4] = {};
      if (!INJ && drop_on && !eval_item) draw(r0, 4, ks, static_cast<uint32_t>(dp.pass_offset + t0), 0u, cb);
#pragma unroll 1
      for (int pi = 0; pi < n_pass; ++pi) {
        const bool eval_pass = MC && eval_item && pi == 0;
        const int tl = MC ? (eval_item ? pi - 1 : pi) : 0;
        const int t = t0 + tl;
        const bool active = drop_on && !eval_pass && (!INJ || valid);
        ks.pass = static_cast<uint32_t>(dp.pass_offset + t);
        ks.mrow = INJ ? dp.masks + (static_cast<size_t>(t) * dp.mask_n + (valid ? s : 0)) * Dm : nullptr;
        // ---- stage layer-0 activations (masked) as the first A operand
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 f0 = park[(8 * half + 2 * c) * k3Tile], f1 = park[(8 * half + 2 * c + 1) * k3Tile];
          const float t8[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          float v[8];
          select8(r0[c], ks, active, 0u, cb + 8 * c, t8, v);
          store8(v, c);
        }
        // ---- hidden layers on the tensor cores
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
          signal_ready();
          uint4 rl[4] = {};
          if (TC3_PREDRAW && !INJ && active) draw(rl, 4, ks, ks.pass, static_cast<uint32_t>(l), cb);
          wait_done();
          const float* bl = fl(lay.b[l]) + cb;
          // ALL of this thread's accumulator columns come out before its first hand-over: the next layer's first product
          // (issued as soon as every thread has handed over chunk 0) overwrites the accumulator
          float z[HH];
          tc::tmem_ld16(d_lane + cb, z);
          tc::tmem_ld16(d_lane + cb + 16, z + 16);
          tc::tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 bA = *reinterpret_cast<const float4*>(bl + 8 * c), bB = *reinterpret_cast<const float4*>(bl + 8 * c + 4);
            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
            float t8[8], v[8];
            tanh8_prescaled<TC3_TANH_PAIR != 0>(z + 8 * c, bb, t8);
            if (!TC3_PREDRAW && !INJ && active)
              rl[c] = Philox::gen_rk(dp.rk, ks.s_lo, ks.s_hi, ks.pass, (static_cast<uint32_t>(l) << 16) | static_cast<uint32_t>((cb >> 3) + c));
            select8(rl[c], ks, active, static_cast<uint32_t>(l), cb + 8 * c, t8, v);
            store8(v, c);
          }
        }
        // ---- heads: [Wv0; Wp] in one N = 48 product
        signal_ready();
        uint4 rv[2] = {};
        if (!INJ && active) draw(rv, 2, ks, ks.pass, static_cast<uint32_t>(L), 16 * half);
        if (!INJ && drop_on && pi + 1 < n_pass) draw(r0, 4, ks, static_cast<uint32_t>(dp.pass_offset + t + 1), 0u, cb);
        wait_done();
        {
          float v0[16], part16[16];
          float u = 0.f;
          tc::tmem_ld16(d_lane + 16 * half, v0);
          if (half == 0) { float zz[4]; tc::tmem_ld4(d_lane + 32, zz); tc::tmem_wait_ld(); u = zz[0] + fl(lay.bp)[0]; }
          else tc::tmem_wait_ld();
          const float* bv0 = fl(lay.bv0) + 16 * half;
#pragma unroll
          for (int g = 0; g < 16; g += 8) {
            const float4 bA = *reinterpret_cast<const float4*>(bv0 + g), bB = *reinterpret_cast<const float4*>(bv0 + g + 4);
            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
            float t8[8], v[8];
            tanh8_prescaled<TC3_TANH_PAIR != 0>(v0 + g, bb, t8);
            select8(rv[g / 8], ks, active, static_cast<uint32_t>(L), 16 * half + g, t8, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) v0[g + q] = v[q];
          }
          const float* Wv1 = fl(lay.Wv1) + 16 * half;     // pre-scaled by 2 log2(e) / (1-p)
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float2 acc = make_float2(0.f, 0.f), acc2 = acc;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 w = *reinterpret_cast<const float4*>(Wv1 + k * HH + 4 * i4);
              acc = ffma2(make_float2(w.x, w.y), make_float2(v0[4 * i4], v0[4 * i4 + 1]), acc);
              acc2 = ffma2(make_float2(w.z, w.w), make_float2(v0[4 * i4 + 2], v0[4 * i4 + 3]), acc2);
            }
            part16[k] = (acc.x + acc.y) + (acc2.x + acc2.y);
          }
          if (half == 1) {
            tc::tmem_st16(x_lane, part16);
            tc::tmem_wait_st();
            tc::fence_before_sync();
            bar3_arrive(1 + grp, 256);
          } else {
            bar3_sync(1 + grp, 256);
            tc::fence_after_sync();
            float p1[16];
            tc::tmem_ld16(x_lane, p1);
            tc::tmem_wait_ld();
            float vraw = fl(lay.bv2)[0];
            const float* bv1 = fl(lay.bv1);
            const float* Wv2 = fl(lay.Wv2);
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
              const float2 a1 = tanh_pre2<true>(make_float2((part16[k] + p1[k]) + bv1[k], (part16[k + 1] + p1[k + 1]) + bv1[k + 1]));
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
          }
        }
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

// 1: launched; 0: shape not covered / opted out (mlp_tc.cu takes the call); -1: error in *err.
int launch_tc3(bool mc, const pinn_net_t* net, const float* x, int64_t n, int T, const DropParams& dp, const TcOut& out,
               cudaStream_t st, int* err, void* workspace, size_t workspace_bytes) {
  *err = 0;
  if ((net->flags & (PINN_NET_NO_TC_FWD | PINN_NET_NO_TC3)) || net->width != k3H || net->n_hidden < 2 || n <= 0) return 0;
  for (int l = 0; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return 0;
  if (!aligned16(net->Wv0) || !aligned16(net->Wp) || !aligned16(x)) return 0;
  const Tc3Layout lay = make_tc3_layout(net->n_hidden);
  if (lay.total > 226 * 1024) return 0;          // up to six hidden layers next to three layer-0 parks
  const int64_t tiles = (n + k3Tile - 1) / k3Tile;
  const int C = mc ? mc_pass_chunks(T) : 1;
  const int64_t want = tiles * C;
  if (want >= (static_cast<int64_t>(1) << 31) - 4 * sm_count()) return 0;      // the kernel indexes work items with 32 bits
  const int grid = static_cast<int>(want < sm_count() ? (want > 0 ? want : 1) : sm_count());
  const bool inj = dp.p > 0.f && dp.masks != nullptr;
  if (C > 1 && (workspace == nullptr || workspace_bytes < static_cast<size_t>(C) * 3 * n * sizeof(float))) { *err = PINN_E_WORKSPACE; return -1; }
  alignas(64) CUtensorMap xmap;
  memset(&xmap, 0, sizeof(xmap));
  const int use_tma = (net->flags & PINN_NET_NO_TMA_INPUT) ? 0 : (make_x_tensor_map(&xmap, x, n) ? 1 : 0);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total);
    if (e != cudaSuccess) return e;
    kern<<<grid, k3Threads, lay.total, st>>>(*net, lay, x, n, T, dp, out, C, static_cast<float*>(workspace), xmap, use_tma);
    return cudaSuccess;
  };
  cudaError_t e;
  if (mc && C > 1) e = inj ? go(mlp_tc3_kernel<true, true, true>) : go(mlp_tc3_kernel<true, false, true>);
  else if (mc) e = inj ? go(mlp_tc3_kernel<true, true, false>) : go(mlp_tc3_kernel<true, false, false>);
  else e = inj ? go(mlp_tc3_kernel<false, true, false>) : go(mlp_tc3_kernel<false, false, false>);
  if (e != cudaSuccess) { *err = static_cast<int>(e); return -1; }
  if (C > 1) launch_mc_merge(static_cast<const float*>(workspace), n, T, C, out, st);
  *err = static_cast<int>(cudaGetLastError());
  return *err == 0 ? 1 : -1;
}

}  // namespace pinn
