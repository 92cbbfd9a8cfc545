// mlp_tc_bwd.cu -- K2 for the 64-wide net on the tensor cores, as two kernels:
//
//  K2a  mlp_tc_bwd_kernel : per 128-sample tile (two independent 256-thread groups per CTA,
//       as in mlp_tc.cu) forward recompute AND dgrad on tcgen05 (3xTF32, fp32-accurate):
//         stage a0 -> [W1] -> a1 -> [W2] -> a2 -> [Wv0;Wp] -> heads + loss gradients
//         -> [ (Wv0;Wp)^T ] -> d a2 -> [W2^T] -> d a1 -> [W1^T] -> d a0
//       Every contraction is a K-major x K-major MMA: the B operand (one layer's weights, or
//       their transpose for dgrad) is re-split into the group's 32 KB B planes right before
//       use -- 16 values per thread -- instead of keeping 2 x 88 KB of planes resident, so
//       two groups fit in 192 KB and overlap each other's MMAs and epilogues.  Masked
//       activations and pre-activation deltas of every layer go to HBM ([N][width] fp32 rows,
//       128 B per thread per layer); keep-bits ride in registers so Philox runs once.
//  K2b  wgrad_kernel : dW_l = delta_l^T a_{l-1} for all seven weight tensors and the bias
//       column sums, contraction over the batch, 8x8 FFMA2 register tiles over 16-sample stages
//       that TMA bulk copies (cp.async.bulk + mbarrier transaction bytes) double-buffer into
//       shared memory; per-CTA partials -> grad_reduce_kernel (fixed-order,
//       deterministic).
//
// (MN-major TF32 operands, which would allow an all-on-chip variant, require CUTLASS's
// SW128_32B swizzled layout; with the plain interleaved layout the MMA is silently dropped --
// tests/cuda/tc_mn_test.cu documents that.)
#include "net.cuh"
#include "tc.cuh"
#include "tc_api.cuh"

namespace pinn {

constexpr int kBH = 64, kBTile = 128;

// Global scratch: per-sample rows written by K2a, read by K2b (float offsets per sample).
struct BwdScratch {
  float* act[PINN_MAX_HIDDEN];   // masked activations a_l            [n][64]
  float* del[PINN_MAX_HIDDEN];   // pre-activation deltas dz_l        [n][64]
  float* av0; float* dv0;        // variance head layer 0             [n][32]
  float* av1; float* dv1;        // variance head layer 1             [n][16]
  float* du;  float* dvs;        // d loss / d u, d loss / d v        [n]
};
PINN_HD size_t bwd_scratch_floats_per_sample(int L) { return static_cast<size_t>(L) * 2 * kBH + 2 * 32 + 2 * 16 + 2; }

inline BwdScratch carve_scratch(float* base, int64_t n, int L) {
  BwdScratch s{};
  float* p = base;
  const size_t N = (static_cast<size_t>(n) + 3) & ~static_cast<size_t>(3);   // every array 16-byte aligned (TMA sources)
  for (int l = 0; l < L; ++l) { s.act[l] = p; p += N * kBH; }
  for (int l = 0; l < L; ++l) { s.del[l] = p; p += N * kBH; }
  s.av0 = p; p += N * 32; s.dv0 = p; p += N * 32;
  s.av1 = p; p += N * 16; s.dv1 = p; p += N * 16;
  s.du = p; p += N; s.dvs = p; p += N;
  return s;
}

// ------------------------------------------------------------------------------- K2b
// One CTA accumulates every gradient over its samples; thread (jt, kt) of a 16 x 16 grid owns
// the 4 x 4 block (j0 = 4 jt, k0 = 4 kt) of each 64 x 64 product and sub-blocks of the rest.
constexpr int kWgS = 16;   // samples per shared-memory stage (two stages in flight)

struct WgradArgs {
  const float* x; int64_t n; int L;
  BwdScratch sc;
  float* partial;          // [grid][lay.total]
  float act_scale;         // dropout scale 1/(1-p) missing from the stored trunk activations (K2a)
};

PINN_D void cp_async16(float* dst_smem, const float* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))),
               "l"(src_gmem) : "memory");
}
PINN_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PINN_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// floats per stage: D[L][S][64], A[L][S][64], X[S][8], DV0[S][32], AV0[S][32], DV1[S][16], AV1[S][16], DU[S], DVS[S]
PINN_HD constexpr int wg_stage_floats(int L) { return 2 * L * kWgS * 64 + kWgS * (8 + 32 + 32 + 16 + 16 + 2); }

// Issue the asynchronous copies of one stage (rows beyond `cnt` are zero-filled with plain stores).
template <int L>
PINN_D void wg_issue_stage(float* st, const WgradArgs& a, int64_t s0, int cnt, int tid) {
  float* sD = st;
  float* sA = sD + L * kWgS * 64;
  float* sX = sA + L * kWgS * 64;
  float* sDV0 = sX + kWgS * 8;
  float* sAV0 = sDV0 + kWgS * 32;
  float* sDV1 = sAV0 + kWgS * 32;
  float* sAV1 = sDV1 + kWgS * 16;
  float* sDU = sAV1 + kWgS * 16;
  float* sDVS = sDU + kWgS;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int l = 0; l < L; ++l)
    for (int i = tid; i < kWgS * 16; i += blockDim.x) {
      const int r = i >> 4, c4 = i & 15;
      float* dD = sD + (l * kWgS + r) * 64 + 4 * c4;
      float* dA = sA + (l * kWgS + r) * 64 + 4 * c4;
      if (r < cnt) {
        cp_async16(dD, a.sc.del[l] + (s0 + r) * 64 + 4 * c4);
        cp_async16(dA, a.sc.act[l] + (s0 + r) * 64 + 4 * c4);
      } else {
        *reinterpret_cast<float4*>(dD) = zero;
        *reinterpret_cast<float4*>(dA) = zero;
      }
    }
  for (int i = tid; i < kWgS * 8; i += blockDim.x) {
    const int r = i >> 3, c4 = i & 7;
    if (r < cnt) {
      cp_async16(sDV0 + r * 32 + 4 * c4, a.sc.dv0 + (s0 + r) * 32 + 4 * c4);
      cp_async16(sAV0 + r * 32 + 4 * c4, a.sc.av0 + (s0 + r) * 32 + 4 * c4);
    } else {
      *reinterpret_cast<float4*>(sDV0 + r * 32 + 4 * c4) = zero;
      *reinterpret_cast<float4*>(sAV0 + r * 32 + 4 * c4) = zero;
    }
  }
  if (tid < kWgS * 4) {
    const int r = tid >> 2, c4 = tid & 3;
    if (r < cnt) {
      cp_async16(sDV1 + r * 16 + 4 * c4, a.sc.dv1 + (s0 + r) * 16 + 4 * c4);
      cp_async16(sAV1 + r * 16 + 4 * c4, a.sc.av1 + (s0 + r) * 16 + 4 * c4);
    } else {
      *reinterpret_cast<float4*>(sDV1 + r * 16 + 4 * c4) = zero;
      *reinterpret_cast<float4*>(sAV1 + r * 16 + 4 * c4) = zero;
    }
  } else if (tid < kWgS * 4 + kWgS * 2) {
    const int t = tid - kWgS * 4, r = t >> 1, c4 = t & 1;
    if (r < cnt) cp_async16(sX + r * 8 + 4 * c4, a.x + (s0 + r) * 8 + 4 * c4);
    else *reinterpret_cast<float4*>(sX + r * 8 + 4 * c4) = zero;
  } else if (tid < kWgS * 4 + kWgS * 2 + kWgS) {
    const int r = tid - kWgS * 6;
    sDU[r] = r < cnt ? a.sc.du[s0 + r] : 0.f;
    sDVS[r] = r < cnt ? a.sc.dvs[s0 + r] : 0.f;
  }
}

// 8x8 register tile of  out[j0..j0+8][k0..k0+8] += sum_r D[r][j0..] * A[r][k0..]  over one stage.
PINN_D void wg_tile8x8(const float* __restrict__ sDl, const float* __restrict__ sAl, int j0, int k0, float (&acc)[64]) {
#pragma unroll 2
  for (int r = 0; r < kWgS; ++r) {
    const float4 d0 = *reinterpret_cast<const float4*>(sDl + r * 64 + j0), d1 = *reinterpret_cast<const float4*>(sDl + r * 64 + j0 + 4);
    const float4 v0 = *reinterpret_cast<const float4*>(sAl + r * 64 + k0), v1 = *reinterpret_cast<const float4*>(sAl + r * 64 + k0 + 4);
    const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    const float2 vv[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float2 dp = make_float2(dd[p], dd[p]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 c = make_float2(acc[8 * p + 2 * q], acc[8 * p + 2 * q + 1]);
        c = ffma2(dp, vv[q], c);
        acc[8 * p + 2 * q] = c.x; acc[8 * p + 2 * q + 1] = c.y;
      }
    }
  }
}

// Contraction over the batch with uniform warps.  NBT = 64 (L-1) + 32 threads each own one 8x8
// register tile of the big products (dW_l 64x64 for l >= 1: 64 threads each; dWv0 32x64: 32
// threads); the small outputs are spread evenly on top:
//   dW0 (64x8): thread t < 128 -> row t/2, 4 columns;   dWv1 (16x32): t < 128 -> row t/8, 4 columns
//   trunk bias sums db_l[c]: entry e = t, t + blockDim (< 64 L);   dWp[c]: t < 64;   dWv2[k]: 64 <= t < 80
//   head bias sums (dbv0 32, dbv1 16, dbp, dbv2): 80 <= t < 130
// One FULL stage (kWgS samples) by TMA: every array's slab is contiguous in HBM, so the stage is
// 2L + 7 bulk copies issued by a single thread; the mbarrier counts the bytes as they land.
template <int L>
PINN_D void wg_issue_stage_bulk(float* st, uint64_t* bar, const WgradArgs& a, int64_t s0) {
  float* sD = st;
  float* sA = sD + L * kWgS * 64;
  float* sX = sA + L * kWgS * 64;
  float* sDV0 = sX + kWgS * 8;
  float* sAV0 = sDV0 + kWgS * 32;
  float* sDV1 = sAV0 + kWgS * 32;
  float* sAV1 = sDV1 + kWgS * 16;
  float* sDU = sAV1 + kWgS * 16;
  float* sDVS = sDU + kWgS;
  tc::mbar_expect_tx(bar, static_cast<uint32_t>(wg_stage_floats(L) * sizeof(float)));
#pragma unroll
  for (int l = 0; l < L; ++l) {
    tc::bulk_g2s(sD + l * kWgS * 64, a.sc.del[l] + s0 * 64, kWgS * 64 * 4, bar);
    tc::bulk_g2s(sA + l * kWgS * 64, a.sc.act[l] + s0 * 64, kWgS * 64 * 4, bar);
  }
  tc::bulk_g2s(sX, a.x + s0 * 8, kWgS * 8 * 4, bar);
  tc::bulk_g2s(sDV0, a.sc.dv0 + s0 * 32, kWgS * 32 * 4, bar);
  tc::bulk_g2s(sAV0, a.sc.av0 + s0 * 32, kWgS * 32 * 4, bar);
  tc::bulk_g2s(sDV1, a.sc.dv1 + s0 * 16, kWgS * 16 * 4, bar);
  tc::bulk_g2s(sAV1, a.sc.av1 + s0 * 16, kWgS * 16 * 4, bar);
  tc::bulk_g2s(sDU, a.sc.du + s0, kWgS * 4, bar);
  tc::bulk_g2s(sDVS, a.sc.dvs + s0, kWgS * 4, bar);
}

PINN_HD constexpr int wg_threads(int L) { return 64 * (L - 1) + 32 < 160 ? 160 : 64 * (L - 1) + 32; }   // >= 130 needed by the small outputs
template <int L>
__global__ void __launch_bounds__(wg_threads(L), (L >= 4 ? 2 : 3))
wgrad_kernel(WgradArgs a, ParamLayout lay) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NBT = 64 * (L - 1) + 32, NT = wg_threads(L);
  constexpr int SF = wg_stage_floats(L);
  const int tid = threadIdx.x;
  // big tile of this thread
  const bool is_v0 = tid >= 64 * (L - 1);
  const int lbig = 1 + (tid >> 6), tt = is_v0 ? tid - 64 * (L - 1) : (tid & 63);
  const int j0 = 8 * (tt >> 3), k0 = 8 * (tt & 7);
  float acc[64];
#pragma unroll
  for (int q = 0; q < 64; ++q) acc[q] = 0.f;
  float aW0[4] = {0.f, 0.f, 0.f, 0.f}, aV1[4] = {0.f, 0.f, 0.f, 0.f}, aB[2] = {0.f, 0.f}, aP = 0.f;
  const int e0 = tid, e1 = tid + NT;                        // trunk-bias entries (flat over [L][64])
  const bool has_big = tid < NBT;

  // sample range of this CTA: a multiple of the stage size (keeps every TMA source 16-byte aligned)
  int64_t per = (a.n + gridDim.x - 1) / gridDim.x;
  per = (per + kWgS - 1) / kWgS * kWgS;
  const int64_t s_begin = static_cast<int64_t>(blockIdx.x) * per < a.n ? static_cast<int64_t>(blockIdx.x) * per : a.n;
  const int64_t s_end = s_begin + per < a.n ? s_begin + per : a.n;
  const int n_stage = s_end > s_begin ? static_cast<int>((s_end - s_begin + kWgS - 1) / kWgS) : 0;
  auto cnt_of = [&](int it) { const int64_t s0 = s_begin + static_cast<int64_t>(it) * kWgS; return static_cast<int>(s_end - s0 < kWgS ? s_end - s0 : kWgS); };
  __shared__ __align__(8) uint64_t full[2];
  if (tid == 0) { tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1); tc::fence_mbar_init(); }
  __syncthreads();
  uint32_t ph[2] = {0u, 0u};
  // full stages arrive by TMA (one thread issues), a ragged last stage by per-thread cp.async
  auto issue = [&](int it) {
    float* st = sm + (it & 1) * SF;
    const int64_t s0 = s_begin + static_cast<int64_t>(it) * kWgS;
    if (cnt_of(it) == kWgS) { if (tid == 0) wg_issue_stage_bulk<L>(st, &full[it & 1], a, s0); }
    else { wg_issue_stage<L>(st, a, s0, cnt_of(it), tid); cp_async_commit(); }
  };
  if (n_stage > 0) issue(0);
  for (int it = 0; it < n_stage; ++it) {
    if (it + 1 < n_stage) issue(it + 1);
    if (cnt_of(it) == kWgS) { tc::mbar_wait(&full[it & 1], ph[it & 1]); ph[it & 1] ^= 1u; }
    else { cp_async_wait<0>(); __syncthreads(); }
    const float* st = sm + (it & 1) * SF;
    const float* sD = st;
    const float* sA = sD + L * kWgS * 64;
    const float* sX = sA + L * kWgS * 64;
    const float* sDV0 = sX + kWgS * 8;
    const float* sAV0 = sDV0 + kWgS * 32;
    const float* sDV1 = sAV0 + kWgS * 32;
    const float* sAV1 = sDV1 + kWgS * 16;
    const float* sDU = sAV1 + kWgS * 16;
    const float* sDVS = sDU + kWgS;
    // ---- big tile
    if (!has_big) {
    } else if (!is_v0) {
      wg_tile8x8(sD + lbig * kWgS * 64, sA + (lbig - 1) * kWgS * 64, j0, k0, acc);
    } else {
      const float* sAl = sA + (L - 1) * kWgS * 64;
#pragma unroll 2
      for (int r = 0; r < kWgS; ++r) {
        const float4 d0 = *reinterpret_cast<const float4*>(sDV0 + r * 32 + j0), d1 = *reinterpret_cast<const float4*>(sDV0 + r * 32 + j0 + 4);
        const float4 v0 = *reinterpret_cast<const float4*>(sAl + r * 64 + k0), v1 = *reinterpret_cast<const float4*>(sAl + r * 64 + k0 + 4);
        const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const float2 vv[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const float2 dp = make_float2(dd[p], dd[p]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 c = make_float2(acc[8 * p + 2 * q], acc[8 * p + 2 * q + 1]);
            c = ffma2(dp, vv[q], c);
            acc[8 * p + 2 * q] = c.x; acc[8 * p + 2 * q + 1] = c.y;
          }
        }
      }
    }
    // ---- small outputs
#pragma unroll 4
    for (int r = 0; r < kWgS; ++r) {
      if (tid < 128) {
        const float d0 = sD[r * 64 + (tid >> 1)];
        const float4 xv = *reinterpret_cast<const float4*>(sX + r * 8 + 4 * (tid & 1));
        aW0[0] = fmaf(d0, xv.x, aW0[0]); aW0[1] = fmaf(d0, xv.y, aW0[1]); aW0[2] = fmaf(d0, xv.z, aW0[2]); aW0[3] = fmaf(d0, xv.w, aW0[3]);
        const float d1 = sDV1[r * 16 + (tid >> 3)];
        const float4 av = *reinterpret_cast<const float4*>(sAV0 + r * 32 + 4 * (tid & 7));
        aV1[0] = fmaf(d1, av.x, aV1[0]); aV1[1] = fmaf(d1, av.y, aV1[1]); aV1[2] = fmaf(d1, av.z, aV1[2]); aV1[3] = fmaf(d1, av.w, aV1[3]);
      }
      if (e0 < 64 * L) aB[0] += sD[((e0 >> 6) * kWgS + r) * 64 + (e0 & 63)];
      if (e1 < 64 * L) aB[1] += sD[((e1 >> 6) * kWgS + r) * 64 + (e1 & 63)];
      if (tid < 64) aP = fmaf(sDU[r], sA[((L - 1) * kWgS + r) * 64 + tid], aP);
      else if (tid < 80) aP = fmaf(sDVS[r], sAV1[r * 16 + tid - 64], aP);
      else if (tid < 112) aP += sDV0[r * 32 + tid - 80];
      else if (tid < 128) aP += sDV1[r * 16 + tid - 112];
      else if (tid == 128) aP += sDU[r];
      else if (tid == 129) aP += sDVS[r];
    }
    __syncthreads();             // everyone is done with this stage before it is refilled
  }
  // ---------------------------------------------------------------- write this CTA's partial
  float* part = a.partial + static_cast<size_t>(blockIdx.x) * lay.total;
  if (has_big) {
    const int64_t base = is_v0 ? lay.offWv0 : lay.offW[lbig];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) part[base + (j0 + p) * 64 + k0 + q] = acc[8 * p + q] * a.act_scale;
  }
  if (tid < 128) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      part[lay.offW[0] + (tid >> 1) * 8 + 4 * (tid & 1) + q] = aW0[q];
      part[lay.offWv1 + (tid >> 3) * 32 + 4 * (tid & 7) + q] = aV1[q];
    }
  }
  if (e0 < 64 * L) part[lay.offb[e0 >> 6] + (e0 & 63)] = aB[0];
  if (e1 < 64 * L) part[lay.offb[e1 >> 6] + (e1 & 63)] = aB[1];
  if (tid < 64) part[lay.offWp + tid] = aP * a.act_scale;
  else if (tid < 80) part[lay.offWv2 + tid - 64] = aP;
  else if (tid < 112) part[lay.offbv0 + tid - 80] = aP;
  else if (tid < 128) part[lay.offbv1 + tid - 112] = aP;
  else if (tid == 128) part[lay.offbp] = aP;
  else if (tid == 129) part[lay.offbv2] = aP;
}

// ------------------------------------------------------------------------------- K2a
struct TcbLayout {   // float offsets into dynamic shared memory
  int pn_hi[2], pn_lo[2];      // per group: activation / delta planes, 128 rows x 64 cols
  int b_hi[2], b_lo[2];        // streaming mode (L > 3), per group: one layer's weight planes
  int wf_hi[PINN_MAX_HIDDEN + 1], wf_lo[PINN_MAX_HIDDEN + 1];   // resident mode: forward planes of W_l (l >= 1); index L = heads
  int wt_hi[PINN_MAX_HIDDEN + 1], wt_lo[PINN_MAX_HIDDEN + 1];   // resident mode: transposed planes (dgrad)
  int W0, b0, b[PINN_MAX_HIDDEN], bv0, bp, Wv1, bv1, Wv2, bv2;
  int total;
};
constexpr int kBPlaneFloats = 16 * 1040 / 4;    // 16 chunks x 1040 B (padded LBO, see wcommit_transposed)
PINN_HD TcbLayout make_tcb_layout(int L) {
  TcbLayout t;
  int o = 0;
  for (int g = 0; g < 2; ++g) t.pn_hi[g] = t.pn_lo[g] = 0;     // activation / delta planes live in tensor memory
  for (int l = 0; l <= PINN_MAX_HIDDEN; ++l) t.wf_hi[l] = t.wf_lo[l] = t.wt_hi[l] = t.wt_lo[l] = 0;
  for (int g = 0; g < 2; ++g) t.b_hi[g] = t.b_lo[g] = 0;
  if (L <= 3) {     // every weight plane resident: L x 66 KB
    for (int l = 1; l <= L; ++l) {
      t.wf_hi[l] = o; o += kBH * kBH; t.wf_lo[l] = o; o += kBH * kBH;
      t.wt_hi[l] = o; o += kBPlaneFloats; t.wt_lo[l] = o; o += kBPlaneFloats;
    }
  } else {
    for (int g = 0; g < 2; ++g) { t.b_hi[g] = o; o += kBPlaneFloats; t.b_lo[g] = o; o += kBPlaneFloats; }
  }
  t.W0 = o; o += kBH * PINN_N_IN;
  t.b0 = o; o += kBH;
  for (int l = 0; l < PINN_MAX_HIDDEN; ++l) t.b[l] = 0;
  for (int l = 1; l < L; ++l) { t.b[l] = o; o += kBH; }
  t.bv0 = o; o += 32; t.bp = o; o += 4;
  t.Wv1 = o; o += 16 * 32; t.bv1 = o; o += 16; t.Wv2 = o; o += 16; t.bv2 = o; o += 4;
  t.total = o;
  return t;
}

PINN_D void grp_sync256(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

// Weight staging is split in two so the L2 latency hides behind the previous MMA:
//   wprefetch : 4 x 128-bit global loads per thread into registers (item idx = t256 + 256 it ->
//               source row j = idx % 64, 4-column chunk kc = idx / 64; rows >= J read as zero,
//               row J optionally comes from `extra_row`, e.g. the mean head stacked under Wv0);
//   wcommit_rows       : B[n = j][k]  (K-major rows, forward:  z = a W^T)
//   wcommit_transposed : B[n = k][col j] = src[j][k]  (dgrad:  d a = d z W); scalar stores, the
//               plane LBO is padded to 1040 B so a warp's eight 4-column chunks hit distinct banks.
constexpr uint32_t kLboT = 1040;
PINN_D void wprefetch(float4 (&w)[4], const float* __restrict__ src, const float* __restrict__ extra_row, int J, int t256) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = t256 + 256 * it, j = idx & 63, kc = idx >> 6;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < J) v = __ldg(reinterpret_cast<const float4*>(src + j * 64) + kc);
    else if (j == J && extra_row != nullptr) v = __ldg(reinterpret_cast<const float4*>(extra_row) + kc);
    w[it] = v;
  }
}
// `c` = the dropout scale 1/(1-p), folded into every plane (see mlp_tc.cu): masked activations and
// masked deltas are then plain selects.
PINN_D void wcommit_rows(float* hi, float* lo, const float4 (&w)[4], int t256, float c) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = t256 + 256 * it;
    tc::store_split4(hi, lo, 64 * 16, idx & 63, idx >> 6, make_float4(w[it].x * c, w[it].y * c, w[it].z * c, w[it].w * c));
  }
}
PINN_D void wcommit_transposed(float* hi, float* lo, const float4 (&w)[4], int t256, float c) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = t256 + 256 * it, j = idx & 63, kc = idx >> 6;
    const float vv[4] = {w[it].x * c, w[it].y * c, w[it].z * c, w[it].w * c};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float h = tc::tf32_hi(vv[r]);
      const size_t off = (static_cast<size_t>(j >> 2) * kLboT + static_cast<size_t>(4 * kc + r) * 16 + (j & 3) * 4) / 4;
      hi[off] = h;
      lo[off] = vv[r] - h;
    }
  }
}

// Resident mode (L <= 3): all forward and transposed planes are built once per CTA by all threads.
PINN_D void stage_plane_rows(float* hi, float* lo, const float* __restrict__ src, const float* __restrict__ extra_row, int J, float c) {
  for (int idx = threadIdx.x; idx < 64 * 16; idx += blockDim.x) {
    const int j = idx & 63, kc = idx >> 6;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < J) v = __ldg(reinterpret_cast<const float4*>(src + j * 64) + kc);
    else if (j == J && extra_row != nullptr) v = __ldg(reinterpret_cast<const float4*>(extra_row) + kc);
    tc::store_split4(hi, lo, 64 * 16, j, kc, make_float4(v.x * c, v.y * c, v.z * c, v.w * c));
  }
}
PINN_D void stage_plane_transposed(float* hi, float* lo, const float* __restrict__ src, const float* __restrict__ extra_row, int J, float c) {
  for (int idx = threadIdx.x; idx < 64 * 16; idx += blockDim.x) {
    const int j = idx & 63, kc = idx >> 6;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < J) v = __ldg(reinterpret_cast<const float4*>(src + j * 64) + kc);
    else if (j == J && extra_row != nullptr) v = __ldg(reinterpret_cast<const float4*>(extra_row) + kc);
    const float vv[4] = {v.x * c, v.y * c, v.z * c, v.w * c};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float h = tc::tf32_hi(vv[r]);
      const size_t off = (static_cast<size_t>(j >> 2) * kLboT + static_cast<size_t>(4 * kc + r) * 16 + (j & 3) * 4) / 4;
      hi[off] = h;
      lo[off] = vv[r] - h;
    }
  }
}

struct TcbArgs {
  const float* x; int64_t n;
  const float* grad_u; const float* grad_s; const float* y; float inv_n_global;
  BwdScratch sc;
  double* loss_partial;     // [2 * grid][4]
};

// Trunk activations are written to the scratch (and staged as MMA operands) as  keep ? tanh : 0,
// WITHOUT the dropout scale: the scale lives in the weight planes here and is applied once to the
// finished dW sums in K2b.  The variance head's CUDA-core tail keeps scaled activations.
template <int L, bool INJ>
__global__ void __launch_bounds__(512, 1)
mlp_tc_bwd_kernel(pinn_net_t net, TcbLayout lay, const __grid_constant__ DropParams dp, TcbArgs a) {
  constexpr int H = kBH, HH = 32;
  constexpr bool RES = L <= 3;     // weight planes resident in shared memory (else re-staged per phase)
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double lred[2][4][4];
  const int tid = threadIdx.x, row = tid & 127, t256 = tid & 255;
  const int warp = tc::uniform_warp_idx(), grp = warp >> 3, half = (warp >> 2) & 1;
  const int Dm = L * H + H / 2;
  const int cb = half * HH;

  if (tid == 0) { tc::mbar_init(&mbar[0], 1); tc::mbar_init(&mbar[1], 1); tc::fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;
  stage_tensor_scaled(smem + lay.W0, net.W[0], H * PINN_N_IN, kTanhArg);    // tanh_pre arguments (common.cuh)
  stage_tensor_scaled(smem + lay.b0, net.b[0], H, kTanhArg);
  for (int l = 1; l < L; ++l) stage_tensor_scaled(smem + lay.b[l], net.b[l], H, kTanhArg);
  stage_tensor_scaled(smem + lay.bv0, net.bv0, 32, kTanhArg);
  stage_tensor(smem + lay.bp, net.bp, 1);
  stage_tensor(smem + lay.Wv1, net.Wv1, 16 * 32);
  stage_tensor_scaled(smem + lay.bv1, net.bv1, 16, kTanhArg);
  stage_tensor(smem + lay.Wv2, net.Wv2, 16);
  stage_tensor(smem + lay.bv2, net.bv2, 1);
  if constexpr (RES) {
    for (int l = 1; l < L; ++l) {
      stage_plane_rows(smem + lay.wf_hi[l], smem + lay.wf_lo[l], net.W[l], nullptr, H, wscale);
      stage_plane_transposed(smem + lay.wt_hi[l], smem + lay.wt_lo[l], net.W[l], nullptr, H, wscale);
    }
    stage_plane_rows(smem + lay.wf_hi[L], smem + lay.wf_lo[L], net.Wv0, net.Wp, 32, wscale);
    stage_plane_transposed(smem + lay.wt_hi[L], smem + lay.wt_lo[L], net.Wv0, net.Wp, 32, wscale);
    tc::fence_proxy_async();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const uint32_t d_tmem = tmem_base_s + static_cast<uint32_t>(grp * 64);
  const uint32_t d_lane = d_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // A operand (activations forward, deltas backward) in tensor memory: hi plane [128 + 128 g, +64), lo plane +64
  // (lane = sample row), see mlp_tc.cu.
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t a_hi_t = tmem_base_s + static_cast<uint32_t>(128 + grp * 128), a_lo_t = a_hi_t + 64u;
  float* b_hi = smem + lay.b_hi[grp];
  float* b_lo = smem + lay.b_lo[grp];
  const uint32_t bh_u = tc::smem_u32(b_hi), bl_u = tc::smem_u32(b_lo);
  const uint32_t idesc64 = tc::make_idesc_tf32(kBTile, 64), idesc48 = tc::make_idesc_tf32(kBTile, 48);
  const bool issuer_warp = (warp & 7) == 0;
  uint32_t phase = 0;
  double l_nll = 0.0, l_abs = 0.0, l_mse = 0.0, l_cnt = 0.0;

  // publish PN + B (generic-proxy writes) to the async proxy, run one 3xTF32 product, wait for it
  auto run_mma = [&](int res_hi, int res_lo, uint32_t lbo_b, uint32_t idesc, auto&& prefetch) {
    const uint32_t bh_a = RES ? tc::smem_u32(smem + res_hi) : bh_u, bl_a = RES ? tc::smem_u32(smem + res_lo) : bl_u;
    tc::tmem_wait_st();
    tc::fence_proxy_async();
    tc::fence_before_sync();
    grp_sync256(grp);
    if (issuer_warp) {
      const uint64_t bhd = tc::make_desc(bh_a, lbo_b, 128), bld = tc::make_desc(bl_a, lbo_b, 128);
      if (tc::elect_one()) {
        tc::fence_after_sync();
        tc::issue_3xtf32_ts<64>(d_tmem, a_hi_t, a_lo_t, bhd, bld, lbo_b, idesc);
        tc::umma_commit(&mbar[grp]);
      }
      __syncwarp();
    }
    prefetch();                 // global loads for the NEXT phase fly while the tensor core works
    tc::mbar_wait(&mbar[grp], phase);
    phase ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
  };
  auto store_pn8 = [&](int c0, const float (&v)[8]) {   // 8 consecutive columns of this thread's row
    float h[8], lo[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { h[q] = tc::tf32_hi_fast(v[q]); lo[q] = v[q] - h[q]; }
    tc::tmem_st8(a_hi_t + lane_sel + static_cast<uint32_t>(c0), h);
    tc::tmem_st8(a_lo_t + lane_sel + static_cast<uint32_t>(c0), lo);
  };

  const int64_t n_tiles = (a.n + kBTile - 1) / kBTile;
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * 2 + grp; tile < n_tiles; tile += static_cast<int64_t>(gridDim.x) * 2) {
    const int64_t s = tile * kBTile + row;
    const bool valid = s < a.n;
    const bool active = drop_on && (!INJ || valid);     // injected masks: tail rows of the last tile have no mask row
    const uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s);
    KeepSrc<INJ> ks;
    ks.s_lo = static_cast<uint32_t>(sg); ks.s_hi = static_cast<uint32_t>(sg >> 32);
    ks.pass = static_cast<uint32_t>(dp.pass_offset);
    ks.mrow = INJ ? dp.masks + static_cast<size_t>(valid ? s : 0) * Dm : nullptr;
    uint32_t kb[L + 1];      // keep bits of this thread's 32 columns, per dropout layer (bit q = column cb + q)
    float4 wpre[4];
    if constexpr (!RES) wprefetch(wpre, net.W[1], nullptr, H, t256);
    // ============================ forward ============================
    {
      float xr[PINN_N_IN];
      if (valid) {
        const float4* px = reinterpret_cast<const float4*>(a.x + s * PINN_N_IN);
        float4 q0 = __ldg(px), q1 = __ldg(px + 1);
        xr[0] = q0.x; xr[1] = q0.y; xr[2] = q0.z; xr[3] = q0.w; xr[4] = q1.x; xr[5] = q1.y; xr[6] = q1.z; xr[7] = q1.w;
      } else {
#pragma unroll
        for (int i = 0; i < PINN_N_IN; ++i) xr[i] = 0.f;
      }
      const float* W0 = smem + lay.W0 + cb * PINN_N_IN;
      const float* b0 = smem + lay.b0 + cb;
      kb[0] = 0u;
#pragma unroll 1
      for (int g = 0; g < HH; g += 8) {     // rolled (code size): 8 columns per trip
        bool k[8] = {true, true, true, true, true, true, true, true};
        if (active) ks.get8(dp, 0u, static_cast<uint32_t>(cb + g), 0u, k);
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 w0 = *reinterpret_cast<const float4*>(W0 + (g + q) * PINN_N_IN);
          const float4 w1 = *reinterpret_cast<const float4*>(W0 + (g + q) * PINN_N_IN + 4);
          float z = b0[g + q];
          z = fmaf(w0.x, xr[0], z); z = fmaf(w0.y, xr[1], z); z = fmaf(w0.z, xr[2], z); z = fmaf(w0.w, xr[3], z);
          z = fmaf(w1.x, xr[4], z); z = fmaf(w1.y, xr[5], z); z = fmaf(w1.z, xr[6], z); z = fmaf(w1.w, xr[7], z);
          v[q] = k[q] ? tanh_pre(z) : 0.f;
          kb[0] |= (k[q] ? 1u : 0u) << (g + q);
        }
        store_pn8(cb + g, v);
        if (valid) {
          float4* o = reinterpret_cast<float4*>(a.sc.act[0] + s * H + cb + g);
          o[0] = make_float4(v[0], v[1], v[2], v[3]); o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
#pragma unroll 1
    for (int l = 1; l < L; ++l) {       // rolled: one copy of the layer body keeps the kernel inside the I-cache
      if constexpr (!RES) wcommit_rows(b_hi, b_lo, wpre, t256, wscale);
      run_mma(lay.wf_hi[l], lay.wf_lo[l], H * 16, idesc64, [&] {
        if constexpr (!RES) {
          if (l + 1 < L) wprefetch(wpre, net.W[l + 1], nullptr, H, t256);
          else wprefetch(wpre, net.Wv0, net.Wp, 32, t256);
        }
      });
      const float* bl = smem + lay.b[l] + cb;
      uint32_t bits = 0u;
#pragma unroll 1
      for (int g = 0; g < HH; g += 8) {
        float z[8];
        tc::tmem_ld8(d_lane + cb + g, z);
        tc::tmem_wait_ld();
        bool k[8] = {true, true, true, true, true, true, true, true};
        if (active) ks.get8(dp, static_cast<uint32_t>(l), static_cast<uint32_t>(cb + g), static_cast<uint32_t>(l * H), k);
        const float4 bA = *reinterpret_cast<const float4*>(bl + g), bB = *reinterpret_cast<const float4*>(bl + g + 4);
        const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[q] = k[q] ? tanh_pre(fmaf(z[q], kTanhArg, bb[q])) : 0.f;
          bits |= (k[q] ? 1u : 0u) << (g + q);
        }
        store_pn8(cb + g, v);
        if (valid) {
          float4* o = reinterpret_cast<float4*>(a.sc.act[l] + s * H + cb + g);
          o[0] = make_float4(v[0], v[1], v[2], v[3]); o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      kb[l] = bits;
    }
    // ---- heads: rows 0..31 = Wv0, row 32 = Wp, rows 33.. = 0 (N = 48 of the 64 staged rows are read)
    if constexpr (!RES) wcommit_rows(b_hi, b_lo, wpre, t256, wscale);
    run_mma(lay.wf_hi[L], lay.wf_lo[L], H * 16, idesc48, [&] { if constexpr (!RES) wprefetch(wpre, net.Wv0, net.Wp, 32, t256); });
    float du = 0.f;
    float dzv0[HH];                      // half 0: d z of the variance head's first layer; half 1: unused
    kb[L] = 0u;
    if (half == 0) {
      float v0[HH], zz[16];
      tc::tmem_ld16(d_lane, v0);
      tc::tmem_ld16(d_lane + 16, v0 + 16);
      tc::tmem_ld16(d_lane + 32, zz);
      tc::tmem_wait_ld();
      const float u = zz[0] + smem[lay.bp];
      const float* bv0 = smem + lay.bv0;
#pragma unroll
      for (int g = 0; g < HH; g += 8) {
        bool k[8] = {true, true, true, true, true, true, true, true};
        if (active) ks.get8(dp, static_cast<uint32_t>(L), static_cast<uint32_t>(g), static_cast<uint32_t>(L * H), k);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v0[g + q] = k[q] ? tanh_pre(fmaf(v0[g + q], kTanhArg, bv0[g + q])) * wscale : 0.f;   // scaled, as K2b expects
          kb[L] |= (k[q] ? 1u : 0u) << (g + q);
        }
      }
      const float* Wv1 = smem + lay.Wv1;
      const float* bv1 = smem + lay.bv1;
      const float* Wv2 = smem + lay.Wv2;
      float v1[16];
      float vraw = smem[lay.bv2];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float2 acc = make_float2(0.f, 0.f), acc2 = acc;
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 w = *reinterpret_cast<const float4*>(Wv1 + k * 32 + 4 * i4);
          acc = ffma2(make_float2(w.x, w.y), make_float2(v0[4 * i4], v0[4 * i4 + 1]), acc);
          acc2 = ffma2(make_float2(w.z, w.w), make_float2(v0[4 * i4 + 2], v0[4 * i4 + 3]), acc2);
        }
        v1[k] = tanh_pre(fmaf((acc.x + acc.y) + (acc2.x + acc2.y), kTanhArg, bv1[k]));
        vraw = fmaf(Wv2[k], v1[k], vraw);
      }
      const float lv = logvar_from_v(vraw);
      float ds = 0.f;
      if (valid) {
        if (a.grad_u != nullptr) {
          du = __ldg(a.grad_u + s);
          ds = a.grad_s ? __ldg(a.grad_s + s) : 0.f;
        } else {
          const float yv = __ldg(a.y + s);
          const float e = expf(-lv), diff = yv - u;
          du = -e * diff * a.inv_n_global;
          const float sg = lv > 0.f ? 1.f : (lv < 0.f ? -1.f : 0.f);
          ds = (-0.5f * e * diff * diff + 0.5f + 0.01f * sg) * a.inv_n_global;
          l_nll += static_cast<double>(0.5f * e * diff * diff + 0.5f * lv);
          l_abs += static_cast<double>(fabsf(lv));
          l_mse += static_cast<double>(diff * diff);
          l_cnt += 1.0;
        }
      }
      const float dv = ds * dlogvar_dv(vraw);
      float dz1[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) dz1[k] = dv * Wv2[k] * (1.0f - v1[k] * v1[k]);
      // d v0[i] = sum_k Wv1[k][i] dz1[k];  dz_v0 = d v0 * keep-mask * (1 - a^2)
#pragma unroll
      for (int i = 0; i < HH; ++i) dzv0[i] = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 w = *reinterpret_cast<const float4*>(Wv1 + k * 32 + 4 * i4);
          dzv0[4 * i4] = fmaf(w.x, dz1[k], dzv0[4 * i4]);         dzv0[4 * i4 + 1] = fmaf(w.y, dz1[k], dzv0[4 * i4 + 1]);
          dzv0[4 * i4 + 2] = fmaf(w.z, dz1[k], dzv0[4 * i4 + 2]); dzv0[4 * i4 + 3] = fmaf(w.w, dz1[k], dzv0[4 * i4 + 3]);
        }
#pragma unroll
      for (int i = 0; i < HH; ++i) {
        const float av = v0[i] * (drop_on ? dp.keep : 1.0f);
        const float mk = ((kb[L] >> i) & 1u) ? wscale : 0.f;
        dzv0[i] = dzv0[i] * mk * (1.0f - av * av);
      }
      if (valid) {
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          reinterpret_cast<float4*>(a.sc.av0 + s * 32)[i4] = make_float4(v0[4 * i4], v0[4 * i4 + 1], v0[4 * i4 + 2], v0[4 * i4 + 3]);
          reinterpret_cast<float4*>(a.sc.dv0 + s * 32)[i4] = make_float4(dzv0[4 * i4], dzv0[4 * i4 + 1], dzv0[4 * i4 + 2], dzv0[4 * i4 + 3]);
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          reinterpret_cast<float4*>(a.sc.av1 + s * 16)[i4] = make_float4(v1[4 * i4], v1[4 * i4 + 1], v1[4 * i4 + 2], v1[4 * i4 + 3]);
          reinterpret_cast<float4*>(a.sc.dv1 + s * 16)[i4] = make_float4(dz1[4 * i4], dz1[4 * i4 + 1], dz1[4 * i4 + 2], dz1[4 * i4 + 3]);
        }
        a.sc.du[s] = du;
        a.sc.dvs[s] = dv;
      }
    }
    // ============================ backward ============================
    // A operand = [dz_v0 (32 cols) | du | 0 ...]; B = ([Wv0; Wp])^T  ->  d a_{L-1}
    if (half == 0) {
#pragma unroll
      for (int g = 0; g < HH; g += 8) {
        const float d8[8] = {dzv0[g], dzv0[g + 1], dzv0[g + 2], dzv0[g + 3], dzv0[g + 4], dzv0[g + 5], dzv0[g + 6], dzv0[g + 7]};
        store_pn8(g, d8);
      }
      const float d8[8] = {du, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store_pn8(32, d8);
    } else {
      const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c0 = 40; c0 < 64; c0 += 8) store_pn8(c0, z8);
    }
    if constexpr (!RES) wcommit_transposed(b_hi, b_lo, wpre, t256, wscale);
#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
      float4 apre[HH / 4];       // this thread's masked activations of layer l, prefetched during the MMA
      run_mma(lay.wt_hi[l + 1], lay.wt_lo[l + 1], kLboT, idesc64, [&] {
#pragma unroll
        for (int g4 = 0; g4 < HH / 4; ++g4)
          apre[g4] = valid ? *reinterpret_cast<const float4*>(a.sc.act[l] + s * H + cb + 4 * g4) : make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (!RES) { if (l > 0) wprefetch(wpre, net.W[l], nullptr, H, t256); }
      });
      const uint32_t kbl = kb[l];
#pragma unroll
      for (int g = 0; g < HH; g += 8) {
        float z[8], dz[8];
        tc::tmem_ld8(d_lane + cb + g, z);
        tc::tmem_wait_ld();
        const float aa[8] = {apre[g / 4].x, apre[g / 4].y, apre[g / 4].z, apre[g / 4].w,
                             apre[g / 4 + 1].x, apre[g / 4 + 1].y, apre[g / 4 + 1].z, apre[g / 4 + 1].w};
#pragma unroll
        for (int q = 0; q < 8; ++q)    // z already carries the dropout scale (folded into W^T); dropped units have a = 0
          dz[q] = ((kbl >> (g + q)) & 1u) ? z[q] * fmaf(-aa[q], aa[q], 1.0f) : 0.f;
        if (valid) {
          float4* o = reinterpret_cast<float4*>(a.sc.del[l] + s * H + cb + g);
          o[0] = make_float4(dz[0], dz[1], dz[2], dz[3]); o[1] = make_float4(dz[4], dz[5], dz[6], dz[7]);
        }
        if (l > 0) store_pn8(cb + g, dz);
      }
      if constexpr (!RES) { if (l > 0) wcommit_transposed(b_hi, b_lo, wpre, t256, wscale); }
    }
  }
  // ---------------------------------------------------------------- loss partials per group
  {
    double vals[4] = {l_nll, l_abs, l_mse, l_cnt};
    const int lane = tid & 31, wq = warp & 7;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double t = vals[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0 && wq < 4) lred[grp][wq][k] = t;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 8) {
    const int g = tid >> 2, k = tid & 3;
    double t = 0.0;
    for (int wq = 0; wq < 4; ++wq) t += lred[g][wq][k];
    a.loss_partial[(static_cast<size_t>(blockIdx.x) * 2 + g) * 4 + k] = t;
  }
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

__global__ void grad_reduce2_kernel(const float* __restrict__ partial, const double* __restrict__ loss_partial, int nblk,
                                    int nloss, int64_t total, float* __restrict__ grad, double* __restrict__ loss) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) {
    double acc = 0.0;
    for (int b = 0; b < nblk; ++b) acc += static_cast<double>(partial[static_cast<size_t>(b) * total + i]);
    grad[i] = static_cast<float>(acc);
  }
  if (loss != nullptr && blockIdx.x == 0 && threadIdx.x < 4) {
    double acc = 0.0;
    for (int b = 0; b < nloss; ++b) acc += loss_partial[static_cast<size_t>(b) * 4 + threadIdx.x];
    loss[threadIdx.x] = acc;
  }
}

static int g_tc_bwd_enabled = 1;

struct TcBwdPlan { int grid_a, grid_b; size_t smem_a, smem_b, off_partial, off_scratch, bytes; };
static TcBwdPlan plan_tc_bwd(int L, int64_t n) {
  TcBwdPlan p{};
  ParamLayout lay = make_layout(kBH, L);
  const int sms = sm_count();
  const int64_t tiles = (n + kBTile - 1) / kBTile;
  int64_t want = (tiles + 1) / 2;
  p.grid_a = static_cast<int>(want < sms ? (want > 0 ? want : 1) : sms);
  int64_t wb = (n + 1023) / 1024;
  p.grid_b = static_cast<int>(wb < 3 * sms ? (wb > 0 ? wb : 1) : 3 * sms);
  p.smem_a = static_cast<size_t>(make_tcb_layout(L).total) * sizeof(float);
  p.smem_b = static_cast<size_t>(2) * wg_stage_floats(L) * sizeof(float);
  size_t off = static_cast<size_t>(2 * p.grid_a) * 4 * sizeof(double);
  p.off_partial = off;
  off += static_cast<size_t>(p.grid_b) * lay.total * sizeof(float);
  off = (off + 255) & ~static_cast<size_t>(255);
  p.off_scratch = off;
  off += bwd_scratch_floats_per_sample(L) * ((static_cast<size_t>(n > 0 ? n : 1) + 3) & ~static_cast<size_t>(3)) * sizeof(float);
  p.bytes = off;
  return p;
}

bool tc_bwd_covers(const pinn_net_t* net) {
  if (!g_tc_bwd_enabled || net->width != kBH || net->n_hidden < 2 || net->n_hidden > 4) return false;
  for (int l = 0; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return false;
  return aligned16(net->Wv0) && aligned16(net->Wp);
}
size_t tc_bwd_workspace_bytes(int L, int64_t n) { return plan_tc_bwd(L, n).bytes; }

int launch_tc_bwd(const pinn_net_t* net, const float* x, int64_t n, const DropParams& dp, const float* grad_u,
                  const float* grad_s, const float* y, int64_t n_global, float* grad_flat, double* loss_sums, void* workspace,
                  size_t workspace_bytes, cudaStream_t st) {
  const int L = net->n_hidden;
  TcBwdPlan p = plan_tc_bwd(L, n);
  if (workspace_bytes < p.bytes) return PINN_E_WORKSPACE;
  ParamLayout lay = make_layout(kBH, L);
  char* ws = static_cast<char*>(workspace);
  TcbArgs a{};
  a.x = x; a.n = n; a.grad_u = grad_u; a.grad_s = grad_s; a.y = y;
  a.inv_n_global = grad_u ? 0.f : static_cast<float>(1.0 / static_cast<double>(n_global));
  a.sc = carve_scratch(reinterpret_cast<float*>(ws + p.off_scratch), n, L);
  a.loss_partial = reinterpret_cast<double*>(ws);
  TcbLayout tl = make_tcb_layout(L);
  const bool inj = dp.p > 0.f && dp.masks != nullptr;
#define LAUNCH_A(LL)                                                                                              \
  {                                                                                                               \
    auto kern = inj ? mlp_tc_bwd_kernel<LL, true> : mlp_tc_bwd_kernel<LL, false>;                                 \
    PINN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                         \
                                       static_cast<int>(p.smem_a)));                                              \
    kern<<<p.grid_a, 512, p.smem_a, st>>>(*net, tl, dp, a);                                                       \
  }
  switch (L) {
    case 2: LAUNCH_A(2) break;
    case 3: LAUNCH_A(3) break;
    case 4: LAUNCH_A(4) break;
    default: return PINN_E_SHAPE;
  }
#undef LAUNCH_A
  PINN_CUDA_TRY(cudaGetLastError());
  WgradArgs w{};
  w.x = x; w.n = n; w.L = L; w.sc = a.sc;
  w.act_scale = dp.p > 0.f ? dp.scale : 1.0f;
  w.partial = reinterpret_cast<float*>(ws + p.off_partial);
#define LAUNCH_B(LL)                                                                                              \
  {                                                                                                               \
    PINN_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                       static_cast<int>(p.smem_b)));                                              \
    wgrad_kernel<LL><<<p.grid_b, wg_threads(LL), p.smem_b, st>>>(w, lay);                                                                  \
  }
  switch (L) {
    case 2: LAUNCH_B(2) break;
    case 3: LAUNCH_B(3) break;
    default: LAUNCH_B(4) break;
  }
#undef LAUNCH_B
  PINN_CUDA_TRY(cudaGetLastError());
  const int rg = static_cast<int>((lay.total + 255) / 256);
  grad_reduce2_kernel<<<rg, 256, 0, st>>>(w.partial, a.loss_partial, p.grid_b, 2 * p.grid_a, lay.total, grad_flat, loss_sums);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pinn

// Ablation / test switch for the tensor-core backward path (1 = on).
extern "C" int pinn_set_tensor_core_bwd(int enable) {
  int prev = pinn::g_tc_bwd_enabled;
  pinn::g_tc_bwd_enabled = enable ? 1 : 0;
  return prev;
}
