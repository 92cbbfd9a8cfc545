// mlp_tc_bwd.cu -- K2 for the 64-wide net on the tensor cores, as two kernels:
//
//  K2a  mlp_tc_bwd_kernel : per 128-sample tile (two independent 256-thread groups per CTA,
//       as in mlp_tc.cu) forward recompute AND dgrad on tcgen05 (3xTF32, fp32-accurate):
//         stage a0 -> [W1] -> a1 -> [W2] -> a2 -> [Wv0;Wp] -> heads + loss gradients
//         -> [ (Wv0;Wp)^T ] -> d a2 -> [W2^T] -> d a1 -> [W1^T] -> d a0
//       The A operand (activations forward, deltas backward) lives in tensor memory; every B
//       operand is a K-major weight plane pair (forward: W rows; dgrad: W^T), resident in
//       shared memory for L <= 3 and re-split per phase for deeper nets.  Masked activations
//       and pre-activation deltas of every layer go to HBM as a TRANSPOSED per-tile row table
//       (RowMap below); keep-bits ride in registers so Philox runs once.
//  K2b  wgrad_tc_kernel : every weight and bias gradient as one 3xTF32 accumulation over
//       samples, reading that row table (K-major: the contraction index is the sample);
//       accumulators resident in tensor memory; per-CTA partials -> grad_reduce2_kernel
//       (fixed-order, deterministic).
//
// (MN-major TF32 operands, which would allow an all-on-chip variant, require CUTLASS's
// SW128_32B swizzled layout; with the plain interleaved layout the MMA is silently dropped --
// tests/cuda/tc_mn_test.cu documents that.)
#include "net.cuh"
#include "tc.cuh"
#include "tc_api.cuh"

namespace pinn {

constexpr int kBH = 64, kBTile = 128;

// Global scratch written by K2a and read by K2b: per 128-sample tile a table of ROWS, one row per
// feature, 128 samples (512 B) per row -- i.e. every array is stored TRANSPOSED.  Two reasons:
//  * the weight-gradient contraction runs over SAMPLES, so sample-contiguous rows are K-major
//    tensor-core operands (the plain interleaved layout cannot feed MN-major TF32 operands,
//    tests/cuda/tc_mn_test.cu);
//  * a warp of K2a (32 consecutive samples) writes / re-reads one 128-byte line per feature:
//    fully coalesced.  The former [sample][feature] rows cost 32 sectors per warp access and made
//    K2a's stores and re-loads 0.48 ms of its 1.23 ms (profiles/README.md, ablate_k2a).
// Inside a tile the table is stored as 8 slabs of 16 samples ([slab][row][16]): the 16-sample stage K2b
// streams is one contiguous 31 KB block, and a K2a warp still touches two full 64-byte segments per feature.
// Row order inside a tile (L hidden layers, nbig = ceil((L-1)/2) blocks of 128 rows):
//   [0, 128 nbig)          : deltas of layers 1..L-1, 64 rows each          (A operands "big")
//   RA + [0, 64)           : deltas of layer 0                               (A operand "small", 128 rows)
//   RA + [64, 96), +96     : variance-head layer-0 deltas, d loss / d u
//   RA + [97, 113), +113   : variance-head layer-1 deltas, d loss / d v      (rows 114..127 unused)
//   RB + 64 l + [0, 64)    : masked activations of layer l (WITHOUT the dropout scale)   (B operands)
//   RV0 + [0, 32)          : variance-head layer-0 activations (scaled),   RV1 + [0, 16): layer 1
constexpr int kRowStride = 16;       // floats between consecutive feature rows inside a slab (= samples per slab = K2b stage)
struct RowMap {
  int L, nbig, RA, RB, RV0, RV1, rows;
};
PINN_HD RowMap make_rowmap(int L) {
  RowMap m;
  m.L = L;
  m.nbig = (L - 1 + 1) / 2;
  m.RA = 128 * m.nbig;
  m.RB = m.RA + 128;
  m.RV0 = m.RB + 64 * L;
  m.RV1 = m.RV0 + 32;
  m.rows = m.RV1 + 16;
  return m;
}
PINN_HD int row_del(const RowMap& m, int l) { return l == 0 ? m.RA : 64 * (l - 1); }
PINN_HD int row_act(const RowMap& m, int l) { return m.RB + 64 * l; }
PINN_HD size_t bwd_scratch_floats(int L, int64_t n) {
  const size_t tiles = static_cast<size_t>((n + kBTile - 1) / kBTile);
  return (tiles > 0 ? tiles : 1) * static_cast<size_t>(make_rowmap(L).rows) * kBTile;
}

// ------------------------------------------------------------------------------- K2b
// Every weight / bias gradient as ONE accumulation over samples on the tensor cores (3xTF32):
//   D[m][n] += sum_s A[m][s] * B[n][s],   A = delta rows, B = activation rows + a row of ones
// (the ones row turns every bias column sum into one more column of the same product).
// M is always 128: delta arrays are stacked in A blocks of 128 rows, and the B rows of everything
// those deltas pair with are concatenated into ONE B block per A block, so each A slab (4 KB) is
// fetched once per MMA instead of once per product -- with both operands in shared memory the
// operand fetch, not the math, paces tcgen05.mma.  Cross terms are computed and never read.  L = 3:
//   big block   [delta_1 ; delta_2]                 x  [a_0 | a_1 | 1]              (N = 144)
//        lanes   0..63 , columns   0..63  = dW1      column 128 = db1
//        lanes  64..127, columns  64..127 = dW2      column 128 = db2
//   small block [delta_0 ; dv0 ; du ; dv1 ; dvs]    x  [a_2 | 1 | x^T | av0 | av1]  (N = 128)
//        lanes   0..63  x columns  65..72  = dW0     lanes 64..95 x columns 0..63 = dWv0    lane 96 = dWp
//        lanes  97..112 x columns  73..104 = dWv1    lane 113 x columns 105..120 = dWv2     column 64 = every bias
// Accumulators stay in tensor memory for the CTA's whole sample range.  Pipeline: 256 loader
// threads bring 16-sample stages (64 B per row, two stages in flight) from HBM into registers,
// split them into tf32 hi / lo planes (K-major canonical layout, LBO padded so that the 4 x 8
// chunk pattern of a warp is bank-conflict free) in one of two shared-memory buffers and arrive
// on `full`; a dedicated MMA warp issues the stage's products and commits to `done`, which frees
// the buffer.
constexpr int kWgStage = 16;                 // samples per stage
constexpr int kWgLoaders = 256;
constexpr int kSmOnes = 64, kSmX = 65, kSmV0 = 73, kSmV1 = 105, kSmN = 128;     // rows of the small B block
PINN_HD constexpr int wg_lbo(int rows) { return (rows + 2) * 16; }     // bytes; (rows + 2) % 8 == 2 for 80, 128 and 144 rows
PINN_HD constexpr int wg_big_n(int L, int i) { return 2 * i + 2 <= L - 1 ? 144 : 80; }     // N of big block i: two activation arrays + ones, or one

struct WgLayout {      // byte offsets of the hi plane of every operand block inside ONE stage buffer; lo plane = hi + plane_bytes
  int a_big[2], a_small, b_big[2], b_small;
  int plane_bytes, buffer_bytes;
};
PINN_HD WgLayout make_wg_layout(int L) {
  WgLayout w{};
  const int nbig = L / 2;          // ceil((L - 1) / 2)
  int o = 0;
  for (int i = 0; i < 2; ++i) { w.a_big[i] = o; if (i < nbig) o += 4 * wg_lbo(128); }
  w.a_small = o; o += 4 * wg_lbo(128);
  for (int i = 0; i < 2; ++i) { w.b_big[i] = o; if (i < nbig) o += 4 * wg_lbo(wg_big_n(L, i)); }
  w.b_small = o; o += 4 * wg_lbo(kSmN);
  w.plane_bytes = o;
  w.buffer_bytes = 2 * o;
  return w;
}

#ifdef PINN_TIMELINE
__device__ long long g_tlw[2][64][4];     // [0]: MMA warp {wake, issued}; [1]: loader thread 0 {pre-wait, post-wait, post-store, post-arrive}
__device__ long long g_tlp[2][8];         // coarse phase stamps of CTA 0: [0] = K2b {entry, prologue done, streaming done, partial written}, [1] = K2a {entry, weights staged, tiles done, end}
#define TLP(k, i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tlp[k][i] = clock64(); } while (0)
#else
#define TLP(k, i) do { } while (0)
#endif
struct WgradArgs {
  const float* x; int64_t n; int64_t n_tiles;
  const float* rows;       // scratch, [tile][RowMap.rows][128]
  float* partial;          // [grid][lay.total]
  float act_scale;         // dropout scale 1/(1-p) missing from the stored trunk activations (K2a)
};

template <int L>
__global__ void __launch_bounds__(kWgLoaders + 32, 1)
wgrad_tc_kernel(WgradArgs a, ParamLayout lay, WgLayout wl, RowMap rm) {
  constexpr int NBIG = L / 2;
  constexpr int COL_SMALL = (NBIG > 0 ? wg_big_n(L, 0) : 0) + (NBIG > 1 ? wg_big_n(L, 1) : 0);      // TMEM column of the small product
  static_assert(COL_SMALL + kSmN <= 512, "TMEM");
  extern __shared__ __align__(1024) unsigned char wsm[];
  __shared__ __align__(8) uint64_t full[2], done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tc::uniform_warp_idx();
  const bool mma_warp = warp == kWgLoaders / 32;
  TLP(0, 0);

  if (tid == 0) {
    tc::mbar_init(&full[0], kWgLoaders); tc::mbar_init(&full[1], kWgLoaders);
    tc::mbar_init(&done[0], 1); tc::mbar_init(&done[1], 1);
    tc::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  // constant rows of the B blocks (the ones row, zero padding): written once into both buffers, both planes
  {
    auto fill = [&](int blk_off, int n_rows, int row0, int ones_row) {
      const int lbo = wg_lbo(n_rows), cnt = n_rows - row0;
      for (int i = tid; i < 2 * 2 * 4 * cnt; i += blockDim.x) {
        const int r = row0 + i % cnt, rest = i / cnt;
        const int kc = rest & 3, plane = (rest >> 2) & 1, buf = rest >> 3;
        const float v = (r == ones_row && plane == 0) ? 1.0f : 0.0f;
        *reinterpret_cast<float4*>(wsm + buf * wl.buffer_bytes + plane * wl.plane_bytes + blk_off + kc * lbo + r * 16) = make_float4(v, v, v, v);
      }
    };
#pragma unroll
    for (int i = 0; i < NBIG; ++i) { const int nb = wg_big_n(L, i); fill(wl.b_big[i], nb, nb - 16, nb - 16); }
    fill(wl.b_small, kSmN, kSmOnes, kSmOnes);       // rows 64..127: ones + (overwritten every stage) x^T, av0, av1 + padding
    // unused A rows (114..127 of the small block; 64..127 of a half-filled big block) only feed lanes nobody reads
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  griddep_wait();          // everything above is independent of K2a's output (the row table)
  TLP(0, 1);
  // stage range of this CTA: the work unit is one 16-sample slab, not a tile, so that small batches spread over
  // every SM (N = 20 000 is 157 tiles: by tiles 79 CTAs would take two each and 69 SMs none)
  const int64_t total_st = a.n_tiles * (kBTile / kWgStage);
  const int64_t per = (total_st + gridDim.x - 1) / gridDim.x;
  const int64_t g_begin = static_cast<int64_t>(blockIdx.x) * per < total_st ? static_cast<int64_t>(blockIdx.x) * per : total_st;
  const int64_t g_end = g_begin + per < total_st ? g_begin + per : total_st;
  const int64_t n_stage = g_end - g_begin;

  if (mma_warp) {
    // ================================================================== MMA warp
    const uint32_t base = tc::smem_u32(wsm);
    uint32_t par[2] = {0u, 0u};
    for (int64_t st = 0; st < n_stage; ++st) {
      const int buf = static_cast<int>(st & 1);
      tc::mbar_wait(&full[buf], par[buf]);
      par[buf] ^= 1u;
      __syncwarp();
#ifdef PINN_TIMELINE
      if (blockIdx.x == 0 && (tid & 31) == 0 && st >= 16 && st < 80) g_tlw[0][st - 16][0] = clock64();
#endif
      if (tc::elect_one()) {
        tc::fence_after_sync();
        const uint32_t hi = base + buf * wl.buffer_bytes, lo = hi + wl.plane_bytes;
        const uint32_t first = st == 0 ? 0u : 1u;
        // one 3xTF32 product over the stage's two K = 8 slabs
        auto prod = [&](int a_off, int b_off, int n_rows, int col) {
          const uint32_t lbo_a = wg_lbo(128), lbo_b = wg_lbo(n_rows), idesc = tc::make_idesc_tf32(128, n_rows);
          const uint64_t a_hi = tc::make_desc(hi + a_off, lbo_a, 128), a_lo = tc::make_desc(lo + a_off, lbo_a, 128);
          const uint64_t b_hi = tc::make_desc(hi + b_off, lbo_b, 128), b_lo = tc::make_desc(lo + b_off, lbo_b, 128);
          const uint64_t as = (2u * lbo_a) >> 4, bs = (2u * lbo_b) >> 4;
          const uint32_t d = tmem_base_s + static_cast<uint32_t>(col);
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t aa = term == 0 ? a_lo : a_hi, bb = term == 1 ? b_lo : b_hi;
#pragma unroll
            for (int k8 = 0; k8 < kWgStage / 8; ++k8)
              tc::umma_tf32(d, aa + k8 * as, bb + k8 * bs, idesc, (term | k8) != 0 ? 1u : first);
          }
        };
        if (NBIG > 0) prod(wl.a_big[0], wl.b_big[0], wg_big_n(L, 0), 0);
        if (NBIG > 1) prod(wl.a_big[1], wl.b_big[1], wg_big_n(L, 1), wg_big_n(L, 0));
        prod(wl.a_small, wl.b_small, kSmN, COL_SMALL);
        tc::umma_commit(&done[buf]);
      }
      __syncwarp();
#ifdef PINN_TIMELINE
      if (blockIdx.x == 0 && (tid & 31) == 0 && st >= 16 && st < 80) g_tlw[0][st - 16][1] = clock64();
#endif
    }
  } else {
    // ================================================================== loaders
    // work items of a stage: (row, 4-sample chunk), 4 chunks per row; item i -> row i / 4, chunk i % 4 of the
    // tile's row table.  The (row -> operand block, block row) map is the same every stage: resolved once.
    constexpr int kItems = (4 * (128 * (L / 2) + 128 + 64 * L + 48) + kWgLoaders - 1) / kWgLoaders;    // ceil(rows * 4 / 256)
    int dst[kItems];          // byte offset inside a plane, or -1: row not used
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int i = tid + it * kWgLoaders, r = i >> 2, c = i & 3;
      int off = -1;
      if (r < rm.RA) { if (r < 64 * (L - 1)) off = wl.a_big[r >> 7] + c * wg_lbo(128) + (r & 127) * 16; }
      else if (r < rm.RB) { if (r - rm.RA < 114) off = wl.a_small + c * wg_lbo(128) + (r - rm.RA) * 16; }
      else if (r < rm.RV0) {
        const int l = (r - rm.RB) >> 6, j = (r - rm.RB) & 63;
        if (l == L - 1) off = wl.b_small + c * wg_lbo(kSmN) + j * 16;
        else off = wl.b_big[l >> 1] + c * wg_lbo(wg_big_n(L, l >> 1)) + (64 * (l & 1) + j) * 16;
      }
      else if (r < rm.RV1) off = wl.b_small + c * wg_lbo(kSmN) + (kSmV0 + r - rm.RV0) * 16;
      else if (r < rm.rows) off = wl.b_small + c * wg_lbo(kSmN) + (kSmV1 + r - rm.RV1) * 16;
      dst[it] = off;
    }
    float4 ldA[kItems], ldB[kItems];                            // two stages of loads in flight per thread (HBM latency)
    float4 ldxA = make_float4(0.f, 0.f, 0.f, 0.f), ldxB = ldxA;
    auto load_stage = [&](int64_t st, float4 (&ld)[kItems], float4& ldx) {
      const int64_t gs = g_begin + st;
      const int64_t tile = gs / (kBTile / kWgStage);
      const int s0 = static_cast<int>(gs % (kBTile / kWgStage)) * kWgStage;
      const float* base = a.rows + static_cast<size_t>(tile) * rm.rows * kBTile + static_cast<size_t>(s0 / kWgStage) * rm.rows * kRowStride;   // one contiguous slab
#pragma unroll
      for (int it = 0; it < kItems; ++it) {
        const int i = tid + it * kWgLoaders;
        if (dst[it] >= 0) ld[it] = __ldcs(reinterpret_cast<const float4*>(base) + i);
      }
      if (tid < 32) {             // x[N][8]: thread -> (sample tid / 2, 4 features)
        const int64_t s = tile * kBTile + s0 + (tid >> 1);
        ldx = s < a.n ? __ldg(reinterpret_cast<const float4*>(a.x + s * PINN_N_IN) + (tid & 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_stage = [&](int buf, const float4 (&ld)[kItems], const float4& ldx) {
      unsigned char* hi = wsm + buf * wl.buffer_bytes;
      unsigned char* lo = hi + wl.plane_bytes;
#pragma unroll
      for (int it = 0; it < kItems; ++it) {
        if (dst[it] < 0) continue;
        const float4 v = ld[it];
        const float4 h = make_float4(tc::tf32_hi_fast(v.x), tc::tf32_hi_fast(v.y), tc::tf32_hi_fast(v.z), tc::tf32_hi_fast(v.w));
        *reinterpret_cast<float4*>(hi + dst[it]) = h;
        *reinterpret_cast<float4*>(lo + dst[it]) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
      if (tid < 32) {             // x^T: feature rows 4 (tid & 1) .. +3, sample tid / 2 of the stage
        const int sl = tid >> 1, f0 = 4 * (tid & 1);
        const float vv[4] = {ldx.x, ldx.y, ldx.z, ldx.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float h = tc::tf32_hi_fast(vv[q]);
          const int off = wl.b_small + (sl >> 2) * wg_lbo(kSmN) + (kSmX + f0 + q) * 16 + (sl & 3) * 4;
          *reinterpret_cast<float*>(hi + off) = h;
          *reinterpret_cast<float*>(lo + off) = vv[q] - h;
        }
      }
    };
    uint32_t dpar[2] = {0u, 0u};
    if (n_stage > 0) load_stage(0, ldA, ldxA);
    if (n_stage > 1) load_stage(1, ldB, ldxB);
    auto step = [&](int64_t st, float4 (&ld)[kItems], float4& ldx) {
      const int buf = static_cast<int>(st & 1);
#ifdef PINN_TIMELINE
      const bool tlon = blockIdx.x == 0 && tid == 0 && st >= 16 && st < 80;
      if (tlon) g_tlw[1][st - 16][0] = clock64();
#endif
      if (st >= 2) { tc::mbar_wait(&done[buf], dpar[buf]); dpar[buf] ^= 1u; }     // the MMAs of stage st-2 have drained this buffer
#ifdef PINN_TIMELINE
      if (tlon) g_tlw[1][st - 16][1] = clock64();
#endif
      store_stage(buf, ld, ldx);
#ifdef PINN_TIMELINE
      if (tlon) g_tlw[1][st - 16][2] = clock64();
#endif
      if (st + 2 < n_stage) load_stage(st + 2, ld, ldx);
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[buf]);
#ifdef PINN_TIMELINE
      if (tlon) g_tlw[1][st - 16][3] = clock64();
#endif
    };
    for (int64_t st = 0; st < n_stage; st += 2) {
      step(st, ldA, ldxA);
      if (st + 1 < n_stage) step(st + 1, ldB, ldxB);
    }
    // drain: the last (up to) two commits
    for (int64_t st = n_stage > 2 ? n_stage - 2 : 0; st < n_stage; ++st) {
      const int buf = static_cast<int>(st & 1);
      tc::mbar_wait(&done[buf], dpar[buf]);
      dpar[buf] ^= 1u;
    }
    tc::fence_after_sync();
    TLP(0, 2);
    griddep_launch();        // late trigger: the reduce CTAs would otherwise sit resident (and waiting) through the whole streaming phase
    // ---------------------------------------------------------------- accumulators -> this CTA's partial
    float* part = a.partial + static_cast<size_t>(blockIdx.x) * lay.total;
    {
      // All eight loader warps read out: warp w owns TMEM lanes 32 (w & 3) .. +31 (the hardware's lane quadrant rule),
      // and the two warps of a quadrant take alternate 16-column groups.  Weight rows go out as 16-byte stores.
      const int wq = warp & 3, wh = warp >> 2;
      const int lane_row = wq * 32 + (tid & 31);
      const uint32_t tl = tmem_base_s + (static_cast<uint32_t>(wq * 32) << 16);
      const bool have = n_stage > 0;
      auto read16 = [&](int col, float (&v)[16]) {
        if (have) { tc::tmem_ld16(tl + static_cast<uint32_t>(col), v); tc::tmem_wait_ld(); }
        else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0.f;
        }
      };
      auto store16_scaled = [&](float* dst, const float (&v)[16]) {     // dst 16-byte aligned
#pragma unroll
        for (int q = 0; q < 16; q += 4)
          *reinterpret_cast<float4*>(dst + q) = make_float4(v[q] * a.act_scale, v[q + 1] * a.act_scale, v[q + 2] * a.act_scale, v[q + 3] * a.act_scale);
      };
      float v[16];
      int item = 0;        // compile-time after unrolling: work item counter, item & 1 selects the warp of the quadrant
      // trunk layers l >= 1: delta_l = lanes 64 h + j of big block i, a_{l-1} = columns 64 h + k of its B block (i = (l-1)/2, h = (l-1)%2)
#pragma unroll
      for (int l = 1; l < L; ++l) {
        const int i = (l - 1) >> 1, h = (l - 1) & 1;
        const int col0 = (i == 0 ? 0 : wg_big_n(L, 0)), ones_col = col0 + wg_big_n(L, i) - 16;
        const int j = lane_row - 64 * h;
        const bool mine = j >= 0 && j < 64;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          if ((item++ & 1) == wh) {
            read16(col0 + 64 * h + c, v);
            if (mine) store16_scaled(part + lay.offW[l] + j * 64 + c, v);
          }
        }
        if ((item++ & 1) == wh) {
          read16(ones_col, v);
          if (mine) part[lay.offb[l] + j] = v[0];
        }
      }
      // small product
#pragma unroll
      for (int c = 0; c < 64; c += 16) {        // columns 0..63 = a_{L-1}: lanes 64..95 -> dWv0, lane 96 -> dWp
        if ((item++ & 1) == wh) {
          read16(COL_SMALL + c, v);
          if (lane_row >= 64 && lane_row < 96) store16_scaled(part + lay.offWv0 + (lane_row - 64) * 64 + c, v);
          else if (lane_row == 96) store16_scaled(part + lay.offWp + c, v);
        }
      }
      if ((item++ & 1) == wh) {
        read16(COL_SMALL + 64, v);                // column 64 = ones (every bias); 65..72 = x^T; 73..79 = av0[0..7)
        if (lane_row < 64) {
          part[lay.offb[0] + lane_row] = v[0];
#pragma unroll
          for (int q = 0; q < 8; ++q) part[lay.offW[0] + lane_row * 8 + q] = v[1 + q];
        } else if (lane_row < 96) part[lay.offbv0 + lane_row - 64] = v[0];
        else if (lane_row == 96) part[lay.offbp] = v[0];
        else if (lane_row < 113) part[lay.offbv1 + lane_row - 97] = v[0];
        else if (lane_row == 113) part[lay.offbv2] = v[0];
        if (lane_row >= 97 && lane_row < 113) {
#pragma unroll
          for (int q = 0; q < 7; ++q) part[lay.offWv1 + (lane_row - 97) * 32 + q] = v[9 + q];
        }
      }
#pragma unroll
      for (int c = 80; c < 128; c += 16) {      // columns 80..104 = av0[7..32), 105..120 = av1
        if ((item++ & 1) == wh) {
          read16(COL_SMALL + c, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int col = c + q;
            if (col < kSmV1) { if (lane_row >= 97 && lane_row < 113) part[lay.offWv1 + (lane_row - 97) * 32 + col - kSmV0] = v[q]; }
            else if (col < kSmV1 + 16) { if (lane_row == 113) part[lay.offWv2 + col - kSmV1] = v[q]; }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  TLP(0, 3);
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
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
  float* rows;              // transposed scratch, [tile][rm.rows][128] (see RowMap)
  RowMap rm;
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
  TLP(1, 0);
  griddep_launch();

  if (tid == 0) { tc::mbar_init(&mbar[0], 1); tc::mbar_init(&mbar[1], 1); tc::fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  griddep_wait();          // the weights come from the previous step's optimiser launch
  const bool drop_on = dp.p > 0.f;
  const float wscale = drop_on ? dp.scale : 1.0f;
  // Weight staging with every global load issued before the first use: the former sequence of 15 load -> store loops
  // paid one L2 round trip each (6.9 us per launch, a tenth of the whole step at N = 20 000).
  {
    // small tensors: one element per thread.  W0 [64 x 8] and Wv1 [16 x 32] are 512 floats each; the biases, Wv2, bp, bv2
    // are laid over the thread index: [64 l, 64 l + 64) = b[l], [256, 288) = bv0, [288, 304) = bv1, [304, 320) = Wv2, 320 = bp, 321 = bv2
    const float w0 = __ldg(net.W[0] + tid), wv1 = __ldg(net.Wv1 + tid);
    const float* sp = nullptr;
    float* dstp = nullptr;
    float sc = 1.0f;
    if (tid < 64 * L) { const int l = tid >> 6, j = tid & 63; sp = net.b[l] + j; dstp = smem + (l == 0 ? lay.b0 : lay.b[l]) + j; sc = kTanhArg; }
    else if (tid >= 256 && tid < 288) { sp = net.bv0 + (tid - 256); dstp = smem + lay.bv0 + (tid - 256); sc = kTanhArg; }
    else if (tid >= 288 && tid < 304) { sp = net.bv1 + (tid - 288); dstp = smem + lay.bv1 + (tid - 288); sc = kTanhArg; }
    else if (tid >= 304 && tid < 320) { sp = net.Wv2 + (tid - 304); dstp = smem + lay.Wv2 + (tid - 304); }
    else if (tid == 320) { sp = net.bp; dstp = smem + lay.bp; }
    else if (tid == 321) { sp = net.bv2; dstp = smem + lay.bv2; }
    const float sv = sp != nullptr ? __ldg(sp) : 0.f;
    float4 wv[RES ? L : 1][2];
    if constexpr (RES) {
#pragma unroll
      for (int l = 1; l <= L; ++l) {
        const float* src = l < L ? net.W[l] : net.Wv0;
        const int J = l < L ? H : 32;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int idx = tid + it * 512, j = idx & 63, kc = idx >> 6;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < J) v = __ldg(reinterpret_cast<const float4*>(src + j * 64) + kc);
          else if (l == L && j == J) v = __ldg(reinterpret_cast<const float4*>(net.Wp) + kc);
          wv[l - 1][it] = v;
        }
      }
    }
    smem[lay.W0 + tid] = w0 * kTanhArg;      // tanh_pre arguments (common.cuh)
    smem[lay.Wv1 + tid] = wv1;
    if (dstp != nullptr) *dstp = sv * sc;
    if constexpr (RES) {
#pragma unroll
      for (int l = 1; l <= L; ++l) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int idx = tid + it * 512, j = idx & 63, kc = idx >> 6;
          const float4 v = wv[l - 1][it];
          const float vv[4] = {v.x * wscale, v.y * wscale, v.z * wscale, v.w * wscale};
          tc::store_split4(smem + lay.wf_hi[l], smem + lay.wf_lo[l], 64 * 16, j, kc, make_float4(vv[0], vv[1], vv[2], vv[3]));
          float* thi = smem + lay.wt_hi[l];
          float* tlo = smem + lay.wt_lo[l];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float h = tc::tf32_hi(vv[r]);
            const size_t off = (static_cast<size_t>(j >> 2) * kLboT + static_cast<size_t>(4 * kc + r) * 16 + (j & 3) * 4) / 4;
            thi[off] = h;
            tlo[off] = vv[r] - h;
          }
        }
      }
      tc::fence_proxy_async();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  TLP(1, 1);
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
    for (int q = 0; q < 8; q += 2) {
      h[q] = tc::tf32_hi_fast(v[q]); h[q + 1] = tc::tf32_hi_fast(v[q + 1]);
      const float2 l2 = __ffma2_rn(make_float2(h[q], h[q + 1]), make_float2(-1.0f, -1.0f), make_float2(v[q], v[q + 1]));   // v - h, exact
      lo[q] = l2.x; lo[q + 1] = l2.y;
    }
    tc::tmem_st8(a_hi_t + lane_sel + static_cast<uint32_t>(c0), h);
    tc::tmem_st8(a_lo_t + lane_sel + static_cast<uint32_t>(c0), lo);
  };

  // keep bits of this thread's 32 units [c0, c0 + 32) of dropout layer `layer` (bit q = unit c0 + q); computed
  // while the tensor core works (they depend on nothing it produces)
  auto keep_bits32 = [&](const KeepSrc<INJ>& ks, bool active, uint32_t layer, int c0) {
    uint32_t bits = 0xffffffffu;
    if (active) {
      bits = 0u;
#pragma unroll
      for (int g = 0; g < 32; g += 8) {
        bool k[8];
        ks.get8(dp, layer, static_cast<uint32_t>(c0 + g), layer * H, k);
#pragma unroll
        for (int q = 0; q < 8; ++q) bits |= (k[q] ? 1u : 0u) << (g + q);
      }
    }
    return bits;
  };
  const int64_t n_tiles = (a.n + kBTile - 1) / kBTile;
  // tile -> (CTA, group): group 0 of every CTA first, then group 1 -- with no more tiles than SMs every tile gets an SM
  // to itself (a lone tile is faster than an interleaved pair; only throughput, not latency, gains from pairing)
#ifdef PINN_K2A_PAIR_FIRST     // A/B build: the former mapping (tiles 2b, 2b+1 on CTA b)
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * 2 + grp; tile < n_tiles; tile += static_cast<int64_t>(gridDim.x) * 2) {
#else
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(grp) * gridDim.x; tile < n_tiles; tile += static_cast<int64_t>(gridDim.x) * 2) {
#endif
    const int64_t s = tile * kBTile + row;
    const bool valid = s < a.n;
    // this sample's column of the tile's row table: slab (row / 16) of the tile, position row % 16 inside each 16-sample row
    float* const trow = a.rows + static_cast<size_t>(tile) * a.rm.rows * kBTile + static_cast<size_t>(row >> 4) * a.rm.rows * kRowStride + (row & 15);
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
#pragma unroll
        for (int q = 0; q < 8; ++q) trow[static_cast<size_t>(row_act(a.rm, 0) + cb + g + q) * kRowStride] = valid ? v[q] : 0.f;
      }
    }
#pragma unroll 1
    for (int l = 1; l < L; ++l) {       // rolled: one copy of the layer body keeps the kernel inside the I-cache
      if constexpr (!RES) wcommit_rows(b_hi, b_lo, wpre, t256, wscale);
      uint32_t bits = 0u;
      run_mma(lay.wf_hi[l], lay.wf_lo[l], H * 16, idesc64, [&] {
        if constexpr (!RES) {
          if (l + 1 < L) wprefetch(wpre, net.W[l + 1], nullptr, H, t256);
          else wprefetch(wpre, net.Wv0, net.Wp, 32, t256);
        }
        bits = keep_bits32(ks, active, static_cast<uint32_t>(l), cb);
      });
      const float* bl = smem + lay.b[l] + cb;
#pragma unroll 1
      for (int g = 0; g < HH; g += 8) {
        float z[8];
        tc::tmem_ld8(d_lane + cb + g, z);
        tc::tmem_wait_ld();
        const float4 bA = *reinterpret_cast<const float4*>(bl + g), bB = *reinterpret_cast<const float4*>(bl + g + 4);
        const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
        const uint32_t kbg = bits >> g;
        float t8[8], v[8];
        tanh8_prescaled(z, bb, t8);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = ((kbg >> q) & 1u) ? t8[q] : 0.f;
        store_pn8(cb + g, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) trow[static_cast<size_t>(row_act(a.rm, l) + cb + g + q) * kRowStride] = valid ? v[q] : 0.f;
      }
      kb[l] = bits;
    }
    // ---- heads: rows 0..31 = Wv0, row 32 = Wp, rows 33.. = 0 (N = 48 of the 64 staged rows are read)
    if constexpr (!RES) wcommit_rows(b_hi, b_lo, wpre, t256, wscale);
    kb[L] = 0u;
    run_mma(lay.wf_hi[L], lay.wf_lo[L], H * 16, idesc48, [&] {
      if constexpr (!RES) wprefetch(wpre, net.Wv0, net.Wp, 32, t256);
      if (half == 0) kb[L] = keep_bits32(ks, active, static_cast<uint32_t>(L), 0);
    });
    float du = 0.f;
    float dzv0[HH];                      // half 0: d z of the variance head's first layer; half 1: unused
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
#pragma unroll
        for (int q = 0; q < 8; ++q)
          v0[g + q] = ((kb[L] >> (g + q)) & 1u) ? tanh_pre(fmaf(v0[g + q], kTanhArg, bv0[g + q])) * wscale : 0.f;   // scaled, as K2b expects
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
      const bool no_lv = (net.flags & PINN_NET_NO_LOGVAR) != 0;
      const float lv = logvar_out(vraw, no_lv);
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
      const float dv = no_lv ? 0.f : ds * dlogvar_dv(vraw);
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
      {   // rows of an invalid sample (tail of the last tile) are written as zeros: K2b sums whole tiles
        float* const pa = trow + static_cast<size_t>(a.rm.RA) * kRowStride;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          trow[static_cast<size_t>(a.rm.RV0 + i) * kRowStride] = valid ? v0[i] : 0.f;
          pa[static_cast<size_t>(64 + i) * kRowStride] = dzv0[i];
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          trow[static_cast<size_t>(a.rm.RV1 + k) * kRowStride] = valid ? v1[k] : 0.f;
          pa[static_cast<size_t>(97 + k) * kRowStride] = dz1[k];
        }
        pa[static_cast<size_t>(96) * kRowStride] = du;
        pa[static_cast<size_t>(113) * kRowStride] = dv;
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
      float apre[HH];            // this thread's masked activations of layer l, prefetched during the MMA
      run_mma(lay.wt_hi[l + 1], lay.wt_lo[l + 1], kLboT, idesc64, [&] {
#pragma unroll
        for (int q = 0; q < HH; ++q) apre[q] = trow[static_cast<size_t>(row_act(a.rm, l) + cb + q) * kRowStride];
        if constexpr (!RES) { if (l > 0) wprefetch(wpre, net.W[l], nullptr, H, t256); }
      });
      const uint32_t kbl = kb[l];
#pragma unroll
      for (int g = 0; g < HH; g += 8) {
        float z[8], dz[8];
        tc::tmem_ld8(d_lane + cb + g, z);
        tc::tmem_wait_ld();
        const float* aa = apre + g;
#pragma unroll
        for (int q = 0; q < 8; ++q)    // z already carries the dropout scale (folded into W^T); dropped units have a = 0
          dz[q] = ((kbl >> (g + q)) & 1u) ? z[q] * fmaf(-aa[q], aa[q], 1.0f) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) trow[static_cast<size_t>(row_del(a.rm, l) + cb + g + q) * kRowStride] = dz[q];
        if (l > 0) store_pn8(cb + g, dz);
      }
      if constexpr (!RES) { if (l > 0) wcommit_transposed(b_hi, b_lo, wpre, t256, wscale); }
    }
  }
  TLP(1, 2);
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
  TLP(1, 3);
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace pinn
#include "mlp_tc_fused.cuh"
namespace pinn {

// Per-CTA partials -> gradient bucket, fixed order (deterministic): a CTA owns 32 consecutive bucket entries (one
// 128-byte line per partial), its 8 warps sum every 8th partial with four loads in flight, the 8 sub-sums are
// folded in warp order.  (The first version walked all partials serially in one thread per entry -- 24 us at
// N = 20 000, a quarter of the whole train_dnn step there.)  With `fa.params` set the same launch applies
// Adam + StepLR to the bucket entry it just reduced (single-GPU train_dnn: no all-reduce sits in between).
#ifndef PINN_RED_GROUPS
#define PINN_RED_GROUPS 16      // 8: +0.7..2 us per step at N = 19..20 k, 32: +8 us (profiles/c1_train_dnn_launches.py)
#endif
constexpr int kRedCols = 32, kRedGroups = PINN_RED_GROUPS;
__global__ void __launch_bounds__(kRedCols * kRedGroups)
grad_reduce2_kernel(const float* __restrict__ partial, const double* __restrict__ loss_partial, int nblk, int nloss, int64_t total,
                    float* __restrict__ grad, double* __restrict__ loss, const FusedAdam fa, const ParamLayout lay) {
  __shared__ double fold[kRedGroups][kRedCols];
  __shared__ float consts[2];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kRedCols + c;
  griddep_launch();
  griddep_wait();
  int64_t t0 = 0;
  if (fa.params != nullptr) {
    t0 = *fa.step_counter;
    if (threadIdx.x == 0) {
      const double lr = fa.h.lr0 * pow(fa.h.gamma, static_cast<double>(t0 / fa.h.step_size));
      adam_consts(lr, t0 + 1, consts[0], consts[1]);
    }
  }
  double acc = 0.0;
  if (i < total) {
    const float* p = partial + i;
    int b = g;
    for (; b + 3 * kRedGroups < nblk; b += 4 * kRedGroups) {
      const float v0 = p[static_cast<size_t>(b) * total], v1 = p[static_cast<size_t>(b + kRedGroups) * total];
      const float v2 = p[static_cast<size_t>(b + 2 * kRedGroups) * total], v3 = p[static_cast<size_t>(b + 3 * kRedGroups) * total];
      acc += static_cast<double>(v0); acc += static_cast<double>(v1); acc += static_cast<double>(v2); acc += static_cast<double>(v3);
    }
    for (; b < nblk; b += kRedGroups) acc += static_cast<double>(p[static_cast<size_t>(b) * total]);
  }
  fold[g][c] = acc;
  __syncthreads();
  if (g == 0) {
    float gr = 0.f;
    if (i < total) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < kRedGroups; ++k) t += fold[k][c];
      gr = layout_is_padding(lay, i) ? 0.0f : static_cast<float>(t);     // padding slots hold no partial sums
    }
    if (fa.peers != nullptr) {
      // Data parallel: this CTA owns one 128-byte line of the bucket on EVERY rank.  Push the local sums of the line into
      // slot (tag & 1), row `rank`, of every rank's symmetric buffer (posted NVLink stores), raise the line's flag there,
      // wait for the same line of every rank in the LOCAL buffer, add the rows in rank order (bit-identical replicas).
      // No grid-wide rendezvous, no NVLink loads; two slots make it race-free without a second barrier (a rank can be at
      // most one step ahead: step t+1 cannot complete anywhere before every rank has pushed -- hence finished reading -- t).
      const int64_t lines = dp_lines(total), fw = dp_flag_words(total, fa.world);
      const size_t row = (static_cast<size_t>(fa.tag & 1u) * fa.world + fa.rank) * total;
      if (i < total) {
        for (int r = 0; r < fa.world; ++r)
          __stcg(reinterpret_cast<float*>(fa.peers[r]) + fw + row + i, gr);
      }
      __threadfence_system();
      __syncwarp();
      if (c < fa.world) {
        unsigned int* flag = reinterpret_cast<unsigned int*>(fa.peers[c]) + static_cast<size_t>(fa.rank) * lines + blockIdx.x;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(fa.tag) : "memory");
        const unsigned int* mine = reinterpret_cast<const unsigned int*>(fa.peers[fa.rank]) + static_cast<size_t>(c) * lines + blockIdx.x;
        unsigned int seen;
        do {
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        } while (static_cast<int>(seen - fa.tag) < 0);
      }
      __syncwarp();
      if (i < total) {
        const float* base = reinterpret_cast<const float*>(fa.peers[fa.rank]) + fw + static_cast<size_t>(fa.tag & 1u) * fa.world * total + i;
        float sum = 0.f;
        for (int r = 0; r < fa.world; ++r) sum += __ldcv(base + static_cast<size_t>(r) * total);
        gr = sum;
      }
    }
    if (i < total) {
      if (grad != nullptr) grad[i] = gr;
      if (fa.params != nullptr) {
        float pp = fa.params[i], mm = fa.m[i], vv = fa.v[i];
        adam_apply(pp, static_cast<float>(static_cast<double>(gr) * fa.h.grad_scale), mm, vv, consts[0], consts[1], 0.f, 0.f, false);
        fa.params[i] = pp; fa.m[i] = mm; fa.v[i] = vv;
        if (fa.images != nullptr) image_write_entry(lay, i, pp, fa.wscale, fa.images);
      }
    }
  }
  if (loss != nullptr && blockIdx.x == 0 && g == 1) {
    // lane l sums entries l, l+32, ...; xor tree over the lanes (fixed order)
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = c; b < nloss; b += 32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) a4[k] += loss_partial[static_cast<size_t>(b) * 4 + k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a4[k] += __shfl_xor_sync(0xffffffffu, a4[k], o);
      if (c == 0) loss[k] = a4[k];
    }
  }
  if (fa.params != nullptr) {
    // every thread has read t0; the last CTA to arrive bumps the counter (same protocol as adam_kernel)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int* ticket = reinterpret_cast<unsigned int*>(fa.step_counter + 1);
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0u; *fa.step_counter = t0 + 1; }
    }
  }
}


struct TcBwdPlan { int grid_a, grid_b; size_t smem_a, smem_b, off_partial, off_scratch, bytes; };
static TcBwdPlan plan_tc_bwd(int L, int64_t n) {
  TcBwdPlan p{};
  ParamLayout lay = make_layout(kBH, L);
  const int sms = sm_count();
  const int64_t tiles = (n + kBTile - 1) / kBTile;
#ifdef PINN_K2A_PAIR_FIRST
  int64_t want = (tiles + 1) / 2;
  p.grid_a = static_cast<int>(want < sms ? (want > 0 ? want : 1) : sms);
#else
  p.grid_a = static_cast<int>(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
#endif
  const int64_t want_b = (tiles * (kBTile / kWgStage) + 3) / 4;       // K2b splits by 16-sample stages: at least four per CTA
  p.grid_b = static_cast<int>(want_b < sms ? (want_b > 0 ? want_b : 1) : sms);
  p.smem_a = static_cast<size_t>(make_tcb_layout(L).total) * sizeof(float);
  p.smem_b = static_cast<size_t>(2) * make_wg_layout(L).buffer_bytes;
  size_t off = static_cast<size_t>(2 * p.grid_a) * 4 * sizeof(double);
  p.off_partial = off;
  off += static_cast<size_t>(p.grid_b) * lay.total * sizeof(float);
  off = (off + 255) & ~static_cast<size_t>(255);
  p.off_scratch = off;
  off += bwd_scratch_floats(L, n) * sizeof(float);
  p.bytes = off;
  return p;
}

// programmatic dependent launch between the launches of a step: 0 never, 1 small batches only (default), 2 always -- chosen per call
// by pinn_net_t.flags (PINN_NET_PDL_NEVER / PINN_NET_PDL_ALWAYS); there is no process-global switch
int dependent_launch_mode(const pinn_net_t* net) { return (net->flags & PINN_NET_PDL_NEVER) ? 0 : ((net->flags & PINN_NET_PDL_ALWAYS) ? 2 : 1); }
bool tc_bwd_covers(const pinn_net_t* net) {
  if ((net->flags & PINN_NET_NO_TC_BWD) || net->width != kBH || net->n_hidden < 2 || net->n_hidden > 4) return false;
  for (int l = 0; l < net->n_hidden; ++l)
    if (!aligned16(net->W[l])) return false;
  return aligned16(net->Wv0) && aligned16(net->Wp);
}
// One-kernel form unless the caller opts out.  (A first version lost to the two-kernel form between one and 1.6 tiles per SM,
// where a few CTAs run two tiles back to back; since the tile got down to ~14 us it wins at every batch size:
// N = 20 000, 157 tiles: 51 vs 56 us per step, profiles/c1_train_dnn_launches.py.)
bool tc_bwd_fused(int L, int flags, int64_t n) {
  (void)n;
  return (L == 2 || L == 3) && !(flags & PINN_NET_NO_FUSED_BWD);
}
size_t tc_bwd_workspace_bytes(int L, int64_t n, int flags) {
  if (flags >= 0 && tc_bwd_fused(L, flags, n)) return plan_fused(L, n).bytes;
  const size_t two = plan_tc_bwd(L, n).bytes;      // flags < 0: any path
  const size_t one = (L == 2 || L == 3) ? plan_fused(L, n).bytes : 0;
  return two > one ? two : one;
}

// One-kernel form (mlp_tc_fused.cuh): weight images, the fused tile kernel, the fixed-order reduce (+ Adam).
static int launch_tc_fused(const pinn_net_t* net, const float* x, int64_t n, const DropParams& dp, const float* grad_u,
                           const float* grad_s, const float* y, int64_t n_global, float* grad_flat, double* loss_sums, void* workspace,
                           size_t workspace_bytes, cudaStream_t st, const FusedAdam* fused) {
  const int L = net->n_hidden;
  const FzPlan p = plan_fused(L, n);
  if (workspace_bytes < p.bytes) return PINN_E_WORKSPACE;
  const ParamLayout lay = make_layout(kBH, L);
  char* ws = static_cast<char*>(workspace);
  FzArgs a{};
  a.x = x; a.n = n; a.grad_u = grad_u; a.grad_s = grad_s; a.y = y;
  a.inv_n_global = grad_u ? 0.f : static_cast<float>(1.0 / static_cast<double>(n_global));
  a.images = reinterpret_cast<unsigned char*>(ws + p.off_images);
  a.partial = reinterpret_cast<float*>(ws + p.off_partial);
  a.loss_partial = reinterpret_cast<double*>(ws);
  a.park = reinterpret_cast<float*>(ws + p.off_park);
  const FzSmall sl = make_fz_small(L);
  const bool inj = dp.p > 0.f && dp.masks != nullptr;
  const int pdl_mode = dependent_launch_mode(net);
  const bool pdl = pdl_mode == 2 || (pdl_mode == 1 && (n + kBTile - 1) / kBTile <= static_cast<int64_t>(2) * sm_count());
  int devi = 0;
  if (cudaGetDevice(&devi) != cudaSuccess || devi < 0 || devi >= 64) devi = 0;
  static bool carve_set[64] = {false};
  if (!carve_set[devi]) {
    cudaFuncSetAttribute(grad_reduce2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(weight_image_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carve_set[devi] = true;
  }
  const float wscale = dp.p > 0.f ? dp.scale : 1.0f;
  if (!(fused != nullptr && fused->images_valid))      // else: written by the previous step's optimiser launch
    PINN_CUDA_TRY(launch_pdl(weight_image_kernel, dim3(4 * L), dim3(256), 0, st, pdl, *net, wscale,
                             reinterpret_cast<unsigned char*>(ws + p.off_images)));
#define LAUNCH_F(LL)                                                                                              \
  {                                                                                                               \
    auto kern = inj ? mlp_tc_fused_kernel<LL, true> : mlp_tc_fused_kernel<LL, false>;                             \
    static bool attr_f[64][2] = {};                                                                               \
    if (!attr_f[devi][inj ? 1 : 0]) {                                                                             \
      PINN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                         static_cast<int>(p.smem)));                                              \
      attr_f[devi][inj ? 1 : 0] = true;                                                                           \
    }                                                                                                             \
    PINN_CUDA_TRY(launch_pdl(kern, dim3(p.grid), dim3(kFzThreads + 32), p.smem, st, pdl, *net, sl, dp, a, lay));  \
  }
  if (L == 2) LAUNCH_F(2) else LAUNCH_F(3)
#undef LAUNCH_F
  PINN_CUDA_TRY(cudaGetLastError());
  const int rg = static_cast<int>((lay.total + kRedCols - 1) / kRedCols);
  FusedAdam fa{};
  if (fused != nullptr) { fa = *fused; fa.images = reinterpret_cast<unsigned char*>(ws + p.off_images); fa.wscale = wscale; }
  PINN_CUDA_TRY(launch_pdl(grad_reduce2_kernel, dim3(rg), dim3(kRedCols * kRedGroups), 0, st, pdl,
                           static_cast<const float*>(a.partial), static_cast<const double*>(a.loss_partial), p.grid, p.grid,
                           lay.total, grad_flat, loss_sums, fa, lay));
  return static_cast<int>(cudaGetLastError());
}

int launch_tc_bwd(const pinn_net_t* net, const float* x, int64_t n, const DropParams& dp, const float* grad_u,
                  const float* grad_s, const float* y, int64_t n_global, float* grad_flat, double* loss_sums, void* workspace,
                  size_t workspace_bytes, cudaStream_t st, const FusedAdam* fused) {
  const int L = net->n_hidden;
  if (tc_bwd_fused(L, net->flags, n))
    return launch_tc_fused(net, x, n, dp, grad_u, grad_s, y, n_global, grad_flat, loss_sums, workspace, workspace_bytes, st, fused);
  TcBwdPlan p = plan_tc_bwd(L, n);
  if (workspace_bytes < p.bytes) return PINN_E_WORKSPACE;
  ParamLayout lay = make_layout(kBH, L);
  char* ws = static_cast<char*>(workspace);
  TcbArgs a{};
  a.x = x; a.n = n; a.grad_u = grad_u; a.grad_s = grad_s; a.y = y;
  a.inv_n_global = grad_u ? 0.f : static_cast<float>(1.0 / static_cast<double>(n_global));
  a.rows = reinterpret_cast<float*>(ws + p.off_scratch);
  a.rm = make_rowmap(L);
  a.loss_partial = reinterpret_cast<double*>(ws);
  TcbLayout tl = make_tcb_layout(L);
  const bool inj = dp.p > 0.f && dp.masks != nullptr;
  // Dependent launch pays where launch latency and prologues are a visible share of the step (N = 20 000: 69 -> 66 us,
  // N = 5 000: 64 -> 57 us); with early-resident successors it costs 1 % at N = 100 000, 2 % at 200 000 and 3-7 % at 1M
  // (profiles/ab_pdl.py), so only batches of up to two tiles per SM use it by default.
  const int pdl_mode = dependent_launch_mode(net);
  const bool pdl = pdl_mode == 2 || (pdl_mode == 1 && (n + kBTile - 1) / kBTile <= static_cast<int64_t>(2) * sm_count());
  static bool carve_set[64] = {false};
  int devi = 0;
  if (cudaGetDevice(&devi) == cudaSuccess && devi >= 0 && devi < 64 && !carve_set[devi]) {
    // keep the SMs in the max-shared-memory configuration across the three launches of a step
    cudaFuncSetAttribute(grad_reduce2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carve_set[devi] = true;
  }
#define LAUNCH_A(LL)                                                                                              \
  {                                                                                                               \
    auto kern = inj ? mlp_tc_bwd_kernel<LL, true> : mlp_tc_bwd_kernel<LL, false>;                                 \
    static bool attr_a[64][2] = {};      /* the attribute is per device and function: set it once, not per step */ \
    if (!attr_a[devi & 63][inj ? 1 : 0]) {                                                                         \
      PINN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                         static_cast<int>(p.smem_a)));                                            \
      attr_a[devi & 63][inj ? 1 : 0] = true;                                                                       \
    }                                                                                                             \
    PINN_CUDA_TRY(launch_pdl(kern, dim3(p.grid_a), dim3(512), p.smem_a, st, pdl, *net, tl, dp, a));               \
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
  w.x = x; w.n = n; w.n_tiles = (n + kBTile - 1) / kBTile; w.rows = a.rows;
  w.act_scale = dp.p > 0.f ? dp.scale : 1.0f;
  w.partial = reinterpret_cast<float*>(ws + p.off_partial);
  const WgLayout wl = make_wg_layout(L);
#define LAUNCH_B(LL)                                                                                              \
  {                                                                                                               \
    static bool attr_b[64] = {};                                                                                  \
    if (!attr_b[devi & 63]) {                                                                                     \
      PINN_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                         static_cast<int>(p.smem_b)));                                            \
      attr_b[devi & 63] = true;                                                                                   \
    }                                                                                                             \
    PINN_CUDA_TRY(launch_pdl(wgrad_tc_kernel<LL>, dim3(p.grid_b), dim3(kWgLoaders + 32), p.smem_b, st,           \
                             pdl, w, lay, wl, a.rm));                                                             \
  }
  switch (L) {
    case 2: LAUNCH_B(2) break;
    case 3: LAUNCH_B(3) break;
    default: LAUNCH_B(4) break;
  }
#undef LAUNCH_B
  PINN_CUDA_TRY(cudaGetLastError());
  const int rg = static_cast<int>((lay.total + kRedCols - 1) / kRedCols);
  FusedAdam fa{};
  if (fused != nullptr) { fa = *fused; fa.images = nullptr; }
  PINN_CUDA_TRY(launch_pdl(grad_reduce2_kernel, dim3(rg), dim3(kRedCols * kRedGroups), 0, st, pdl,
                           static_cast<const float*>(w.partial), static_cast<const double*>(a.loss_partial), p.grid_b, 2 * p.grid_a,
                           lay.total, grad_flat, loss_sums, fa, lay));
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pinn

#ifdef PINN_TIMELINE
extern "C" int pinn_debug_timeline_wgrad(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, pinn::g_tlw, sizeof(pinn::g_tlw)));
}
extern "C" int pinn_debug_timeline_fused(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, pinn::g_tlf, sizeof(pinn::g_tlf)));
}
extern "C" int pinn_debug_timeline_phases(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, pinn::g_tlp, sizeof(pinn::g_tlp)));
}
#endif
