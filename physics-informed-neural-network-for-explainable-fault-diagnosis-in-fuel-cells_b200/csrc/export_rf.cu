// export_rf.cu -- K5: the callers downstream of the hot path.
//
//  * pinn_export_rows : the 22-column float64 `comprehensive_results` row of
//    create_comprehensive_results_array_v2 (01:1907-2010): un-scaled inputs / label /
//    prediction, uncertainty un-scaled and smoothed per segment with pandas' centred moving
//    average (01:1830-1872, window span [i-w/2, i+w/2-1], min_periods=1), prediction residual,
//    the four physics residuals, segment label, physical-model outputs.
//  * pinn_rf_stats / pinn_rf_series : estimate_mu_sigma_normal (04:181-197) and
//    compute_rf_time_series (04:201-285): z-score -> dead zone -> per-layer 2-norms ->
//    C_t = lambda C_{t-1} + S_t (first-order linear recurrence = associative scan, 3 phases)
//    -> logistic map -> EMA; plus the first-alarm index (04:289-300).  All float64 like the
//    reference; many independent series (stacks) per launch.
#include "common.cuh"

namespace pinn {

constexpr int kRowCols = 22;
constexpr int kRfCols = 6;        // compact RF input: res, pV, pT, pH, pO, label (= columns 12..17 of a full row)

__global__ void __launch_bounds__(256)
export_rows_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ pm,
                   const float* __restrict__ au, const float* __restrict__ eu, const float* __restrict__ cols,
                   const int64_t* __restrict__ seg_ends, int n_seg, int n_labeled, int window, pinn_export_scalers_t sc, int64_t n,
                   double* __restrict__ out, double* __restrict__ rf_cols /* optional [n][6]: columns 12..17, dense */) {
  __shared__ int64_t ends[64];
  for (int i = threadIdx.x; i < n_seg && i < 64; i += blockDim.x) ends[i] = seg_ends[i];
  __syncthreads();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // segment of row i (labels: 0 = first segment, 01:2013-2031)
  int seg = 0;
  int64_t s0 = 0, s1 = n;
  if (n_seg > 0) {
    while (seg < n_seg - 1 && i >= ends[seg]) ++seg;
    s0 = seg == 0 ? 0 : ends[seg - 1];
    s1 = ends[seg] < n ? ends[seg] : n;
    if (i >= s1) { s0 = s1; s1 = n; }       // rows past the last boundary: one trailing segment
  }
  double* o = out + i * kRowCols;
  // sklearn inverse_transform on an fp32 array: each in-place op in fp64, rounded to fp32
#pragma unroll
  for (int j = 0; j < PINN_N_IN; ++j) {
    const float t = static_cast<float>(static_cast<double>(x[i * PINN_N_IN + j]) - sc.x_min[j]);
    o[j] = static_cast<double>(static_cast<float>(static_cast<double>(t) / sc.x_scale[j]));
  }
  const float ty = static_cast<float>(static_cast<double>(y[i]) - sc.y_min);
  const double yr = static_cast<double>(static_cast<float>(static_cast<double>(ty) / sc.y_scale));
  const double den = sc.scale_y + 1e-12;                               // 01:1928-1932
  const double pr = (static_cast<double>(pm[i]) - sc.min_y) / den;
  const int half = window / 2;
  const int64_t lo = (i - half > s0) ? i - half : s0;
  const int64_t hi = (i + (window - half) < s1) ? i + (window - half) : s1;
  double sa = 0.0, se = 0.0;
  for (int64_t k = lo; k < hi; ++k) {
    sa += static_cast<double>(au[k]) / den;
    se += static_cast<double>(eu[k]) / den;
  }
  const double cnt = static_cast<double>(hi - lo);
  o[8] = yr; o[9] = pr; o[10] = sa / cnt; o[11] = se / cnt; o[12] = yr - pr;
  auto c = [&](int col) { return static_cast<double>(cols[static_cast<size_t>(col) * n + i]); };
  o[13] = c(PINN_C_FV); o[14] = c(PINN_C_FTS); o[15] = c(PINN_C_FH); o[16] = c(PINN_C_FO);
  o[17] = static_cast<double>(seg <= n_labeled ? seg : 0);   // 01:2026-2030: only listed fault segments get a label
  o[18] = c(PINN_C_VEST5); o[19] = c(PINN_C_TS_PRED); o[20] = c(PINN_C_H_ACT); o[21] = c(PINN_C_O_ACT);
  if (rf_cols != nullptr) {      // the six columns RF(t) consumes, gathered while they are in registers (48 dense bytes per row)
    double2* q = reinterpret_cast<double2*>(rf_cols + i * kRfCols);
    q[0] = make_double2(o[12], o[13]); q[1] = make_double2(o[14], o[15]); q[2] = make_double2(o[16], o[17]);
  }
}

// ------------------------------------------------------------------------------- RF(t)
// res, pV, pT, pH, pO, label = columns 12..17 of a comprehensive_results row (04:58-62,80), or columns 0..5 of the compact
// [n][6] form the row writer can emit next to it.  The layout travels as (doubles per row, first column).
struct RfLayout { int row_cols, col0; };
constexpr int kRfChunk = 2048, kRfThreads = 256, kRfPerThread = kRfChunk / kRfThreads;

// Columns 12..17 (five residual scores + the label) of a row: bytes 96..143 of its 176, 16-byte aligned -> three 128-bit
// loads, two 32-byte sectors of DRAM traffic per row instead of the whole 176-byte row.
struct RfRow { double r[5]; double label; };
PINN_D RfRow rf_load_row(const double* __restrict__ row /* already at the first RF column */) {
  const double2* p = reinterpret_cast<const double2*>(row);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  RfRow q;
  q.r[0] = a.x; q.r[1] = a.y; q.r[2] = b.x; q.r[3] = b.y; q.r[4] = c.x; q.label = c.y;
  return q;
}

// ONE pass over the rows (the first version read them twice: means, then squared deviations): per column the count and
// the shifted sums  S1 = sum (v - K),  S2 = sum (v - K)^2  over label-0, non-NaN rows, with the shift K = the column's
// value in the series' first row (any value within a few sigma of the mean keeps  S2 - S1^2 / n  well conditioned in
// float64; 0 if that value is NaN).  mean = K + S1 / n,  var = (S2 - S1^2 / n) / (n - 1)  (ddof 1, 04:195).
__global__ void __launch_bounds__(256)
rf_stats_partial_kernel(const double* __restrict__ res, int64_t n, RfLayout lay, double* __restrict__ partial /* [series][chunks][15] */) {
  const int series = blockIdx.y, nchunk = gridDim.x;
  const double* R = res + static_cast<size_t>(series) * n * lay.row_cols + lay.col0;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * kRfChunk;
  double K[5];
#pragma unroll
  for (int d = 0; d < 5; ++d) { const double k = __ldg(R + d); K[d] = k == k ? k : 0.0; }
  double s1[5] = {0, 0, 0, 0, 0}, s2[5] = {0, 0, 0, 0, 0}, cnt[5] = {0, 0, 0, 0, 0};
  for (int64_t t = c0 + threadIdx.x; t < c0 + kRfChunk && t < n; t += blockDim.x) {
    const RfRow q = rf_load_row(R + t * lay.row_cols);
    if (static_cast<int>(q.label) != 0) continue;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const double v = q.r[d];
      if (v == v) { const double e = v - K[d]; s1[d] += e; s2[d] = fma(e, e, s2[d]); cnt[d] += 1.0; }
    }
  }
  __shared__ double red[8][15];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    double a = s1[d], b = s2[d], c = cnt[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { red[warp][d] = a; red[warp][5 + d] = b; red[warp][10 + d] = c; }
  }
  __syncthreads();
  if (threadIdx.x < 15) {
    double a = 0.0;
    for (int wq = 0; wq < 8; ++wq) a += red[wq][threadIdx.x];
    partial[(static_cast<size_t>(series) * nchunk + blockIdx.x) * 15 + threadIdx.x] = a;
  }
}
// mu_sigma[series][0..5) = mean, [5..10) = sigma (ddof 1, 0 -> 1e-6).  One 256-thread block per series: thread (value v,
// part p) sums chunks p, p + 16, ... (independent loads), the 16 parts are folded in order -- a fixed summation order.
// (The first version walked the ~500 chunk partials of a series serially in one thread per value: 140 us of pure load
// latency per call, as long as the pass over the rows itself.)
__global__ void __launch_bounds__(256)
rf_stats_final_kernel(const double* __restrict__ res, int64_t n, RfLayout lay, const double* __restrict__ partial, int nchunk,
                      double* __restrict__ mu_sigma) {
  __shared__ double fold[16][16];
  const int series = blockIdx.x, v = threadIdx.x & 15, part = threadIdx.x >> 4;
  double acc = 0.0;
  if (v < 15) {
    const double* q = partial + static_cast<size_t>(series) * nchunk * 15 + v;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int k = part;
    for (; k + 48 < nchunk; k += 64) {
      a0 += q[static_cast<size_t>(k) * 15]; a1 += q[static_cast<size_t>(k + 16) * 15];
      a2 += q[static_cast<size_t>(k + 32) * 15]; a3 += q[static_cast<size_t>(k + 48) * 15];
    }
    for (; k < nchunk; k += 16) a0 += q[static_cast<size_t>(k) * 15];
    acc = (a0 + a1) + (a2 + a3);
  }
  fold[part][v] = acc;
  __syncthreads();
  const int d = threadIdx.x;
  if (d >= 5) return;
  double s1 = 0.0, s2 = 0.0, c = 0.0;
  for (int p = 0; p < 16; ++p) { s1 += fold[p][d]; s2 += fold[p][5 + d]; c += fold[p][10 + d]; }
  const double k0 = res[static_cast<size_t>(series) * n * lay.row_cols + lay.col0 + d];
  const double K = k0 == k0 ? k0 : 0.0;
  mu_sigma[series * 10 + d] = K + s1 / c;
  double sg = sqrt(fmax(s2 - s1 * s1 / c, 0.0) / (c - 1.0));
  if (sg == 0.0) sg = 1e-6;
  mu_sigma[series * 10 + 5 + d] = sg;
}

PINN_D double rf_strength(const double* __restrict__ row, const double* __restrict__ ms, double z_safe) {
  const RfRow q = rf_load_row(row);
  double a[5];
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    const double z = fabs((q.r[d] - ms[d]) / ms[5 + d]);
    a[d] = z != z ? z : fmax(0.0, z - z_safe);       // np.maximum propagates NaN (04:238), fmax would drop it
  }
  // layers {res,pV}, {pH,pO}, {pT} with p = 2, unit weights (04:84-96)
  return sqrt(a[0] * a[0] + a[1] * a[1]) + sqrt(a[3] * a[3] + a[4] * a[4]) + sqrt(a[2] * a[2]);
}

// First-order recurrence v_t = a_t v_{t-1} + b_t as an associative scan on pairs (a, b):
// (a1,b1) then (a2,b2)  ==  (a1 a2, a2 b1 + b2).
struct Lin { double a, b; };
PINN_D Lin lin_then(Lin f, Lin g) { return Lin{f.a * g.a, g.a * f.b + g.b}; }

// Exclusive scan of the per-thread maps across the block (warp shuffles, then warp totals): returns the composition of
// every map before this thread's; *total = the composition of the whole block (valid in every thread).
PINN_D Lin block_excl_scan(Lin loc, Lin* wtot /* shared, kRfThreads / 32 entries */, Lin* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Lin inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Lin prev{__shfl_up_sync(0xffffffffu, inc.a, o), __shfl_up_sync(0xffffffffu, inc.b, o)};
    if (lane >= o) inc = lin_then(prev, inc);
  }
  __syncthreads();                       // wtot may still be read from a previous scan
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  Lin before{1.0, 0.0}, all{1.0, 0.0};
  for (int wq = 0; wq < kRfThreads / 32; ++wq) {
    if (wq == warp) before = all;
    all = lin_then(all, wtot[wq]);
  }
  Lin excl{__shfl_up_sync(0xffffffffu, inc.a, 1), __shfl_up_sync(0xffffffffu, inc.b, 1)};
  if (lane == 0) excl = Lin{1.0, 0.0};
  *total = all;
  return lin_then(before, excl);
}

// phase 0: strengths S_t from the rows (written to `S_buf`, 8 bytes per row) + per-chunk aggregate; phase 1: re-reads S_t
// (not the rows: 8 instead of 64+ bytes per row), applies the carry and emits C, RF_inst.
template <int PHASE>
__global__ void __launch_bounds__(kRfThreads)
rf_scan_kernel(const double* __restrict__ res, int64_t n, RfLayout lay, const double* __restrict__ mu_sigma, pinn_rf_params_t prm,
               Lin* __restrict__ agg /* [series][chunks] */, const double* __restrict__ carry /* [series][chunks] */,
               double* __restrict__ rf_inst, double* __restrict__ C_out, double* __restrict__ S_buf /* [series][n] */,
               Lin* __restrict__ agg_ema /* phase 1: [series][chunks] aggregate of the EMA recurrence over RF_inst */) {
  const int series = blockIdx.y, nchunk = gridDim.x;
  const double* R = res + static_cast<size_t>(series) * n * lay.row_cols + lay.col0;
  const double* ms = mu_sigma + series * 10;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * kRfChunk + static_cast<int64_t>(threadIdx.x) * kRfPerThread;
  double S[kRfPerThread];
  Lin loc{1.0, 0.0};
#pragma unroll
  for (int k = 0; k < kRfPerThread; ++k) {
    const int64_t t = t0 + k;
    S[k] = 0.0;
    if (t < n) {
      const size_t si = static_cast<size_t>(series) * n + t;
      if (PHASE == 0) { S[k] = rf_strength(R + t * lay.row_cols, ms, prm.z_safe); S_buf[si] = S[k]; }
      else S[k] = S_buf[si];
      // C[0] = 0 regardless of S[0] (04:262-264): element 0 is the constant map v -> 0
      loc = lin_then(loc, t == 0 ? Lin{0.0, 0.0} : Lin{prm.lambda_decay, S[k]});
    }
  }
  __shared__ Lin wtot[kRfThreads / 32];
  Lin total;
  const Lin excl = block_excl_scan(loc, wtot, &total);
  if (PHASE == 0) {
    if (threadIdx.x == 0) agg[static_cast<size_t>(series) * nchunk + blockIdx.x] = total;
    return;
  }
  double v = excl.a * carry[static_cast<size_t>(series) * nchunk + blockIdx.x] + excl.b;   // value just before t0
  const double L0 = 1.0 / (1.0 + exp(-prm.k_logistic * (0.0 - prm.c0_logistic)));
  const double Lm = 1.0 / (1.0 + exp(-prm.k_logistic * (prm.c_max - prm.c0_logistic)));
  const double den = (Lm - L0) != 0.0 ? (Lm - L0) : 1e-6;
  Lin ema{1.0, 0.0};
#pragma unroll
  for (int k = 0; k < kRfPerThread; ++k) {
    const int64_t t = t0 + k;
    if (t >= n) break;
    v = t == 0 ? 0.0 : prm.lambda_decay * v + S[k];
    const double cc = fmin(fmax(v, 0.0), prm.c_max);
    double rf = (1.0 / (1.0 + exp(-prm.k_logistic * (cc - prm.c0_logistic))) - L0) / den;
    rf = fmin(fmax(rf, 0.0), 1.0);
    const size_t idx = static_cast<size_t>(series) * n + t;
    rf_inst[idx] = rf;
    if (C_out) C_out[idx] = v;
    // RF_smooth[t] = alpha RF[t] + (1 - alpha) RF_smooth[t-1], RF_smooth[0] = RF[0] (04:276-279): the same kind of
    // first-order recurrence -- its per-chunk aggregate is formed here, while RF[t] is still in registers
    ema = lin_then(ema, t == 0 ? Lin{0.0, rf} : Lin{1.0 - prm.alpha_smooth, prm.alpha_smooth * rf});
  }
  Lin tot_e;
  block_excl_scan(ema, wtot, &tot_e);
  if (threadIdx.x == 0) agg_ema[static_cast<size_t>(series) * nchunk + blockIdx.x] = tot_e;
}
// Chunk carries: value of the recurrence just before each chunk.  One warp per series: lane l composes the maps of its
// contiguous run of chunks (independent loads), a warp scan of the 32 compositions gives each lane its starting value,
// then it walks its run.  (The first version walked all ~500 chunks in one thread: 80 us of dependent load latency.)
__global__ void __launch_bounds__(32)
rf_carry_kernel(const Lin* __restrict__ agg, int nchunk, double* __restrict__ carry) {
  const int series = blockIdx.x, lane = threadIdx.x;
  const int per = (nchunk + 31) / 32;
  const int c0 = lane * per, c1 = c0 + per < nchunk ? c0 + per : nchunk;
  const Lin* A = agg + static_cast<size_t>(series) * nchunk;
  Lin loc{1.0, 0.0};
  for (int c = c0; c < c1; ++c) loc = lin_then(loc, A[c]);
  Lin inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Lin prev{__shfl_up_sync(0xffffffffu, inc.a, o), __shfl_up_sync(0xffffffffu, inc.b, o)};
    if (lane >= o) inc = lin_then(prev, inc);
  }
  Lin excl{__shfl_up_sync(0xffffffffu, inc.a, 1), __shfl_up_sync(0xffffffffu, inc.b, 1)};
  if (lane == 0) excl = Lin{1.0, 0.0};
  double v = excl.b;                       // the recurrence starts from 0: value before chunk c0 = excl applied to 0
  for (int c = c0; c < c1; ++c) {
    carry[static_cast<size_t>(series) * nchunk + c] = v;
    const Lin f = A[c];
    v = f.a * v + f.b;
  }
}
// Third phase: RF_smooth as a scan over RF_inst with the chunk carries of the EMA recurrence, and the first index with
// RF_smooth >= threshold.  (The first version evaluated the closed form over a 194-tap window per element: 194 loads and
// FMAs per row, the longest of the five launches.)
__global__ void __launch_bounds__(kRfThreads)
rf_smooth_kernel(const double* __restrict__ rf_inst, int64_t n, pinn_rf_params_t prm, const double* __restrict__ carry,
                 double* __restrict__ rf_smooth, unsigned long long* __restrict__ first_alarm) {
  const int series = blockIdx.y, nchunk = gridDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * kRfChunk + static_cast<int64_t>(threadIdx.x) * kRfPerThread;
  const double alpha = prm.alpha_smooth, beta = 1.0 - alpha;
  double r[kRfPerThread];
  Lin loc{1.0, 0.0};
#pragma unroll
  for (int k = 0; k < kRfPerThread; ++k) {
    const int64_t t = t0 + k;
    r[k] = 0.0;
    if (t < n) {
      r[k] = rf_inst[static_cast<size_t>(series) * n + t];
      loc = lin_then(loc, t == 0 ? Lin{0.0, r[k]} : Lin{beta, alpha * r[k]});
    }
  }
  __shared__ Lin wtot[kRfThreads / 32];
  Lin total;
  const Lin excl = block_excl_scan(loc, wtot, &total);
  double v = excl.a * carry[static_cast<size_t>(series) * nchunk + blockIdx.x] + excl.b;
  long long first = -1;
#pragma unroll
  for (int k = 0; k < kRfPerThread; ++k) {
    const int64_t t = t0 + k;
    if (t >= n) break;
    v = t == 0 ? r[k] : fma(beta, v, alpha * r[k]);
    rf_smooth[static_cast<size_t>(series) * n + t] = v;
    if (first < 0 && v >= prm.warn_threshold) first = t;
  }
  if (first_alarm && first >= 0) atomicMin(first_alarm + series, static_cast<unsigned long long>(first));
}
__global__ void rf_alarm_init_kernel(unsigned long long* a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = 0xFFFFFFFFFFFFFFFFull;     // reads back as int64 -1: "never reached"
}

}  // namespace pinn

using namespace pinn;

extern "C" int pinn_export_rows(const float* x, const float* y, const float* pred_mean, const float* a_u, const float* e_u,
                                const float* cols, const int64_t* seg_ends, int32_t n_seg, int32_t n_labeled,
                                int32_t window, const pinn_export_scalers_t* sc, int64_t n, double* out, double* rf_cols,
                                void* stream) {
  if (n < 0 || !sc || window < 1 || n_seg < 0 || n_seg > 64) return PINN_E_ARG;
  if (n == 0) return 0;
  if (!x || !y || !pred_mean || !a_u || !e_u || !cols || !out || (n_seg > 0 && !seg_ends)) return PINN_E_ARG;
  const int grid = static_cast<int>((n + 255) / 256);
  export_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, pred_mean, a_u, e_u, cols, seg_ends, n_seg,
                                                                         n_labeled, window, *sc, n, out, rf_cols);
  return static_cast<int>(cudaGetLastError());
}

static int rf_chunks(int64_t n) { return static_cast<int>((n + kRfChunk - 1) / kRfChunk); }

extern "C" size_t pinn_rf_workspace_bytes(int64_t n, int32_t n_series) {
  const size_t nc = static_cast<size_t>(rf_chunks(n > 0 ? n : 1)) * (n_series > 0 ? n_series : 1);
  // [chunk partials: 15 doubles | scan aggregates | carries] + the strengths S_t of every row (phase 0 -> phase 1)
  return nc * (15 * sizeof(double) + 2 * sizeof(Lin) + 2 * sizeof(double)) + 512 +
         static_cast<size_t>(n > 0 ? n : 1) * (n_series > 0 ? n_series : 1) * sizeof(double);
}

static bool rf_layout_ok(int32_t row_cols, int32_t first_col) {      // 16-byte aligned rows and first column, six columns inside the row
  return row_cols >= 6 && first_col >= 0 && first_col + 6 <= row_cols && row_cols % 2 == 0 && first_col % 2 == 0;
}

extern "C" int pinn_rf_stats(const double* results, int64_t n, int32_t n_series, int32_t row_cols, int32_t first_col,
                             double* mu_sigma, void* workspace, size_t workspace_bytes, void* stream) {
  if (n <= 0 || n_series <= 0 || !results || !mu_sigma || !workspace || !rf_layout_ok(row_cols, first_col)) return PINN_E_ARG;
  if (!aligned16(results)) return PINN_E_ALIGN;
  const RfLayout lay{row_cols, first_col};
  if (workspace_bytes < pinn_rf_workspace_bytes(n, n_series)) return PINN_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = rf_chunks(n);
  double* partial = static_cast<double*>(workspace);
  dim3 grid(nc, n_series);
  rf_stats_partial_kernel<<<grid, 256, 0, st>>>(results, n, lay, partial);
  rf_stats_final_kernel<<<n_series, 256, 0, st>>>(results, n, lay, partial, nc, mu_sigma);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int pinn_rf_series(const double* results, int64_t n, int32_t n_series, int32_t row_cols, int32_t first_col,
                              const double* mu_sigma, const pinn_rf_params_t* prm, double* rf_inst, double* rf_smooth,
                              double* c_out, double* s_out, int64_t* first_alarm, void* workspace, size_t workspace_bytes,
                              void* stream) {
  if (n <= 0 || n_series <= 0 || n_series > 1024 || !results || !mu_sigma || !prm || !rf_inst || !rf_smooth || !workspace ||
      !rf_layout_ok(row_cols, first_col))
    return PINN_E_ARG;
  if (!aligned16(results)) return PINN_E_ALIGN;
  const RfLayout lay{row_cols, first_col};
  if (!(prm->alpha_smooth > 0.0 && prm->alpha_smooth <= 1.0)) return PINN_E_ARG;
  if (workspace_bytes < pinn_rf_workspace_bytes(n, n_series)) return PINN_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = rf_chunks(n);
  const size_t ncs = static_cast<size_t>(nc) * n_series;
  char* ws = static_cast<char*>(workspace);
  Lin* agg = reinterpret_cast<Lin*>(ws + ncs * 15 * sizeof(double));
  double* carry = reinterpret_cast<double*>(ws + ncs * (15 * sizeof(double) + sizeof(Lin)));
  Lin* agg_e = reinterpret_cast<Lin*>(ws + ncs * (15 * sizeof(double) + sizeof(Lin) + sizeof(double)));
  double* carry_e = reinterpret_cast<double*>(ws + ncs * (15 * sizeof(double) + 2 * sizeof(Lin) + sizeof(double)));
  // strengths: the caller's S output when wanted, else scratch behind the chunk tables (256-byte aligned)
  size_t s_off = ncs * (15 * sizeof(double) + 2 * sizeof(Lin) + 2 * sizeof(double));
  s_off = (s_off + 255) & ~static_cast<size_t>(255);
  double* s_buf = s_out != nullptr ? s_out : reinterpret_cast<double*>(ws + s_off);
  dim3 grid(nc, n_series);
  rf_scan_kernel<0><<<grid, kRfThreads, 0, st>>>(results, n, lay, mu_sigma, *prm, agg, nullptr, nullptr, nullptr, s_buf, nullptr);
  rf_carry_kernel<<<n_series, 32, 0, st>>>(agg, nc, carry);
  rf_scan_kernel<1><<<grid, kRfThreads, 0, st>>>(results, n, lay, mu_sigma, *prm, agg, carry, rf_inst, c_out, s_buf, agg_e);
  rf_carry_kernel<<<n_series, 32, 0, st>>>(agg_e, nc, carry_e);
  if (first_alarm) rf_alarm_init_kernel<<<(n_series + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned long long*>(first_alarm), n_series);
  rf_smooth_kernel<<<grid, kRfThreads, 0, st>>>(rf_inst, n, *prm, carry_e, rf_smooth, reinterpret_cast<unsigned long long*>(first_alarm));
  return static_cast<int>(cudaGetLastError());
}
