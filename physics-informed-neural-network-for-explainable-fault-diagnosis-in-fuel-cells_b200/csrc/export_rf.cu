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

__global__ void __launch_bounds__(256)
export_rows_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ pm,
                   const float* __restrict__ au, const float* __restrict__ eu, const float* __restrict__ cols,
                   const int64_t* __restrict__ seg_ends, int n_seg, int n_labeled, int window, pinn_export_scalers_t sc, int64_t n,
                   double* __restrict__ out) {
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
}

// ------------------------------------------------------------------------------- RF(t)
PINN_HD constexpr int rf_col(int d) { return 12 + d; }   // res, pV, pT, pH, pO = columns 12..16 (04:58-62,80)
constexpr int kRfChunk = 2048, kRfThreads = 256, kRfPerThread = kRfChunk / kRfThreads;

// pass = 0: partial sums and counts over label-0, non-NaN rows; pass = 1: squared deviations.
__global__ void __launch_bounds__(256)
rf_stats_partial_kernel(const double* __restrict__ res, int64_t n, int pass, const double* __restrict__ mean,
                        double* __restrict__ partial /* [series][chunks][10] */) {
  const int series = blockIdx.y, nchunk = gridDim.x;
  const double* R = res + static_cast<size_t>(series) * n * kRowCols;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * kRfChunk;
  double s[5] = {0, 0, 0, 0, 0}, cnt[5] = {0, 0, 0, 0, 0};
  for (int64_t t = c0 + threadIdx.x; t < c0 + kRfChunk && t < n; t += blockDim.x) {
    const double* r = R + t * kRowCols;
    if (static_cast<int>(r[17]) != 0) continue;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const double v = r[rf_col(d)];
      if (v == v) {
        const double q = pass == 0 ? v : (v - mean[series * 10 + d]) * (v - mean[series * 10 + d]);
        s[d] += q; cnt[d] += 1.0;
      }
    }
  }
  __shared__ double red[8][10];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    double a = s[d], b = cnt[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { red[warp][d] = a; red[warp][5 + d] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double a = 0.0;
    for (int wq = 0; wq < 8; ++wq) a += red[wq][threadIdx.x];
    partial[(static_cast<size_t>(series) * nchunk + blockIdx.x) * 10 + threadIdx.x] = a;
  }
}
// pass = 0 -> mu_sigma[series][0..5) = mean;  pass = 1 -> [5..10) = sigma (ddof 1, 0 -> 1e-6)
__global__ void rf_stats_final_kernel(const double* __restrict__ partial, int nchunk, int pass, double* __restrict__ mu_sigma) {
  const int series = blockIdx.x, d = threadIdx.x;
  if (d >= 5) return;
  double s = 0.0, c = 0.0;
  for (int k = 0; k < nchunk; ++k) {
    s += partial[(static_cast<size_t>(series) * nchunk + k) * 10 + d];
    c += partial[(static_cast<size_t>(series) * nchunk + k) * 10 + 5 + d];
  }
  if (pass == 0) {
    mu_sigma[series * 10 + d] = s / c;
  } else {
    double sg = sqrt(s / (c - 1.0));
    if (sg == 0.0) sg = 1e-6;
    mu_sigma[series * 10 + 5 + d] = sg;
  }
}

PINN_D double rf_strength(const double* __restrict__ r, const double* __restrict__ ms, double z_safe) {
  double a[5];
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    const double z = fabs((r[rf_col(d)] - ms[d]) / ms[5 + d]);
    a[d] = z != z ? z : fmax(0.0, z - z_safe);       // np.maximum propagates NaN (04:238), fmax would drop it
  }
  // layers {res,pV}, {pH,pO}, {pT} with p = 2, unit weights (04:84-96)
  return sqrt(a[0] * a[0] + a[1] * a[1]) + sqrt(a[3] * a[3] + a[4] * a[4]) + sqrt(a[2] * a[2]);
}

// First-order recurrence v_t = a_t v_{t-1} + b_t as an associative scan on pairs (a, b):
// (a1,b1) then (a2,b2)  ==  (a1 a2, a2 b1 + b2).
struct Lin { double a, b; };
PINN_D Lin lin_then(Lin f, Lin g) { return Lin{f.a * g.a, g.a * f.b + g.b}; }

// phase 0: per-chunk aggregate; phase 1: apply the carry and emit C, RF_inst (and S).
template <int PHASE>
__global__ void __launch_bounds__(kRfThreads)
rf_scan_kernel(const double* __restrict__ res, int64_t n, const double* __restrict__ mu_sigma, pinn_rf_params_t prm,
               Lin* __restrict__ agg /* [series][chunks] */, const double* __restrict__ carry /* [series][chunks] */,
               double* __restrict__ rf_inst, double* __restrict__ C_out, double* __restrict__ S_out) {
  const int series = blockIdx.y, nchunk = gridDim.x;
  const double* R = res + static_cast<size_t>(series) * n * kRowCols;
  const double* ms = mu_sigma + series * 10;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * kRfChunk + static_cast<int64_t>(threadIdx.x) * kRfPerThread;
  double S[kRfPerThread];
  Lin loc{1.0, 0.0};
#pragma unroll
  for (int k = 0; k < kRfPerThread; ++k) {
    const int64_t t = t0 + k;
    S[k] = 0.0;
    if (t < n) {
      S[k] = rf_strength(R + t * kRowCols, ms, prm.z_safe);
      // C[0] = 0 regardless of S[0] (04:262-264): element 0 is the constant map v -> 0
      loc = lin_then(loc, t == 0 ? Lin{0.0, 0.0} : Lin{prm.lambda_decay, S[k]});
    }
  }
  // exclusive scan of the per-thread maps across the block (warp shuffles, then warp totals)
  __shared__ Lin wtot[kRfThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Lin inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Lin prev{__shfl_up_sync(0xffffffffu, inc.a, o), __shfl_up_sync(0xffffffffu, inc.b, o)};
    if (lane >= o) inc = lin_then(prev, inc);
  }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  Lin before{1.0, 0.0};
  for (int wq = 0; wq < warp; ++wq) before = lin_then(before, wtot[wq]);
  Lin excl{__shfl_up_sync(0xffffffffu, inc.a, 1), __shfl_up_sync(0xffffffffu, inc.b, 1)};
  if (lane == 0) excl = Lin{1.0, 0.0};
  excl = lin_then(before, excl);
  if (PHASE == 0) {
    if (threadIdx.x == kRfThreads - 1) agg[static_cast<size_t>(series) * nchunk + blockIdx.x] = lin_then(excl, loc);
    return;
  }
  double v = excl.a * carry[static_cast<size_t>(series) * nchunk + blockIdx.x] + excl.b;   // value just before t0
  const double L0 = 1.0 / (1.0 + exp(-prm.k_logistic * (0.0 - prm.c0_logistic)));
  const double Lm = 1.0 / (1.0 + exp(-prm.k_logistic * (prm.c_max - prm.c0_logistic)));
  const double den = (Lm - L0) != 0.0 ? (Lm - L0) : 1e-6;
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
    if (S_out) S_out[idx] = S[k];
  }
}
__global__ void rf_carry_kernel(const Lin* __restrict__ agg, int nchunk, double* __restrict__ carry) {
  const int series = threadIdx.x;
  double v = 0.0;
  for (int c = 0; c < nchunk; ++c) {
    carry[static_cast<size_t>(series) * nchunk + c] = v;
    const Lin f = agg[static_cast<size_t>(series) * nchunk + c];
    v = f.a * v + f.b;
  }
}
// RF_smooth[t] = alpha RF[t] + (1-alpha) RF_smooth[t-1], RF_smooth[0] = RF[0] (04:276-279):
// closed form over a 192-tap window ((1-alpha)^192 ~ 2.5e-19 for alpha = 0.2, below fp64 eps);
// the tap count is derived from alpha.  Also the first index with RF_smooth >= threshold.
__global__ void __launch_bounds__(256)
rf_smooth_kernel(const double* __restrict__ rf_inst, int64_t n, pinn_rf_params_t prm, int taps, double* __restrict__ rf_smooth,
                 unsigned long long* __restrict__ first_alarm) {
  const int series = blockIdx.y;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* r = rf_inst + static_cast<size_t>(series) * n;
  const double alpha = prm.alpha_smooth, beta = 1.0 - alpha;
  double acc = 0.0, w = alpha;
  const int64_t kmax = t < taps ? t : taps;
  for (int64_t k = 0; k < kmax; ++k) { acc += w * r[t - k]; w *= beta; }
  if (t < taps) acc += (w / alpha) * r[0];      // beta^t * RF[0]
  rf_smooth[static_cast<size_t>(series) * n + t] = acc;
  if (first_alarm && acc >= prm.warn_threshold) atomicMin(first_alarm + series, static_cast<unsigned long long>(t));
}
__global__ void rf_alarm_init_kernel(unsigned long long* a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = 0xFFFFFFFFFFFFFFFFull;     // reads back as int64 -1: "never reached"
}

}  // namespace pinn

using namespace pinn;

extern "C" int pinn_export_rows(const float* x, const float* y, const float* pred_mean, const float* a_u, const float* e_u,
                                const float* cols, const int64_t* seg_ends, int32_t n_seg, int32_t n_labeled,
                                int32_t window, const pinn_export_scalers_t* sc, int64_t n, double* out, void* stream) {
  if (n < 0 || !sc || window < 1 || n_seg < 0 || n_seg > 64) return PINN_E_ARG;
  if (n == 0) return 0;
  if (!x || !y || !pred_mean || !a_u || !e_u || !cols || !out || (n_seg > 0 && !seg_ends)) return PINN_E_ARG;
  const int grid = static_cast<int>((n + 255) / 256);
  export_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, pred_mean, a_u, e_u, cols, seg_ends, n_seg,
                                                                         n_labeled, window, *sc, n, out);
  return static_cast<int>(cudaGetLastError());
}

static int rf_chunks(int64_t n) { return static_cast<int>((n + kRfChunk - 1) / kRfChunk); }

extern "C" size_t pinn_rf_workspace_bytes(int64_t n, int32_t n_series) {
  const size_t nc = static_cast<size_t>(rf_chunks(n > 0 ? n : 1)) * (n_series > 0 ? n_series : 1);
  return nc * (10 * sizeof(double) + sizeof(Lin) + sizeof(double)) + 256;
}

extern "C" int pinn_rf_stats(const double* results, int64_t n, int32_t n_series, double* mu_sigma, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (n <= 0 || n_series <= 0 || !results || !mu_sigma || !workspace) return PINN_E_ARG;
  if (workspace_bytes < pinn_rf_workspace_bytes(n, n_series)) return PINN_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = rf_chunks(n);
  double* partial = static_cast<double*>(workspace);
  dim3 grid(nc, n_series);
  for (int pass = 0; pass < 2; ++pass) {
    rf_stats_partial_kernel<<<grid, 256, 0, st>>>(results, n, pass, mu_sigma, partial);
    rf_stats_final_kernel<<<n_series, 32, 0, st>>>(partial, nc, pass, mu_sigma);
  }
  return static_cast<int>(cudaGetLastError());
}

extern "C" int pinn_rf_series(const double* results, int64_t n, int32_t n_series, const double* mu_sigma,
                              const pinn_rf_params_t* prm, double* rf_inst, double* rf_smooth, double* c_out,
                              double* s_out, int64_t* first_alarm, void* workspace, size_t workspace_bytes, void* stream) {
  if (n <= 0 || n_series <= 0 || n_series > 1024 || !results || !mu_sigma || !prm || !rf_inst || !rf_smooth || !workspace)
    return PINN_E_ARG;
  if (!(prm->alpha_smooth > 0.0 && prm->alpha_smooth <= 1.0)) return PINN_E_ARG;
  if (workspace_bytes < pinn_rf_workspace_bytes(n, n_series)) return PINN_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = rf_chunks(n);
  const size_t ncs = static_cast<size_t>(nc) * n_series;
  char* ws = static_cast<char*>(workspace);
  Lin* agg = reinterpret_cast<Lin*>(ws + ncs * 10 * sizeof(double));
  double* carry = reinterpret_cast<double*>(ws + ncs * (10 * sizeof(double) + sizeof(Lin)));
  dim3 grid(nc, n_series);
  rf_scan_kernel<0><<<grid, kRfThreads, 0, st>>>(results, n, mu_sigma, *prm, agg, nullptr, nullptr, nullptr, nullptr);
  rf_carry_kernel<<<1, n_series, 0, st>>>(agg, nc, carry);
  rf_scan_kernel<1><<<grid, kRfThreads, 0, st>>>(results, n, mu_sigma, *prm, agg, carry, rf_inst, c_out, s_out);
  if (first_alarm) rf_alarm_init_kernel<<<(n_series + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned long long*>(first_alarm), n_series);
  int taps = 1;
  if (prm->alpha_smooth < 1.0) taps = static_cast<int>(ceil(-43.0 / log(1.0 - prm->alpha_smooth))) + 1;   // beta^taps < 2e-19
  if (taps > 1 << 20) taps = 1 << 20;
  dim3 g2(static_cast<unsigned>((n + 255) / 256), n_series);
  rf_smooth_kernel<<<g2, 256, 0, st>>>(rf_inst, n, *prm, taps, rf_smooth, reinterpret_cast<unsigned long long*>(first_alarm));
  return static_cast<int>(cudaGetLastError());
}
