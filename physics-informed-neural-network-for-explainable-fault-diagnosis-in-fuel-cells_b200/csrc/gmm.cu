// gmm.cu -- f4: the Gaussian-mixture posterior pass behind `fit_gmm_and_get_probabilities`
// (03_unsupervised_gmm_fault_diagnosis 03:360-426): full-covariance GaussianMixture over a handful of
// per-sample features (the reference uses pV, pT, pH, pO: d = 4, 20 components), float64 like sklearn.
//
// One pass over X[n][d] evaluates, per row, what sklearn's `_estimate_log_prob_resp` does
//   y = (x - mu_c) @ precisions_cholesky_c ;  log p_c = -0.5 (d log 2pi + |y|^2) + sum log diag(chol_c) + log w_c
//   log_prob_norm = logsumexp_c ;  resp_c = exp(log p_c - log_prob_norm)
// and then whatever the caller asked for:
//   * EM sufficient statistics per component (sum resp, sum resp (x - mu_c), sum resp (x - mu_c)(x - mu_c)^T: the M-step
//     of `GaussianMixture.fit`, finished on the host from C x (1 + d + d(d+1)/2) numbers),
//   * the label calibration sums  W[c][k] = sum_i resp[i][c] [y_i == k]      (03:394-412),
//   * y_prob = clip(resp @ P, 1e-12, 1) row-normalised and its argmax          (03:415-423),
//   * resp itself.
// Layout of the work: a warp takes 32 rows.  Stage 1 (lane = row) computes the responsibilities into a per-warp
// shared-memory tile; stage 2 (lane = component) walks the 32 rows and accumulates ITS component's statistics in
// registers -- no shuffles, no atomics, fixed order; stage 3 (lane = row) maps responsibilities to class
// probabilities.  The pass is bound by fp64 arithmetic (~C (d^2 + 2d + 40) DFMA-class instructions per row against
// 8 d bytes), not by HBM.
#include "common.cuh"

namespace pinn {

constexpr int kGmmThreads = 512;     // 16 warps per SM: the pass is a chain of fp64 latencies, occupancy is what hides them
constexpr int kGmmWarps = kGmmThreads / 32;
constexpr int kGmmMaxC = 32;      // components (a lane each in stage 2)
constexpr int kGmmMaxK = 16;      // fault classes
constexpr int kGmmMaxD = 8;

PINN_HD constexpr int gmm_nstat(int d) { return 1 + d + d * (d + 1) / 2; }

struct GmmArgs {
  const double* X; int64_t n;
  int C, K;
  const double* weights; const double* means; const double* prec_chol;
  const int32_t* labels;
  const double* comp_class_prob;
  double* resp; double* y_prob; int32_t* y_pred;
  double* partials;        // [grid][C * NS + C * K + 1]
  int want_stats, want_cal;
};

// shared-memory map (offsets in doubles), sized for the actual component / class counts so that 16 warps fit
struct GmmSmem {
  int rs;                 // doubles per row of a warp's responsibility tile: C rounded up to odd (no bank conflicts)
  int kMeans, kChol, kConst, kP, kResp, kX, kCal, kAcc, kLab, total;
};
PINN_HD GmmSmem gmm_smem(int D, int C, int K, int nwarps) {
  GmmSmem m;
  const int NS = gmm_nstat(D), Kp = K > 0 ? K : 1;
  m.rs = C | 1;
  m.kMeans = 0;                                   // [C][D]
  m.kChol = m.kMeans + C * D;                     // [C][D][D]
  m.kConst = m.kChol + C * D * D;                 // [C]  log w + log det - 0.5 d log 2pi
  m.kP = m.kConst + C;                            // [C][K]
  m.kResp = m.kP + C * Kp;                        // [warps][32][rs]
  m.kX = m.kResp + nwarps * 32 * m.rs;         // [warps][32][D]
  m.kCal = m.kX + nwarps * 32 * D;             // [warps][C][K]
  m.kAcc = m.kCal + nwarps * C * Kp;           // [C][NS] + [C][K] + 1 : CTA accumulators
  m.kLab = (m.kAcc + C * NS + C * Kp + 2) & ~1;   // int32 [warps][32]
  m.total = m.kLab + nwarps * 32 / 2;
  return m;
}

template <int D>
__global__ void __launch_bounds__(kGmmThreads, 1) gmm_pass_kernel(const GmmArgs a) {
  constexpr int NS = gmm_nstat(D);
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = a.C, K = a.K, Kp = K > 0 ? K : 1;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;       // 512 threads, or 256 when the tiles of 16 warps do not fit
  const GmmSmem S = gmm_smem(D, C, K, nwarps);
  const int RS = S.rs;
  // ---- parameters -> shared memory
  for (int i = tid; i < C * D; i += nthreads) sm[S.kMeans + i] = a.means[i];
  for (int i = tid; i < C * D * D; i += nthreads) sm[S.kChol + i] = a.prec_chol[i];
  for (int c = tid; c < C; c += nthreads) {
    double ld = 0.0;
    for (int j = 0; j < D; ++j) ld += log(a.prec_chol[(static_cast<size_t>(c) * D + j) * D + j]);
    sm[S.kConst + c] = log(a.weights[c]) + ld - 0.5 * D * 1.8378770664093453;     // log(2 pi)
  }
  if (a.comp_class_prob != nullptr)
    for (int i = tid; i < C * K; i += nthreads) sm[S.kP + i] = a.comp_class_prob[i];
  for (int i = tid; i < nwarps * C * Kp; i += nthreads) sm[S.kCal + i] = 0.0;
  __syncthreads();

  double* resp_w = sm + S.kResp + warp * 32 * RS;
  double* x_w = sm + S.kX + warp * 32 * D;
  double* cal_w = sm + S.kCal + warp * C * Kp;
  int32_t* lab_w = reinterpret_cast<int32_t*>(sm + S.kLab) + warp * 32;

  // stage-2 state of lane c: its component's mean and running statistics
  double mu[D], st[NS];
#pragma unroll
  for (int i = 0; i < D; ++i) mu[i] = lane < C ? sm[S.kMeans + lane * D + i] : 0.0;
#pragma unroll
  for (int i = 0; i < NS; ++i) st[i] = 0.0;
  double lpn = 0.0;

  const int64_t n_batches = (a.n + 31) / 32;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * nwarps + warp; b < n_batches; b += static_cast<int64_t>(gridDim.x) * nwarps) {
    const int64_t row = b * 32 + lane;
    const bool valid = row < a.n;
    // ---------------------------------------------------------------- stage 1: lane = row
    double x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = valid ? __ldg(a.X + row * D + i) : 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) x_w[lane * D + i] = x[i];
    if (a.labels != nullptr) lab_w[lane] = valid ? __ldg(a.labels + row) : -1;
    double mx = -1.0e300;
    for (int c = 0; c < C; ++c) {
      const double* m = sm + S.kMeans + c * D;
      const double* P = sm + S.kChol + c * D * D;
      double dx[D];
#pragma unroll
      for (int i = 0; i < D; ++i) dx[i] = x[i] - m[i];
      double q = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double y = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) y = fma(dx[i], P[i * D + j], y);
        q = fma(y, y, q);
      }
      const double lp = sm[S.kConst + c] - 0.5 * q;
      resp_w[lane * RS + c] = lp;
      mx = fmax(mx, lp);
    }
    double se = 0.0;
    for (int c = 0; c < C; ++c) {
      const double e = exp(resp_w[lane * RS + c] - mx);
      resp_w[lane * RS + c] = e;
      se += e;
    }
    const double inv = 1.0 / se;
    for (int c = 0; c < C; ++c) resp_w[lane * RS + c] *= inv;
    if (valid) lpn += mx + log(se);
    __syncwarp();
    // ---------------------------------------------------------------- stage 2: lane = component
    const int rows_here = a.n - b * 32 < 32 ? static_cast<int>(a.n - b * 32) : 32;
    if (lane < C) {
      if (a.want_stats || a.want_cal || a.resp != nullptr) {
        for (int r = 0; r < rows_here; ++r) {
          const double w = resp_w[r * RS + lane];
          if (a.resp != nullptr) a.resp[(b * 32 + r) * C + lane] = w;
          if (a.want_cal) {
            const int lab = lab_w[r];
            if (lab >= 0 && lab < K) cal_w[lane * Kp + lab] += w;
          }
          if (a.want_stats) {
            double dx[D];
#pragma unroll
            for (int i = 0; i < D; ++i) dx[i] = x_w[r * D + i] - mu[i];
            st[0] += w;
            int s = 1 + D;
#pragma unroll
            for (int i = 0; i < D; ++i) {
              const double wd = w * dx[i];
              st[1 + i] += wd;
#pragma unroll
              for (int j = i; j < D; ++j) { st[s] = fma(wd, dx[j], st[s]); ++s; }
            }
          }
        }
      }
    }
    // ---------------------------------------------------------------- stage 3: lane = row
    if (a.comp_class_prob != nullptr && valid) {
      double yk[kGmmMaxK];
#pragma unroll
      for (int k = 0; k < kGmmMaxK; ++k) yk[k] = 0.0;
      for (int c = 0; c < C; ++c) {
        const double w = resp_w[lane * RS + c];
        const double* Pc = sm + S.kP + c * Kp;
#pragma unroll
        for (int k = 0; k < kGmmMaxK; ++k)
          if (k < K) yk[k] = fma(w, Pc[k], yk[k]);
      }
      double tot = 0.0;
#pragma unroll
      for (int k = 0; k < kGmmMaxK; ++k)
        if (k < K) { yk[k] = fmin(fmax(yk[k], 1e-12), 1.0); tot += yk[k]; }
      int best = 0;
      double bv = -1.0;
#pragma unroll
      for (int k = 0; k < kGmmMaxK; ++k)
        if (k < K) {
          const double v = yk[k] / tot;
          if (a.y_prob != nullptr) a.y_prob[row * K + k] = v;
          if (v > bv) { bv = v; best = k; }
        }
      if (a.y_pred != nullptr) a.y_pred[row] = best;
    }
    __syncwarp();
  }

  // ---- CTA reduction, fixed order: warps add their registers / tiles into the CTA accumulators one after another
  double* acc = sm + S.kAcc;
  const int n_acc = C * NS + C * K + 1;
  for (int i = tid; i < n_acc; i += nthreads) acc[i] = 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lpn += __shfl_xor_sync(0xffffffffu, lpn, o);
  __syncthreads();
  for (int w = 0; w < nwarps; ++w) {
    if (warp == w) {
      if (lane < C) {
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[lane * NS + s] += st[s];
        for (int k = 0; k < K; ++k) acc[C * NS + lane * K + k] += cal_w[lane * Kp + k];
      }
      if (lane == 0) acc[C * NS + C * K] += lpn;
    }
    __syncthreads();
  }
  double* part = a.partials + static_cast<size_t>(blockIdx.x) * n_acc;
  for (int i = tid; i < n_acc; i += nthreads) part[i] = acc[i];
}

// CTA partials -> totals, fixed order: a CTA owns 32 entries, its 8 warps add every 8th partial, folded in warp order.
__global__ void __launch_bounds__(256) gmm_reduce_kernel(const double* __restrict__ partials, int nblk, int n_acc, int C, int NS, int K,
                                                        double* __restrict__ stats, double* __restrict__ cal, double* __restrict__ lpn) {
  __shared__ double fold[8][32];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;
  double v = 0.0;
  if (i < n_acc)
    for (int b = g; b < nblk; b += 8) v += partials[static_cast<size_t>(b) * n_acc + i];
  fold[g][c] = v;
  __syncthreads();
  if (g == 0 && i < n_acc) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += fold[k][c];
    if (i < C * NS) { if (stats != nullptr) stats[i] = t; }
    else if (i < C * NS + C * K) { if (cal != nullptr) cal[i - C * NS] = t; }
    else if (lpn != nullptr) *lpn = t;
  }
}

static int gmm_grid(int64_t n, int nwarps) {
  const int64_t want = (n + 32 * nwarps - 1) / (32 * nwarps);
  const int64_t cap = sm_count();
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

template <int D>
static int launch_gmm(GmmArgs& a, double* stats, double* cal, double* lpn, cudaStream_t st) {
  int threads = kGmmThreads;
  size_t smem = static_cast<size_t>(gmm_smem(D, a.C, a.K, threads / 32).total) * sizeof(double);
  if (smem > 227 * 1024) {      // the largest d / component / class counts: eight warps per CTA
    threads = 256;
    smem = static_cast<size_t>(gmm_smem(D, a.C, a.K, threads / 32).total) * sizeof(double);
  }
  if (smem > 227 * 1024) return PINN_E_SHAPE;
  PINN_CUDA_TRY(cudaFuncSetAttribute(gmm_pass_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = gmm_grid(a.n, threads / 32);
  gmm_pass_kernel<D><<<grid, threads, smem, st>>>(a);
  PINN_CUDA_TRY(cudaGetLastError());
  const int NS = gmm_nstat(D), n_acc = a.C * NS + a.C * a.K + 1;
  gmm_reduce_kernel<<<(n_acc + 31) / 32, 256, 0, st>>>(a.partials, grid, n_acc, a.C, NS, a.K, stats, cal, lpn);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pinn

using namespace pinn;

extern "C" size_t pinn_gmm_workspace_bytes(int32_t d, int32_t n_components, int32_t n_classes) {
  if (d < 1 || d > kGmmMaxD || n_components < 1 || n_components > kGmmMaxC || n_classes < 0 || n_classes > kGmmMaxK) return 0;
  const size_t n_acc = static_cast<size_t>(n_components) * gmm_nstat(d) + static_cast<size_t>(n_components) * n_classes + 1;
  return static_cast<size_t>(sm_count()) * n_acc * sizeof(double) + 16;
}

extern "C" int pinn_gmm_pass(const double* X, int64_t n, int32_t d, int32_t n_components, const double* weights, const double* means,
                             const double* prec_chol, const int32_t* labels, int32_t n_classes, const double* comp_class_prob,
                             double* resp, double* y_prob, int32_t* y_pred, double* stats, double* comp_class_weight,
                             double* log_prob_norm_sum, void* workspace, size_t workspace_bytes, void* stream) {
  if (n <= 0 || !X || !weights || !means || !prec_chol || !workspace) return PINN_E_ARG;
  if (d < 1 || d > kGmmMaxD || n_components < 1 || n_components > kGmmMaxC || n_classes < 0 || n_classes > kGmmMaxK) return PINN_E_SHAPE;
  if ((comp_class_weight != nullptr) != (labels != nullptr)) return PINN_E_ARG;
  if ((labels != nullptr || comp_class_prob != nullptr) && n_classes < 1) return PINN_E_ARG;
  if ((y_prob != nullptr || y_pred != nullptr) && comp_class_prob == nullptr) return PINN_E_ARG;
  if (workspace_bytes < pinn_gmm_workspace_bytes(d, n_components, n_classes)) return PINN_E_WORKSPACE;
  GmmArgs a{};
  a.X = X; a.n = n; a.C = n_components; a.K = n_classes;
  a.weights = weights; a.means = means; a.prec_chol = prec_chol;
  a.labels = labels; a.comp_class_prob = comp_class_prob;
  a.resp = resp; a.y_prob = y_prob; a.y_pred = y_pred;
  a.partials = static_cast<double*>(workspace);
  a.want_stats = stats != nullptr; a.want_cal = labels != nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d) {
    case 1: return launch_gmm<1>(a, stats, comp_class_weight, log_prob_norm_sum, st);
    case 2: return launch_gmm<2>(a, stats, comp_class_weight, log_prob_norm_sum, st);
    case 3: return launch_gmm<3>(a, stats, comp_class_weight, log_prob_norm_sum, st);
    case 4: return launch_gmm<4>(a, stats, comp_class_weight, log_prob_norm_sum, st);
    case 5: return launch_gmm<5>(a, stats, comp_class_weight, log_prob_norm_sum, st);
    case 6: return launch_gmm<6>(a, stats, comp_class_weight, log_prob_norm_sum, st);
    case 7: return launch_gmm<7>(a, stats, comp_class_weight, log_prob_norm_sum, st);
    default: return launch_gmm<8>(a, stats, comp_class_weight, log_prob_norm_sum, st);
  }
}
