// residuals.cu -- K3: one streaming pass over x_norm[N,8] (+u, +y) that evaluates every
// physics residual of the reference and reduces the losses and analytic lambda-gradients.
//
//   net_f_V 01:724-765 | net_f_T_simple 01:869-914 | net_f_T 01:767-867
//   net_f_H 01:621-722 | net_f_O 01:535-619 | losses 01:1029-1034,1112,1222,1360
//
// The reference crosses PCIe twice per residual call (sklearn inverse_transform on host
// numpy, 01:726-737) and launches ~40 elementwise kernels; here the scaler affine runs
// in-kernel and the row is read once: 32 B x-row + 4 B u (+4 B y) per sample, HBM-bound.
// Loads are 128-bit, grid = a multiple of the SM count, reductions go warp-shuffle ->
// shared -> per-CTA partial (double) -> last-CTA fixed-order sum (deterministic).
#include "common.cuh"
#include <cooperative_groups.h>

namespace pinn {

constexpr int kResThreads = 256;
constexpr int kResCtasPerSm = 3;

template <bool ACC> PINN_D float f_log(float v) { return ACC ? logf(v) : __logf(v); }
template <bool ACC> PINN_D float f_exp(float v) { return ACC ? expf(v) : __expf(v); }
template <bool ACC> PINN_D float f_div(float a, float b) { return ACC ? a / b : __fdividef(a, b); }
template <bool ACC> PINN_D float f_pow(float a, float b) { return ACC ? powf(a, b) : __powf(a, b); }

struct Row { float r[PINN_N_IN]; };

PINN_D Row load_phys(const float* __restrict__ x, int64_t s, const pinn_scalers_t& sc) {
  const float4* p = reinterpret_cast<const float4*>(x + s * PINN_N_IN);
  float4 a = __ldg(p), b = __ldg(p + 1);
  Row o;
  o.r[0] = fmaf(a.x, sc.x_inv_scale[0], -sc.x_off[0]);
  o.r[1] = fmaf(a.y, sc.x_inv_scale[1], -sc.x_off[1]);
  o.r[2] = fmaf(a.z, sc.x_inv_scale[2], -sc.x_off[2]);
  o.r[3] = fmaf(a.w, sc.x_inv_scale[3], -sc.x_off[3]);
  o.r[4] = fmaf(b.x, sc.x_inv_scale[4], -sc.x_off[4]);
  o.r[5] = fmaf(b.y, sc.x_inv_scale[5], -sc.x_off[5]);
  o.r[6] = fmaf(b.z, sc.x_inv_scale[6], -sc.x_off[6]);
  o.r[7] = fmaf(b.w, sc.x_inv_scale[7], -sc.x_off[7]);
  return o;
}

PINN_D float mufu_lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
PINN_D float mufu_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
PINN_D float mufu_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

struct Lam {  // 01:453-517 order
  float l1, l2, l3, l4, T1, T2, T3, T4, T5, H1, H2, H3, H4, O1, O2, O3, O4;
};

// Loop-invariant constants of the fast voltage path: scaler affines folded with the physical
// constants of 01:729-761, and everything that depends only on lambda_1..3 hoisted.
struct VConst {
  float i_a, i_b;        // i   = x0*i_a + i_b            (I/270 + 1e-5)
  float tk_a, tk_b;      // Tk  = x5*tk_a + tk_b          (T_out + 273.15)
  float ph_a, ph_b;      // PH2 = x3*ph_a + ph_b          (P_H/101 + 1)
  float pa_a, pa_b;      // Pair likewise with x4
  float vo_a, vo_b;      // V_out = u*vo_a + vo_b         (stack volts / 5)
  float lg_ph2o;         // log2(P_H2O)
  float inv_l2, inv_l3, l1, l3;
  float ea_a, ea_b;      // (5 V_est) scale_y + min_y = V_est*ea_a + ea_b
  float ga;              // -2 * 5 * scale_y
};
PINN_D VConst make_vconst(const pinn_scalers_t& sc, const Lam& L) {
  VConst c;
  c.i_a = sc.x_inv_scale[0] / 270.0f; c.i_b = 1e-5f - sc.x_off[0] / 270.0f;
  c.tk_a = sc.x_inv_scale[5];         c.tk_b = 273.15f - sc.x_off[5];
  c.ph_a = sc.x_inv_scale[3] / 101.0f; c.ph_b = 1.0f - sc.x_off[3] / 101.0f;
  c.pa_a = sc.x_inv_scale[4] / 101.0f; c.pa_b = 1.0f - sc.x_off[4] / 101.0f;
  c.vo_a = sc.y_inv_scale / 5.0f;     c.vo_b = -sc.y_off / 5.0f;
  c.lg_ph2o = log2f(sc.p_h2o);
  c.inv_l2 = 1.0f / L.l2; c.inv_l3 = 1.0f / L.l3; c.l1 = L.l1; c.l3 = L.l3;
  c.ea_a = 5.0f * sc.scale_y; c.ea_b = sc.min_y; c.ga = -10.0f * sc.scale_y;
  return c;
}
// ~70 instructions and 9 MUFU per sample (the generic path is ~230): same formulas as
// net_f_V with logs taken base 2 and products folded; absolute deviations ~1e-8 V.
PINN_D void eval_V_fast(const VConst& c, float p_h2o, float x0, float x3, float x4, float x5, float us, float ys, bool has_y,
                        bool mode_a, bool mode_b, float* acc, float* __restrict__ cols, int64_t n, int64_t s) {
  constexpr float LN2 = 0.6931471805599453f, LOG2E = 1.4426950408889634f;
  constexpr float CB = 8.314f / 96485.0f;                 // b = R Tk / (2 alpha F), alpha = 0.5
  const float i = fmaf(x0, c.i_a, c.i_b);
  const float Tk = fmaf(x5, c.tk_a, c.tk_b);
  const float PH2 = fmaf(x3, c.ph_a, c.ph_b), Pair = fmaf(x4, c.pa_a, c.pa_b);
  const float z = i * mufu_ex2(-1.334f * mufu_lg2(Tk));    // i / Tk^1.334
  const float ppH2 = 0.5f * fmaf(PH2, mufu_ex2(-1.653f * LOG2E * z), -p_h2o);
  const float ppO2 = fmaf(Pair, mufu_ex2(-4.192f * LOG2E * z), -p_h2o);
  const float b = CB * Tk;
  const float bl = b * LN2;
  const float lg_i = mufu_lg2(i * c.inv_l2);
  const float lg_c = mufu_lg2(fmaf(-i, c.inv_l3, 1.0f));
  const float lg_n = c.lg_ph2o - mufu_lg2(ppH2) - 0.5f * mufu_lg2(ppO2);
  const float Vact = -bl * lg_i;
  const float Vconc = 0.5f * bl * lg_c;
  const float E = fmaf(-0.5f * bl, lg_n, 220170.0f / (2.0f * 96485.0f));
  const float Vohm = -i * c.l1;
  const float Vest = (E + Vact) + (Vohm + Vconc);
  const float Vout = fmaf(us, c.vo_a, c.vo_b);
  const float fV = Vest - Vout;
  const float d2 = b * c.inv_l2;
  const float d3 = 0.5f * b * i * mufu_rcp(c.l3 * (c.l3 - i));
  if (mode_b) {
    const float g = 2.0f * fV;
    acc[PINN_S_FV2] = fmaf(fV, fV, acc[PINN_S_FV2]);
    acc[PINN_S_GB1] = fmaf(-g, i, acc[PINN_S_GB1]);
    acc[PINN_S_GB2] = fmaf(g, d2, acc[PINN_S_GB2]);
    acc[PINN_S_GB3] = fmaf(g, d3, acc[PINN_S_GB3]);
  }
  if (mode_a && has_y) {
    const float eA = ys - fmaf(Vest, c.ea_a, c.ea_b);
    const float g = c.ga * eA;
    acc[PINN_S_EA2] = fmaf(eA, eA, acc[PINN_S_EA2]);
    acc[PINN_S_GA1] = fmaf(-g, i, acc[PINN_S_GA1]);
    acc[PINN_S_GA2] = fmaf(g, d2, acc[PINN_S_GA2]);
    acc[PINN_S_GA3] = fmaf(g, d3, acc[PINN_S_GA3]);
  }
  if (cols) {
    auto put = [&](int col, float v) { cols[static_cast<size_t>(col) * n + s] = v; };
    put(PINN_C_FV, fV); put(PINN_C_VACT, Vact); put(PINN_C_VOHM, Vohm); put(PINN_C_VCONC, Vconc);
    put(PINN_C_ENERNST, E); put(PINN_C_VEST5, Vest * 5.0f); put(PINN_C_I, i); put(PINN_C_VOUT5, Vout * 5.0f);
  }
}

template <uint32_t FAMC, bool ACC>
PINN_D void eval_sample(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ y,
                        int64_t s, int64_t n, const pinn_scalers_t& sc, const Lam& L, uint32_t fam,
                        const float* halo_x, const float* halo_u, float* __restrict__ cols, float* acc) {
  constexpr float A = 270.0f, F = 96485.0f, R = 8.314f, NC = 5.0f, ALPHA = 0.5f;
  const Row row = load_phys(x, s, sc);
  const float* r = row.r;
  acc[PINN_S_N] += 1.0f;
  auto put = [&](int c, float v) { if (cols) cols[static_cast<size_t>(c) * n + s] = v; };
  float us = 0.f, ys = 0.f;
  if ((FAMC & (PINN_FAM_V | PINN_FAM_DATA | PINN_FAM_T)) && (fam & (PINN_FAM_V | PINN_FAM_DATA | PINN_FAM_T)) && u) us = __ldg(u + s);
  if ((FAMC & (PINN_FAM_V | PINN_FAM_DATA)) && y) ys = __ldg(y + s);

  if ((FAMC & PINN_FAM_V) && (fam & PINN_FAM_V)) {
    const float i = f_div<ACC>(r[0], A) + 1e-5f;
    const float Tk = r[5] + 273.15f;
    const float PH2 = f_div<ACC>(r[3], 101.0f) + 1.0f;
    const float Pair = f_div<ACC>(r[4], 101.0f) + 1.0f;
    const float tkp = f_pow<ACC>(Tk, 1.334f);
    const float z = f_div<ACC>(i, tkp);
    float ppH2, ppO2, lnterm;
    if (ACC) {
      ppH2 = 0.5f * (PH2 / expf(1.653f * z) - sc.p_h2o);
      ppO2 = Pair / expf(4.192f * z) - sc.p_h2o;
      lnterm = logf(sc.p_h2o / (ppH2 * sqrtf(ppO2)));
    } else {
      ppH2 = 0.5f * (PH2 * __expf(-1.653f * z) - sc.p_h2o);
      ppO2 = Pair * __expf(-4.192f * z) - sc.p_h2o;
      lnterm = __logf(sc.p_h2o) - __logf(ppH2) - 0.5f * __logf(ppO2);
    }
    const float b = R * Tk / (2.0f * ALPHA * F);
    const float Vact = -b * f_log<ACC>(f_div<ACC>(i, L.l2));
    const float Vohm = -(i * L.l1);
    const float Vconc = ALPHA * b * f_log<ACC>(1.0f - f_div<ACC>(i, L.l3));
    const float E = 220170.0f / (2.0f * F) - (R * Tk) * lnterm / (2.0f * F);
    const float Vest = E + Vact + Vohm + Vconc;
    const float Vout = fmaf(us, sc.y_inv_scale, -sc.y_off) / NC;
    const float fV = Vest - Vout;
    const float d1 = -i, d2 = f_div<ACC>(b, L.l2), d3 = f_div<ACC>(ALPHA * b * i, L.l3 * (L.l3 - i));
    acc[PINN_S_FV2] = fmaf(fV, fV, acc[PINN_S_FV2]);
    const float gB = 2.0f * fV;
    acc[PINN_S_GB1] = fmaf(gB, d1, acc[PINN_S_GB1]);
    acc[PINN_S_GB2] = fmaf(gB, d2, acc[PINN_S_GB2]);
    acc[PINN_S_GB3] = fmaf(gB, d3, acc[PINN_S_GB3]);
    if (y) {
      const float eA = ys - fmaf(Vest * NC, sc.scale_y, sc.min_y);
      const float gA = -2.0f * NC * sc.scale_y * eA;
      acc[PINN_S_EA2] = fmaf(eA, eA, acc[PINN_S_EA2]);
      acc[PINN_S_GA1] = fmaf(gA, d1, acc[PINN_S_GA1]);
      acc[PINN_S_GA2] = fmaf(gA, d2, acc[PINN_S_GA2]);
      acc[PINN_S_GA3] = fmaf(gA, d3, acc[PINN_S_GA3]);
    }
    put(PINN_C_FV, fV); put(PINN_C_VACT, Vact); put(PINN_C_VOHM, Vohm); put(PINN_C_VCONC, Vconc);
    put(PINN_C_ENERNST, E); put(PINN_C_VEST5, Vest * NC); put(PINN_C_I, i); put(PINN_C_VOUT5, Vout * NC);
  }
  if ((FAMC & PINN_FAM_DATA) && (fam & PINN_FAM_DATA) && y && u) {
    const float e = ys - us;
    acc[PINN_S_DATA2] = fmaf(e, e, acc[PINN_S_DATA2]);
  }
  if ((FAMC & PINN_FAM_TS) && (fam & PINN_FAM_TS)) {
    const float i = f_div<ACC>(r[0], A) + 1e-6f;
    const float It = i * A, m = r[1] + 1e-6f;
    const float Tp = L.T1 * It + L.T3 * m + 0.5f * r[2] + L.T5;
    const float fT = r[5] - Tp;
    acc[PINN_S_FT2] = fmaf(fT, fT, acc[PINN_S_FT2]);
    acc[PINN_S_FTABS] += fabsf(fT);
    const float g = -2.0f * fT;
    acc[PINN_S_GT1] = fmaf(g, It, acc[PINN_S_GT1]);
    acc[PINN_S_GT3] = fmaf(g, m, acc[PINN_S_GT3]);
    acc[PINN_S_GT5] += g;
    put(PINN_C_FTS, fT); put(PINN_C_TS_PRED, Tp); put(PINN_C_T_REAL, r[5]);
  }
  if ((FAMC & PINN_FAM_T) && (fam & PINN_FAM_T)) {
    float Tp = r[5];
    const bool first = (s == 0);
    if (!first || (halo_x && halo_u)) {
      Row pr;
      float up;
      if (first) {
        for (int j = 0; j < PINN_N_IN; ++j) pr.r[j] = fmaf(__ldg(halo_x + j), sc.x_inv_scale[j], -sc.x_off[j]);
        up = __ldg(halo_u);
      } else {
        pr = load_phys(x, s - 1, sc);
        up = __ldg(u + s - 1);
      }
      const float ip = f_div<ACC>(pr.r[0], A) + 1e-5f;
      const float It = ip * A, mp = pr.r[1] + 1e-6f;
      const float Vrev = 1.229f - 0.0009f * ((pr.r[5] + 273.15f) - 298.15f);
      const float Vcell = fmaf(up, sc.y_inv_scale, -sc.y_off) / NC;
      const float Qe = (It * Vrev - It * Vcell) * L.T4;
      const float Qc = mp * 4180.0f * (pr.r[5] - pr.r[2]) * L.T1;
      const float Qr = 4.0f * (pr.r[5] - 25.0f) * L.T3;
      Tp = pr.r[5] + ((Qe - Qc - Qr) / L.T2) * 0.1f;
    }
    const float fT = r[5] - Tp;
    acc[PINN_S_FTE2] = fmaf(fT, fT, acc[PINN_S_FTE2]);
    put(PINN_C_FT, fT); put(PINN_C_T_PRED, Tp);
    if (!((FAMC & PINN_FAM_TS) && (fam & PINN_FAM_TS))) put(PINN_C_T_REAL, r[5]);
  }
  if ((FAMC & (PINN_FAM_H | PINN_FAM_O)) && (fam & (PINN_FAM_H | PINN_FAM_O))) {
    const float i = f_div<ACC>(r[0], A) + 1e-5f;
    const float It = i * A;
    if ((FAMC & PINN_FAM_H) && (fam & PINN_FAM_H)) {
      float Q = It / (2.0f * F) * NC * 22.4f * 60.0f;
      Q = fmaxf(Q, 1e-8f);
      const bool lin = It <= L.H3;
      const float sel = lin ? It : L.H3;
      const float tgt = L.H1 + L.H2 * (sel / 100.0f);
      const float act = f_div<ACC>(r[6] + 1e-6f, Q);
      const float fH = act - tgt;
      const float g = -2.0f * fH;
      acc[PINN_S_FH2] = fmaf(fH, fH, acc[PINN_S_FH2]);
      acc[PINN_S_GH1] += g;
      acc[PINN_S_GH2] = fmaf(g, sel / 100.0f, acc[PINN_S_GH2]);
      acc[PINN_S_GH3] += lin ? 0.0f : g * (L.H2 / 100.0f);
      acc[PINN_S_HACT] += act;
      acc[PINN_S_HTGT] += tgt;
      put(PINN_C_FH, fH); put(PINN_C_H_ACT, act); put(PINN_C_H_TGT, tgt); put(PINN_C_I_TOTAL, It);
    }
    if ((FAMC & PINN_FAM_O) && (fam & PINN_FAM_O)) {
      float Q = (It * NC) / (4.0f * F) * 22.4f * 60.0f;
      Q = fmaxf(Q, 1e-8f);
      const float th = fabsf(L.O3);
      const bool lin = It <= th;
      const float sel = lin ? It : th;
      const float raw = L.O1 + L.O2 * (sel / 100.0f);
      const float tgt = fminf(fmaxf(raw, 1.05f), 15.0f);
      const float gate = (raw >= 1.05f && raw <= 15.0f) ? 1.0f : 0.0f;
      const float o2 = (r[7] + 1e-6f) * 0.21f;
      const float act = f_div<ACC>(o2, Q);
      const float fO = act - tgt + fmaxf(1.0f - act, 0.0f) * 10.0f;
      const float g = -2.0f * fO * gate;
      const float sgn = L.O3 > 0.f ? 1.0f : (L.O3 < 0.f ? -1.0f : 0.0f);
      acc[PINN_S_FO2] = fmaf(fO, fO, acc[PINN_S_FO2]);
      acc[PINN_S_GO1] += g;
      acc[PINN_S_GO2] = fmaf(g, sel / 100.0f, acc[PINN_S_GO2]);
      acc[PINN_S_GO3] += lin ? 0.0f : g * (L.O2 * sgn / 100.0f);
      acc[PINN_S_OACT] += act;
      acc[PINN_S_OTGT] += tgt;
      put(PINN_C_FO, fO); put(PINN_C_O_ACT, act); put(PINN_C_O_TGT, tgt); put(PINN_C_O_Q, Q); put(PINN_C_O2, o2);
    }
  }
}

// Block reduce of the per-thread accumulators -> this CTA's double partial; the last CTA to
// arrive sums all partials in a fixed order (deterministic) with the whole block: thread
// (slot = t % 32, lane-group = t / 32) adds every 8th CTA's partial, then the 8 groups are
// folded in order.
// Only the slots [R0, R1) plus PINN_S_N are reduced (the others are written as zeros): the voltage phase owns 10 of
// the 28 slots, and at N = 1M the reduction tail is a visible share of the launch.
template <int R0 = 1, int R1 = PINN_S_COUNT>
PINN_D void finish_sums(const float* acc, double* __restrict__ partials, unsigned int* ticket, double* __restrict__ sums) {
  constexpr int NR = R1 - R0 + 1;                       // slot 0 (N) + the range
  __shared__ double red[kResThreads / 32][NR];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const int k = j == 0 ? PINN_S_N : R0 + j - 1;
    double v = static_cast<double>(acc[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < NR) {
    double v = 0.0;
    for (int wdx = 0; wdx < kResThreads / 32; ++wdx) v += red[wdx][threadIdx.x];
    partials[static_cast<size_t>(blockIdx.x) * NR + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed order: thread (slot, part) sums CTAs part, part + P, ... with eight independent loads in flight, then the P parts
    // are folded in order.  (16 slots x 16 parts when the family owns <= 16 slots: at N = 1M this fold was the longest
    // serial piece of the launch -- 444 partials walked by 8 warps, ~14 dependent L2 round trips.)
    constexpr int SL = NR <= 16 ? 16 : 32, P = kResThreads / SL;
    __shared__ double fold[P][SL];
    const int slot = threadIdx.x % SL, part = threadIdx.x / SL;
    double v = 0.0;
    if (slot < NR) {
      double q[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      unsigned int b = part;
      for (; b + 7 * P < gridDim.x; b += 8 * P) {
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] += partials[static_cast<size_t>(b + k * P) * NR + slot];
      }
      for (int k = 0; b < gridDim.x; b += P, ++k) q[k] += partials[static_cast<size_t>(b) * NR + slot];
      v = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
    }
    fold[part][slot] = v;
    __syncthreads();
    if (threadIdx.x < PINN_S_COUNT) {
      const int k = threadIdx.x;
      const int j = k == PINN_S_N ? 0 : ((k >= R0 && k < R1) ? k - R0 + 1 : -1);
      double t = 0.0;
      if (j >= 0)
        for (int g = 0; g < P; ++g) t += fold[g][j];
      sums[k] = t;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// Training-form kernel of the voltage phase (train_lambda, 01:1008-1055), MUFU math: consumes
// columns 0,3,4,5 of the row plus u and y -- 40 algorithmic bytes per sample; two samples per
// loop trip so four 128-bit loads are in flight per thread.
__global__ void __launch_bounds__(kResThreads, 3)
residual_v_fast_kernel(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ y, int64_t n,
                       pinn_scalers_t sc, const float* __restrict__ lam, uint32_t fam, uint32_t flags,
                       double* __restrict__ partials, unsigned int* ticket, double* __restrict__ sums) {
  Lam L;
  L = Lam{lam[0], lam[1], lam[2], lam[3], lam[4], lam[5], lam[6], lam[7], lam[8], lam[9], lam[10], lam[11], lam[12],
          lam[13], lam[14], lam[15], lam[16]};
  const VConst c = make_vconst(sc, L);
  const bool do_v = (fam & PINN_FAM_V) != 0, do_d = (fam & PINN_FAM_DATA) != 0 && y != nullptr;
  const bool mode_a = !(flags & PINN_RES_NO_MODE_A), mode_b = !(flags & PINN_RES_NO_MODE_B), has_y = y != nullptr;
  float acc[PINN_S_COUNT];
#pragma unroll
  for (int k = 0; k < PINN_S_COUNT; ++k) acc[k] = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  auto one = [&](float4 a, float4 b, float us, float ys, int64_t idx) {
    acc[PINN_S_N] += 1.0f;
    if (do_v) eval_V_fast(c, sc.p_h2o, a.x, a.w, b.x, b.y, us, ys, has_y, mode_a, mode_b, acc, nullptr, n, idx);
    if (do_d) { const float e = ys - us; acc[PINN_S_DATA2] = fmaf(e, e, acc[PINN_S_DATA2]); }
  };
  // two samples per trip, the NEXT trip's loads issued before this trip's math (each thread only makes a few trips at
  // N = 1M: without the prefetch every trip pays a full DRAM round trip)
  struct Pair { float4 a0, b0, a1, b1; float u0, u1, y0, y1; };
  auto load2 = [&](int64_t q) {
    Pair p;
    const float4* p0 = reinterpret_cast<const float4*>(x + q * PINN_N_IN);
    const float4* p1 = reinterpret_cast<const float4*>(x + (q + stride) * PINN_N_IN);
    p.a0 = __ldg(p0); p.b0 = __ldg(p0 + 1); p.a1 = __ldg(p1); p.b1 = __ldg(p1 + 1);
    p.u0 = __ldg(u + q); p.u1 = __ldg(u + q + stride);
    p.y0 = has_y ? __ldg(y + q) : 0.f; p.y1 = has_y ? __ldg(y + q + stride) : 0.f;
    return p;
  };
  if (s + stride < n) {
    Pair cur = load2(s);
    for (;;) {
      const int64_t nx = s + 2 * stride;
      const bool more = nx + stride < n;
      Pair nxt = cur;
      if (more) nxt = load2(nx);
      one(cur.a0, cur.b0, cur.u0, cur.y0, s);
      one(cur.a1, cur.b1, cur.u1, cur.y1, s + stride);
      s = nx;
      if (!more) break;
      cur = nxt;
    }
  }
  if (s < n) {
    const float4* p0 = reinterpret_cast<const float4*>(x + s * PINN_N_IN);
    one(__ldg(p0), __ldg(p0 + 1), __ldg(u + s), has_y ? __ldg(y + s) : 0.f, s);
  }
  finish_sums<PINN_S_FV2, PINN_S_GB3 + 1>(acc, partials, ticket, sums);
}

// Workspace: double partials[grid][PINN_S_COUNT] followed by one uint32 ticket (zeroed
// once by the caller; the last CTA resets it).
template <uint32_t FAMC, bool ACC>
__global__ void __launch_bounds__(kResThreads, kResCtasPerSm)
residual_kernel(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ y, int64_t n,
                pinn_scalers_t sc, const float* __restrict__ lam, uint32_t fam, const float* halo_x,
                const float* halo_u, float* __restrict__ cols, double* __restrict__ partials,
                unsigned int* ticket, double* __restrict__ sums) {
  Lam L;
  {
    const float* lp = lam;
    L = Lam{lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], lp[6], lp[7], lp[8], lp[9], lp[10], lp[11], lp[12],
            lp[13], lp[14], lp[15], lp[16]};
  }
  float acc[PINN_S_COUNT];
#pragma unroll
  for (int k = 0; k < PINN_S_COUNT; ++k) acc[k] = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < n; s += stride)
    eval_sample<FAMC, ACC>(x, u, y, s, n, sc, L, fam, halo_x, halo_u, cols, acc);

  finish_sums(acc, partials, ticket, sums);
}

static int res_grid(int64_t n, int ctas_per_sm = kResCtasPerSm, int samples_per_thread = 1) {
  int64_t want = (n + static_cast<int64_t>(kResThreads) * samples_per_thread - 1) / (static_cast<int64_t>(kResThreads) * samples_per_thread);
  int64_t cap = static_cast<int64_t>(sm_count()) * ctas_per_sm;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}


// ------------------------------------------------------------------------------------------------
// Persistent scalar-phase trainer: a whole block of optimiser steps of train_lambda / train_thermal /
// train_hydrogen / train_oxygen (01:1008-1055, 1107-1151, 1354-1391, 1204-1274) in ONE cooperative
// launch.  The per-step form (pinn_residuals + pinn_adam_step_from_sums) is bound by launch latency
// at the reference's own data sizes (N ~ 2e4: 25 us per step for ~2 us of work).  Here every CTA keeps
// the 17 scalars in shared memory and runs, per step:
//   residual sums of its samples -> per-CTA double partial (double-buffered by step parity)
//   -> ONE grid barrier -> every CTA sums all partials in the same fixed order (identical totals
//   everywhere, so no broadcast is needed) -> Adam + StepLR + clamp on its private copy.
// x / u / y never change during a phase, so after the first step they are served from L1 / L2.
// Two partial buffers make one barrier per step enough: a CTA can be at most one step ahead of the
// slowest one, so the buffer it writes is never the one a straggler still reads.
constexpr int kPhaseMaxThreads = 1024;
constexpr int kPhaseMaxParams = 8;
constexpr int kPhaseFold = 16;

struct PhaseArgs {
  const float* x; const float* u; const float* y;
  int64_t n;
  pinn_scalers_t sc;
  float* lam;                    // all 17 scalars (read once, slice written back at the end)
  uint32_t fam, flags;
  int first, count;              // the optimiser's slice lam[first .. first+count)
  int slot[kPhaseMaxParams];     // PINN_S_* gradient slot per scalar, < 0: no gradient (clamp only)
  float lo[kPhaseMaxParams], hi[kPhaseMaxParams];
  float* m; float* v;            // Adam moments of the slice
  int64_t* step_counter;
  AdamHyper h;
  int64_t n_steps;
  double* partials;              // [2][grid][R]
  unsigned int* barrier;         // zeroed by the host before the launch
  double* sums;                  // PINN_S_COUNT totals of the LAST step (its loss is the one reported)
};

PINN_D unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Block size 512 (n <= 512 x SMs: one sample per thread, few CTAs -> cheap barrier) or 1024 (one CTA per SM,
// grid-stride).  What is on the critical path of a step besides the samples themselves: one global store +
// fence + atomic per CTA, the barrier, one round of L2 loads for the fold, and the optimiser arithmetic --
// so the bias corrections 1 - beta^t are carried as running products (one multiply per step; `pow` only at
// launch start and at StepLR boundaries; three double-precision `pow` calls cost more than everything else
// in the step) and every thread folds at most four partials with all loads in flight at once.
// CLUSTER = true: the whole batch is one thread-block cluster (<= 8 CTAs, n <= 8 192): partials stay in each CTA's shared
// memory, the grid barrier becomes the hardware cluster barrier and the fold reads its peers through distributed shared
// memory -- no global-memory round trip is left on the step (N = 5 000: 3.3-4.6 us per step).
template <uint32_t FAMC, int R0, int R1, bool CLUSTER>
__global__ void __launch_bounds__(kPhaseMaxThreads, 1) scalar_phase_kernel(const PhaseArgs a) {
  constexpr int R = R1 - R0;
  static_assert(R <= kPhaseFold, "fold layout holds 16 slots");
  __shared__ double red[kPhaseMaxThreads / 32][kPhaseFold];
  __shared__ double tot[kPhaseFold];
  __shared__ double part_s[2][kPhaseFold];     // CLUSTER: this CTA's partial of the current / previous step
  __shared__ float adam_c[2][kPhaseMaxThreads];   // per-step optimiser constants of the next <= 1024 steps: {lr / (1 - b1^t), sqrt(1 - b2^t)}
  __shared__ float lam_s[PINN_N_LAMBDA];
  __shared__ int slot_s[kPhaseMaxParams];
  __shared__ float lo_s[kPhaseMaxParams], hi_s[kPhaseMaxParams];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (tid < PINN_N_LAMBDA) lam_s[tid] = a.lam[tid];
  if (tid < kPhaseMaxParams) { slot_s[tid] = a.slot[tid]; lo_s[tid] = a.lo[tid]; hi_s[tid] = a.hi[tid]; }
  float mm = 0.f, vv = 0.f;
  if (tid < a.count) { mm = a.m[tid]; vv = a.v[tid]; }
  int64_t t0 = *a.step_counter;
  const bool has_y = a.y != nullptr;
  const bool do_v = (a.fam & PINN_FAM_V) != 0, do_d = (a.fam & PINN_FAM_DATA) != 0 && has_y;
  const bool mode_a = !(a.flags & PINN_RES_NO_MODE_A), mode_b = !(a.flags & PINN_RES_NO_MODE_B);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t s0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + tid;
  const double cnt = a.n > 0 ? static_cast<double>(a.n) : 1.0;
  const int fk = tid & (kPhaseFold - 1), fg = tid >> 4, nfg = blockDim.x >> 4;     // fold: slot fk, CTAs fg, fg + nfg, ...
  __syncthreads();

  for (int64_t step = 0; step < a.n_steps; ++step) {
    if ((step & (kPhaseMaxThreads - 1)) == 0) {
      // table of the optimiser's per-step constants, one step per thread (the same double-precision expressions as
      // adam_update: three `pow` per step cost more than the rest of a step, here they are off the critical path)
      for (int i = tid; i < kPhaseMaxThreads && step + i < a.n_steps; i += blockDim.x) {
        const int64_t t = t0 + i;       // steps taken before that step
        const double lr = a.h.lr0 * pow(a.h.gamma, static_cast<double>(t / a.h.step_size));
        adam_consts(lr, t + 1, adam_c[0][i], adam_c[1][i]);
      }
      __syncthreads();
    }
    const Lam L{lam_s[0], lam_s[1], lam_s[2], lam_s[3], lam_s[4], lam_s[5], lam_s[6], lam_s[7], lam_s[8],
                lam_s[9], lam_s[10], lam_s[11], lam_s[12], lam_s[13], lam_s[14], lam_s[15], lam_s[16]};
    float acc[PINN_S_COUNT];
#pragma unroll
    for (int k = 0; k < PINN_S_COUNT; ++k) acc[k] = 0.f;
    if constexpr ((FAMC & PINN_FAM_V) != 0) {
      const VConst c = make_vconst(a.sc, L);
      for (int64_t s = s0; s < a.n; s += stride) {
        const float4* p0 = reinterpret_cast<const float4*>(a.x + s * PINN_N_IN);
        const float4 r0 = __ldg(p0), r1 = __ldg(p0 + 1);
        const float us = __ldg(a.u + s), ys = has_y ? __ldg(a.y + s) : 0.f;
        if (do_v) eval_V_fast(c, a.sc.p_h2o, r0.x, r0.w, r1.x, r1.y, us, ys, has_y, mode_a, mode_b, acc, nullptr, a.n, s);
        if (do_d) { const float e = ys - us; acc[PINN_S_DATA2] = fmaf(e, e, acc[PINN_S_DATA2]); }
      }
    } else {
      for (int64_t s = s0; s < a.n; s += stride)
        eval_sample<FAMC, false>(a.x, a.u, a.y, s, a.n, a.sc, L, a.fam, nullptr, nullptr, nullptr, acc);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      double v = static_cast<double>(acc[R0 + k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if constexpr (CLUSTER) {
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      const int par = static_cast<int>(step & 1);
      if (tid < R) {
        double v = 0.0;
        for (int w = 0; w < nwarp; ++w) v += red[w][tid];
        part_s[par][tid] = v;
      }
      cluster.sync();        // release / acquire over the cluster: every CTA's partial of this step is visible
      if (tid < kPhaseFold) {
        double t = 0.0;
        if (tid < R) {
          const unsigned int nr = cluster.num_blocks();
          for (unsigned int r = 0; r < nr; ++r) t += cluster.map_shared_rank(&part_s[par][0], r)[tid];     // rank order: identical in every CTA
        }
        tot[tid] = t;
      }
      __syncthreads();
    } else {
    const size_t buf = static_cast<size_t>(step & 1) * gridDim.x * R;
    if (tid < R) {
      double v = 0.0;
      for (int w = 0; w < nwarp; ++w) v += red[w][tid];
      a.partials[buf + static_cast<size_t>(blockIdx.x) * R + tid] = v;
    }
    // grid barrier: arrivals are counted monotonically, step k completes at (k+1) * gridDim.x
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(a.barrier, 1u);
      const unsigned int target = static_cast<unsigned int>(step + 1) * gridDim.x;
      while (ld_acquire_gpu_u32(a.barrier) < target) {}
      __threadfence();
    }
    __syncthreads();
    // fold: identical order in every CTA.  Thread (fk, fg) adds CTAs fg, fg + nfg, ... (<= 4 per batch, loads issued
    // together), the two groups of a warp meet by shuffle, warp sub-sums go through shared memory, warp 0 finishes.
    {
      double v = 0.0;
      if (fk < R) {
        const double* base = a.partials + buf + fk;
        for (unsigned int b = fg; b < gridDim.x; b += 4 * nfg) {
          const unsigned int b1 = b + nfg, b2 = b + 2 * nfg, b3 = b + 3 * nfg;
          const double v0 = __ldcg(base + static_cast<size_t>(b) * R);
          const double v1 = b1 < gridDim.x ? __ldcg(base + static_cast<size_t>(b1) * R) : 0.0;
          const double v2 = b2 < gridDim.x ? __ldcg(base + static_cast<size_t>(b2) * R) : 0.0;
          const double v3 = b3 < gridDim.x ? __ldcg(base + static_cast<size_t>(b3) * R) : 0.0;
          v += v0; v += v1; v += v2; v += v3;
        }
      }
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < kPhaseFold) red[warp][lane] = v;
    }
    __syncthreads();
    if (warp == 0) {
      double t = 0.0;
      const int half = lane >> 4, k = lane & 15, per = (nwarp + 1) >> 1;
      for (int w = half * per; w < (half + 1) * per && w < nwarp; ++w) t += red[w][k];
      t += __shfl_xor_sync(0xffffffffu, t, 16);
      if (lane < kPhaseFold) tot[lane] = t;
    }
    __syncthreads();
    }
    if (tid < a.count) {
      const int sl = slot_s[tid];
      float p = lam_s[a.first + tid];
      if (sl >= 0) {
        const float g = static_cast<float>(tot[sl - R0] / cnt);
        const int ci = static_cast<int>(step & (kPhaseMaxThreads - 1));
        adam_apply(p, g, mm, vv, adam_c[0][ci], adam_c[1][ci], lo_s[tid], hi_s[tid], true);
      } else {
        p = fminf(fmaxf(p, lo_s[tid]), hi_s[tid]);   // the reference clamps every listed scalar each step
      }
      lam_s[a.first + tid] = p;
    }
    ++t0;
    __syncthreads();
  }
  if constexpr (CLUSTER) cooperative_groups::this_cluster().sync();     // nobody leaves while a peer may still read its partial
  if (blockIdx.x == 0) {
    if (tid < a.count) { a.lam[a.first + tid] = lam_s[a.first + tid]; a.m[tid] = mm; a.v[tid] = vv; }
    if (tid < PINN_S_COUNT) {
      double v = 0.0;
      if (tid == PINN_S_N) v = static_cast<double>(a.n);
      else if (tid >= R0 && tid < R1 && a.n_steps > 0) v = tot[tid - R0];
      a.sums[tid] = v;
    }
    if (tid == 0) *a.step_counter = t0;
  }
}

template <uint32_t FAMC, int R0, int R1>
static int launch_phase(PhaseArgs& a, size_t workspace_bytes, void* workspace, cudaStream_t st) {
  for (int i = 0; i < a.count; ++i)
    if (a.slot[i] >= 0 && (a.slot[i] < R0 || a.slot[i] >= R1)) return PINN_E_ARG;
  const int sms = sm_count();
  // ---- one cluster (<= 8 CTAs x 1024 threads): batches of up to one sample per thread (with more work per thread the
  // 40-CTA grid form wins again: N = 20 000, voltage phase 5.6 us per step as a cluster vs 4.9 us as a grid)
  if (!(a.flags & PINN_RES_NO_CLUSTER) && a.n <= static_cast<int64_t>(8) * kPhaseMaxThreads) {
    const int threads = a.n <= 4096 ? 512 : kPhaseMaxThreads;
    int csize = 1;
    while (csize < 8 && static_cast<int64_t>(csize) * threads < a.n) csize *= 2;
    auto kern = scalar_phase_kernel<FAMC, R0, R1, true>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(csize); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters >= 1)
      return static_cast<int>(cudaLaunchKernelEx(&cfg, kern, static_cast<const PhaseArgs>(a)));
    (void)cudaGetLastError();      // no room for the cluster: the grid-barrier form below
  }
  // ---- cooperative grid with a global-memory barrier
  const int threads = a.n <= static_cast<int64_t>(512) * sms ? 512 : kPhaseMaxThreads;
  auto kern = scalar_phase_kernel<FAMC, R0, R1, false>;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (occ < 1) return PINN_E_ARG;
  const int64_t want = (a.n + threads - 1) / threads;
  const int grid = static_cast<int>(want < sms ? (want > 0 ? want : 1) : sms);      // never more than one CTA per SM
  if (static_cast<uint64_t>(a.n_steps) * static_cast<uint64_t>(grid) >= 0x7fffffffull) return PINN_E_ARG;
  const size_t part_bytes = static_cast<size_t>(2) * sms * kPhaseFold * sizeof(double);
  if (workspace_bytes < part_bytes + 16) return PINN_E_WORKSPACE;
  a.partials = static_cast<double*>(workspace);
  a.barrier = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + part_bytes);
  e = cudaMemsetAsync(a.barrier, 0, 16, st);
  if (e != cudaSuccess) return static_cast<int>(e);
  void* params[] = {&a};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(threads), params, 0, st);
  return static_cast<int>(e);
}

}  // namespace pinn

using namespace pinn;

extern "C" size_t pinn_residuals_workspace_bytes(int64_t n) {
  (void)n;
  size_t cap = static_cast<size_t>(sm_count()) * 4;
  return cap * PINN_S_COUNT * sizeof(double) + 16;
}

extern "C" int pinn_residuals(const float* x, const float* u, const float* y, int64_t n,
                              const pinn_scalers_t* scalers, const float* lambdas, uint32_t families,
                              uint32_t flags, const float* halo_x, const float* halo_u, float* cols,
                              double* sums, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || !scalers || !lambdas || !sums || !workspace) return PINN_E_ARG;
  if (n > 0 && !x) return PINN_E_ARG;
  if ((families & (PINN_FAM_V | PINN_FAM_T | PINN_FAM_DATA)) && n > 0 && !u) return PINN_E_ARG;
  if ((families & PINN_FAM_DATA) && n > 0 && !y) return PINN_E_ARG;
  if (workspace_bytes < pinn_residuals_workspace_bytes(n)) return PINN_E_WORKSPACE;
  if (!aligned16(x) || !aligned16(workspace)) return PINN_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = res_grid(n);
  double* partials = static_cast<double*>(workspace);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(
      static_cast<char*>(workspace) + static_cast<size_t>(sm_count()) * 4 * PINN_S_COUNT * sizeof(double));
  const bool accm = (flags & PINN_RES_ACCURATE_MATH) != 0;
  const uint32_t f = families;
#define LAUNCH(FAMC)                                                                                          \
  do {                                                                                                        \
    if (accm)                                                                                                 \
      residual_kernel<FAMC, true><<<grid, kResThreads, 0, st>>>(x, u, y, n, *scalers, lambdas, f, halo_x,   \
                                                                halo_u, cols, partials, ticket, sums);       \
    else                                                                                                      \
      residual_kernel<FAMC, false><<<grid, kResThreads, 0, st>>>(x, u, y, n, *scalers, lambdas, f, halo_x,  \
                                                                 halo_u, cols, partials, ticket, sums);      \
  } while (0)
  constexpr uint32_t VD = PINN_FAM_V | PINN_FAM_DATA;
  if ((f & ~VD) == 0 && !accm && cols == nullptr && (n == 0 || u != nullptr)) {
    const int g4 = res_grid(n, 3, 2);
    residual_v_fast_kernel<<<g4, kResThreads, 0, st>>>(x, u, y, n, *scalers, lambdas, f, flags, partials, ticket, sums);
    return static_cast<int>(cudaGetLastError());
  }
  constexpr uint32_t ALL = PINN_FAM_V | PINN_FAM_TS | PINN_FAM_T | PINN_FAM_H | PINN_FAM_O | PINN_FAM_DATA;
  if ((f & ~VD) == 0) LAUNCH(VD);
  else if (f == PINN_FAM_TS) LAUNCH(PINN_FAM_TS);
  else if (f == PINN_FAM_H) LAUNCH(PINN_FAM_H);
  else if (f == PINN_FAM_O) LAUNCH(PINN_FAM_O);
  else LAUNCH(ALL);
#undef LAUNCH
  return static_cast<int>(cudaGetLastError());
}

extern "C" size_t pinn_scalar_phase_workspace_bytes(void) {
  return static_cast<size_t>(2) * sm_count() * kPhaseFold * sizeof(double) + 16;
}

extern "C" int pinn_scalar_phase(const float* x, const float* u, const float* y, int64_t n,
                                 const pinn_scalers_t* scalers, float* lambdas, uint32_t families, uint32_t flags,
                                 int32_t first, int32_t count, const int32_t* grad_slot, const float* lo,
                                 const float* hi, float* exp_avg, float* exp_avg_sq, int64_t* step_counter,
                                 double lr0, double gamma, int64_t step_size, int64_t n_steps, double* sums,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (n <= 0 || !x || !scalers || !lambdas || !grad_slot || !lo || !hi || !exp_avg || !exp_avg_sq || !step_counter ||
      !sums || !workspace || step_size <= 0 || n_steps < 0)
    return PINN_E_ARG;
  if (first < 0 || count < 1 || count > kPhaseMaxParams || first + count > PINN_N_LAMBDA) return PINN_E_ARG;
  if (!aligned16(x) || !aligned16(workspace)) return PINN_E_ALIGN;
  PhaseArgs a{};
  a.x = x; a.u = u; a.y = y; a.n = n; a.sc = *scalers; a.lam = lambdas; a.fam = families; a.flags = flags;
  a.first = first; a.count = count;
  for (int i = 0; i < kPhaseMaxParams; ++i) {
    a.slot[i] = i < count ? grad_slot[i] : -1;
    a.lo[i] = i < count ? lo[i] : 0.f;
    a.hi[i] = i < count ? hi[i] : 0.f;
  }
  a.m = exp_avg; a.v = exp_avg_sq; a.step_counter = step_counter;
  a.h = AdamHyper{lr0, gamma, 1.0, step_size};
  a.n_steps = n_steps; a.sums = sums;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr uint32_t VD = PINN_FAM_V | PINN_FAM_DATA;
  if (families != 0 && (families & ~VD) == 0) {
    if (!u) return PINN_E_ARG;
    if ((families & PINN_FAM_DATA) && !y) return PINN_E_ARG;
    return launch_phase<VD, PINN_S_FV2, PINN_S_GB3 + 1>(a, workspace_bytes, workspace, st);
  }
  if (families == PINN_FAM_TS) return launch_phase<PINN_FAM_TS, PINN_S_FT2, PINN_S_GT5 + 1>(a, workspace_bytes, workspace, st);
  if (families == PINN_FAM_H) return launch_phase<PINN_FAM_H, PINN_S_FH2, PINN_S_HTGT + 1>(a, workspace_bytes, workspace, st);
  if (families == PINN_FAM_O) return launch_phase<PINN_FAM_O, PINN_S_FO2, PINN_S_OTGT + 1>(a, workspace_bytes, workspace, st);
  return PINN_E_ARG;
}

