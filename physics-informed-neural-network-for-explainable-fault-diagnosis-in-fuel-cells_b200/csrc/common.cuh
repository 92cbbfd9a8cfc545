// common.cuh -- shared device/host helpers for libb200pinn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/b200pinn.h"

#define PINN_HD __host__ __device__ __forceinline__
#define PINN_D __device__ __forceinline__

#define PINN_CUDA_TRY(expr)                          \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return static_cast<int>(_e); \
  } while (0)

namespace pinn {

// ------------------------------------------------------------------ Philox4x32-7
// Counter-based RNG (Salmon et al. 2011).  counter = (sample_lo, sample_hi, pass,
// layer<<16 | unit/4), key = seed: a mask bit depends only on global indices, so
// results are identical for any grid, GPU count or sharding.
#ifndef PINN_PHILOX_ROUNDS
#define PINN_PHILOX_ROUNDS 7
#endif
struct Philox {
  // Philox4x32-R.  R = 7 is the fewest rounds that pass BigCrush ("Crush-resistant", Salmon et al. 2011, table 2); the
  // customary 10 adds safety margin a dropout mask does not need, and the generator is a quarter of the MC kernel's
  // instructions.  Every kernel draws through this one definition, so the mask stream is identical on all paths.
  static constexpr int kRounds = PINN_PHILOX_ROUNDS;
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

  static PINN_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = static_cast<uint64_t>(a) * b;
    lo = static_cast<uint32_t>(p);
    hi = static_cast<uint32_t>(p >> 32);
#endif
  }

  static PINN_HD uint4 gen(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                           uint32_t c3) {
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(M0, c0, hi0, lo0);
      mulhilo(M1, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
  }

  // Same generator with the ten round keys precomputed on the host (rk[2r] = k0 + r W0,
  // rk[2r+1] = k1 + r W1).  `rk` lives in the kernel-parameter constant bank and is indexed
  // with compile-time constants, so every key is a c[0x0][..] operand of the LOP3 that consumes
  // it: no key-schedule adds and no key registers in the hot loop.
  static PINN_HD uint4 gen_rk(const uint32_t (&rk)[20], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#if defined(PINN_ABL) && (PINN_ABL & 2)     // ablation build (profiles/ablate_mc.py): no Philox rounds
    return make_uint4(c0 * M0 + c3, c1 ^ (c3 * M1), c2 + c3 * W0, c3 * W1 ^ rk[0]);
#endif
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(M0, c0, hi0, lo0);
      mulhilo(M1, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// Dropout draw context for one sample and one pass.
struct DropCtx {
  uint32_t k0, k1;       // key
  uint32_t s_lo, s_hi;   // global sample index
  uint32_t pass;         // global pass / step index
  uint32_t thresh;       // drop iff 16-bit draw < thresh  (thresh = round(p * 2^16))
  float scale;           // 1/(1-p), fp32 like torch
  float keep;            // (1-p) in fp32
  const uint8_t* mrow;   // injected keep bits of this (pass, sample): D bytes, or nullptr
  bool active;           // p > 0
};

// Draws are 16 bits wide (8 per Philox call): a unit is dropped iff its 16-bit field is
// below thresh = round(p * 2^16), i.e. the realised drop rate is p rounded to 1/65536
// (|error| <= 7.7e-6, far below the 1/sqrt(T) Monte-Carlo noise); the scale stays 1/(1-p).
PINN_HD uint32_t drop_threshold(float p) {
  double t = static_cast<double>(p) * 65536.0 + 0.5;
  if (t <= 0.0) return 0u;
  if (t >= 65535.0) return 65535u;     // p < 1 always; keeps thresh << 16 representable
  return static_cast<uint32_t>(t);
}
PINN_HD float drop_scale(float p) {
  float keep = static_cast<float>(1.0 - static_cast<double>(p));  // torch: double 1-p, cast to fp32
  return 1.0f / keep;
}

// Keep-multipliers ({0, scale}) of 8 consecutive units [j0, j0+8), j0 % 8 == 0, of dropout
// layer `layer`; `unit_base` is the byte offset of that layer inside an injected mask row.
PINN_HD void drop8(const DropCtx& c, uint32_t layer, uint32_t j0, uint32_t unit_base, float m[8]) {
  if (c.mrow != nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q) m[q] = c.mrow[unit_base + j0 + q] ? c.scale : 0.0f;
    return;
  }
  const uint4 r = Philox::gen(c.k0, c.k1, c.s_lo, c.s_hi, c.pass, (layer << 16) | (j0 >> 3));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    m[2 * q] = (w[q] & 0xFFFFu) < c.thresh ? 0.0f : c.scale;
    m[2 * q + 1] = (w[q] >> 16) < c.thresh ? 0.0f : c.scale;
  }
}

// Host-built, kernel-parameter form of pinn_dropout_t.
struct DropParams {
  float p;
  uint32_t thresh;
  float scale, keep;
  uint32_t k0, k1;
  int64_t sample_offset, pass_offset, mask_n;
  const uint8_t* masks;
  uint32_t rk[20];       // Philox round keys (Philox::gen_rk)
  uint32_t thresh_hi;    // thresh << 16: "high 16-bit field >= thresh" is one unsigned compare
};
inline DropParams make_drop_params(const pinn_dropout_t* d) {
  DropParams q{};
  q.scale = 1.f;
  q.keep = 1.f;
  if (d && d->p > 0.f) {
    q.p = d->p;
    q.thresh = drop_threshold(d->p);
    q.scale = drop_scale(d->p);
    q.keep = static_cast<float>(1.0 - static_cast<double>(d->p));
    q.k0 = static_cast<uint32_t>(d->seed);
    q.k1 = static_cast<uint32_t>(d->seed >> 32);
    q.sample_offset = d->sample_offset;
    q.pass_offset = d->pass_offset;
    q.mask_n = d->mask_sample_stride_n;
    q.masks = d->masks;
  }
  for (int r = 0; r < 10; ++r) {
    q.rk[2 * r] = q.k0 + static_cast<uint32_t>(r) * Philox::W0;
    q.rk[2 * r + 1] = q.k1 + static_cast<uint32_t>(r) * Philox::W1;
  }
  q.thresh_hi = q.thresh << 16;
  return q;
}
// Context of (shard-local sample s_local, shard-local pass pass_local); D = mask row bytes.
PINN_HD DropCtx make_ctx(const DropParams& dp, int64_t s_local, int64_t pass_local, int D, bool active,
                         bool has_mask_row = true) {
  DropCtx c;
  c.k0 = dp.k0; c.k1 = dp.k1;
  uint64_t sg = static_cast<uint64_t>(dp.sample_offset + s_local);
  c.s_lo = static_cast<uint32_t>(sg); c.s_hi = static_cast<uint32_t>(sg >> 32);
  c.pass = static_cast<uint32_t>(dp.pass_offset + pass_local);
  c.thresh = dp.thresh; c.scale = dp.scale; c.keep = dp.keep;
  c.active = active && dp.p > 0.f;
  c.mrow = nullptr;
  if (c.active && dp.masks) {
    if (has_mask_row) c.mrow = dp.masks + (static_cast<size_t>(pass_local) * dp.mask_n + s_local) * D;
    else c.active = false;  // tail slot of a tile: no mask row exists, contributes nothing
  }
  return c;
}

// Keep decisions of 8 consecutive units [j0, j0+8) of dropout layer `layer` for one (sample, pass):
// Philox mode draws the same 16-bit fields as drop8() (common.cuh), compared without extracting
// them: the high field is kept iff  w >= thresh<<16,  the low field iff  (w<<16) >= thresh<<16.
// keep decisions of the eight 16-bit fields of one Philox block (field 2q = low half of word q)
PINN_HD void keep8_from(const uint4& r, uint32_t th, bool (&k)[8]) {
  k[0] = (r.x << 16) >= th; k[1] = r.x >= th;
  k[2] = (r.y << 16) >= th; k[3] = r.y >= th;
  k[4] = (r.z << 16) >= th; k[5] = r.z >= th;
  k[6] = (r.w << 16) >= th; k[7] = r.w >= th;
}
template <bool INJ>
struct KeepSrc {
  uint32_t s_lo, s_hi, pass;
  const uint8_t* mrow;     // INJ: keep bytes of this (pass, sample)
  PINN_D void get8(const DropParams& dp, uint32_t layer, uint32_t j0, uint32_t unit_base, bool (&k)[8]) const {
    if constexpr (INJ) {
#pragma unroll
      for (int q = 0; q < 8; ++q) k[q] = mrow[unit_base + j0 + q] != 0;
    } else {
      const uint4 r = Philox::gen_rk(dp.rk, s_lo, s_hi, pass, (layer << 16) | (j0 >> 3));
      keep8_from(r, dp.thresh_hi, k);
    }
  }
};

// ------------------------------------------------------------------------ math
// tanh(x) = sign(x) (1 - e)/(1 + e), e = exp(-2|x|) from MUFU.EX2 + MUFU.RCP: 8 instructions,
// branch-free.  Its ABSOLUTE error is <= ~3e-7 everywhere (the relative error grows as
// |x| -> 0, where the value itself vanishes); activations enter O(1)-weighted dot products,
// so absolute error is what the 1e-5 parity bar sees.  libdevice tanhf costs 20 issue slots
// per call and was a quarter of the MC kernel's instructions (profiles/r1_*).
PINN_HD float tanh_act(float x) {
#ifdef __CUDA_ARCH__
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.885390082f * fabsf(x)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));     // 1+e in [1,2]: no range fix-ups needed
  return copysignf((1.0f - e) * r, x);
#else
  return tanhf(x);
#endif
}

// The tensor-core kernels use the sign-free form  tanh(x) = 1 - 2 / (1 + 2^(c x)),  c = 2 log2(e):
// with biases (and the layer-0 weights) pre-scaled by c the argument is one FFMA (or comes out
// of the dot product directly) and the whole activation is FFMA, EX2, FADD, RCP, FFMA.  Same
// absolute error (<= ~3e-7); saturates cleanly (2^a -> inf gives rcp -> 0 -> 1; 2^a -> 0 gives -1).
constexpr float kTanhArg = 2.8853900817779268f;
PINN_D float tanh_pre(float a) {       // a = kTanhArg * x
#if defined(PINN_ABL) && (PINN_ABL & 1)     // ablation build: no MUFU
  return a * 0.25f;
#endif
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(-2.0f, r, 1.0f);
}

// Two activations per instruction where the ISA allows it: sm_100 has packed fp32x2 FFMA2 / FADD2 (same FLOP rate,
// half the issue slots -- and issue slots, not FLOPs, bound the tensor-core kernels' epilogues).
// PAIR: ONE reciprocal for the two activations, 1/d0 = d1 / (d0 d1): 3 MUFU operations per pair instead of 4 (the XU pipe
// runs 16 lanes per clock and SM and is the busiest pipe of the 64-wide MC kernel's epilogue) for three more FMA-pipe
// instructions.  The arguments are clamped at 2^40 (tanh is 1.0f to the last bit from 2^25 on) so that the product of the
// two denominators cannot overflow; the extra product and multiply leave the absolute error within the same 3e-7.
// Measured (profiles/r2_wide_res_ab4.log): 64-wide sweep T = 1000 x N = 1M 201.5 -> 187.4 ms; no gain for the fused
// training kernel or the 256-wide kernel (neither is bound by the XU pipe), which keep the two-reciprocal form; four
// activations on one reciprocal (5 MUFU per four, five more multiplies) came out 1.5 % SLOWER than the pair form
// (profiles/r2_ab_k4_quad.log).
template <bool PAIR = false>
PINN_D float2 tanh_pre2(float2 a) {     // a = kTanhArg * x, two lanes
  if constexpr (PAIR) {
    float e0, e1, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(a.x, 40.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(a.y, 40.0f)));
    const float2 dd = __fadd2_rn(make_float2(e0, e1), make_float2(1.0f, 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd.x * dd.y));
    const float2 rr = __fmul2_rn(make_float2(r, r), make_float2(dd.y, dd.x));
    return __ffma2_rn(rr, make_float2(-2.0f, -2.0f), make_float2(1.0f, 1.0f));
  }
  float ex, ey, rx, ry;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(a.y));
  const float2 d = __fadd2_rn(make_float2(ex, ey), make_float2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(d.y));
  return __ffma2_rn(make_float2(rx, ry), make_float2(-2.0f, -2.0f), make_float2(1.0f, 1.0f));
}
// t[q] = tanh(z[q] + b[q]) for eight units, biases pre-scaled by kTanhArg
template <bool PAIR = false>
PINN_D void tanh8_prescaled(const float* z, const float (&bs)[8], float (&t)[8]) {
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    const float2 a2 = __ffma2_rn(make_float2(z[q], z[q + 1]), make_float2(kTanhArg, kTanhArg), make_float2(bs[q], bs[q + 1]));
    const float2 t2 = tanh_pre2<PAIR>(a2);
    t[q] = t2.x; t[q + 1] = t2.y;
  }
}

// log(softplus(v) + 1e-6), softplus with torch's threshold 20 (01:432-434).
PINN_HD float softplus_f(float v) { return v > 20.0f ? v : log1pf(expf(v)); }
PINN_HD float logvar_from_v(float v) { return logf(softplus_f(v) + 1e-6f); }
// DNN(logvar=False) (01:436, PINN_NET_NO_LOGVAR): the log-variance output is identically zero and nothing flows
// back into the variance head
PINN_HD float logvar_out(float v, bool no_logvar) { return no_logvar ? 0.0f : logvar_from_v(v); }
// d logvar / d v  
PINN_HD float dlogvar_dv(float v) {
  float sp = softplus_f(v);
  float sg = v > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-v));
  return sg / (sp + 1e-6f);
}

// ------------------------------------------------------------- parameter layout
// Canonical flat order == dnn.parameters() order (01:399-419).
struct ParamLayout {
  int H, L;
  int64_t offW[PINN_MAX_HIDDEN], offb[PINN_MAX_HIDDEN];
  int64_t offWp, offbp, offWv0, offbv0, offWv1, offbv1, offWv2, offbv2, total;
};
PINN_HD ParamLayout make_layout(int H, int L) {
  ParamLayout p;
  p.H = H; p.L = L;
  int64_t o = 0;
  for (int l = 0; l < PINN_MAX_HIDDEN; ++l) { p.offW[l] = 0; p.offb[l] = 0; }
  // every tensor starts on a 16-byte boundary (128-bit weight loads); the two
  // single-element biases are padded to 4 floats.
  auto pad4 = [](int64_t v) { return (v + 3) & ~static_cast<int64_t>(3); };
  for (int l = 0; l < L; ++l) {
    int in = l == 0 ? PINN_N_IN : H;
    p.offW[l] = o; o = pad4(o + static_cast<int64_t>(H) * in);
    p.offb[l] = o; o = pad4(o + H);
  }
  p.offWp = o; o = pad4(o + H);
  p.offbp = o; o = pad4(o + 1);
  p.offWv0 = o; o = pad4(o + static_cast<int64_t>(H / 2) * H);
  p.offbv0 = o; o = pad4(o + H / 2);
  p.offWv1 = o; o = pad4(o + static_cast<int64_t>(H / 4) * (H / 2));
  p.offbv1 = o; o = pad4(o + H / 4);
  p.offWv2 = o; o = pad4(o + H / 4);
  p.offbv2 = o; o = pad4(o + 1);
  p.total = o;
  return p;
}

// Entry i of the flat bucket is alignment padding (behind the two single-element biases bp / bv2, or behind any tensor
// whose size is not a multiple of four): the gradient reduces write 0 there instead of summing uninitialised partials.
PINN_HD bool layout_is_padding(const ParamLayout& p, int64_t i) {
  auto in = [&](int64_t off, int64_t cnt) { return i >= off && i < off + cnt; };
  const int64_t H = p.H;
  for (int l = 0; l < p.L; ++l)
    if (in(p.offW[l], H * (l == 0 ? PINN_N_IN : H)) || in(p.offb[l], H)) return false;
  return !(in(p.offWp, H) || in(p.offbp, 1) || in(p.offWv0, (H / 2) * H) || in(p.offbv0, H / 2) ||
           in(p.offWv1, (H / 4) * (H / 2)) || in(p.offbv1, H / 4) || in(p.offWv2, H / 4) || in(p.offbv2, 1));
}

inline int validate_net(const pinn_net_t* net) {
  if (!net) return PINN_E_ARG;
  if (net->n_in != PINN_N_IN) return PINN_E_SHAPE;
  if (net->n_hidden < 1 || net->n_hidden > PINN_MAX_HIDDEN) return PINN_E_SHAPE;
  int H = net->width;
  if (!(H == 32 || H == 64 || H == 128 || H == 256)) return PINN_E_SHAPE;
  for (int l = 0; l < net->n_hidden; ++l)
    if (!net->W[l] || !net->b[l]) return PINN_E_ARG;
  if (!net->Wp || !net->bp || !net->Wv0 || !net->bv0 || !net->Wv1 || !net->bv1 || !net->Wv2 ||
      !net->bv2)
    return PINN_E_ARG;
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start (and run its prologue) while its predecessor in the stream is still running; `griddep_wait` blocks until the
// predecessor grid has completed and its memory is visible, `griddep_launch` lets the successor start launching.  Both
// are no-ops in a kernel launched the ordinary way.
PINN_D void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PINN_D void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Launch helper: <<<grid, block, smem, st>>> with (pdl = true) or without the programmatic-serialization attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

struct AdamHyper {
  double lr0, gamma, grad_scale;
  int64_t step_size;
};

// torch/optim/adam.py (_single_tensor_adam): exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2*v + (1-b2) g^2;
// step_size = lr / (1-b1^t); denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= step_size * m/denom.
// The per-step constants are split off so a kernel can compute them once per CTA.
PINN_D void adam_consts(double lr, int64_t t, float& step, float& sqrt_bc2) {
  const double bc1 = 1.0 - pow(0.9, static_cast<double>(t));
  const double bc2 = 1.0 - pow(0.999, static_cast<double>(t));
  step = static_cast<float>(lr / bc1);
  sqrt_bc2 = static_cast<float>(sqrt(bc2));
}
PINN_D void adam_apply(float& p, float g, float& m, float& v, float step, float sqrt_bc2, float lo, float hi, bool clamp) {
  // explicit roundings / fusions: every kernel that inlines this produces the same bits (the compiler's own
  // contraction choices depend on the surrounding code)
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  m = __fmaf_rn(__fsub_rn(g, m), 1.0f - b1, m);
  v = __fmaf_rn(v, b2, __fmul_rn(__fmul_rn(1.0f - b2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), sqrt_bc2), eps);
  float q = __fmaf_rn(-step, __fdiv_rn(m, denom), p);
  if (clamp) q = fminf(fmaxf(q, lo), hi);
  p = q;
}
PINN_D void adam_update(float& p, float g, float& m, float& v, double lr, int64_t t, float lo, float hi,
                        bool clamp) {
  float step, sqrt_bc2;
  adam_consts(lr, t, step, sqrt_bc2);
  adam_apply(p, g, m, v, step, sqrt_bc2, lo, hi, clamp);
}

// Optimiser state handed to a gradient-reduce kernel that applies Adam + StepLR in the same launch
// (params == nullptr: plain reduce).  step_counter as in pinn_adam_step: [0] steps so far, [1] ticket.
struct FusedAdam {
  float* params; float* m; float* v;
  int64_t* step_counter;
  AdamHyper h;
  // one-kernel 64-wide backward (mlp_tc_fused.cuh): the optimiser launch also rewrites the UMMA weight images of the
  // entries it updates, so the next step of the same call needs no image launch (`images_valid`: this step's are in place)
  unsigned char* images; float wscale; int images_valid;
  // data-parallel step: the gradient bucket is exchanged INSIDE the reduce launch over NVLink peer memory (see
  // grad_reduce2_kernel); peers == nullptr: single GPU
  const unsigned long long* peers; int rank, world; unsigned int tag;
};
// Symmetric buffer of the data-parallel reduce (32-bit words): [world x lines flag words, padded to 64] then, per slot
// (step parity) and source rank, one copy of the gradient bucket.
PINN_HD int64_t dp_lines(int64_t total) { return (total + 31) / 32; }
PINN_HD int64_t dp_flag_words(int64_t total, int world) { return (static_cast<int64_t>(world) * dp_lines(total) + 63) / 64 * 64; }
PINN_HD int64_t dp_bucket_words(int64_t total, int world) { return dp_flag_words(total, world) + 2 * static_cast<int64_t>(world) * total; }

}  // namespace pinn
